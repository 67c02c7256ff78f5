#!/usr/bin/env python
"""Low-level evidence for the hot kernels, taken from the built library without a GPU: per-kernel resource usage
(`cuobjdump -res-usage`: registers, stack, shared memory) and a SASS mnemonic census (`cuobjdump -sass`): vector reductions
(REDG.E.ADD.F32x4 ...), MUFU.EX2 / MUFU.RCP, votes / shuffles, local-memory instructions (LDL / STL = spills), and the absence
of TMA / tensor-core instructions the design does not use.  Writes profiles/r2_sass_blend.txt.

Usage: python tools/sass_evidence.py [substring of kernel name ...]   (default: the F = 10 blend kernels and the per-Gaussian kernels)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gs-2m_b200", "lib", "libgs2m_rasterizer.so")
DEFAULT = ["blend_forward_kernelILi10E", "blend_backward_kernelILi10E", "preprocess_forward_kernelILb1E",
           "preprocess_backward_staged_kernelILi0E", "preprocess_backward_views_kernelILb0E", "ranges_and_masks_kernelIjE",
           "dense_fill_kernel", "rs_onesweep_kernelIjE"]
WATCH = ["REDG", "RED.", "ATOMG", "ATOMS", "MUFU.EX2", "MUFU.RCP", "MUFU.RSQ", "MUFU.LG2", "VOTE", "SHFL", "MATCH", "LDL", "STL",
         "LDS", "STS", "LDG", "STG", "FFMA2", "FMUL2", "FFMA", "FMUL", "FADD", "BAR", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "UTCHMMA", "REDUX"]


def sh(cmd):
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout


def main():
    wanted = sys.argv[1:] or DEFAULT
    res = sh(["cuobjdump", "-res-usage", LIB])
    usage = {}
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            usage[name] = line.strip()
            name = None
    sass = sh(["cuobjdump", "-sass", LIB])
    blocks = re.split(r"\n\s*Function : ", sass)
    out = ["# SASS / resource evidence of the hot kernels (sm_100a cubins inside gs-2m_b200/lib/libgs2m_rasterizer.so)",
           "# produced by tools/sass_evidence.py with `cuobjdump -res-usage` and `cuobjdump -sass`; no GPU involved", ""]
    for w in wanted:
        for b in blocks[1:]:
            fn = b.split("\n", 1)[0].strip()
            if w not in fn:
                continue
            demangled = sh(["cu++filt", fn]).strip().replace("(int)", "").replace("(bool)", "").split("(")[0]
            ops = collections.Counter()
            full = collections.Counter()
            n = 0
            for line in b.splitlines():
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
                if not m:
                    continue
                n += 1
                op = m.group(1)
                full[op] += 1
                base = op.split(".")[0]
                for k in WATCH:
                    if (base == k) if k in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2") else op.startswith(k):
                        ops[k] += 1
            out.append("## %s" % demangled)
            out.append("mangled: %s" % fn)
            out.append("resources: %s" % usage.get(fn, "?"))
            out.append("SASS instructions: %d" % n)
            out.append("census: " + ", ".join("%s %d" % (k, ops[k]) for k in WATCH if ops[k]))
            red = {k: v for k, v in full.items() if k.startswith(("RED", "ATOM"))}
            out.append("reductions / atomics by exact mnemonic: %s" % (dict(sorted(red.items())) or "none"))
            out.append("local memory (LDL/STL): %d   TMA (UTMALDG/UTMASTG/LDGSTS): %d   tensor core (HMMA/UTCHMMA): %d" % (
                ops["LDL"] + ops["STL"], ops["UTMALDG"] + ops["UTMASTG"] + ops["LDGSTS"], ops["HMMA"] + ops["UTCHMMA"]))
            out.append("")
    # resource table of every kernel in the library
    rows = []
    for fn, u in usage.items():
        d = dict(kv.split(":") for kv in u.split() if ":" in kv)
        rows.append((sh(["cu++filt", fn]).strip().replace("(int)", "").replace("(bool)", "").split("(")[0]
                     .replace("void gs2m::<unnamed>::", "").replace("void gs2m::", ""),
                     d.get("REG"), d.get("STACK"), d.get("SHARED"), d.get("LOCAL")))
    rows.sort()
    with open(os.path.join(ROOT, "profiles", "r2_kernel_resources.md"), "w") as f:
        f.write("# Registers / stack / static shared memory / local memory of every kernel (`cuobjdump -res-usage` of the built library)\n\n")
        spills = ", ".join("`%s` (%s B)" % (r[0], r[2]) for r in rows if r[2] not in ("0", None)) or "none"
        f.write("Kernels with a stack frame (register spills): " + spills + ".  Dynamic shared memory is not listed here: "
                "`blend_backward_kernel<F>` takes sizeof(WarpSmemB<F>) x 2 warps (18.7 KB per CTA at F = 10), "
                "`preprocess_backward_staged_kernel` 19 KB and `preprocess_backward_views_kernel` 28 KB per 64-thread block.\n\n")
        f.write("| kernel | registers | stack | static smem | local |\n|---|---|---|---|---|\n")
        for r in rows:
            f.write("| `%s` | %s | %s | %s | %s |\n" % r)
    path = os.path.join(ROOT, "profiles", "r2_sass_blend.txt")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print("\n".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
