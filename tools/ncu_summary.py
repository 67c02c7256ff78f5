#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the per-kernel markdown table kept under profiles/.
Usage: python tools/ncu_summary.py <report.ncu-rep> <out.md> "<title>" "<command>" """
import collections
import csv
import subprocess
import sys

rep, out_md, title, command = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
TS = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}
BS = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}


def val(r, k):
    try:
        return float(r[idx[k]].replace(",", ""))
    except Exception:
        return 0.0


stall_keys = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
seen = collections.OrderedDict()
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("gs2m::<unnamed>::", "")[-60:]
    ms = val(r, "gpu__time_duration.sum") * TS[units[idx["gpu__time_duration.sum"]]]
    rd = val(r, "dram__bytes_read.sum") * BS[units[idx["dram__bytes_read.sum"]]]
    wr = val(r, "dram__bytes_write.sum") * BS[units[idx["dram__bytes_write.sum"]]]
    st = sorted([(val(r, k), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in stall_keys], reverse=True)
    tot = sum(v for v, _ in st) or 1
    sig = name.split("(", 1)[1] if "(" in name else ""
    if "renderCUDA" in short or "preprocessCUDA" in short:      # forward and backward kernels share a name in the reference
        short += " [bwd]" if (sig.startswith("const uint2 *, const unsigned int *, int, int, float, float") or
                               sig.startswith("int, int, int, const float3 *")) else " [fwd]"
    d = seen.get(short)
    if d is None:
        seen[short] = dict(n=1, ms=ms, rd=rd, wr=wr,
                           dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                           issue=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           warps=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                           regs=val(r, "launch__registers_per_thread"), inst=val(r, "smsp__inst_executed.sum"),
                           thr=val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                           l1=val(r, "l1tex__t_sector_hit_rate.pct"), l2=val(r, "lts__t_sector_hit_rate.pct"),
                           l2_gb=val(r, "lts__t_bytes.sum") * BS.get(units[idx["lts__t_bytes.sum"]], 1e-6) / 1e3 if "lts__t_bytes.sum" in idx else 0,
                           smem=val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                           red=val(r, "smsp__inst_executed_op_global_red.sum") + val(r, "smsp__inst_executed_op_global_atom.sum"),
                           stalls=", ".join("%s %.0f%%" % (k, 100 * v / tot) for v, k in st[:5]))
    else:   # same kernel launched again (e.g. one radix pass per digit): accumulate time and traffic
        d["n"] += 1; d["ms"] += ms; d["rd"] += rd; d["wr"] += wr; d["inst"] += val(r, "smsp__inst_executed.sum")
with open(out_md, "w") as f:
    f.write("# %s\n\nCommand (under gpurun): `%s`; report read with `ncu -i ... --page raw --csv` by tools/ncu_summary.py.\n"
            "Times are ncu replays (cold cache, serialised).  `issue` = smsp__issue_active %% of peak, `warps` = sm__warps_active %% of peak, "
            "`thr/inst` = active threads per warp instruction, `RED/ATOM` = global reduction/atomic warp instructions.\n\n" % (title, command))
    f.write("| kernel | launches | ms | DRAM rd MB | DRAM wr MB | DRAM % | L2 traffic GB | issue % | warps % | regs | warp instr | thr/inst | L1 hit % | L2 hit % | smem wavefronts | RED/ATOM instr | top stalls |\n")
    f.write("|" + "---|" * 17 + "\n")
    for k, d in seen.items():
        f.write("| `%s` | %d | %.3f | %.0f | %.0f | %.1f | %.2f | %.1f | %.1f | %d | %.3g | %.1f | %.1f | %.1f | %.3g | %.3g | %s |\n" % (
            k, d["n"], d["ms"], d["rd"], d["wr"], d["dram_pct"], d["l2_gb"], d["issue"], d["warps"], d["regs"], d["inst"], d["thr"],
            d["l1"], d["l2"], d["smem"], d["red"], d["stalls"]))
print(open(out_md).read())
