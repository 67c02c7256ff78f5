#!/usr/bin/env python
"""Builds the C-ABI CUDA library for sm_100a in-tree: gs-2m_b200/lib/libgs2m_rasterizer.so

Plain nvcc, no torch / pybind in the library (seconds per file).  No fast-math: the geometry and per-pair alpha
arithmetic must round exactly like the reference build.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libgs2m_rasterizer.so")
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "binning_depthfirst.cu", "footprint_masks.cu", "blend_fwd.cu", "blend_bwd.cu", "preprocess_bwd.cu", "feature_pack.cu", "postblend.cu", "photometric_loss.cu", "adam.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False, extra_flags=(), lib_path=None, obj_subdir="build"):
    global LIB
    if lib_path:
        LIB = lib_path
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    obj_dir = os.path.join(HERE, obj_subdir)
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "gs2m_rasterizer.h"))
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(SRC, s)
        obj = os.path.join(obj_dir, s + ".o")
        objs.append(obj)
        if force or _newer(src, obj) or any(_newer(h, obj) for h in headers):
            jobs.append(["nvcc", "-c"] + NVCC_FLAGS + list(extra_flags) + [src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stdout))
        return r.stdout

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            logs = list(ex.map(run, jobs))
        if verbose:
            print("\n".join(logs))
        with open(os.path.join(obj_dir, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    if jobs or not os.path.exists(LIB):
        run(["nvcc", "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    # experiment builds:  python build.py --variant NAME -DFOO=1 ...  ->  lib/variants/NAME.so
    if "--variant" in sys.argv:
        name = sys.argv[sys.argv.index("--variant") + 1]
        flags = [a for a in sys.argv[1:] if a.startswith("-D")]
        print(build(force=True, extra_flags=flags, lib_path=os.path.join(LIB_DIR, "variants", name + ".so"),
                    obj_subdir=os.path.join("build", "variant_" + name)))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
