#!/usr/bin/env python
"""Build recipe for the *unmodified* reference rasterizer (test infrastructure, not product).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
baseline legs may use it.

What it does
------------
Compiles the reference's own sources *where they lie* under
``/root/reference/submodules/diff-gaussian-rasterization`` (never copied into
the repo) with plain ``nvcc``/``g++`` commands — not the reference's
``setup.py``/CMake — for ``compute_100/sm_100`` with exactly the flag set torch's
``CUDAExtension`` would use (no fast-math), and links
them into ``oracle/_ref/diff_gaussian_rasterization/_C*.so``.  The reference's
Python binding (``diff_gaussian_rasterization/__init__.py``) is byte-compiled
(``py_compile``) to ``binding_bytecode.bin`` beside it — a build artefact, no reference
source is copied — so the unmodified binding imports under its own package.
``oracle/_ref`` is git-ignored (build output) but travels to the GPU box.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/submodules/diff-gaussian-rasterization"
OUT = os.path.join(HERE, "_ref")
PKG = os.path.join(OUT, "diff_gaussian_rasterization")
OBJ = os.path.join(OUT, "obj")

CU_SOURCES = [
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "rasterize_points.cu",
]
CPP_SOURCES = ["ext.cpp"]
BINDING = "binding_bytecode.bin"   # byte-compiled reference __init__.py (not named *.pyc: those are not shipped to the GPU box)


def so_path():
    return os.path.join(PKG, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def available():
    return os.path.exists(so_path()) and os.path.exists(os.path.join(PKG, BINDING))


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=True):
    if available() and not force:
        return so_path()
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s not present (GPU box?) and no prebuilt oracle/_ref" % REF)
    import torch  # noqa: F401  (for include paths)
    from torch.utils import cpp_extension as ce

    os.makedirs(PKG, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    incs = ce.include_paths("cuda") + [sysconfig.get_paths()["include"], os.path.join(REF, "third_party/glm")]
    inc_flags = [f"-I{p}" for p in incs]
    defs = [
        "-DTORCH_EXTENSION_NAME=_C",
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=1",
    ]
    nvcc_flags = (
        ["-c", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
         "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
         "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
         "-gencode=arch=compute_100,code=sm_100", "-lineinfo",
         "--compiler-options", "-fPIC"]
        + defs + inc_flags
    )
    cxx_flags = ["-c", "-O3", "-std=c++17", "-fPIC"] + defs + inc_flags

    jobs = []
    objs = []
    for s in CU_SOURCES:
        o = os.path.join(OBJ, s.replace("/", "_") + ".o")
        objs.append(o)
        jobs.append(["nvcc"] + nvcc_flags + [os.path.join(REF, s), "-o", o])
    for s in CPP_SOURCES:
        o = os.path.join(OBJ, s.replace("/", "_") + ".o")
        objs.append(o)
        jobs.append(["g++"] + cxx_flags + [os.path.join(REF, s), "-o", o])
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out.strip():
                print(out)
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = (["g++", "-shared"] + objs + ["-o", so_path(),
            f"-L{torch_lib}", "-L/usr/local/cuda/lib64",
            "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart",
            f"-Wl,-rpath,{torch_lib}"])
    _run(link)
    # the reference's Python binding, byte-compiled next to the built module (build output, git-ignored)
    import py_compile
    py_compile.compile(os.path.join(REF, "diff_gaussian_rasterization/__init__.py"), cfile=os.path.join(PKG, BINDING),
                       doraise=True)
    stale = os.path.join(PKG, "__init__.py")
    if os.path.exists(stale):
        os.remove(stale)
    shutil.rmtree(OBJ, ignore_errors=True)
    return so_path()


def load():
    """Import the compiled reference as module object (without polluting ``sys.modules['diff_gaussian_rasterization']``
    which is the name of the product's drop-in package too)."""
    import importlib.util
    if not available():
        raise ImportError("oracle/_ref not built")
    name = "_gs2m_reference_dgr"
    if name in sys.modules:
        return sys.modules[name]
    import importlib.machinery
    import torch  # noqa: F401
    pyc = os.path.join(PKG, BINDING)
    spec = importlib.util.spec_from_file_location(name, pyc, loader=importlib.machinery.SourcelessFileLoader(name, pyc),
                                                  submodule_search_locations=[PKG])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("built", p)
