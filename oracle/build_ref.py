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

The same is done for the Python modules of the render facade (``gaussian_renderer/__init__.py`` = SURVEY 2.1 #9, "the
drop-in's acceptance harness", and the pure-torch helpers it and the caller-side oracles lean on): byte-compiled into
``oracle/_ref/facade/*.bin`` so that the *unmodified* ``render()`` can be driven on the GPU box, once per rasterizer
(``load_facade``).

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

# the render facade and the modules it needs, relative to the reference root: module name -> source file
REF_ROOT = "/root/reference"
FACADE_DIR = os.path.join(OUT, "facade")
FACADE_MODULES = {
    "gaussian_renderer": "gaussian_renderer/__init__.py",
    "utils.sh_utils": "utils/sh_utils.py",
    "utils.normal_utils": "utils/normal_utils.py",
    "utils.general_utils": "utils/general_utils.py",
    "utils.graphics_utils": "utils/graphics_utils.py",
    "utils.system_utils": "utils/system_utils.py",
    "utils.loss_utils": "utils/loss_utils.py",
    "scene.gaussian_model": "scene/gaussian_model.py",
    "scene.cameras": "scene/cameras.py",
}


def so_path():
    return os.path.join(PKG, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def available():
    return os.path.exists(so_path()) and os.path.exists(os.path.join(PKG, BINDING))


def facade_available():
    return all(os.path.exists(os.path.join(FACADE_DIR, n + ".bin")) for n in FACADE_MODULES)


def build_facade(force=False):
    """Byte-compile the facade modules from where they lie (no source is copied)."""
    if facade_available() and not force:
        return FACADE_DIR
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree %s not present and no prebuilt oracle/_ref/facade" % REF_ROOT)
    import py_compile
    os.makedirs(FACADE_DIR, exist_ok=True)
    for name, rel in FACADE_MODULES.items():
        py_compile.compile(os.path.join(REF_ROOT, rel), cfile=os.path.join(FACADE_DIR, name + ".bin"), doraise=True)
    return FACADE_DIR


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


def _load_bytecode(name, as_name=None, package_path=None):
    import importlib.machinery
    import importlib.util
    as_name = as_name or name
    path = os.path.join(FACADE_DIR, name + ".bin")
    spec = importlib.util.spec_from_file_location(as_name, path, loader=importlib.machinery.SourcelessFileLoader(as_name, path),
                                                  submodule_search_locations=package_path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[as_name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_facade(rasterizer_module, tag):
    """The reference's unmodified ``gaussian_renderer`` module bound to ``rasterizer_module`` (the compiled reference binding or
    the drop-in package) as ``sys.modules['gaussian_renderer_<tag>']``, plus the shared helper modules under their own names
    (``utils.*``, ``scene.gaussian_model``, ``scene.cameras``).  Third-party modules the reference imports at module level but
    the facade never calls (plyfile, simple_knn, matplotlib) are replaced by empty stubs."""
    import types
    if not facade_available():
        raise ImportError("oracle/_ref/facade not built")

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
        return sys.modules[name]

    stub("plyfile", PlyData=object, PlyElement=object)
    stub("simple_knn")
    stub("simple_knn._C", distCUDA2=None)
    for pkg in ("utils", "scene"):          # package shells: the reference's scene/__init__.py (dataset loading, PBR) is not run
        if pkg not in sys.modules or not hasattr(sys.modules[pkg], "__path__"):
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
    stub("utils.image_utils", process_input_image=None, erode=None)      # matplotlib-dependent, unused by render()
    for name in ("utils.sh_utils", "utils.normal_utils", "utils.general_utils", "utils.graphics_utils", "utils.system_utils",
                 "scene.gaussian_model", "scene.cameras"):
        if name not in sys.modules:
            mod = _load_bytecode(name)
            setattr(sys.modules[name.split(".")[0]], name.split(".")[1], mod)
    as_name = "gaussian_renderer_" + tag
    if as_name in sys.modules:
        return sys.modules[as_name]
    saved = sys.modules.get("diff_gaussian_rasterization")
    sys.modules["diff_gaussian_rasterization"] = rasterizer_module
    try:
        return _load_bytecode("gaussian_renderer", as_name=as_name, package_path=[])
    finally:
        if saved is not None:
            sys.modules["diff_gaussian_rasterization"] = saved
        else:
            del sys.modules["diff_gaussian_rasterization"]


class cuda_literals_on_cpu:
    """Context manager for running the reference's pure-torch helpers in a container without a GPU: a few of them create
    their scratch tensors with a literal ``device="cuda"`` (utils/general_utils.py:59,77,96).  While active, ``torch.zeros`` /
    ``torch.ones`` ignore that literal when no CUDA device exists; values are unaffected."""

    def __enter__(self):
        import torch
        self._saved = (torch.zeros, torch.ones)
        if torch.cuda.is_available():
            return self

        def wrap(fn):
            def inner(*a, **k):
                if str(k.get("device", "")).startswith("cuda"):
                    k = dict(k, device="cpu")
                return fn(*a, **k)
            return inner
        torch.zeros, torch.ones = wrap(torch.zeros), wrap(torch.ones)
        return self

    def __exit__(self, *exc):
        import torch
        torch.zeros, torch.ones = self._saved
        return False


def load_loss_utils(facade_module):
    """``utils.loss_utils`` of the reference (l1_loss, ssim, the regularisers); it imports ``render`` from ``gaussian_renderer``."""
    if "utils.loss_utils" in sys.modules:
        return sys.modules["utils.loss_utils"]
    saved = sys.modules.get("gaussian_renderer")
    sys.modules["gaussian_renderer"] = facade_module
    try:
        mod = _load_bytecode("utils.loss_utils")
        sys.modules["utils"].loss_utils = mod
        return mod
    finally:
        if saved is not None:
            sys.modules["gaussian_renderer"] = saved
        else:
            del sys.modules["gaussian_renderer"]


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("built", p)
    print("facade bytecode in", build_facade(force="--force" in sys.argv))
