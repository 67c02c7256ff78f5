"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares, the
ctypes structs match the C layouts, and the Python surface keeps the reference binding's names and error behaviour
(diff_gaussian_rasterization/__init__.py:17-218).  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gs2m_rasterizer.h")


@pytest.fixture(scope="module")
def lib_path():
    sys.path.insert(0, os.path.join(ROOT, "gs-2m_b200"))
    import build as gs2m_build  # gs-2m_b200/build.py
    return gs2m_build.build()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(gs2m_[a-z0-9_]+)\s*\(", src))
    names -= {"gs2m_resize_fn"}
    return sorted(names)


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for required in ("gs2m_rasterize_forward", "gs2m_rasterize_backward", "gs2m_mark_visible", "gs2m_state_view_get",
                     "gs2m_sort_pairs_u64", "gs2m_inclusive_sum_u32", "gs2m_last_error", "gs2m_abi_version"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), "symbol %s declared in the header but missing from the library" % name
    lib.gs2m_abi_version.restype = ctypes.c_int
    assert lib.gs2m_abi_version() == 3


def test_ctypes_binding_covers_the_header(lib_path):
    from diff_gaussian_rasterization import _native
    bound = {n for n, _, _ in _native.EXPORTS}
    assert bound == set(declared_functions())
    assert _native.load() is not None


def test_struct_layouts_match_c(lib_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from diff_gaussian_rasterization import _native
    structs = {"gs2m_forward_args": _native.ForwardArgs, "gs2m_backward_args": _native.BackwardArgs,
               "gs2m_adam_group": _native.AdamGroup,
               "gs2m_state_view": _native.StateView}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gs2m_rasterizer.h"', "int main(void){"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "layout.c")
        open(c, "w").write("\n".join(lines))
        exe = os.path.join(d, "layout")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, "%s.%s" % (cname, fname)


def test_arena_size_queries(lib_path):
    from diff_gaussian_rasterization import _native
    lib = _native.load()
    assert lib.gs2m_geometry_bytes(0) >= 128
    g1, g2 = lib.gs2m_geometry_bytes(1000), lib.gs2m_geometry_bytes(2000)
    assert g2 > g1 > 1000 * (4 + 16 + 16 + 16 + 24 + 4 + 4 + 4 + 96)
    assert lib.gs2m_image_bytes(1959, 1090) >= 1959 * 1090 * 8 + 123 * 69 * 8
    assert lib.gs2m_binning_bytes(1 << 20) >= (1 << 20) * 24
    assert lib.gs2m_sort_temp_bytes(1) > 0 and lib.gs2m_scan_temp_bytes(1) > 0


def test_python_surface_and_error_behaviour(lib_path):
    import diff_gaussian_rasterization as dgr
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "feature_count")
    s = dgr.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                          torch.zeros(3), False, 1)
    rast = dgr.GaussianRasterizer(raster_settings=s)
    assert isinstance(rast, torch.nn.Module) and hasattr(rast, "markVisible")
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        rast(x, x, x[:, :1], shs=None, colors_precomp=None, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        rast(x, x, x[:, :1], shs=torch.zeros(4, 1, 3), colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="precomputed 3D covariance"):
        rast(x, x, x[:, :1], colors_precomp=x, scales=x, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))
    # no CPU fallback: a CPU tensor is an error, not a slow path
    with pytest.raises(RuntimeError, match="no CPU path"):
        rast(x, x, x[:, :1], colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        rast(torch.zeros(4, 2), x, x[:, :1], colors_precomp=x, scales=x, rotations=torch.zeros(4, 4))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from diff_gaussian_rasterization import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _native.load()


def test_capacity_hint_is_quantized():
    """The automatic instance-capacity hint is rounded up to a small set of sizes (8-16 per octave, >= 64 Ki apart): arenas whose
    size keeps changing miss torch's caching allocator and cost a cudaMalloc in the middle of a step."""
    import diff_gaussian_rasterization as dgr
    q = dgr._quantize_capacity
    assert q(0) == 0 and q(-5) == 0 and q(1) == 65536
    seen = set()
    for c in range(6_500_000, 13_000_000, 9_973):
        v = q(c)
        assert c <= v <= c + max(c // 8, 65536) and q(v) == v
        seen.add(v)
    assert len(seen) <= 12                       # one octave of counts -> a handful of arena sizes
    assert q((1 << 30) + 5) == (1 << 30) - 1 and all(q(c) <= q(c + 1) for c in range(1 << 20, (1 << 20) + 70000, 997))
