"""Facade-level drop-in acceptance (VERDICT r1 "missing #1"): the reference's UNMODIFIED render()
(gaussian_renderer/__init__.py:21-175) is driven with real GaussianModel / Camera objects once around the compiled reference
rasterizer and once around this package, for every stage / pipe variant; every entry of the returned dict and every raw
parameter gradient must agree.  A second set of tests needs no reference on the box: the committed facade_*.npz vectors (the
real facade around the reference rasterizer, tests/golden/make_facade_golden.py) against the restated facade of
oracle/pack_reference.py and against the fused caller-side kernels, both around this package's rasterizer."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import build_ref
import golden_io
import pack_reference
import synthetic_scenes as syn

pytestmark = pytest.mark.gpu
ILL = ("_scaling", "_rotation")          # chain through the ill-conditioned conic backward


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def modules():
    import diff_gaussian_rasterization as dgr
    if not (build_ref.available() and build_ref.facade_available()):
        pytest.skip("oracle/_ref (compiled reference + facade bytecode) not present; build with `python oracle/build_ref.py`")
    import facade_harness as fh
    return fh, fh.facade(build_ref.load(), "ref"), fh.facade(dgr, "ours")


@pytest.mark.parametrize("variant", ["plain", "geometry", "geometry_sobel", "material", "material_metallic_sobel",
                                     "geometry_zdepth", "material_python_sh", "geometry_python_cov"])
def test_render_facade_is_identical_around_both_rasterizers(modules, variant):
    fh, facade_ref, facade_ours = modules
    P, W, H = 40_000, 400, 300
    scene = syn.make_scene(P, shell_fraction=0.7)
    cam = syn.make_cameras(3, W, H)[2]
    bg = torch.tensor([0.2, 0.4, 0.1], device="cuda")
    camera = fh.make_camera(cam)
    pc = fh.make_model(scene)
    out_r, g_r, vs_r, weights = fh.run_variant(facade_ref, pc, camera, bg, variant)
    out_r2, g_r2, vs_r2, _ = fh.run_variant(facade_ref, pc, camera, bg, variant, weights)     # the reference's own atomic noise
    out_o, g_o, vs_o, _ = fh.run_variant(facade_ours, pc, camera, bg, variant, weights)
    assert set(out_o) == set(out_r)
    for k in fh.EXACT:
        assert out_o[k].dtype == out_r[k].dtype and torch.equal(out_o[k], out_r[k]), k
    for k in fh.FLOAT_MAPS:
        if out_r.get(k) is None:
            assert out_o.get(k) is None, k
            continue
        assert out_o[k].shape == out_r[k].shape, k
        torch.testing.assert_close(out_o[k], out_r[k], rtol=1e-5, atol=1e-7, msg=lambda m: "%s: %s" % (k, m))
    assert out_o["viewspace_points"].shape == (P, 4)
    noise_vs = _rel(vs_r2, vs_r)
    assert _rel(vs_o, vs_r) <= max(1e-5, 4 * noise_vs), "viewspace_points.grad %.3e (reference noise %.3e)" % (_rel(vs_o, vs_r), noise_vs)
    for n in fh.PARAM_NAMES:
        if g_r[n] is None:
            assert g_o[n] is None or float(g_o[n].abs().max()) == 0.0, n
            continue
        err, noise = _rel(g_o[n], g_r[n]), _rel(g_r2[n], g_r[n])
        floor = 5e-4 if (n in ILL or variant == "geometry_python_cov") else 1e-5
        assert err <= max(floor, 4 * noise), "%s %s: %.3e (reference noise %.3e)" % (variant, n, err, noise)


# ---- golden vectors: no reference needed on the box ----
def _golden_case(path):
    spec = importlib.util.spec_from_file_location("make_facade_golden", os.path.join(golden_io.GOLDEN_DIR, "make_facade_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    name = os.path.basename(path)[:-4]
    c = mk.CASES[name]
    z = np.load(path)
    assert [int(v) for v in z["in_meta"]] == [c["P"], c["W"], c["H"], c["seed"], c["D"]]
    scene, cam, bg = mk.case_inputs(c)
    return c, z, scene, cam, bg


def _check_against_golden(out, grads, vs_grad, z, tag):
    import facade_harness as fh
    # the restated facade runs the reference's eager ops in the reference's order; the fused kernels evaluate the same formulas
    # in one pass with their own rounding (their own parity gate against the fp64 oracle is 2e-5, tests/test_feature_pack.py)
    rtol, atol = (1e-5, 1e-6) if tag == "restated" else (1e-4, 1e-5)
    for k in fh.EXACT:
        assert np.array_equal(out[k].cpu().numpy(), z["out_" + k]), "%s %s" % (tag, k)
    for k in fh.FLOAT_MAPS:
        if "out_" + k not in z.files:
            continue
        a, b = out[k].detach().cpu(), torch.from_numpy(z["out_" + k])
        if k in ("depth_map", "sobel_map"):
            # plane depth divides by (n . ray); a restated / fused evaluation differs in the last bits where that is ~0
            ok = torch.from_numpy(z["out_depth_map"]).abs() < 50.0
            ok = ok.expand_as(b) if k == "depth_map" else ok.expand(3, -1, -1)
            torch.testing.assert_close(a[ok], b[ok], rtol=2e-4, atol=2e-4, msg=lambda m: "%s %s: %s" % (tag, k, m))
        else:
            torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: "%s %s: %s" % (tag, k, m))
    assert _rel(vs_grad.cpu(), torch.from_numpy(z["grad_viewspace_points"])) <= 1e-4, tag
    for n in fh.PARAM_NAMES:
        if "grad" + n not in z.files:
            continue
        err = _rel(grads[n].cpu(), torch.from_numpy(z["grad" + n]))
        assert err <= (2e-3 if n in ILL else 1e-4), "%s %s: %.3e" % (tag, n, err)


@pytest.mark.parametrize("path", golden_io.facade_golden_files(), ids=[os.path.basename(p) for p in golden_io.facade_golden_files()])
@pytest.mark.parametrize("fused", [False, True], ids=["restated-facade", "fused-kernels"])
def test_facade_golden_vectors(path, fused):
    """render() of the reference around the reference rasterizer (committed vectors) vs (a) the restated facade of
    oracle/pack_reference.py and (b) the fused caller-side kernels of this package, both around this package's rasterizer."""
    import diff_gaussian_rasterization as dgr
    import facade_harness as fh
    from diff_gaussian_rasterization import packing
    c, z, scene, cam, bg = _golden_case(path)
    kw = dict(fh.VARIANTS[c["variant"]])
    z_depth = kw.pop("pipe", {}).get("z_depth", False)
    raw = {k: v.cuda().requires_grad_(True) for k, v in syn.raw_parameters(scene).items()}
    shs = scene.shs.cuda().requires_grad_(True)
    cam = syn.camera_to(cam, "cuda")
    tanx, tany = float(z["in_tanfov"][0]), float(z["in_tanfov"][1])
    W, H = c["W"], c["H"]
    if not fused:
        out = pack_reference.render_like(dgr, raw, shs, cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                         tanx, tany, W, H, bg, active_sh_degree=c["D"], z_depth=z_depth, **kw)
    else:
        saved = (pack_reference.activate_and_pack, pack_reference.derive_maps, pack_reference.sobel_normal_map)
        try:     # same composition, the three eager stages replaced by the library's fused kernels
            pack_reference.activate_and_pack = packing.activate_and_pack
            pack_reference.derive_maps = packing.derive_maps
            pack_reference.sobel_normal_map = lambda d, a, b, w, fx, fy, cx, cy: packing.sobel_normal_map(d, a, b, w, fx, fy, cx, cy)
            out = pack_reference.render_like(dgr, raw, shs, cam.world_view_transform, cam.full_proj_transform,
                                             cam.camera_center, tanx, tany, W, H, bg, active_sh_degree=c["D"], z_depth=z_depth, **kw)
        finally:
            pack_reference.activate_and_pack, pack_reference.derive_maps, pack_reference.sobel_normal_map = saved
    weights = {k[2:]: torch.from_numpy(z[k]).cuda() for k in z.files if k.startswith("w_")}
    fh.scalar_loss(out, weights).backward()
    grads = {"_xyz": raw["xyz"].grad, "_features_dc": shs.grad[:, :1], "_features_rest": shs.grad[:, 1:],
             "_scaling": raw["scaling"].grad, "_rotation": raw["rotation"].grad, "_opacity": raw["opacity"].grad,
             "_albedo": raw["albedo"].grad, "_roughness": raw["roughness"].grad, "_metallic": raw["metallic"].grad}
    grads = {k: (v if v is not None else torch.zeros(1)) for k, v in grads.items()}
    _check_against_golden(out, grads, out["viewspace_points"].grad, z, "fused" if fused else "restated")
