"""GPU parity tests (run with `-m gpu` on the B200 box): the CUDA implementation, called through the C-ABI, against
(1) the compiled reference rasterizer (oracle/_ref) on identical seeded inputs — sort keys, sorted indices, tile
ranges, per-pixel contributor counts, radii, observe and final transmittance BIT-EXACT; rendered channels within 1e-5
relative; gradients within 1e-4 (max|d|/max|ref| per tensor, or 4x the reference's own run-to-run atomic noise where
that is larger: dL/dcov3D, dL/dscale and dL/drot are ill-conditioned and the reference itself moves by up to 3e-4 there) — (2) the committed golden vectors, and (3) the CPU oracle."""
import numpy as np
import pytest
import torch

import helpers
import golden_io
import synthetic_scenes as syn

pytestmark = pytest.mark.gpu

GRAD_NAMES = ["dL_dmeans2D", "dL_dcolor", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscale", "dL_drot",
              "dL_dfeatures"]
RENDER_RTOL, RENDER_ATOL = 1e-5, 1e-7   # north_star: rendered channels within 1e-5 relative
GRAD_TOL = 1e-4                         # north_star: gradients within 1e-4 relative (per-tensor max norm)
ILL_CONDITIONED = ("dL_dcov3D", "dL_dscale", "dL_drot")   # conic backward amplifies input rounding ~1e3
# The gates actually enforced are tighter than north_star's 1e-4 wherever the measurements allow: the six well-conditioned
# tensors measure 1-2e-6 against the reference (its own run-to-run atomic noise), so a 10x regression must fail.  The three
# ill-conditioned ones are gated against the reference's noise floor and, in test_ill_conditioned_gradients_against_fp64,
# against an fp64 evaluation of the same formulas.
WELL_TOL = 1e-5


def grad_gate(name):
    return 5e-4 if name in ILL_CONDITIONED else WELL_TOL



@pytest.fixture(scope="module")
def dgr():
    import diff_gaussian_rasterization as m
    return m


@pytest.fixture(scope="module")
def ref():
    import build_ref
    if not build_ref.available():
        pytest.skip("compiled reference (oracle/_ref) not present; build with `python oracle/build_ref.py`")
    return build_ref.load()


def assert_forward_bit_exact(o, r, P):
    assert o["R"] == r["R"]
    assert torch.equal(o["radii"], r["radii"])
    vis = r["radii"] > 0
    assert torch.equal(o["tiles_touched"], r["tiles_touched"])
    assert torch.equal(o["point_offsets"], r["point_offsets"])
    for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
        assert torch.equal(helpers.bits(o[k][vis]), helpers.bits(r[k][vis])), k
    assert torch.equal(o["clamped"].view(P, 3)[vis], r["clamped"].view(P, 3)[vis])
    assert torch.equal(o["keys_sorted"], r["keys_sorted"])
    assert torch.equal(o["point_list"], r["point_list"])
    assert torch.equal(o["ranges"], r["ranges"])
    assert torch.equal(o["n_contrib"], r["n_contrib"])
    assert torch.equal(helpers.bits(o["final_T"]), helpers.bits(r["final_T"]))
    assert torch.equal(o["observe"], r["observe"])
    torch.testing.assert_close(o["color"], r["color"], rtol=RENDER_RTOL, atol=RENDER_ATOL)
    torch.testing.assert_close(o["buffer"], r["buffer"], rtol=RENDER_RTOL, atol=RENDER_ATOL)


CASES = [
    # P, W, H, F, cam_radius, shell
    pytest.param(20_000, 320, 240, 10, 3.0, 0.7, id="20k-320x240-F10"),
    pytest.param(100_000, 800, 800, 5, 3.0, 0.7, id="config1-100k-800x800-F5"),
    pytest.param(300_000, 800, 600, 5, 3.0, 0.7, id="config2-300k-800x600-F5"),
    pytest.param(500_000, 800, 800, 9, 3.0, 0.7, id="config3-500k-800x800-F9"),
    pytest.param(500_000, 800, 800, 10, 3.0, 0.7, id="config3-500k-800x800-F10"),
    pytest.param(3_000_000, 1959, 1090, 10, 2.2, 0.0, id="config4-3M-1959x1090-F10"),
    pytest.param(6_000_000, 1959, 1090, 10, 2.2, 0.0, id="config5-6M-1959x1090-F10"),
]


@pytest.mark.parametrize("P,W,H,F,rad,shell", CASES)
def test_forward_and_backward_vs_reference(dgr, ref, P, W, H, F, rad, shell):
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=shell, cam_radius=rad)
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    r2 = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)   # second run: the reference's own atomic noise
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    assert_forward_bit_exact(o, r, P)
    for k in GRAD_NAMES:
        err, l2 = helpers.grad_errors(o[k], r[k])
        noise, _ = helpers.grad_errors(r2[k], r[k])
        # dL/dcov3D, dL/dscale, dL/drot: the reference's own two runs differ by 1e-4..4e-4 in the max norm at 3 M
        # Gaussians (one noise sample is itself noisy), so their max-norm floor is 5e-4; the relative L2 norm, which is
        # stable, must still meet 1e-4 for every tensor.
        tol = max(grad_gate(k), 4.0 * noise)
        assert err <= tol, "%s: max|d|/max|ref| %.3e (l2 %.3e) > %.3e (reference self-noise %.3e)" % (k, err, l2, tol, noise)
        assert l2 <= GRAD_TOL, "%s: relative L2 %.3e" % (k, l2)


def _random_cases(n, seed=20261017):
    """Seeded random (P, W, H, F, camera radius, shell fraction, scene seed, view index): ragged image sizes (tile rows / columns
    cut anywhere, images smaller than a tile, one-pixel strips), every feature count, cameras from inside the cloud to far away."""
    import random
    rng = random.Random(seed)
    cases = []
    for k in range(n):
        P = rng.choice([1, 7, 200, 3_000, 25_000, 60_000])
        W, H = rng.choice([(1, 1), (5, 300), (300, 5), (17, 33), (251, 190), (640, 361), (1023, 65)])
        F = rng.randint(0, 10)
        cases.append(pytest.param(P, W, H, F, rng.choice([0.3, 1.5, 3.0, 8.0]), rng.choice([0.0, 0.6, 1.0]),
                                  rng.randint(1, 10_000), rng.randint(0, 3), id="rand%02d-P%d-%dx%d-F%d" % (k, P, W, H, F)))
    return cases


@pytest.mark.parametrize("P,W,H,F,rad,shell,scene_seed,view", _random_cases(16))
def test_random_shapes_vs_reference(dgr, ref, P, W, H, F, rad, shell, scene_seed, view):
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=shell, cam_radius=rad, view=view, n_views=4,
                                                  scene_seed=scene_seed)
    if F == 0:
        gb = torch.zeros_like(gb)
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    assert_forward_bit_exact(o, r, P)
    r2 = None
    for k in GRAD_NAMES:
        if float(r[k].abs().max()) == 0.0:
            assert float(o[k].abs().max()) == 0.0, k
            continue
        err, l2 = helpers.grad_errors(o[k], r[k])
        # with a handful of Gaussians the max norm is one Gaussian's own cancellation-heavy pixel sum (signed terms of either
        # implementation's fp32 summation order, and the order of the reference's per-pixel atomics changes from run to run):
        # north_star's 1e-4 there, widened to 4 x the reference's own run-to-run distance when that is larger; the tight
        # gate from 1000 Gaussians on
        well = WELL_TOL
        if P < 1000:
            if r2 is None:
                r2 = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
            well = max(GRAD_TOL, 4.0 * helpers.grad_errors(r2[k], r[k])[0])
        assert err <= (2e-3 if k in ILL_CONDITIONED else well), "%s: max %.3e l2 %.3e (gate %.3e)" % (k, err, l2, well)


def test_ill_conditioned_gradients_against_fp64(dgr, ref):
    """dL/dcov3D, dL/dscale, dL/drot come out of the conic backward (backward.cu:153-281), which amplifies the rounding of its
    inputs by ~1e3; the reference itself moves by 1e-4..4e-4 between two runs there.  "Inside the reference's noise" is grounded
    here against a ground truth: the fp64 CPU oracle evaluates the same formulas on the same inputs and the same per-pixel
    forward state (final_T, n_contrib are bit-identical between the two implementations).  Ours must be as close to it as the
    reference is (several reference runs give its spread)."""
    import cpu_rasterizer as cr
    P, W, H, F = 12_000, 256, 192, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.7)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    refs = [helpers.run_reference(ref, scene, cam, feats, F, gc, gb) for _ in range(3)]
    assert torch.equal(o["n_contrib"], refs[0]["n_contrib"]) and torch.equal(helpers.bits(o["final_T"]), helpers.bits(refs[0]["final_T"]))
    settings = syn.raster_settings_for(syn.camera_to(cam, "cpu"), F, golden_io.Settings)
    orc = cr.CpuRasterizer(torch.float64)
    s = syn.scene_to(scene, "cpu")
    orc.forward(s.means3D, s.shs, None, s.opacities, s.scales, s.rotations, None, feats.cpu(), settings)
    truth = orc.backward(gc.cpu(), gb.cpu(), final_T=o["final_T"].cpu(), n_contrib=o["n_contrib"].cpu())
    for k in GRAD_NAMES:
        mine = helpers.grad_errors(o[k].cpu(), truth[k])[0]
        theirs = [helpers.grad_errors(r[k].cpu(), truth[k])[0] for r in refs]
        # as good as the reference: within its own spread of distances to the truth (25 % slack for the three-sample estimate)
        assert mine <= 1.25 * max(theirs) + 1e-6, "%s: ours %.3e vs reference %s from the fp64 truth" % (k, mine, ["%.3e" % t for t in theirs])
        if k not in ILL_CONDITIONED:
            # (the oracle also evaluates the FORWARD in fp64, so this distance contains the fp32 rounding of the projected
            # means / conics both GPU implementations share, ~1e-4 of the largest gradient; the gate above is the sharp one)
            assert mine <= 5e-4, "%s: %.3e from the fp64 truth" % (k, mine)


@pytest.mark.parametrize("F", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10])
def test_every_feature_count(dgr, ref, F):
    P, W, H = 6000, 200, 136
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    if F == 0:
        gb = torch.zeros_like(gb)
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    assert_forward_bit_exact(o, r, P)
    assert float(o["buffer"][F:].abs().max()) == 0.0 if F < 10 else True
    for k in GRAD_NAMES:
        err, _ = helpers.grad_errors(o[k], r[k])
        assert err <= grad_gate(k), "%s F=%d err %.3e" % (k, F, err)
    assert float(o["dL_dfeatures"][:, F:].abs().max()) == 0.0 if F < 10 else True


@pytest.mark.parametrize("deg,M", [(0, 16), (1, 16), (2, 16), (3, 16), (1, 4), (0, 1), (2, 9)])
def test_sh_degrees_and_coefficient_counts(dgr, ref, deg, M):
    P, W, H, F = 4000, 160, 120, 5
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    scene = scene._replace(shs=scene.shs[:, :M].contiguous())
    bg = torch.tensor([0.3, 0.1, 0.9], device="cuda")
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb, sh_degree=deg, bg=bg)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb, sh_degree=deg, bg=bg)
    assert_forward_bit_exact(o, r, P)
    for k in GRAD_NAMES:
        err, _ = helpers.grad_errors(o[k], r[k])
        assert err <= grad_gate(k), "%s deg=%d M=%d err %.3e" % (k, deg, M, err)


@pytest.mark.parametrize("path,P,W,H", [("depthfirst", 20_000, 321, 200), ("depthfirst-exact", 3_000, 97, 50), ("sort64", 20_000, 321, 200),
                                        ("depthfirst", 150_000, 640, 361)])
def test_warp_block_lists_are_the_stable_compaction_of_the_tile_lists(dgr, path, P, W, H, monkeypatch):
    """footprint_masks.cu compacts the instance list once per warp-block position w by mask bit w (dense_gid / dense_pos /
    block_ranges); the blend kernels walk those lists instead of filtering the tile's list.  Checked against the definition."""
    monkeypatch.setenv("GS2M_BINNING", "sort64" if path == "sort64" else "depthfirst")
    if path == "depthfirst-exact":
        monkeypatch.setenv("GS2M_EXACT_BINNING", "1")
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, 3, shell=0.7)
    for _ in range(2):                       # second call: speculative capacity on the default path
        o = helpers.run_ours(dgr, scene, cam, feats, 3)
    R = o["R"]
    lists, masks, ranges = o["point_list"][:R].cpu(), o["masks"][:R].cpu().to(torch.int32), o["ranges"].cpu()
    tile_of = torch.repeat_interleave(torch.arange(ranges.shape[0]), (ranges[:, 1] - ranges[:, 0]).to(torch.int64))
    assert tile_of.numel() == R
    pos = torch.arange(R) - ranges[tile_of, 0].to(torch.int64)
    br, dg, dp = o["block_ranges"].cpu().to(torch.int64), o["dense_gid"].cpu(), o["dense_pos"].cpu()
    for w in range(8):
        sel = ((masks >> w) & 1).bool()
        n = int(sel.sum())
        assert torch.equal(dg[w, :n], lists[sel]) and torch.equal(dp[w, :n].to(torch.int64), pos[sel]), w
        counts = torch.zeros(ranges.shape[0], dtype=torch.int64).index_add_(0, tile_of[sel], torch.ones(n, dtype=torch.int64))
        ends = torch.cumsum(counts, 0)
        nonempty = ranges[:, 1] > ranges[:, 0]
        assert torch.equal(br[nonempty, w, 1], ends[nonempty]) and torch.equal(br[nonempty, w, 0], (ends - counts)[nonempty]), w
        assert int((br[~nonempty, w, 1] - br[~nonempty, w, 0]).abs().sum()) == 0
    # the forward's last-contributor index in list coordinates maps back to the reference's through dense_pos
    ncd, nc = o["n_contrib_dense"].cpu().to(torch.int64), o["n_contrib"].cpu().to(torch.int64)
    assert torch.equal(ncd == 0, nc == 0)


@pytest.mark.parametrize("path", ["depthfirst", "depthfirst-exact", "sort64"])
def test_all_binning_paths_are_bit_exact(dgr, ref, path, monkeypatch):
    """GS2M_BINNING selects depth sort + emission in depth order + tile sort (default; speculative from the second call on,
    or exact with GS2M_EXACT_BINNING=1) or duplicate + 64-bit onesweep sort; keys, lists and ranges must be bit-identical
    to the reference with each (includes huge rectangles, a one-tile image = 1 tile-key bit, a 256-tile image = one digit
    pass, and depth ties)."""
    monkeypatch.setenv("GS2M_BINNING", path.split("-")[0])
    if path.endswith("exact"):
        monkeypatch.setenv("GS2M_EXACT_BINNING", "1")
    for P, W, H, F, scale_big in ((60_000, 640, 400, 10, 1.0), (4_000, 330, 210, 5, 60.0), (900, 16, 16, 3, 1.0),
                                  (20_000, 256, 256, 2, 1.0)):
        scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
        if scale_big != 1.0:
            sc = scene.scales.clone()
            sc[:50] *= scale_big
            scene = scene._replace(scales=sc.contiguous())
        if P == 20_000:   # exact depth ties between different Gaussians: the order inside a tile must fall back to the index
            m = scene.means3D.clone()
            m[1::2] = m[0::2]
            scene = scene._replace(means3D=m.contiguous())
        r = helpers.run_reference(ref, scene, cam, feats, F)
        for _ in range(2):      # the second call of the default path runs speculatively (instance-count hint from the first)
            o = helpers.run_ours(dgr, scene, cam, feats, F)
            assert_forward_bit_exact(o, r, P)


def test_speculative_forward_modes(dgr, ref):
    """The forward without a host round trip: binning arena sized from a capacity, count-dependent kernels on
    capacity-sized grids that read R / V from device memory.  Same bits as the exact mode and the reference for any sufficient
    capacity; a capacity that is too small is reported (explicit capacity) or transparently re-run (automatic mode);
    no_wait never touches the host and leaves the verdict in bin_info."""
    from diff_gaussian_rasterization import _native
    P, W, H, F = 30_000, 400, 300, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    settings = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    args = (scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, feats, settings)
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    R = int(r["R"])
    exact = dgr.forward_raw(*args, capacity=0)
    assert exact[4].num_rendered == R and exact[4].capacity == 0
    g_exact = dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, exact[1],
                               settings, exact[4])
    for cap in (R, R + 1, 2 * R + 12345, 40 * R):
        color, radii, observe, buffer, state = dgr.forward_raw(*args, capacity=cap)
        assert state.num_rendered == R and state.capacity == cap
        sv = dgr.state_view(P, settings, state)
        assert sv["bin_info"][:5].tolist() == [R, int((r["radii"] > 0).sum()), 0, R, int((r["radii"] > 0).sum())]
        assert torch.equal(sv["keys_sorted"], r["keys_sorted"]) and torch.equal(sv["point_list"], r["point_list"])
        assert torch.equal(sv["ranges"], r["ranges"]) and torch.equal(sv["n_contrib"], r["n_contrib"])
        assert torch.equal(helpers.bits(color), helpers.bits(exact[0])) and torch.equal(helpers.bits(buffer), helpers.bits(exact[3]))
        assert torch.equal(observe, r["observe"])
        g = dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii,
                             settings, state)
        for k in GRAD_NAMES:    # two runs of the same kernels: only the order of the vector reductions differs
            err, _ = helpers.grad_errors(g[k], g_exact[k])
            assert err <= (5e-4 if k in ILL_CONDITIONED else 1e-5), "%s capacity %d: %.3e" % (k, cap, err)
    # explicit capacity that is too small: checked error, nothing written past the arena
    with pytest.raises(dgr.RasterizerError) as ei:
        dgr.forward_raw(*args, capacity=R - 1)
    assert ei.value.code == _native.ERR_CAPACITY and _native.load().gs2m_last_instance_count() == R
    # automatic mode recovers by itself
    key = (torch.cuda.current_device(), W, H)
    dgr._R_HINT[key] = 10
    color, radii, observe, buffer, state = dgr.forward_raw(*args)
    assert state.num_rendered == R and state.capacity == 0 and dgr._R_HINT[key] == R
    assert torch.equal(helpers.bits(color), helpers.bits(exact[0]))
    color, radii, observe, buffer, state = dgr.forward_raw(*args)
    assert state.num_rendered == R and state.capacity >= R and torch.equal(helpers.bits(color), helpers.bits(exact[0]))
    # no_wait: capturable, verdict on the device
    color, radii, observe, buffer, state = dgr.forward_raw(*args, capacity=R + 1000, no_wait=True)
    sv = dgr.state_view(P, settings, state)
    assert state.num_rendered == R + 1000 and sv["bin_info"][:4].tolist() == [R, int((radii > 0).sum()), 0, R]
    assert torch.equal(helpers.bits(color), helpers.bits(exact[0]))
    color, radii, observe, buffer, state = dgr.forward_raw(*args, capacity=R // 2, no_wait=True)
    sv = dgr.state_view(P, settings, state)
    assert sv["bin_info"][0].item() == R and sv["bin_info"][2].item() == _native.BIN_OVERFLOW and sv["bin_info"][3].item() == 0
    assert float(buffer.abs().max()) == 0.0      # discarded result: nothing was blended


def test_forward_under_cuda_graph(dgr, ref):
    """no_wait forward + backward captured in a CUDA graph and replayed with new camera matrices in the static buffers."""
    P, W, H, F = 20_000, 320, 240, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6, n_views=2)
    _, cam1, feats1, _, _ = helpers.make_view(P, W, H, F, shell=0.6, view=1, n_views=2)
    st_cam = syn.Camera(H, W, cam.tanfovx, cam.tanfovy, cam.world_view_transform.clone(), cam.full_proj_transform.clone(),
                        cam.camera_center.clone())
    st_feats = feats.clone()
    settings = syn.raster_settings_for(st_cam, F, dgr.GaussianRasterizationSettings)
    args = (scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, st_feats, settings)
    cap = 4 * int(dgr.forward_raw(*args, capacity=0)[4].num_rendered)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):      # warm-up on the capture stream (one-time function attributes, allocator pools)
        out = dgr.forward_raw(*args, capacity=cap, no_wait=True)
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, st_feats, out[1], settings, out[4])
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        color, radii, observe, buffer, state = dgr.forward_raw(*args, capacity=cap, no_wait=True)
        g = dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, st_feats, radii,
                             settings, state)
    for c, f in ((cam1, feats1), (cam, feats)):
        st_cam.world_view_transform.copy_(c.world_view_transform)
        st_cam.full_proj_transform.copy_(c.full_proj_transform)
        st_cam.camera_center.copy_(c.camera_center)
        st_feats.copy_(f)
        graph.replay()
        torch.cuda.synchronize()
        r = helpers.run_reference(ref, scene, c, f, F, gc, gb)
        assert torch.equal(radii, r["radii"]) and torch.equal(observe, r["observe"])
        torch.testing.assert_close(color, r["color"], rtol=RENDER_RTOL, atol=RENDER_ATOL)
        torch.testing.assert_close(buffer, r["buffer"], rtol=RENDER_RTOL, atol=RENDER_ATOL)
        for k in ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dsh", "dL_dfeatures"):
            err, _ = helpers.grad_errors(g[k], r[k])
            assert err <= 1e-5, "%s %.3e" % (k, err)


def test_backward_twice_and_inference_forward(dgr):
    """The forward zeroes the backward accumulator rows of the visible Gaussians; a second backward over the same state, or a
    backward after an inference-only forward, must clear it again (grad_acc_dirty)."""
    P, W, H, F = 9000, 200, 150, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    settings = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    args = (scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, feats, settings)
    bargs = (scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats)
    color, radii, observe, buffer, state = dgr.forward_raw(*args)
    g1 = {k: v.clone() for k, v in dgr.backward_raw(gc, gb, *bargs, radii, settings, state).items()}
    g2 = dgr.backward_raw(gc, gb, *bargs, radii, settings, state)
    color, radii, observe, buffer, state_inf = dgr.forward_raw(*args, for_backward=False)
    g3 = dgr.backward_raw(gc, gb, *bargs, radii, settings, state_inf)
    for k in GRAD_NAMES:
        for other in (g2, g3):
            err, _ = helpers.grad_errors(other[k], g1[k])
            assert err <= (5e-4 if k in ILL_CONDITIONED else 1e-5), "%s %.3e" % (k, err)


def test_prefiltered_is_a_checked_error(dgr):
    """The reference traps the device when `prefiltered` is set and a Gaussian fails the near-plane test
    (auxiliary.h:154-160); here it is a checked error, in exact and in speculative mode."""
    from diff_gaussian_rasterization import _native
    scene, cam, feats, _, _ = helpers.make_view(20_000, 160, 120, 5, cam_radius=0.5)      # camera inside the cloud
    settings = syn.raster_settings_for(cam, 5, dgr.GaussianRasterizationSettings)
    args = (scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, feats)
    R = dgr.forward_raw(*args, settings, capacity=0)[4].num_rendered
    for cap in (0, R + 100):
        with pytest.raises(dgr.RasterizerError) as ei:
            dgr.forward_raw(*args, settings._replace(prefiltered=True), capacity=cap)
        assert ei.value.code == -6
    # everything in front of the camera: the flag is harmless
    scene2, cam2, feats2, _, _ = helpers.make_view(5_000, 160, 120, 5, cam_radius=4.0)
    st2 = syn.raster_settings_for(cam2, 5, dgr.GaussianRasterizationSettings)._replace(prefiltered=True)
    dgr.forward_raw(scene2.means3D, scene2.shs, None, scene2.opacities, scene2.scales, scene2.rotations, None, feats2, st2)


def test_second_device_in_one_process(dgr):
    """Function attributes (dynamic shared-memory opt-in) and the host read-back slot are per device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        scene, cam, feats, gc, gb = helpers.make_view(8000, 200, 150, 10, shell=0.6, device=dev)
        o = helpers.run_ours(dgr, scene, cam, feats, 10, gc, gb)
        outs.append({k: o[k].cpu() for k in ("color", "radii", "dL_dmeans3D", "dL_dsh")})
    for o in outs[1:]:
        assert torch.equal(o["radii"], outs[0]["radii"]) and torch.equal(o["color"], outs[0]["color"])
        assert helpers.grad_errors(o["dL_dsh"], outs[0]["dL_dsh"])[0] <= 1e-5


def test_precomputed_colors_and_covariances(dgr, ref):
    """colors_precomp / cov3D_precomp inputs (binding :186-203; forward.cu:194,227; backward.cu:396,408)."""
    P, W, H, F = 5000, 176, 144, 9
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    gen = torch.Generator().manual_seed(5)
    colors = torch.rand(P, 3, generator=gen).cuda()
    import cpu_rasterizer as cr
    cov = cr._cov3d(scene.scales.cpu(), 1.0, scene.rotations.cpu()).cuda().contiguous()
    settings = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    empty = torch.Tensor([])
    bg = settings.bg
    Rr, color, radii, observe, buffer, geom, binning, img = ref._C.rasterize_gaussians(
        bg, scene.means3D, colors, scene.opacities, empty, empty, 1.0, cov, feats, cam.world_view_transform,
        cam.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, empty, 3, cam.camera_center, False, F)
    g_ref = ref._C.rasterize_gaussians_backward(
        bg, scene.means3D, radii, buffer, colors, empty, empty, 1.0, cov, feats, cam.world_view_transform,
        cam.full_proj_transform, cam.tanfovx, cam.tanfovy, gc, gb, empty, 3, cam.camera_center, geom, Rr, binning, img, F)
    c2, radii2, obs2, buf2, state = dgr.forward_raw(scene.means3D, None, colors, scene.opacities, None, None, cov, feats,
                                                    settings)
    assert state.num_rendered == Rr
    assert torch.equal(radii2, radii) and torch.equal(obs2, observe)
    torch.testing.assert_close(c2, color, rtol=RENDER_RTOL, atol=RENDER_ATOL)
    torch.testing.assert_close(buf2, buffer, rtol=RENDER_RTOL, atol=RENDER_ATOL)
    g = dgr.backward_raw(gc, gb, scene.means3D, None, colors, None, None, cov, feats, radii2, settings, state)
    names = ["dL_dmeans2D", "dL_dcolor", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D"]
    for k, t in zip(names, g_ref[:5]):
        err, _ = helpers.grad_errors(g[k], t)
        assert err <= grad_gate(k), "%s err %.3e" % (k, err)
    err, _ = helpers.grad_errors(g["dL_dfeatures"], g_ref[8])
    assert err <= WELL_TOL
    assert float(g["dL_dscale"].abs().max()) == 0.0 and float(g["dL_drot"].abs().max()) == 0.0


def test_autograd_surface_matches_reference_binding(dgr, ref):
    """GaussianRasterizer(...)(...) + loss.backward() through both Python bindings (same call render() makes,
    gaussian_renderer/__init__.py:98-123)."""
    P, W, H, F = 8000, 240, 160, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)

    def run(mod):
        leaves = dict(means3D=scene.means3D.clone().requires_grad_(True),
                      means2D=torch.zeros(P, 4, device="cuda", requires_grad=True),
                      opacities=scene.opacities.clone().requires_grad_(True),
                      shs=scene.shs.clone().requires_grad_(True), scales=scene.scales.clone().requires_grad_(True),
                      rotations=scene.rotations.clone().requires_grad_(True), features=feats.clone().requires_grad_(True))
        settings = syn.raster_settings_for(cam, F, mod.GaussianRasterizationSettings)
        rast = mod.GaussianRasterizer(raster_settings=settings)
        color, radii, observe, buffer = rast(means3D=leaves["means3D"], means2D=leaves["means2D"],
                                            opacities=leaves["opacities"], shs=leaves["shs"], colors_precomp=None,
                                            scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None,
                                            features=leaves["features"])
        loss = (color * gc).sum() + (buffer * gb).sum()
        loss.backward()
        return color, radii, observe, buffer, {k: v.grad for k, v in leaves.items()}

    c1, r1, o1, b1, g1 = run(ref)
    c2, r2, o2, b2, g2 = run(dgr)
    assert c2.shape == (3, H, W) and b2.shape == (10, H, W) and r2.dtype == torch.int32 and o2.dtype == torch.int32
    assert torch.equal(r1, r2) and torch.equal(o1, o2)
    torch.testing.assert_close(c2, c1, rtol=RENDER_RTOL, atol=RENDER_ATOL)
    torch.testing.assert_close(b2, b1, rtol=RENDER_RTOL, atol=RENDER_ATOL)
    for k in g1:
        assert g2[k] is not None and g2[k].shape == g1[k].shape, k
        err, _ = helpers.grad_errors(g2[k], g1[k])
        assert err <= (5e-4 if k in ("scales", "rotations") else WELL_TOL), "%s err %.3e" % (k, err)


def test_mark_visible(dgr, ref):
    scene, cam, feats, _, _ = helpers.make_view(50_000, 320, 240, 1, cam_radius=0.5)   # camera inside the cloud
    settings = syn.raster_settings_for(cam, 1, dgr.GaussianRasterizationSettings)
    ours = dgr.GaussianRasterizer(settings).markVisible(scene.means3D)
    theirs = ref._C.mark_visible(scene.means3D, cam.world_view_transform, cam.full_proj_transform)
    assert ours.dtype == torch.bool and torch.equal(ours, theirs)
    assert 0 < int(ours.sum()) < 50_000


def test_input_validation_mirrors_reference(dgr):
    scene, cam, feats, _, _ = helpers.make_view(100, 64, 64, 5)
    settings = syn.raster_settings_for(cam, 5, dgr.GaussianRasterizationSettings)
    rast = dgr.GaussianRasterizer(settings)
    m2d = torch.zeros(100, 4, device="cuda")
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        rast(scene.means3D, m2d, scene.opacities, shs=None, colors_precomp=None, scales=scene.scales, rotations=scene.rotations)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(scene.means3D, m2d, scene.opacities, shs=scene.shs, scales=scene.scales, rotations=None)
    with pytest.raises(RuntimeError, match="num_points, 3"):
        rast(scene.means3D.view(-1), m2d, scene.opacities, shs=scene.shs, scales=scene.scales, rotations=scene.rotations)
    with pytest.raises(dgr.RasterizerError):   # feature_count outside 0..10 is rejected by the C-ABI, not UB
        bad = settings._replace(feature_count=11)
        dgr.GaussianRasterizer(bad)(scene.means3D, m2d, scene.opacities, shs=scene.shs, scales=scene.scales,
                                    rotations=scene.rotations, features=feats)


def test_empty_and_fully_culled_inputs(dgr, ref):
    """P == 0 short-circuit (rasterize_points.cu:78,161) and a view that culls everything (ranges stay (0,0))."""
    scene, cam, feats, gc, gb = helpers.make_view(1000, 100, 70, 5)
    settings = syn.raster_settings_for(cam, 5, dgr.GaussianRasterizationSettings,
                                       bg=torch.tensor([0.25, 0.5, 0.75], device="cuda"))
    z = lambda *s: torch.zeros(*s, device="cuda")  # noqa: E731
    color, radii, observe, buffer, state = dgr.forward_raw(z(0, 3), z(0, 16, 3), None, z(0, 1), z(0, 3), z(0, 4), None,
                                                           z(0, 10), settings)
    assert state.num_rendered == 0 and radii.numel() == 0
    assert torch.equal(color, settings.bg[:, None, None].expand(3, 70, 100))
    assert float(buffer.abs().max()) == 0.0
    # everything behind the camera
    behind = scene._replace(means3D=(scene.means3D * 0.01 + cam.camera_center[None] * 3.0).contiguous())
    o = helpers.run_ours(dgr, behind, cam, feats, 5, gc, gb, bg=settings.bg)
    r = helpers.run_reference(ref, behind, cam, feats, 5, gc, gb, bg=settings.bg)
    assert o["R"] == 0 and r["R"] == 0
    assert int(o["radii"].abs().sum()) == 0 and int(o["ranges"].abs().sum()) == 0
    torch.testing.assert_close(o["color"], r["color"], rtol=0, atol=0)
    for k in GRAD_NAMES:
        assert float(o[k].abs().max()) == 0.0, k


def test_huge_and_degenerate_gaussians(dgr, ref):
    """Screen-filling splats (long per-tile lists, every tile touched), needle-like splats and opacity below 1/255."""
    P, W, H, F = 3000, 330, 210, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.5)
    scales = scene.scales.clone()
    scales[:40] *= 80.0                        # screen filling
    scales[40:400, 0] *= 30.0                  # needles
    scales[400:500] *= 1e-3                    # sub-pixel (radius clamps through the 0.1 floor)
    opac = scene.opacities.clone()
    opac[500:700] = 0.003                      # can never reach alpha 1/255
    opac[700:800] = 1.0
    scene = scene._replace(scales=scales.contiguous(), opacities=opac.contiguous())
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    assert_forward_bit_exact(o, r, P)
    # These splats make the conic backward so ill-conditioned that the reference's own dL/drot moves by 10-50 % (max
    # norm) between runs, dL/dscale by 1-3 %, dL/dmeans3D by 0.1-0.8 %: gate against the reference's measured spread
    # (three runs, largest pairwise difference) rather than a fixed number.
    r2 = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    r3 = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    for k in GRAD_NAMES:
        err = min(helpers.grad_errors(o[k], x[k])[0] for x in (r, r2, r3))
        noise = max(helpers.grad_errors(a[k], b[k])[0] for a, b in ((r2, r), (r3, r), (r3, r2)))
        assert err <= max(5e-4, 6 * noise), "%s err %.3e noise %.3e" % (k, err, noise)


@pytest.mark.parametrize("path", golden_io.golden_files(), ids=[p.split("/")[-1] for p in golden_io.golden_files()])
def test_against_golden_vectors(dgr, path):
    """Committed outputs of the compiled reference (tests/golden): needs no reference on the box."""
    inp, gold = golden_io.load(path, device="cuda", settings_cls=dgr.GaussianRasterizationSettings)
    s = inp["scene"]
    color, radii, observe, buffer, state = dgr.forward_raw(s.means3D, s.shs, None, s.opacities, s.scales, s.rotations,
                                                           None, inp["features"], inp["settings"])
    sv = dgr.state_view(inp["P"], inp["settings"], state)
    assert state.num_rendered == int(gold["R"][0])
    for name, mine in (("radii", radii), ("observe", observe), ("keys_sorted", sv["keys_sorted"]),
                       ("point_list", sv["point_list"]), ("ranges", sv["ranges"]), ("n_contrib", sv["n_contrib"])):
        assert torch.equal(mine.cpu(), gold[name]), name
    assert torch.equal(helpers.bits(sv["final_T"]).cpu(), helpers.bits(gold["final_T"]))
    torch.testing.assert_close(color.cpu(), gold["color"], rtol=RENDER_RTOL, atol=RENDER_ATOL)
    torch.testing.assert_close(buffer.cpu(), gold["buffer"], rtol=RENDER_RTOL, atol=RENDER_ATOL)
    g = dgr.backward_raw(inp["grad_color"], inp["grad_buffer"], s.means3D, s.shs, None, s.scales, s.rotations, None,
                         inp["features"], radii, inp["settings"], state)
    for k in GRAD_NAMES:
        err, _ = helpers.grad_errors(g[k].cpu(), gold[k])
        assert err <= grad_gate(k), "%s err %.3e" % (k, err)


def test_against_cpu_oracle(dgr):
    """CUDA path vs the pure-PyTorch CPU restatement on a small seeded scene (the check smoke() also runs)."""
    import cpu_rasterizer as cr
    P, W, H, F = 3000, 128, 96, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.7)
    settings = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    orc = cr.CpuRasterizer(torch.float64)
    color, radii, observe, buffer = orc.forward(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                scene.rotations, None, feats, settings)
    assert (radii != o["radii"].cpu()).sum().item() <= 3
    helpers.assert_close_except_flips(o["color"].cpu(), color, rtol=1e-4, atol=2e-5)
    helpers.assert_close_except_flips(o["buffer"].cpu(), buffer, rtol=1e-4, atol=2e-5)
    g = orc.backward(gc, gb, final_T=o["final_T"], n_contrib=o["n_contrib"])
    # vs the fp64 oracle a handful of borderline alpha >= 1/255 decisions flip (host exp vs device expf), which moves
    # individual Gaussians' gradients; the relative L2 norm is insensitive to that, the max norm gets a looser bound
    for k in GRAD_NAMES:
        err, l2 = helpers.grad_errors(o[k].cpu(), g[k])
        loose = k in ("dL_dcov3D", "dL_dscale", "dL_drot")
        assert l2 <= (1e-3 if loose else 3e-4), "%s l2 %.3e" % (k, l2)
        assert err <= 5e-3, "%s max %.3e" % (k, err)


def test_full_size_properties(dgr):
    """Size-independent properties at BASELINE config 4 (3 M Gaussians, 1959x1090, F=10), no reference needed."""
    cfg = syn.CONFIGS["tnt-3m"]
    P, W, H, F = cfg["P"], cfg["W"], cfg["H"], cfg["F"]
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=cfg["shell"], cam_radius=cfg["cam_radius"])
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    keys = o["keys_sorted"]
    R = o["R"]
    assert R > P // 2
    # sortedness (keys are positive: tile < 2^14, depth > 0) and tie-break by ascending Gaussian index
    assert bool((keys[1:] >= keys[:-1]).all())
    ties = keys[1:] == keys[:-1]
    assert bool((o["point_list"][1:][ties] > o["point_list"][:-1][ties]).all())
    # the sorted list is a permutation of the emitted instances: per-Gaussian counts match tiles_touched
    counts = torch.bincount(o["point_list"].long(), minlength=P)
    assert torch.equal(counts.int(), o["tiles_touched"])
    # ranges partition [0,R) by tile id
    rg = o["ranges"].long()
    lens = rg[:, 1] - rg[:, 0]
    assert int(lens.sum()) == R and bool((lens >= 0).all())
    tile_of = (keys >> 32)
    nz = torch.nonzero(lens > 0).reshape(-1)
    assert torch.equal(tile_of[rg[nz, 0]], nz) and torch.equal(tile_of[rg[nz, 1] - 1], nz)
    # depth part of every key is the Gaussian's depth bits
    depth_bits = o["depths"].view(torch.int32).long()[o["point_list"].long()]
    assert torch.equal(keys & 0xFFFFFFFF, depth_bits)
    # determinism of the forward (idempotence): bit-identical on a second run
    o2 = helpers.run_ours(dgr, scene, cam, feats, F)
    for k in ("color", "buffer", "n_contrib", "observe", "radii", "point_list"):
        assert torch.equal(helpers.bits(o[k]), helpers.bits(o2[k])), k
    # alpha channel of the buffer equals 1 - final_T (features[:,0] == 1, no background on features)
    torch.testing.assert_close(o["buffer"][0], 1.0 - o["final_T"], rtol=0, atol=2e-5)
    # linearity of the backward in the upstream gradient, and accumulate mode == sum of two backward passes
    settings = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    c, radii, obs, buf, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                scene.rotations, None, feats, settings)
    args = (scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii, settings, state)
    g1 = dgr.backward_raw(gc, gb, *args)
    g2 = dgr.backward_raw(2.0 * gc, 2.0 * gb, *args)
    acc = dgr.alloc_grads(P, 16, "cuda", zero=True)
    dgr.backward_raw(gc, gb, *args, grads=acc, accumulate=True)
    dgr.backward_raw(gc, gb, *args, grads=acc, accumulate=True)
    for k in GRAD_NAMES:
        # dL/dcov3D, dL/dscale, dL/drot amplify the (order-dependent) rounding of the blend's float reductions by
        # ~1e3 (ill-conditioned conic backward; the reference moves by up to 3e-4 run to run there), so two runs of
        # the SAME kernel only agree to ~1e-3 on those three tensors
        tol = 2e-3 if k in ("dL_dcov3D", "dL_dscale", "dL_drot") else 1e-4
        e, _ = helpers.grad_errors(g2[k], 2.0 * g1[k])
        assert e <= tol, "linearity %s %.3e" % (k, e)
        e, _ = helpers.grad_errors(acc[k], 2.0 * g1[k])
        assert e <= tol, "accumulate %s %.3e" % (k, e)
    # culled Gaussians receive exactly zero gradient
    culled = o["radii"] == 0
    assert int(culled.sum()) > 0
    for k in GRAD_NAMES:
        assert float(g1[k][culled].abs().max()) == 0.0, k


@pytest.mark.parametrize("n", [0, 1, 31, 3071, 3072, 3073, 100_000, 2_000_003])
@pytest.mark.parametrize("end_bit", [9, 32, 46, 64])
def test_radix_sort_building_block(dgr, n, end_bit):
    """gs2m_sort_pairs_u64 against numpy's stable argsort on the masked key bits (what SortPairs does at
    rasterizer_impl.cu:291-296)."""
    from diff_gaussian_rasterization import _native
    lib = _native.load()
    rng = np.random.default_rng(n * 131 + end_bit)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.int64)
    if n > 10:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)]   # plenty of duplicates: stability matters
        rng.shuffle(keys[: n // 2])
    vals = np.arange(n, dtype=np.int32)
    k_in = torch.from_numpy(keys).cuda()
    v_in = torch.from_numpy(vals).cuda()
    k_out, v_out = torch.empty_like(k_in), torch.empty_like(v_in)
    temp = torch.empty(lib.gs2m_sort_temp_bytes(n), dtype=torch.uint8, device="cuda")
    p = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
    rc = lib.gs2m_sort_pairs_u64(p(k_in), p(k_out), p(v_in), p(v_out), n, end_bit, temp.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ku = keys.view(np.uint64)
    masked = ku & np.uint64((1 << end_bit) - 1) if end_bit < 64 else ku
    order = np.argsort(masked, kind="stable")
    assert np.array_equal(v_out.cpu().numpy(), vals[order])
    assert np.array_equal(k_out.cpu().numpy().view(np.uint64), ku[order])


@pytest.mark.parametrize("n", [1, 255, 2048, 2049, 1_000_003, 6_000_000])
def test_inclusive_scan_building_block(dgr, n):
    from diff_gaussian_rasterization import _native
    lib = _native.load()
    rng = np.random.default_rng(n)
    x = rng.integers(0, 40, size=n, dtype=np.int32)
    xin = torch.from_numpy(x).cuda()
    out = torch.empty_like(xin)
    temp = torch.empty(lib.gs2m_scan_temp_bytes(n), dtype=torch.uint8, device="cuda")
    assert lib.gs2m_inclusive_sum_u32(xin.data_ptr(), out.data_ptr(), n, temp.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream) == 0
    assert np.array_equal(out.cpu().numpy(), np.cumsum(x, dtype=np.int64).astype(np.int32))
