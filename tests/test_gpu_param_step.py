"""Raw-parameter data-parallel step (the collective runs over GS-2M's nine parameter-gradient tensors):
per-view chain  rasterizer backward (accumulate mode 2) -> fused packing backward (+=)  against autograd through
activate_and_pack + the drop-in rasterizer summed over the same views."""
import pytest
import torch

import helpers
import synthetic_scenes as syn
import view_parallel as vp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_streams,deferred,n_views", [(1, False, 4), (2, False, 4), (1, True, 4), (2, True, 4), (1, "fused", 4),
                                                        (2, "fused", 4), (1, "views", 4), (2, "views", 3), (2, "views", 11)])
def test_parameter_step_matches_autograd_over_views(n_streams, deferred, n_views):
    """deferred = "fused": the packing chain runs inside the rasterizer's per-Gaussian backward (gs2m_backward_args::chain);
    "views": one multi-view pass per Gaussian range (gs2m_rasterize_backward_views; 11 views = two launches, the second adding)."""
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization.packing import activate_and_pack
    P, W, H, F, M = 30_000, 320, 240, 10, 16
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=0.6), "cuda")
    cams = [syn.camera_to(c, "cuda") for c in syn.make_cameras(n_views, W, H)]
    settings = [syn.raster_settings_for(c, F, dgr.GaussianRasterizationSettings) for c in cams]
    gc, gb = (t.cuda() for t in syn.make_upstream_grads(W, H, F))
    raw = syn.raw_parameters(scene)
    order = ("xyz", "scaling", "rotation", "opacity", "albedo", "roughness", "metallic")

    # ---- ground truth: autograd through the packing stage and the drop-in rasterizer, summed over the views ----
    req = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    sh = scene.shs.clone().requires_grad_(True)
    total = 0.0
    for cam, st in zip(cams, settings):
        s, q, o, f = activate_and_pack(*[req[k] for k in order], cam.world_view_transform, cam.camera_center, blend_metallic=True)
        means2D = torch.zeros(P, 4, device="cuda", requires_grad=True)
        color, radii, observe, buffer = dgr.GaussianRasterizer(st)(means3D=req["xyz"], means2D=means2D, opacities=o, shs=sh,
                                                                    scales=s, rotations=q, features=f)
        total = total + (color * gc).sum() + (buffer * gb).sum()
    total.backward()
    expect = {k: req[k].grad for k in order}
    expect["sh"] = sh.grad

    # ---- the step: per view  forward -> backward (mode 2) -> chain_view, raw-parameter buckets ----
    def render_view(v, buckets, accumulate):
        cam, st = cams[v], settings[v]
        with torch.no_grad():
            s, q, o, f = activate_and_pack(*[raw[k] for k in order], cam.world_view_transform, cam.camera_center,
                                           blend_metallic=True)
        color, radii, observe, buffer, state = dgr.forward_raw(raw["xyz"], scene.shs, None, o, s, q, None, f, st)
        dgr.backward_raw(gc, gb, raw["xyz"], scene.shs, None, s, q, None, f, radii, st, state, grads=buckets.raster,
                         accumulate=buckets.raster_accumulate_mode(accumulate))
        buckets.chain_view(raw, cam.world_view_transform, cam.camera_center, radii, blend_metallic=True)
        return {"radii": radii, "observe": observe}

    # ---- the deferred step: forward + reverse blend per view, then the per-Gaussian stage + packing chain range by range ----
    def begin_view(v):
        cam, st = cams[v], settings[v]
        with torch.no_grad():
            s, q, o, f = activate_and_pack(*[raw[k] for k in order], cam.world_view_transform, cam.camera_center,
                                           blend_metallic=True)
        color, radii, observe, buffer, state = dgr.forward_raw(raw["xyz"], scene.shs, None, o, s, q, None, f, st)
        h = dict(v=v, s=s, q=q, f=f, radii=radii, observe=observe, state=state, st=st)
        dgr.backward_raw(gc, gb, raw["xyz"], scene.shs, None, s, q, None, f, radii, st, state, grads=step.buckets.raster,
                         phase="blend")
        return h

    def finish_view(h, buckets, accumulate, rows):
        cam, st = cams[h["v"]], settings[h["v"]]
        if deferred == "fused":
            dgr.backward_raw(gc, gb, raw["xyz"], scene.shs, None, h["s"], h["q"], None, h["f"], h["radii"], st, h["state"],
                             grads=buckets.raster, accumulate=2 if accumulate else 0, phase="gaussians", rows=rows,
                             chain=buckets.chain_spec(raw, blend_metallic=True))
            return
        dgr.backward_raw(gc, gb, raw["xyz"], scene.shs, None, h["s"], h["q"], None, h["f"], h["radii"], st, h["state"],
                         grads=buckets.raster, accumulate=2 if accumulate else 0, phase="gaussians", rows=rows)
        buckets.chain_rows(raw, cam.world_view_transform, cam.camera_center, h["radii"], rows[0], rows[1], blend_metallic=True)

    def finish_views(handles, buckets, rows, accumulate=False):
        chain = buckets.chain_spec(raw, blend_metallic=True)
        dgr.backward_views_raw([dict(grad_color=gc, grad_buffer=gb, means3D=raw["xyz"], shs=scene.shs, scales=h["s"], rotations=h["q"],
                                     features=h["f"], radii=h["radii"], raster_settings=h["st"], state=h["state"],
                                     grads=buckets.raster, densify_stats=step.stats.backward_args(), chain=chain)
                                for h in handles], rows=rows, accumulate=accumulate)

    if deferred:
        step = vp.ViewShardedStep(P, M, "cuda", world=1, rank=0, n_streams=n_streams, buckets_cls=vp.ParameterBuckets,
                                  begin_view=begin_view, finish_view=finish_view, n_chunks=5,
                                  finish_views=finish_views if deferred == "views" else None,
                                  max_views_in_flight=4 if n_views == 11 else None)
        assert len(step.chunks) == 5 and len(step.bucket_sets) == 1
        step.buckets.fused_chain = deferred in ("fused", "views")
    else:
        step = vp.ViewShardedStep(P, M, "cuda", render_view, world=1, rank=0, n_streams=n_streams, buckets_cls=vp.ParameterBuckets)
    for b in step.bucket_sets:
        b.flat.fill_(3.0)                                   # stale content must not leak into the step
    got = step.run(n_views, reduce=False)
    torch.cuda.synchronize()
    assert step.buckets.nbytes_reduced() == P * 64 * 4      # 64 floats per Gaussian
    for k in vp.ParameterBuckets.names:
        err, l2 = helpers.grad_errors(got[k], expect[k])
        tol = 2e-3 if k in ("scaling", "rotation") else 1e-4      # the ill-conditioned conic backward feeds these two
        assert err <= tol and l2 <= tol, "%s: max %.3e l2 %.3e" % (k, err, l2)
    if deferred == "views":      # the gradient-norm statistics of every view went through the same pass
        vis_count = torch.zeros(P, device="cuda")
        for v in range(n_views):
            with torch.no_grad():
                s, q, o, f = activate_and_pack(*[raw[k] for k in order], cams[v].world_view_transform, cams[v].camera_center,
                                               blend_metallic=True)
            _c, radii, _o, _b, _s = dgr.forward_raw(raw["xyz"], scene.shs, None, o, s, q, None, f, settings[v])
            vis_count += (radii > 0).float()
        assert torch.equal(step.stats.denom.view(-1), vis_count)
        assert float(step.stats.xyz_gradient_accum.abs().max()) > 0.0


def test_accumulate_mode_2_overwrites_the_view_dependent_tensors():
    import diff_gaussian_rasterization as dgr
    P, W, H, F = 20_000, 256, 192, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    st = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                           scene.rotations, None, feats, st)
    args = (gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii, st, state)
    ref = dgr.backward_raw(*args)
    g = dgr.alloc_grads(P, 16, "cuda")
    base = {k: 0.37 * float(ref[k].abs().max()) for k in g}   # same magnitude as the gradients (2.0 would swamp them in fp32)
    for k, t in g.items():
        t.fill_(base[k])
    dgr.backward_raw(*args, grads=g, accumulate=2)
    vis = radii > 0
    # (two backward runs differ in the last bits: the blend kernel's vector reductions arrive in any order)
    for k in ("dL_dmeans3D", "dL_dsh"):                      # += for the raw parameters, untouched rows for culled Gaussians
        err, _ = helpers.grad_errors(g[k][vis] - base[k], ref[k][vis])
        assert err <= 1e-4, "%s: %.3e" % (k, err)
        assert bool((g[k][~vis] == torch.tensor(base[k], dtype=torch.float32)).all())
    for k in ("dL_dmeans2D", "dL_dopacity", "dL_dscale", "dL_drot", "dL_dfeatures", "dL_dcolor", "dL_dcov3D"):
        err, _ = helpers.grad_errors(g[k], ref[k])          # overwritten ...
        assert err <= (2e-3 if k in ("dL_dscale", "dL_drot", "dL_dcov3D") else 1e-4), "%s: %.3e" % (k, err)
        assert float(g[k][~vis].abs().max()) == 0.0, k      # ... with zeros for culled Gaussians
    with pytest.raises(dgr.RasterizerError):
        dgr.backward_raw(*args, grads=g, accumulate=3)


def test_backward_phases_and_row_ranges_equal_the_whole_backward():
    """phase="blend" followed by phase="gaussians" over 256-aligned row ranges writes exactly what the one-call backward writes
    (the per-Gaussian stage is independent per Gaussian), in overwrite and in accumulate mode."""
    import diff_gaussian_rasterization as dgr
    P, W, H, F = 21_000, 256, 192, 10
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
    st = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                           scene.rotations, None, feats, st)
    args = (gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii, st, state)
    whole = {k: v.clone() for k, v in dgr.backward_raw(*args).items()}
    for acc in (0, 1):
        g = dgr.alloc_grads(P, 16, "cuda", zero=True)
        base = {k: (0.37 * float(whole[k].abs().max()) if acc else 9.0) for k in g}     # same magnitude as the gradients: an
        for k, t in g.items():                                                           # fp32 sum keeps their low bits
            t.fill_(base[k])
        dgr.backward_raw(*args, grads=g, phase="blend")
        for rows in ((0, 5120), (5120, 5376), (5376, 20992), (20992, P)):
            dgr.backward_raw(*args, grads=g, accumulate=acc, phase="gaussians", rows=rows)
        for k in whole:
            if k == "dL_dconic":
                continue
            # same kernels over the same accumulator: only the second blend's reduction order differs from the first
            err, _ = helpers.grad_errors(g[k] - (base[k] if acc else 0.0), whole[k])
            assert err <= (5e-4 if k in ("dL_dscale", "dL_drot", "dL_dcov3D") else 2e-5), "%s acc=%d: %.3e" % (k, acc, err)
    with pytest.raises(dgr.RasterizerError):
        dgr.backward_raw(*args, phase="gaussians", rows=(100, 300))      # range start must be a multiple of 256
