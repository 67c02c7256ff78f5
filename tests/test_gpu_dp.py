"""Multi-GPU check of the view-sharded data-parallel step (needs >= 2 CUDA devices; skipped otherwise): the gradients
every rank holds after the NCCL all-reduce equal the single-GPU sequential sum over the same views, and the MAX / SUM
reduced densification statistics agree."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
import synthetic_scenes as syn

pytestmark = pytest.mark.gpu

P, W, H, F, N_VIEWS = 40_000, 480, 320, 10, 4


def _make_render(dgr, scene, cams, feats, gc, gb):
    settings = {v: syn.raster_settings_for(cams[v], F, dgr.GaussianRasterizationSettings) for v in cams}

    holder = {}

    def render_view(v, buckets, accumulate):
        color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities,
                                                               scene.scales, scene.rotations, None, feats[v], settings[v])
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v], radii,
                         settings[v], state, grads=buckets.tensors, accumulate=accumulate,
                         densify_stats=holder["step"].stats.backward_args())
        return {"radii": radii, "observe": observe}

    def begin_view(v):          # deferred protocol (same arithmetic, the per-Gaussian stage postponed and split into ranges)
        color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities,
                                                               scene.scales, scene.rotations, None, feats[v], settings[v])
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v], radii,
                         settings[v], state, grads=holder["step"].buckets.tensors, phase="blend")
        return {"v": v, "radii": radii, "observe": observe, "state": state}

    def finish_view(h, buckets, accumulate, rows):
        v = h["v"]
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v], h["radii"],
                         settings[v], h["state"], grads=buckets.tensors, accumulate=accumulate, phase="gaussians", rows=rows,
                         densify_stats=holder["step"].stats.backward_args())
    holder["deferred"] = (begin_view, finish_view)
    return render_view, holder


def _setup(device):
    import diff_gaussian_rasterization as dgr
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=0.6), device)
    cams_cpu = syn.make_cameras(N_VIEWS, W, H)
    cams = {v: syn.camera_to(cams_cpu[v], device) for v in range(N_VIEWS)}
    feats = {v: syn.pack_features(scene, cams[v], F) for v in range(N_VIEWS)}
    gc, gb = syn.make_upstream_grads(W, H, F)
    return dgr, scene, cams, feats, gc.to(device), gb.to(device)


def _worker(rank, world, port, out_dir, n_streams, deferred=False):
    import view_parallel as vp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    device = torch.device("cuda", rank)
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    try:
        dgr, scene, cams, feats, gc, gb = _setup(device)
        render, holder = _make_render(dgr, scene, cams, feats, gc, gb)
        if deferred:
            step = holder["step"] = vp.ViewShardedStep(P, 16, device, n_streams=n_streams, begin_view=holder["deferred"][0],
                                                       finish_view=holder["deferred"][1], n_chunks=3)
        else:
            step = holder["step"] = vp.ViewShardedStep(P, 16, device, render, n_streams=n_streams)
        step.buckets.flat.fill_(7.0)                                    # stale content must not leak into the step
        grads = step.run(N_VIEWS)
        torch.cuda.synchronize(device)
        torch.save({"grads": {k: grads[k].cpu() for k in step.buckets.names}, "stats": step.stats.flat.cpu()},
                   os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_streams,deferred", [(1, False), (2, False), (2, True)])
def test_two_gpu_step_equals_single_gpu_sequential_sum(tmp_path, n_streams, deferred):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    import view_parallel as vp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), n_streams, deferred), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    # single-GPU sequential reference: the same four views, one after the other, on cuda:0
    device = torch.device("cuda", 0)
    dgr, scene, cams, feats, gc, gb = _setup(device)
    render, holder = _make_render(dgr, scene, cams, feats, gc, gb)
    ref_step = holder["step"] = vp.ViewShardedStep(P, 16, device, render, world=1, rank=0)
    ref = ref_step.run(N_VIEWS, reduce=False)
    for k in vp.REDUCED:
        for r in range(world):
            tol = 2e-3 if k in ("dL_dscale", "dL_drot") else 1e-4     # atomic-order noise of the ill-conditioned pair
            err, _ = helpers.grad_errors(res[r]["grads"][k], ref[k].cpu())
            assert err <= tol, "%s rank %d: %.3e" % (k, r, err)
        assert torch.equal(res[0]["grads"][k], res[1]["grads"][k])      # all ranks hold identical bits
    ref_stats = ref_step.stats.flat.cpu()
    assert 0 < float(ref_stats[3 * P:4 * P].max()) <= N_VIEWS                    # denom counts the views that saw a Gaussian
    for r in range(world):
        # max_radii2D, denom and observe_cnt are exact; the two gradient-norm sums differ by summation order only
        for lo, hi in ((0, P), (3 * P, 5 * P)):
            assert torch.equal(res[r]["stats"][lo:hi], ref_stats[lo:hi])
        torch.testing.assert_close(res[r]["stats"][P:3 * P], ref_stats[P:3 * P], rtol=1e-4, atol=1e-9)
    assert torch.equal(res[0]["stats"], res[1]["stats"])
