"""Host logic of the view-sharded data-parallel step on CPU: world_size-2 gloo run with a fake per-view renderer
(the real one needs a GPU), plus the sharding arithmetic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import view_parallel as vp


@pytest.mark.parametrize("n_views,world", [(64, 8), (49, 8), (8, 8), (3, 8), (1, 2), (0, 4), (7, 2)])
def test_shard_views_partitions_the_batch(n_views, world):
    got = [list(vp.shard_views(n_views, world, r)) for r in range(world)]
    flat = [v for part in got for v in part]
    assert flat == list(range(n_views))                      # disjoint, complete, ordered
    sizes = [len(p) for p in got]
    assert max(sizes) - min(sizes) <= 1                      # balanced


def test_shard_views_rejects_bad_rank():
    with pytest.raises(ValueError):
        vp.shard_views(4, 2, 2)


def _fake_view_gradient(v, shape):
    g = torch.Generator().manual_seed(1000 + v)
    return torch.randn(shape, generator=g)


def _worker(rank, world, port, n_views, P, M, out_dir, param_buckets=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def render_view(v, buckets, accumulate):
            for name in buckets.names:
                t = buckets.tensors[name]
                g = _fake_view_gradient(v * 31 + len(name), t.shape)
                if accumulate:
                    t += g
                else:
                    t.copy_(g)          # first view of the step overwrites (kernels write every element)
            radii = torch.full((P,), v + 1, dtype=torch.int32)
            radii[(v + 1) % P] = 0                                  # one culled Gaussian per view
            observe = torch.zeros(P, dtype=torch.int32)
            observe[v % P] = 3
            return {"radii": radii, "observe": observe, "means2D_grad": _fake_view_gradient(v * 31 + len("dL_dmeans2D"), (P, 4))}

        step = vp.ViewShardedStep(P, M, "cpu", render_view, buckets_cls=vp.ParameterBuckets if param_buckets else None)
        assert step.world == world and step.rank == rank
        # poison the buckets: a stale gradient from a previous step must not leak into this one
        for t in step.buckets.tensors.values():
            t.fill_(123.0)
        grads = step.run(n_views)
        torch.save({"grads": {k: grads[k].clone() for k in step.buckets.names}, "radii": step.stats.max_radii2D.clone(),
                    "observe": step.stats.observe_cnt.clone(), "accum": step.stats.xyz_gradient_accum.clone(),
                    "accum_abs": step.stats.xyz_gradient_accum_abs.clone(), "denom": step.stats.denom.clone(),
                    "mine": step.local_views(n_views)},
                   os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_views,param_buckets", [(5, False), (1, False), (5, True)])
def test_two_rank_step_equals_sequential_sum(tmp_path, n_views, param_buckets):
    world, P, M = 2, 37, 4
    mp.spawn(_worker, args=(world, _free_port(), n_views, P, M, str(tmp_path), param_buckets), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(n_views))
    names = vp.ParameterBuckets.names if param_buckets else vp.REDUCED
    shapes = {k: res[0]["grads"][k].shape for k in names}
    for k in names:
        expect = torch.zeros(shapes[k])
        for v in range(n_views):
            expect += _fake_view_gradient(v * 31 + len(k), shapes[k])
        for r in range(world):   # every rank holds the batch gradient == single-process sequential sum
            torch.testing.assert_close(res[r]["grads"][k], expect, rtol=1e-6, atol=1e-6)
    # densification statistics, reference semantics (train.py:225-241, scene/gaussian_model.py:569-573), sequentially
    max_r, cnt = torch.zeros(P), torch.zeros(P, 1)
    accum, accum_abs, denom = torch.zeros(P, 1), torch.zeros(P, 1), torch.zeros(P, 1)
    for v in range(n_views):
        radii = torch.full((P,), float(v + 1)); radii[(v + 1) % P] = 0
        observe = torch.zeros(P); observe[v % P] = 3
        mask = (observe > 0) & (radii > 0)
        max_r = torch.where(mask, torch.max(max_r, radii), max_r)
        cnt[observe > 0] += 1
        g = _fake_view_gradient(v * 31 + len("dL_dmeans2D"), (P, 4))
        vis = radii > 0
        accum[vis] += torch.norm(g[vis, :2], dim=-1, keepdim=True)
        accum_abs[vis] += torch.norm(g[vis, 2:], dim=-1, keepdim=True)
        denom[vis] += 1
    for r in range(world):
        assert torch.equal(res[r]["radii"], max_r)                       # MAX over views and ranks
        assert torch.equal(res[r]["observe"], cnt)                       # SUM over views and ranks
        assert torch.equal(res[r]["denom"], denom)
        torch.testing.assert_close(res[r]["accum"], accum, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(res[r]["accum_abs"], accum_abs, rtol=1e-6, atol=1e-6)
