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


def _worker(rank, world, port, n_views, P, M, out_dir, param_buckets=False, deferred=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def view_stats(v):
            radii = torch.full((P,), v + 1, dtype=torch.int32)
            radii[(v + 1) % P] = 0                                  # one culled Gaussian per view
            observe = torch.zeros(P, dtype=torch.int32)
            observe[v % P] = 3
            return radii, observe

        def begin_view(v):                                          # deferred protocol: forward + reverse blend of a view ...
            radii, observe = view_stats(v)
            return {"v": v, "radii": radii, "observe": observe}

        def finish_view(h, buckets, accumulate, rows):              # ... and its per-Gaussian stage for a range of Gaussians
            b, e = rows
            for name in buckets.names:
                t = buckets.tensors[name]
                g = _fake_view_gradient(h["v"] * 31 + len(name), t.shape)[b:e]
                if accumulate or isinstance(buckets, vp.ParameterBuckets) and name not in ("xyz", "sh"):
                    t[b:e] += g         # (the packed groups of ParameterBuckets are zeroed by begin_rows and always added to)
                else:
                    t[b:e] = g
            if b == 0:
                holder["step"].stats.update_backward_eager(_fake_view_gradient(h["v"] * 31 + len("dL_dmeans2D"), (P, 4)), h["radii"])

        holder = {}

        def render_view(v, buckets, accumulate):
            for name in buckets.names:
                t = buckets.tensors[name]
                g = _fake_view_gradient(v * 31 + len(name), t.shape)
                if accumulate:
                    t += g
                else:
                    t.copy_(g)          # first view of the step overwrites (kernels write every element)
            radii, observe = view_stats(v)
            return {"radii": radii, "observe": observe, "means2D_grad": _fake_view_gradient(v * 31 + len("dL_dmeans2D"), (P, 4))}

        cls = vp.ParameterBuckets if param_buckets else None
        if deferred:
            step = vp.ViewShardedStep(P, M, "cpu", buckets_cls=cls, begin_view=begin_view, finish_view=finish_view, n_chunks=3,
                                      max_views_in_flight=2 if deferred == "capped" else None)   # rank 0: groups of 2 + 1 views
            assert len(step.chunks) == 3 and step.chunks[0][0] == 0 and step.chunks[-1][1] == P
        else:
            step = vp.ViewShardedStep(P, M, "cpu", render_view, buckets_cls=cls)
        holder["step"] = step
        assert step.world == world and step.rank == rank
        # poison the buckets: a stale gradient from a previous step must not leak into this one
        step.buckets.flat.fill_(123.0)
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


@pytest.mark.parametrize("n_views,param_buckets,deferred", [(5, False, False), (1, False, False), (5, True, False),
                                                            (5, False, True), (1, True, True), (5, True, True),
                                                            (5, False, "capped"), (5, True, "capped")])
def test_two_rank_step_equals_sequential_sum(tmp_path, n_views, param_buckets, deferred):
    world, P, M = 2, (1500 if deferred else 37), 4         # 1500 Gaussians: three 256-aligned, shrinking ranges in the deferred step
    mp.spawn(_worker, args=(world, _free_port(), n_views, P, M, str(tmp_path), param_buckets, deferred), nprocs=world, join=True)
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


def test_balance_views_equal_counts_and_smaller_spread():
    import random
    rng = random.Random(3)
    for world, n in ((8, 64), (4, 10), (3, 3), (5, 2)):
        costs = [rng.uniform(6.5, 8.0) for _ in range(n)]
        table = vp.balance_views(costs, world)
        assert sorted(v for vs in table for v in vs) == list(range(n))
        assert [len(vs) for vs in table] == [len(vp.shard_views(n, world, r)) for r in range(world)]
        spread = lambda t: max(sum(costs[v] for v in vs) for vs in t) - min(sum(costs[v] for v in vs) for vs in t if vs or True)  # noqa: E731
        contiguous = [list(vp.shard_views(n, world, r)) for r in range(world)]
        assert spread(table) <= spread(contiguous) + 1e-9
        assert vp.balance_views(costs, world) == table           # deterministic: every rank derives the same table
    step = vp.ViewShardedStep(4, 1, "cpu", render_view=lambda v, b, a: None, world=2, rank=1, assignment=[[0, 3], [1, 2]])
    assert step.local_views(4) == [1, 2]
    with pytest.raises(ValueError):
        vp.ViewShardedStep(4, 1, "cpu", render_view=lambda v, b, a: None, world=2, rank=0, assignment=[[0, 1], [1, 2]]).local_views(4)
