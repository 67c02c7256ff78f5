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


def _worker(rank, world, port, n_views, P, M, out_dir):
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
            observe = torch.zeros(P, dtype=torch.int32)
            observe[v % P] = 3
            return {"radii": radii, "observe": observe}

        step = vp.ViewShardedStep(P, M, "cpu", render_view)
        assert step.world == world and step.rank == rank
        # poison the buckets: a stale gradient from a previous step must not leak into this one
        for t in step.buckets.tensors.values():
            t.fill_(123.0)
        grads = step.run(n_views)
        torch.save({"grads": {k: grads[k].clone() for k in step.buckets.names}, "radii": step.radii_max.clone(),
                    "observe": step.observe_count.clone(), "mine": step.local_views(n_views)},
                   os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_views", [5, 1])
def test_two_rank_step_equals_sequential_sum(tmp_path, n_views):
    world, P, M = 2, 37, 4
    mp.spawn(_worker, args=(world, _free_port(), n_views, P, M, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "rank%d.pt" % r)) for r in range(world)]
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(n_views))
    names = vp.REDUCED
    shapes = {k: res[0]["grads"][k].shape for k in names}
    for k in names:
        expect = torch.zeros(shapes[k])
        for v in range(n_views):
            expect += _fake_view_gradient(v * 31 + len(k), shapes[k])
        for r in range(world):   # every rank holds the batch gradient == single-process sequential sum
            torch.testing.assert_close(res[r]["grads"][k], expect, rtol=1e-6, atol=1e-6)
    for r in range(world):
        assert int(res[r]["radii"].max()) == n_views                    # MAX over views and ranks
        assert int(res[r]["observe"].sum()) == n_views                   # one hit per view, SUM over ranks
