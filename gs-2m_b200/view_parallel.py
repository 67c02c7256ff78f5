"""View-sharded data parallelism for the rasterizer (north_star: parameters replicated on the GPUs of one box, each
training batch of camera views sharded across them, parameter gradients summed by NCCL all-reduce over NVLink).

The reference has no distributed code at all; what it does have is autograd summing the
per-Gaussian gradients of the two views it renders per iteration (train.py:95 + utils/loss_utils.py:253).  This
module does the same sum across a *batch* of views and across ranks:

* ``shard_views(n_views, world, rank)``  – contiguous, balanced split of a batch of cameras (64 views / 8 ranks = 8).
* ``GradientBuckets``                    – the dense per-Gaussian gradient tensors (reference shapes,
  rasterize_points.cu:150-159).  A rank's views accumulate into them in place (the C-ABI's ``accumulate`` flag, so
  there is no per-view zero-fill or extra add kernel); one ``all_reduce(SUM)`` per tensor then makes every rank hold
  the batch gradient.  Densification statistics need the same reduction (SUM for the .zw |grad| columns of
  ``dL_dmeans2D``, MAX for radii, SUM>0 for observe; train.py:225-245) and are covered by ``reduce_statistics``.
* ``ViewShardedStep``                    – runs forward+backward for this rank's views through a caller-supplied
  ``render_view`` callable and finishes with the all-reduce.  The callable is the only piece that touches CUDA, so
  the sharding / accumulation / reduction logic is testable on CPU with the gloo backend.

One process per GPU (torchrun); the path shards with no data-path collective other than this one exchange step.
"""
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

# tensors that are summed over views and ranks: (name, trailing shape as a function of M)
# 73 floats per Gaussian.  dL_dcolor / dL_dcov3D are gradients of the *precomputed* colour / covariance inputs, which GS-2M
# does not train (it feeds SHs and scale+rotation), so they are per-view scratch here; pass names=REDUCED + (...) to
# GradientBuckets if a caller does optimise them.
REDUCED = ("dL_dmeans3D", "dL_dmeans2D", "dL_dsh", "dL_dopacity", "dL_dscale", "dL_drot", "dL_dfeatures")


def shard_views(n_views: int, world: int, rank: int) -> range:
    """Contiguous balanced split: the first ``n_views % world`` ranks get one extra view."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank %d/%d" % (world, rank))
    base, extra = divmod(n_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class GradientBuckets:
    """The per-Gaussian gradient tensors of one rank.  The reduced ones are views into ONE flat fp32 buffer (each view
    starts on a 128-byte boundary), so the exchange step is a single large all-reduce — at 8 ranks one 0.9 GB
    collective reaches a much higher bus bandwidth than seven separate ones, three of which are latency-bound."""

    def __init__(self, P: int, M: int, device, names: Sequence[str] = REDUCED):
        shapes = {"dL_dmeans3D": (P, 3), "dL_dmeans2D": (P, 4), "dL_dsh": (P, M, 3), "dL_dopacity": (P, 1),
                  "dL_dscale": (P, 3), "dL_drot": (P, 4), "dL_dfeatures": (P, 10), "dL_dcolor": (P, 3),
                  "dL_dcov3D": (P, 6), "dL_dconic": (P, 4)}
        self.names = tuple(names)

        def numel(shape):
            n = 1
            for d in shape:
                n *= d
            return n
        offsets, total = {}, 0
        for n in self.names:
            offsets[n] = total
            total += (numel(shapes[n]) + 31) // 32 * 32          # 128-byte granularity
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.tensors: Dict[str, torch.Tensor] = {
            n: self.flat[offsets[n]:offsets[n] + numel(shapes[n])].view(shapes[n]) for n in self.names}
        # per-view scratch of the backward (never reduced), but the C-ABI wants a pointer for each
        for n in shapes:
            if n not in self.tensors:
                self.tensors[n] = torch.zeros(shapes[n], dtype=torch.float32, device=device)
        self.views_accumulated = 0

    def zero_(self):
        for t in self.tensors.values():
            t.zero_()
        self.views_accumulated = 0

    def nbytes_reduced(self) -> int:
        return sum(self.tensors[n].numel() * 4 for n in self.names)

    def all_reduce(self, async_op: bool = False):
        """SUM over ranks: one collective over the flat buffer. Returns the work handle when ``async_op``."""
        if not _dist_ready():
            return []
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return [work] if async_op else []


def reduce_statistics(radii_max: torch.Tensor, observe_count: torch.Tensor):
    """Densification bookkeeping that also has to agree on every rank: screen radii are MAX-reduced
    (scene/gaussian_model.py max_radii2D, train.py:226) and the per-view ``observe > 0`` hit counts are SUM-reduced
    (train.py:238-243)."""
    if _dist_ready():
        dist.all_reduce(radii_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(observe_count, op=dist.ReduceOp.SUM)
    return radii_max, observe_count


class ViewShardedStep:
    """One data-parallel rasterization step over a batch of views.

    ``render_view(view_index, buckets, accumulate) -> dict`` must run forward+backward of one view, adding its
    gradients into ``buckets.tensors`` (``accumulate`` is False for the first view that writes a bucket set: the kernels
    then overwrite, which saves zero-filling), and may return per-view statistics
    ``{"radii": int32[P], "observe": int32[P]}``.

    With ``n_streams > 1`` (CUDA only) consecutive views of a rank are issued round-robin on that many streams, each with
    its own bucket set, so the kernels of view k+1 fill the GPU while view k sits in the forward's instance-count
    read-back or in a tail wave; the bucket sets are summed once per step before the all-reduce.
    """

    def __init__(self, P: int, M: int, device, render_view: Callable[[int, GradientBuckets, bool], Optional[dict]],
                 world: Optional[int] = None, rank: Optional[int] = None, n_streams: int = 1):
        self.world = world if world is not None else (dist.get_world_size() if _dist_ready() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if _dist_ready() else 0)
        self.device = torch.device(device)
        use_streams = n_streams > 1 and self.device.type == "cuda"
        self.n_streams = n_streams if use_streams else 1
        self.bucket_sets = [GradientBuckets(P, M, device) for _ in range(self.n_streams)]
        self.buckets = self.bucket_sets[0]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)] if use_streams else [None]
        self.render_view = render_view
        self.radii_max = torch.zeros(P, dtype=torch.int32, device=device)
        self.observe_count = torch.zeros(P, dtype=torch.int32, device=device)

    def local_views(self, n_views: int) -> List[int]:
        return list(shard_views(n_views, self.world, self.rank))

    def _one_view(self, v, buckets, accumulate, radii_max, observe_count):
        stats = self.render_view(v, buckets, accumulate)
        if stats:
            if "radii" in stats:
                torch.maximum(radii_max, stats["radii"], out=radii_max)
            if "observe" in stats:
                observe_count += (stats["observe"] > 0).to(torch.int32)

    def run(self, n_views: int, reduce: bool = True) -> Dict[str, torch.Tensor]:
        mine = self.local_views(n_views)
        self.radii_max.zero_()
        self.observe_count.zero_()
        if not mine:  # more ranks than views: contribute zeros
            self.buckets.zero_()
        if self.n_streams == 1:
            for k, v in enumerate(mine):
                self._one_view(v, self.buckets, k > 0, self.radii_max, self.observe_count)
                self.buckets.views_accumulated = k + 1
        else:
            main = torch.cuda.current_stream(self.device)
            ready = torch.cuda.Event()
            ready.record(main)
            stats = [(torch.zeros_like(self.radii_max), torch.zeros_like(self.observe_count)) for _ in self.streams]
            used = min(self.n_streams, len(mine))
            for k, v in enumerate(mine):
                j = k % self.n_streams
                with torch.cuda.stream(self.streams[j]):
                    if k < self.n_streams:
                        self.streams[j].wait_event(ready)        # inputs / previous step's consumers are done
                    self._one_view(v, self.bucket_sets[j], k >= self.n_streams, stats[j][0], stats[j][1])
            for j in range(used):
                main.wait_stream(self.streams[j])
            for j in range(used):                                   # fold the per-stream partial sums into set 0
                torch.maximum(self.radii_max, stats[j][0], out=self.radii_max)
                self.observe_count += stats[j][1]
                if j > 0:
                    self.buckets.flat += self.bucket_sets[j].flat
            for j in range(used):                                   # later work on the side streams must wait for the fold
                self.streams[j].wait_stream(main)
            self.buckets.views_accumulated = len(mine)
        if reduce:
            self.buckets.all_reduce()
            reduce_statistics(self.radii_max, self.observe_count)
        return self.buckets.tensors
