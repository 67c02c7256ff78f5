"""View-sharded data parallelism for the rasterizer (north_star: parameters replicated on the GPUs of one box, each
training batch of camera views sharded across them, parameter gradients summed by NCCL all-reduce over NVLink).

The reference has no distributed code at all; what it does have is autograd summing the
per-Gaussian gradients of the two views it renders per iteration (train.py:95 + utils/loss_utils.py:253).  This
module does the same sum across a *batch* of views and across ranks:

* ``shard_views(n_views, world, rank)``  – contiguous, balanced split of a batch of cameras (64 views / 8 ranks = 8).
* ``GradientBuckets``                    – the dense per-Gaussian gradient tensors (reference shapes,
  rasterize_points.cu:150-159).  A rank's views accumulate into them in place (the C-ABI's ``accumulate`` flag, so
  there is no per-view zero-fill or extra add kernel); one ``all_reduce(SUM)`` per tensor then makes every rank hold
  the batch gradient.
* ``DensificationStats``                 – GS-2M's densification bookkeeping (max_radii2D, xyz_gradient_accum{,_abs}, denom,
  observe_cnt; train.py:225-245, scene/gaussian_model.py:569-573) with fused per-view updates, MAX / SUM all-reduced.
* ``ViewShardedStep``                    – runs forward+backward for this rank's views through caller-supplied callables and
  reduces the gradients over the ranks.  The callables are the only pieces that touch CUDA, so the sharding / accumulation /
  reduction logic is testable on CPU with the gloo backend.  Two protocols: ``render_view`` (a view's whole forward+backward
  in one call; one all-reduce at the end of the step) and ``begin_view`` / ``finish_view`` (the *deferred* step: a view's
  forward and reverse blend run as soon as possible, the per-Gaussian half of every view's backward runs afterwards, Gaussian
  range by Gaussian range, and the all-reduce of a finished range overlaps the next range's kernels).  With ``finish_views`` a
  range's per-Gaussian stage is one call for ALL of the rank's views (``backward_views_raw``: every output element written once
  with the sum over the views); ``assignment`` replaces the contiguous split by a table (``balance_views``),
  ``max_views_in_flight`` / ``max_steps_ahead`` bound the views that hold their arenas / the steps a non-waiting host may queue.
* ``balance_views(costs, world)``        – equal view counts per rank, per-view costs (instance counts) dealt longest first.

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


def balance_views(costs: List[float], world: int) -> List[List[int]]:
    """Assignment of ``len(costs)`` views to ``world`` ranks with equal view counts (the first ``n % world`` ranks get one more,
    like :func:`shard_views`) and nearly equal summed cost: longest-processing-time-first, ties and order resolved by view index
    so every rank computes the same table from the same costs.  A view's cost is its instance count (tile-list entries), which a
    training loop knows from the camera's previous visit; a step waits for its slowest rank at every collective, so the spread
    of the ranks' sums is lost time."""
    n = len(costs)
    base, extra = divmod(n, world)
    cap = [base + (1 if r < extra else 0) for r in range(world)]
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for v in sorted(range(n), key=lambda v: (-float(costs[v]), v)):
        r = min((r for r in range(world) if len(out[r]) < cap[r]), key=lambda r: (load[r], r))
        out[r].append(v)
        load[r] += float(costs[v])
    return [sorted(vs) for vs in out]


def _dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _row_chunks(P: int, n_chunks: int):
    """[(begin, end)] splitting P rows into n_chunks ranges whose starts are multiples of 256 (the per-Gaussian backward's block
    size).  The ranges SHRINK towards the end (weights n, n-1, ..., 1): the all-reduce of a range runs under the kernels of the
    ranges after it, so only the last range's exchange is exposed — it gets the smallest range (1 / (n (n+1) / 2) of the rows,
    10 % for n = 4), while the large early ranges have the most work left to hide behind."""
    if P <= 0:
        return []
    n = max(n_chunks, 1)
    total = n * (n + 1) // 2
    bounds, acc = [0], 0
    for k in range(n, 0, -1):
        acc += k
        b = min(P, -(-(P * acc // total) // 256) * 256)
        if k == 1:
            b = P
        if b > bounds[-1]:
            bounds.append(b)
    if bounds[-1] != P:
        bounds.append(P)
    return [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)]


def _record_stream(obj, stream):
    """Tell the caching allocator that the tensors of a view handle (dict values, and the arenas of a rasterizer state) are
    also used on `stream`, so that freeing them does not hand their memory out while that stream still reads it."""
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
    elif isinstance(obj, dict):
        for v in obj.values():
            _record_stream(v, stream)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _record_stream(v, stream)
    else:
        for name in ("geom", "binning", "img"):
            t = getattr(obj, name, None)
            if isinstance(t, torch.Tensor):
                _record_stream(t, stream)


def _all_reduce_many(tensors, async_op):
    """SUM all-reduce of several tensors as ONE collective launch where the backend can coalesce them (NCCL group call)."""
    if not tensors:
        return []
    try:
        with dist._coalescing_manager(async_ops=async_op) as cm:
            for t in tensors:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [cm] if async_op else []
    except (AttributeError, RuntimeError, ValueError):
        works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=async_op) for t in tensors]
        return works if async_op else []


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
        # gradients nobody consumes here (those of the precomputed colour / covariance inputs, the conic scratch) are not
        # computed at all: the C-ABI takes NULL for them
        for n in shapes:
            if n not in self.tensors:
                self.tensors[n] = None
        self.views_accumulated = 0

    def zero_(self):
        self.flat.zero_()
        self.views_accumulated = 0

    def nbytes_reduced(self) -> int:
        return sum(self.tensors[n].numel() * 4 for n in self.names)

    def all_reduce(self, async_op: bool = False):
        """SUM over ranks: one collective over the flat buffer. Returns the work handle when ``async_op``."""
        if not _dist_ready():
            return []
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return [work] if async_op else []

    def all_reduce_rows(self, begin: int, end: int, async_op: bool = True):
        """SUM over ranks of the Gaussians [begin, end) of every reduced tensor (one coalesced collective)."""
        if not _dist_ready() or end <= begin:
            return []
        return _all_reduce_many([self.tensors[n][begin:end] for n in self.names], async_op)

    def begin_rows(self):
        """Called once per step before the deferred per-Gaussian stage (nothing to prepare: the first view overwrites)."""

    def zero_rows(self, begin: int, end: int):
        for n in self.names:
            self.tensors[n][begin:end].zero_()


class ParameterBuckets:
    """Gradients of GS-2M's nine trained parameter groups (64 floats per Gaussian at M = 16) as views of
    one flat buffer: ``xyz (P,3)``, ``sh (P,M,3)`` (``_features_dc`` and ``_features_rest`` side by side, as ``get_features``
    concatenates them), ``scaling (P,3)``, ``rotation (P,4)``, ``opacity (P,1)``, ``albedo (P,3)``, ``roughness (P,1)``,
    ``metallic (P,1)`` — all w.r.t. the RAW (pre-activation) tensors.

    ``features``, the normals inside it and the activated scale / rotation / opacity reach the rasterizer through the
    camera-dependent packing stage (gaussian_renderer/__init__.py:49-96), so their gradients cannot be summed over views
    before the chain rule: every view runs  rasterizer backward (``accumulate=2``: ``dL_dmeans3D`` and ``dL_dsh`` add straight
    into ``xyz`` / ``sh``, the view-dependent rest is overwritten in the per-view scratch)  ->  ``chain_view`` (fused packing
    backward with ``+=`` into the other six groups and ``xyz``).  One SUM all-reduce over the flat buffer ends the step.
    """

    names = ("xyz", "sh", "scaling", "rotation", "opacity", "albedo", "roughness", "metallic")

    def __init__(self, P: int, M: int, device, names: Sequence[str] = ()):
        shapes = {"xyz": (P, 3), "sh": (P, M, 3), "scaling": (P, 3), "rotation": (P, 4), "opacity": (P, 1), "albedo": (P, 3),
                  "roughness": (P, 1), "metallic": (P, 1)}
        offsets, total = {}, 0
        for n in self.names:
            offsets[n] = total
            numel = 1
            for d in shapes[n]:
                numel *= d
            total += (numel + 31) // 32 * 32                     # 128-byte granularity
            shapes[n] = (shapes[n], numel)
        self.flat = torch.zeros(max(total, 1), dtype=torch.float32, device=device)
        self.tensors: Dict[str, torch.Tensor] = {
            n: self.flat[offsets[n]:offsets[n] + shapes[n][1]].view(shapes[n][0]) for n in self.names}
        self._packed_from = offsets["scaling"]                   # everything behind xyz and sh comes from chain_view
        # per-view scratch: the rasterizer-side gradients that are consumed by chain_view (or unused by GS-2M)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)  # noqa: E731
        # (dL_dconic / dL_dcolor / dL_dcov3D have no consumer in GS-2M's chain: NULL, the kernel skips them)
        self.raster = {"dL_dmeans3D": self.tensors["xyz"], "dL_dsh": self.tensors["sh"], "dL_dmeans2D": z(P, 4),
                       "dL_dconic": None, "dL_dopacity": z(P, 1), "dL_dcolor": None, "dL_dcov3D": None,
                       "dL_dscale": z(P, 3), "dL_drot": z(P, 4), "dL_dfeatures": z(P, 10)}
        self.views_accumulated = 0

    def zero_(self):
        self.flat.zero_()
        self.views_accumulated = 0

    def nbytes_reduced(self) -> int:
        return sum(self.tensors[n].numel() * 4 for n in self.names)

    def raster_accumulate_mode(self, accumulate: bool) -> int:
        """``accumulate`` argument of ``backward_raw`` for this view; the first view of a step overwrites ``xyz`` / ``sh``
        (the kernel writes every element) and clears the groups ``chain_view`` adds into."""
        if not accumulate:
            self.flat[self._packed_from:].zero_()
            return 0
        return 2

    def chain_view(self, raw: Dict[str, torch.Tensor], world_view_transform, camera_center, radii, z_depth=False,
                   blend_metallic=False):
        """Packing-stage chain rule of one view, ``+=`` into the raw-parameter gradients (Gaussians the view culled are skipped)."""
        from diff_gaussian_rasterization import packing
        t, r = self.tensors, self.raster
        packing.pack_backward_accumulate(
            raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"], raw["albedo"], raw["roughness"], raw["metallic"],
            world_view_transform, camera_center, r["dL_dscale"], r["dL_drot"], r["dL_dopacity"], r["dL_dfeatures"],
            t["xyz"], t["scaling"], t["rotation"], t["opacity"], t["albedo"], t["roughness"], t["metallic"], radii,
            z_depth=z_depth, blend_metallic=blend_metallic)

    def all_reduce(self, async_op: bool = False):
        if not _dist_ready():
            return []
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return [work] if async_op else []

    def all_reduce_rows(self, begin: int, end: int, async_op: bool = True):
        if not _dist_ready() or end <= begin:
            return []
        return _all_reduce_many([self.tensors[n][begin:end] for n in self.names], async_op)

    fused_chain = False      # set when every view goes through ``chain_spec`` (the first view of a range then overwrites all groups)

    def begin_rows(self):
        """Deferred step: the groups ``chain_view`` / ``chain_rows`` add into start the step at zero (``xyz`` / ``sh`` are
        overwritten by the first view's rasterizer backward); nothing to do when the kernel-side chain is used."""
        if not self.fused_chain:
            self.flat[self._packed_from:].zero_()

    def zero_rows(self, begin: int, end: int):
        for n in self.names:
            self.tensors[n][begin:end].zero_()

    def chain_spec(self, raw: Dict[str, torch.Tensor], z_depth=False, blend_metallic=False) -> dict:
        """``chain=`` argument of ``backward_raw``: the rasterizer's per-Gaussian backward then chains its gradients through the
        packing stage of the view's camera itself and adds them to this bucket set (no scratch round trip, no ``chain_rows``)."""
        return {"raw": raw, "grads": self.tensors, "z_depth": z_depth, "blend_metallic": blend_metallic}

    def chain_rows(self, raw: Dict[str, torch.Tensor], world_view_transform, camera_center, radii, begin: int, end: int,
                   z_depth=False, blend_metallic=False):
        """``chain_view`` for the Gaussians [begin, end) only (the packing backward is per Gaussian: row slices of every
        operand)."""
        from diff_gaussian_rasterization import packing
        t, r = self.tensors, self.raster
        sl = slice(begin, end)
        packing.pack_backward_accumulate(
            raw["xyz"][sl], raw["scaling"][sl], raw["rotation"][sl], raw["opacity"][sl], raw["albedo"][sl], raw["roughness"][sl],
            raw["metallic"][sl], world_view_transform, camera_center, r["dL_dscale"][sl], r["dL_drot"][sl], r["dL_dopacity"][sl],
            r["dL_dfeatures"][sl], t["xyz"][sl], t["scaling"][sl], t["rotation"][sl], t["opacity"][sl], t["albedo"][sl],
            t["roughness"][sl], t["metallic"][sl], radii[sl], z_depth=z_depth, blend_metallic=blend_metallic)


class DensificationStats:
    """GS-2M's densification bookkeeping for one step, reference semantics and dtypes (float tensors):

    * ``max_radii2D [P]``            ``where((observe > 0) & (radii > 0), max(., radii), .)``      train.py:225-227   MAX over ranks
    * ``xyz_gradient_accum [P,1]``   ``+= |dL_dmeans2D.xy|`` per view, visible Gaussians          gaussian_model.py:569-571   SUM
    * ``xyz_gradient_accum_abs``     ``+= |dL_dmeans2D.zw|`` (AbsGS)                               :572   SUM
    * ``denom [P,1]``                ``+= 1`` per view, visible Gaussians                          :573   SUM
    * ``observe_cnt [P,1]``          ``[observe > 0] += 1`` per view (multi-view trim)             train.py:238-241   SUM

    All five are views of one flat buffer: one MAX and one SUM all-reduce.  On CUDA the per-view updates are fused into the
    library (``update_view_stats`` after the forward; the gradient norms inside the backward, which sees the view's own
    gradient even when the buckets accumulate) and are atomic, so views on different streams may share the buffers.
    On CPU (gloo tests of the reduction logic) the same updates are plain torch ops.
    """

    def __init__(self, P: int, device):
        self.P = P
        self.flat = torch.zeros(5 * max(P, 1), dtype=torch.float32, device=device)
        self.max_radii2D = self.flat[0:P]
        self.xyz_gradient_accum = self.flat[P:2 * P].view(P, 1)
        self.xyz_gradient_accum_abs = self.flat[2 * P:3 * P].view(P, 1)
        self.denom = self.flat[3 * P:4 * P].view(P, 1)
        self.observe_cnt = self.flat[4 * P:5 * P].view(P, 1)

    def zero_(self):
        self.flat.zero_()

    def backward_args(self):
        """``densify_stats`` argument of ``backward_raw``."""
        return (self.xyz_gradient_accum, self.xyz_gradient_accum_abs, self.denom)

    def update_forward(self, radii: torch.Tensor, observe: torch.Tensor):
        if radii.is_cuda:
            import diff_gaussian_rasterization as dgr
            dgr.update_view_stats(radii, observe, self.max_radii2D, self.observe_cnt.view(-1))
        else:
            seen = observe > 0
            mask = seen & (radii > 0)
            torch.where(mask, torch.maximum(self.max_radii2D, radii.to(torch.float32)), self.max_radii2D, out=self.max_radii2D)
            self.observe_cnt.view(-1)[seen] += 1

    def update_backward_eager(self, means2D_grad: torch.Tensor, radii: torch.Tensor):
        """``add_densification_stats`` with torch ops (callers that do not use the fused path)."""
        vis = radii > 0
        self.xyz_gradient_accum[vis] += torch.norm(means2D_grad[vis, :2], dim=-1, keepdim=True)
        self.xyz_gradient_accum_abs[vis] += torch.norm(means2D_grad[vis, 2:], dim=-1, keepdim=True)
        self.denom[vis] += 1

    def all_reduce(self):
        if _dist_ready():
            dist.all_reduce(self.flat[:self.P], op=dist.ReduceOp.MAX)
            dist.all_reduce(self.flat[self.P:], op=dist.ReduceOp.SUM)

    def all_reduce_rows(self, begin: int, end: int, async_op: bool = True):
        """MAX / SUM over ranks of the Gaussians [begin, end) of the five statistics (the deferred step queues this next to the
        gradients of the same range, so that it too runs under the following range's kernels)."""
        if not _dist_ready() or end <= begin:
            return []
        w = dist.all_reduce(self.flat[begin:end], op=dist.ReduceOp.MAX, async_op=async_op)
        works = [w] if async_op else []
        P = self.P
        return works + _all_reduce_many([self.flat[k * P + begin:k * P + end] for k in range(1, 5)], async_op)


class ViewShardedStep:
    """One data-parallel rasterization step over a batch of views.

    ``render_view(view_index, buckets, accumulate) -> dict`` must run forward+backward of one view, adding its
    gradients into ``buckets.tensors`` (``accumulate`` is False for the first view that writes a bucket set: the kernels
    then overwrite, which saves zero-filling), and may return the view's ``{"radii": int32[P], "observe": int32[P]}`` for
    the densification statistics in ``self.stats`` (:class:`DensificationStats`); to get the gradient-norm statistics it
    passes ``densify_stats=step.stats.backward_args()`` to ``backward_raw`` (or returns ``"means2D_grad"``, the view's own
    ``[P,4]`` gradient, for the eager update).

    With ``n_streams > 1`` (CUDA only) consecutive views of a rank are issued round-robin on that many streams, each with
    its own bucket set, so the kernels of view k+1 fill the GPU while view k sits in the forward's instance-count
    read-back or in a tail wave; the bucket sets are summed once per step before the all-reduce.
    """

    def __init__(self, P: int, M: int, device, render_view: Optional[Callable[[int, GradientBuckets, bool], Optional[dict]]] = None,
                 world: Optional[int] = None, rank: Optional[int] = None, n_streams: int = 1, buckets_cls=None,
                 begin_view: Optional[Callable[[int], dict]] = None,
                 finish_view: Optional[Callable[[dict, object, bool, tuple], None]] = None, n_chunks: int = 4, buckets=None,
                 finish_views: Optional[Callable[[list, object, tuple], None]] = None,
                 assignment: Optional[List[List[int]]] = None, max_views_in_flight: Optional[int] = None,
                 max_steps_ahead: Optional[int] = None):
        self.world = world if world is not None else (dist.get_world_size() if _dist_ready() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if _dist_ready() else 0)
        self.device = torch.device(device)
        use_streams = n_streams > 1 and self.device.type == "cuda"
        self.n_streams = n_streams if use_streams else 1
        self.deferred = begin_view is not None
        if self.deferred == (render_view is not None) or (self.deferred and finish_view is None and finish_views is None):
            raise ValueError("give either render_view, or begin_view together with finish_view / finish_views")
        self.begin_view, self.finish_view, self.P = begin_view, finish_view, P
        # finish_views(handles, buckets, (begin, end)): the per-Gaussian stage of ALL the rank's views for one Gaussian range
        # in one call (``backward_views_raw``: every output element written once with the sum over the views) — preferred
        # over the per-view ``finish_view`` loop when given
        self.finish_views = finish_views
        # optional table of view indices per rank (``balance_views``) replacing the contiguous split
        self.assignment = assignment
        self.max_views_in_flight = max_views_in_flight      # deferred step: views between their two phases at any time
        self.max_steps_ahead = max_steps_ahead              # deferred step: steps the host may queue ahead of the running one
        self._step_done: list = []
        self.chunks = _row_chunks(P, n_chunks)
        buckets_cls = buckets_cls or GradientBuckets        # ParameterBuckets: raw-parameter gradients, chained per view
        # the deferred step sums all of a rank's views in ONE bucket set (its per-Gaussian stage runs on one stream)
        # (`buckets`: an existing bucket set to accumulate into instead of allocating one)
        n_sets = 1 if self.deferred else self.n_streams
        self.bucket_sets = [buckets] if (buckets is not None and n_sets == 1) else [buckets_cls(P, M, device) for _ in range(n_sets)]
        self.buckets = self.bucket_sets[0]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n_streams)] if use_streams else [None]
        self.render_view = render_view
        self.stats = DensificationStats(P, device)

    def local_views(self, n_views: int) -> List[int]:
        if self.assignment is not None:
            if sorted(v for vs in self.assignment for v in vs) != list(range(n_views)) or len(self.assignment) != self.world:
                raise ValueError("the assignment must place each of the %d views on exactly one of the %d ranks" % (n_views, self.world))
            return list(self.assignment[self.rank])
        return list(shard_views(n_views, self.world, self.rank))

    def _one_view(self, v, buckets, accumulate):
        out = self.render_view(v, buckets, accumulate)
        if out and "radii" in out and "observe" in out:
            self.stats.update_forward(out["radii"], out["observe"])
            if "means2D_grad" in out:
                self.stats.update_backward_eager(out["means2D_grad"], out["radii"])

    def _run_deferred(self, mine: List[int], reduce: bool) -> Dict[str, torch.Tensor]:
        """Phase A: every view's forward + reverse blend (``begin_view``), round-robin over the streams.  Phase B: the
        per-Gaussian stage of every view (``finish_views(handles, buckets, (begin, end), accumulate)``, or the per-view
        ``finish_view(handle, buckets, accumulate, (begin, end))``), Gaussian range by Gaussian range on the main stream; as
        soon as a range has received all of the rank's views its all-reduce is queued (NCCL's own stream), so it runs under
        the next range's kernels and only the last range's exchange is exposed.

        A view between its two phases holds its three arenas (a few GB at several million Gaussians).  With
        ``max_views_in_flight`` the rank's views go through both phases in groups of that many, the later groups adding
        to the buckets; the exchange then follows the last group."""
        cuda = self.device.type == "cuda"
        multi = cuda and self.n_streams > 1 and bool(mine)
        cap = self.max_views_in_flight or len(mine) or 1
        groups = [mine[i:i + cap] for i in range(0, len(mine), cap)] or [[]]
        main = torch.cuda.current_stream(self.device) if multi else None
        self.buckets.begin_rows()
        works = []
        for gi, group in enumerate(groups):
            handles = []
            if multi:
                ready = torch.cuda.Event()
                ready.record(main)
                for k, v in enumerate(group):
                    j = k % self.n_streams
                    with torch.cuda.stream(self.streams[j]):
                        if k < self.n_streams:
                            self.streams[j].wait_event(ready)
                        handles.append(self._begin(v))
                for j in range(min(self.n_streams, len(group))):
                    main.wait_stream(self.streams[j])
                for h in handles:                                 # allocated on a side stream, consumed on the main one
                    _record_stream(h, main)
            else:
                handles = [self._begin(v) for v in group]
            last = gi == len(groups) - 1
            for (b, e) in self.chunks:
                if not handles and gi == 0:
                    self.buckets.zero_rows(b, e)                  # more ranks than views: contribute zeros
                if handles and self.finish_views is not None:
                    if gi > 0:
                        self.finish_views(handles, self.buckets, (b, e), True)
                    else:
                        self.finish_views(handles, self.buckets, (b, e))
                else:
                    for k, h in enumerate(handles):
                        self.finish_view(h, self.buckets, k > 0 or gi > 0, (b, e))
                if reduce and last:
                    works += self.buckets.all_reduce_rows(b, e, async_op=True)
                    works += self.stats.all_reduce_rows(b, e, async_op=True)
            del handles
        for w in works:
            w.wait()                                              # stream-level wait: later work sees the reduced gradients
        if multi:
            for j in range(min(self.n_streams, len(mine))):       # the side streams' next step must wait for this one
                self.streams[j].wait_stream(main)
        self.buckets.views_accumulated = len(mine)
        return self.buckets.tensors

    def _begin(self, v):
        h = self.begin_view(v)
        if h and "radii" in h and "observe" in h:
            self.stats.update_forward(h["radii"], h["observe"])
        return h

    def run(self, n_views: int, reduce: bool = True) -> Dict[str, torch.Tensor]:
        mine = self.local_views(n_views)
        if self.max_steps_ahead and self.device.type == "cuda":
            # a host that never waits (no_wait forwards) may queue no more than `max_steps_ahead` steps behind the one the GPU
            # is working on: every queued step holds its views' arenas, and a caching allocator that has to grow in the middle
            # of a run (cudaMalloc synchronises the device) costs far more than the wait
            while len(self._step_done) >= self.max_steps_ahead + 1:
                self._step_done.pop(0).synchronize()
        self.stats.zero_()
        if self.deferred:
            out = self._run_deferred(mine, reduce)
            if self.max_steps_ahead and self.device.type == "cuda":
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                self._step_done.append(ev)
            return out
        if not mine:  # more ranks than views: contribute zeros
            self.buckets.zero_()
        if self.n_streams == 1:
            for k, v in enumerate(mine):
                self._one_view(v, self.buckets, k > 0)
                self.buckets.views_accumulated = k + 1
        else:
            main = torch.cuda.current_stream(self.device)
            ready = torch.cuda.Event()
            ready.record(main)
            used = min(self.n_streams, len(mine))
            for k, v in enumerate(mine):
                j = k % self.n_streams
                with torch.cuda.stream(self.streams[j]):
                    if k < self.n_streams:
                        self.streams[j].wait_event(ready)        # inputs / previous step's consumers are done
                    self._one_view(v, self.bucket_sets[j], k >= self.n_streams)   # statistics updates are atomic
            for j in range(used):
                main.wait_stream(self.streams[j])
            for j in range(1, used):                                # fold the per-stream partial sums into set 0
                self.buckets.flat += self.bucket_sets[j].flat
            for j in range(used):                                   # later work on the side streams must wait for the fold
                self.streams[j].wait_stream(main)
            self.buckets.views_accumulated = len(mine)
        if reduce:
            self.buckets.all_reduce()
            self.stats.all_reduce()
        return self.buckets.tensors
