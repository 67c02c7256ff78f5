"""B200-native drop-in for GS-2M's ``diff_gaussian_rasterization`` package.

Public surface (identical names, argument order, return tuples and gradient slots to the reference binding,
``submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py:17-218``):

* ``GaussianRasterizationSettings``  — 12-field NamedTuple (reference ``:143-155``)
* ``GaussianRasterizer(raster_settings)(means3D, means2D, opacities, shs=, colors_precomp=, scales=, rotations=,
  cov3D_precomp=, features=) -> (color[3,H,W], radii[P] i32, observe[P] i32, buffer[10,H,W])`` and
  ``.markVisible(positions) -> bool[P]``  (reference ``:157-218``)
* ``rasterize_gaussians(...)``  (reference ``:17-40``)

Underneath, tensors are allocated here with torch and handed as raw device pointers to the hand-written sm_100a
CUDA library through the C-ABI of ``include/gs2m_rasterizer.h`` (ctypes, no pybind / libtorch in the library).
Kernels run on torch's *current stream of the inputs' device* (the reference uses the legacy default stream), so
one-process-per-GPU data parallelism works unchanged.  There is no CPU path.

Extras beyond the reference surface (used by the view-sharded data-parallel step and by the parity tests):
``forward_raw`` / ``backward_raw`` (explicit-state calls with in-place gradient accumulation) and ``state_view``.
"""
import os
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _native
from ._native import NUM_CHANNELS, NUM_FEATURES, RasterizerError  # noqa: F401

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "forward_raw", "backward_raw", "state_view", "RasterizerError"]


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    feature_count: int


# ----------------------------------------------------------------------------------------------------------------
# marshalling helpers
# ----------------------------------------------------------------------------------------------------------------
def _absent(t: Optional[torch.Tensor]) -> bool:
    return t is None or t.numel() == 0


def _dev_f32(t: torch.Tensor, device: torch.device, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (got %s)" % (name, t.dtype))
    if t.device != device:
        raise RuntimeError("%s lives on %s but means3D lives on %s" % (name, t.device, device))
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _Arena:
    """One growable byte buffer handed to the library through a resize callback
    (the role of ``resizeFunctional`` in the reference glue, rasterize_points.cu:22-28)."""

    def __init__(self, device):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.callback = _native.RESIZE_FN(self._resize)

    def _resize(self, _user, nbytes):
        try:
            if self.tensor.numel() < nbytes:
                # round large arenas up to 8 MiB so that consecutive views (whose instance count differs a little)
                # hit the same block of torch's caching allocator instead of triggering cudaMalloc
                n = int(nbytes)
                if n > (8 << 20):
                    n = (n + (8 << 20) - 1) & ~((8 << 20) - 1)
                self.tensor = torch.empty(n, dtype=torch.uint8, device=self.device)
            return self.tensor.data_ptr()
        except Exception:  # allocation failure -> NULL -> GS2M_ERR_ALLOC
            return None


class RasterState:
    """What backward needs from forward (the reference keeps the same things in ``ctx``, binding ``:86-88``).
    ``capacity`` is the instance capacity the binning arena was sized for when the forward ran speculatively (0: it was sized
    for ``num_rendered`` itself); ``backward_runs`` counts the backward passes already made over this state (the library
    clears its accumulator again from the second one on)."""
    __slots__ = ("num_rendered", "geom", "binning", "img", "capacity", "backward_runs", "prepared")

    def __init__(self, num_rendered, geom, binning, img, capacity=0, prepared=True):
        self.num_rendered, self.geom, self.binning, self.img = num_rendered, geom, binning, img
        self.capacity, self.backward_runs, self.prepared = capacity, 0, prepared


# Instance-count hints for the speculative forward, per (device, width, height): the largest count of the recent calls.
# The binning arena is sized for the hint plus a margin and nothing waits for the host; a view that needs more is detected
# by the library (GS2M_ERR_CAPACITY) and simply re-run in exact mode.  GS2M_EXACT_BINNING=1 disables the speculation.
_R_HINT = {}
_R_MARGIN = 1.25


def _quantize_capacity(c):
    """Round an instance capacity up to one of 8-16 steps per octave (at least 64 Ki apart).  The hint below moves with every
    view, and an arena whose size is new to torch's caching allocator costs a cudaMalloc of ~1 GB in the middle of a step
    (measured: single 35-45 ms steps among 31 ms ones); a small set of sizes is cached once and for all."""
    c = int(c)
    if c <= 0:
        return 0
    step = 1 << max(c.bit_length() - 4, 16)
    return min(-(-c // step) * step, (1 << 30) - 1)


def _capacity_for(key):
    if os.environ.get("GS2M_EXACT_BINNING") == "1" or os.environ.get("GS2M_BINNING", "depthfirst") != "depthfirst":
        return 0
    hint = _R_HINT.get(key, 0)
    if hint <= 0:
        return 0
    return _quantize_capacity(min(int(hint * _R_MARGIN) + 65536, (1 << 30) - 1))


def _note_instances(key, R):
    # decay slowly so that one outlier view does not pin the arena size forever
    _R_HINT[key] = max(int(R), int(_R_HINT.get(key, 0) * 0.98))


def forward_raw(means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, features,
                raster_settings: GaussianRasterizationSettings, for_backward=True, capacity=None, no_wait=False):
    """Run the forward pipeline. Returns ``(color, radii, observe, buffer, RasterState)``.

    ``capacity`` (instances): None = automatic (speculative once a previous call on this device and image size has told the
    instance count, exact otherwise), 0 = exact mode (host reads the count back before the binning arena is sized, like the
    reference), > 0 = speculative with that capacity.  ``no_wait`` (needs an explicit capacity) never touches the host, so
    the call can be captured in a CUDA graph; ``state.num_rendered`` is then the capacity and the caller checks
    ``state_view(...)["bin_info"]``.  ``for_backward=False`` skips preparing the backward accumulator (inference)."""
    lib = _native.load()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:52-54
    if not means3D.is_cuda:
        raise RuntimeError("diff_gaussian_rasterization (B200) has no CPU path: means3D must be a CUDA tensor")
    dev = means3D.device
    rs = raster_settings
    P = int(means3D.shape[0])
    H, W = int(rs.image_height), int(rs.image_width)
    F = int(rs.feature_count)

    means3D = _dev_f32(means3D, dev, "means3D")
    shs_t = None if _absent(shs) else _dev_f32(shs, dev, "shs")
    col_t = None if _absent(colors_precomp) else _dev_f32(colors_precomp, dev, "colors_precomp")
    opa_t = None if _absent(opacities) else _dev_f32(opacities, dev, "opacities")
    sca_t = None if _absent(scales) else _dev_f32(scales, dev, "scales")
    rot_t = None if _absent(rotations) else _dev_f32(rotations, dev, "rotations")
    cov_t = None if _absent(cov3D_precomp) else _dev_f32(cov3D_precomp, dev, "cov3D_precomp")
    fea_t = None if _absent(features) else _dev_f32(features, dev, "features")
    bg = _dev_f32(rs.bg, dev, "bg")
    vm = _dev_f32(rs.viewmatrix, dev, "viewmatrix")
    pm = _dev_f32(rs.projmatrix, dev, "projmatrix")
    cam = _dev_f32(rs.campos, dev, "campos")
    if fea_t is not None and (fea_t.dim() != 2 or fea_t.shape[1] != NUM_FEATURES):
        raise RuntimeError("features must have dimensions (num_points, %d)" % NUM_FEATURES)
    M = int(shs_t.shape[1]) if shs_t is not None else 0

    key = (dev.index, W, H)
    auto = capacity is None
    if auto:
        capacity = _capacity_for(key) if P > 0 else 0
    if no_wait and capacity <= 0:
        raise RuntimeError("no_wait needs an explicit instance capacity")
    with torch.cuda.device(dev):
        color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
        buffer = torch.empty((NUM_FEATURES, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        observe = torch.empty((P,), dtype=torch.int32, device=dev)
        geom, binning, img = _Arena(dev), _Arena(dev), _Arena(dev)
        a = _native.ForwardArgs()
        a.geometry_buffer, a.binning_buffer, a.image_buffer = geom.callback, binning.callback, img.callback
        a.R_capacity, a.no_wait, a.no_backward = int(capacity), int(bool(no_wait)), int(not for_backward)
        a.P, a.D, a.M = P, int(rs.sh_degree), M
        a.background = _ptr(bg)
        a.width, a.height = W, H
        a.means3D, a.shs, a.colors_precomp, a.opacities = _ptr(means3D), _ptr(shs_t), _ptr(col_t), _ptr(opa_t)
        a.scales, a.scale_modifier, a.rotations = _ptr(sca_t), float(rs.scale_modifier), _ptr(rot_t)
        a.cov3D_precomp, a.features = _ptr(cov_t), _ptr(fea_t)
        a.viewmatrix, a.projmatrix, a.cam_pos = _ptr(vm), _ptr(pm), _ptr(cam)
        a.tan_fovx, a.tan_fovy = float(rs.tanfovx), float(rs.tanfovy)
        a.prefiltered, a.feature_count = int(bool(rs.prefiltered)), F
        a.out_color, a.out_radii, a.out_observe, a.out_buffer = _ptr(color), _ptr(radii), _ptr(observe), _ptr(buffer)
        a.stream = torch.cuda.current_stream(dev).cuda_stream
        try:
            try:
                num_rendered = _native.check(lib.gs2m_rasterize_forward(a), "gs2m_rasterize_forward")
            except RasterizerError as e:
                if not (auto and e.code == _native.ERR_CAPACITY):
                    raise
                # this view has more instances than the hint allowed for: run it again with the exact count
                a.R_capacity = capacity = 0
                num_rendered = _native.check(lib.gs2m_rasterize_forward(a), "gs2m_rasterize_forward")
        finally:
            # the ctypes callback holds a bound method of its arena: break that reference cycle now so the arenas
            # (hundreds of MB each) are released by reference counting, not whenever the cyclic GC next runs
            a.geometry_buffer = a.binning_buffer = a.image_buffer = _native.RESIZE_FN()
            geom.callback = binning.callback = img.callback = None
    if not no_wait and P > 0:
        _note_instances(key, num_rendered)
    state = RasterState(num_rendered, geom.tensor, binning.tensor, img.tensor, capacity=int(capacity), prepared=bool(for_backward))
    return color, radii, observe, buffer, state


_GRAD_SHAPES = (("dL_dmeans2D", 4), ("dL_dconic", 4), ("dL_dopacity", 1), ("dL_dcolor", 3), ("dL_dmeans3D", 3),
                ("dL_dcov3D", 6), ("dL_dscale", 3), ("dL_drot", 4), ("dL_dfeatures", NUM_FEATURES))


_OPTIONAL_GRADS = ("dL_dconic", "dL_dcolor", "dL_dcov3D")     # may be None in `grads`: then the kernel does not write them


def alloc_grads(P, M, device, zero=False, skip=()):
    """Gradient tensors in the reference's shapes (rasterize_points.cu:150-159). The kernels write every element,
    so ``torch.empty`` suffices unless the caller wants to accumulate several views (then start from zeros).  ``skip`` names
    optional tensors to leave out (``dL_dconic``, and ``dL_dcolor`` / ``dL_dcov3D`` when SHs / scale+rotation are the inputs)."""
    mk = torch.zeros if zero else torch.empty
    g = {n: (None if n in skip else mk((P, c), dtype=torch.float32, device=device)) for n, c in _GRAD_SHAPES}
    g["dL_dsh"] = mk((P, M, 3), dtype=torch.float32, device=device)
    return g


def backward_raw(grad_color, grad_buffer, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, features,
                 radii, raster_settings: GaussianRasterizationSettings, state: RasterState, grads=None,
                 accumulate=False, densify_stats=None, phase="all", rows=None, chain=None):
    """Run the backward pipeline into ``grads`` (dict from :func:`alloc_grads`; allocated when None).
    With ``accumulate=True`` (or 1) the nine caller-visible tensors are updated with ``+=``; with ``accumulate=2`` only
    ``dL_dmeans3D`` and ``dL_dsh`` (GS-2M's raw, view-independent parameters) are, the rest is overwritten.
    ``densify_stats = (xyz_gradient_accum, xyz_gradient_accum_abs, denom)`` (float ``[P]`` / ``[P,1]`` CUDA tensors, any may be
    None) are updated like ``GaussianModel.add_densification_stats`` with this view's screen-space gradient.
    ``phase``: "all" (default), "blend" (only the reverse blend, which fills the state's internal accumulator) or "gaussians"
    (only the per-Gaussian stage, for Gaussians ``rows = (begin, end)`` when given; ``begin`` a multiple of 256) — a view-sharded
    step runs "blend" per view and defers "gaussians" range by range so that finished ranges can be all-reduced early.
    ``chain = {"raw": {scaling, rotation, opacity, albedo, roughness, metallic}, "grads": {xyz, scaling, rotation, opacity, albedo,
    roughness, metallic}, "z_depth": bool, "blend_metallic": bool}``: chain the gradients w.r.t. the activated scale / rotation /
    opacity, the features and the 3-D mean through GS-2M's packing stage of this view's camera inside the kernel and add them to the
    raw-parameter gradients in ``grads`` (``accumulate`` 0: overwritten, 2: ``+=``); ``grads["dL_dscale"]`` etc. may then be None."""
    b, keep, grads = _backward_args(grad_color, grad_buffer, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, features,
                                    radii, raster_settings, state, grads, accumulate, densify_stats, phase, rows, chain)
    if b is not None:
        with torch.cuda.device(means3D.device):
            _native.check(_native.load().gs2m_rasterize_backward(b), "gs2m_rasterize_backward")
    return grads


def backward_views_raw(views, rows=None, accumulate=False):
    """The per-Gaussian stage (``phase="gaussians"``) of several views in ONE pass over the Gaussians ``rows`` (C-ABI
    ``gs2m_rasterize_backward_views``).  ``views`` is a list of dicts holding, per view, the arguments of :func:`backward_raw`
    (``grad_color, grad_buffer, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, features, radii, raster_settings,
    state`` and optionally ``grads``, ``densify_stats``) and the same ``chain`` for all of them.  ``dL_dsh`` and the seven
    raw-parameter gradients of the chain receive the SUM over the views, each element written once (``accumulate``: added to
    what is there) — where a loop of per-view calls read-modify-writes them once per view."""
    if not views:
        return
    built, keep_all = [], []
    for k, v in enumerate(views):
        b, keep, _ = _backward_args(v["grad_color"], v["grad_buffer"], v["means3D"], v["shs"], v.get("colors_precomp"), v["scales"],
                                    v["rotations"], v.get("cov3D_precomp"), v["features"], v["radii"], v["raster_settings"],
                                    v["state"], v.get("grads"), 2 if (accumulate or k > 0) else 0, v.get("densify_stats"),
                                    "gaussians", rows, v["chain"])
        if b is None:
            return
        built.append(b)
        keep_all.append(keep)
    arr = (_native.BackwardArgs * len(built))(*built)
    with torch.cuda.device(views[0]["means3D"].device):
        _native.check(_native.load().gs2m_rasterize_backward_views(arr, len(built)), "gs2m_rasterize_backward_views")


def _backward_args(grad_color, grad_buffer, means3D, shs, colors_precomp, scales, rotations, cov3D_precomp, features,
                   radii, raster_settings, state, grads, accumulate, densify_stats, phase, rows, chain):
    """Validates one backward call and fills its C argument block.  Returns ``(block or None when P == 0, objects the block
    points into, grads)``."""
    lib = _native.load()
    dev = means3D.device
    rs = raster_settings
    P = int(means3D.shape[0])
    H, W = int(rs.image_height), int(rs.image_width)
    means3D = _dev_f32(means3D, dev, "means3D")
    shs_t = None if _absent(shs) else _dev_f32(shs, dev, "shs")
    col_t = None if _absent(colors_precomp) else _dev_f32(colors_precomp, dev, "colors_precomp")
    sca_t = None if _absent(scales) else _dev_f32(scales, dev, "scales")
    rot_t = None if _absent(rotations) else _dev_f32(rotations, dev, "rotations")
    cov_t = None if _absent(cov3D_precomp) else _dev_f32(cov3D_precomp, dev, "cov3D_precomp")
    fea_t = None if _absent(features) else _dev_f32(features, dev, "features")
    bg = _dev_f32(rs.bg, dev, "bg")
    vm = _dev_f32(rs.viewmatrix, dev, "viewmatrix")
    pm = _dev_f32(rs.projmatrix, dev, "projmatrix")
    cam = _dev_f32(rs.campos, dev, "campos")
    M = int(shs_t.shape[1]) if shs_t is not None else 0
    if grad_color is None:
        grad_color = torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=dev)
    if grad_buffer is None:
        grad_buffer = torch.zeros((NUM_FEATURES, H, W), dtype=torch.float32, device=dev)
    grad_color = _dev_f32(grad_color, dev, "grad_color")
    grad_buffer = _dev_f32(grad_buffer, dev, "grad_buffer")
    if tuple(grad_color.shape) != (NUM_CHANNELS, H, W) or tuple(grad_buffer.shape) != (NUM_FEATURES, H, W):
        raise RuntimeError("upstream gradients must be (3,H,W) and (10,H,W)")
    if grads is None:
        if accumulate:
            raise RuntimeError("accumulate=True needs caller-provided gradient tensors")
        grads = alloc_grads(P, M, dev)
    for name, c in _GRAD_SHAPES + (("dL_dsh", 3 * M),):
        t = grads.get(name)
        if name == "dL_dsh" and M == 0:
            continue
        if t is None and (name == "dL_dconic" or (name == "dL_dcolor" and col_t is None) or (name == "dL_dcov3D" and cov_t is None)):
            continue
        if t is None and chain is not None and name in ("dL_dmeans2D", "dL_dopacity", "dL_dmeans3D", "dL_dscale", "dL_drot", "dL_dfeatures"):
            continue
        if t is None or t.dtype != torch.float32 or t.device != dev or not t.is_contiguous() or t.numel() != P * c:
            raise RuntimeError("grads[%r] must be a contiguous float32 tensor with %d x %d elements on %s" % (name, P, c, dev))
    if P == 0:
        return None, None, grads
    with torch.cuda.device(dev):
        b = _native.BackwardArgs()
        b.P, b.D, b.M, b.R, b.R_capacity = P, int(rs.sh_degree), M, int(state.num_rendered), int(state.capacity)
        b.phase = {"all": 0, "blend": 1, "gaussians": 2}[phase]
        if rows is not None:
            b.row_begin, b.row_end = int(rows[0]), int(rows[1])
        if phase != "gaussians":
            b.grad_acc_dirty = int(state.backward_runs > 0 or not state.prepared)
            state.backward_runs += 1
        b.background = _ptr(bg)
        b.width, b.height = W, H
        b.means3D, b.shs, b.colors_precomp, b.scales = _ptr(means3D), _ptr(shs_t), _ptr(col_t), _ptr(sca_t)
        b.scale_modifier, b.rotations, b.cov3D_precomp = float(rs.scale_modifier), _ptr(rot_t), _ptr(cov_t)
        b.features = _ptr(fea_t)
        b.viewmatrix, b.projmatrix, b.cam_pos = _ptr(vm), _ptr(pm), _ptr(cam)
        b.tan_fovx, b.tan_fovy = float(rs.tanfovx), float(rs.tanfovy)
        b.radii = _ptr(radii)
        b.geometry_buffer, b.binning_buffer, b.image_buffer = _ptr(state.geom), _ptr(state.binning), _ptr(state.img)
        b.geometry_bytes, b.binning_bytes, b.image_bytes = state.geom.numel(), state.binning.numel(), state.img.numel()
        b.feature_count = int(rs.feature_count)
        b.grad_color, b.grad_buffer = _ptr(grad_color), _ptr(grad_buffer)
        for name, _c in _GRAD_SHAPES:
            setattr(b, name, _ptr(grads[name]))
        b.dL_dsh = _ptr(grads["dL_dsh"]) if M > 0 else None
        b.accumulate = int(accumulate)
        b.stream = torch.cuda.current_stream(dev).cuda_stream
        if chain is not None:
            c = _native.ParamChain()
            raw, out = chain["raw"], chain["grads"]
            widths = {"xyz": 3, "scaling": 3, "rotation": 4, "opacity": 1, "albedo": 3, "roughness": 1, "metallic": 1}
            for n in ("scaling", "rotation", "opacity", "albedo", "roughness", "metallic"):
                t = raw[n]
                if t.dtype != torch.float32 or t.device != dev or not t.is_contiguous() or t.numel() != P * widths[n]:
                    raise RuntimeError("chain raw[%r] must be a contiguous float32 tensor with %d x %d elements on %s" % (n, P, widths[n], dev))
                setattr(c, n + "_raw", t.data_ptr())
            for n, w in widths.items():
                t = out[n]
                if t.dtype != torch.float32 or t.device != dev or not t.is_contiguous() or t.numel() != P * w:
                    raise RuntimeError("chain grads[%r] must be a contiguous float32 tensor with %d x %d elements on %s" % (n, P, w, dev))
                setattr(c, "d_xyz" if n == "xyz" else "d_" + n + "_raw", t.data_ptr())
            c.z_depth, c.blend_metallic = int(bool(chain.get("z_depth", False))), int(bool(chain.get("blend_metallic", False)))
            b.chain = _native.C.pointer(c)
        if densify_stats is not None:
            for t in densify_stats:
                if t is not None and (t.dtype != torch.float32 or t.numel() != P or not t.is_contiguous() or t.device != dev):
                    raise RuntimeError("densify_stats tensors must be contiguous float32 with P elements on the inputs' device")
            b.densify_grad_accum, b.densify_grad_accum_abs, b.densify_denom = (_ptr(t) for t in densify_stats)
    keep = (means3D, shs_t, col_t, sca_t, rot_t, cov_t, fea_t, bg, vm, pm, cam, grad_color, grad_buffer, grads, radii, state,
            chain, densify_stats, c if chain is not None else None)
    return b, keep, grads


def update_view_stats(radii, observe, max_radii2D=None, observe_cnt=None):
    """Per-view densification statistics of the forward outputs, fused (train.py:225-228, 238-241):
    ``max_radii2D = where((observe > 0) & (radii > 0), max(max_radii2D, radii), max_radii2D)``; ``observe_cnt[observe > 0] += 1``."""
    lib = _native.load()
    dev = radii.device
    if not radii.is_cuda:
        raise RuntimeError("update_view_stats has no CPU path: radii must be a CUDA tensor")
    P = int(radii.numel())
    for t in (max_radii2D, observe_cnt):
        if t is not None and (t.dtype != torch.float32 or t.numel() != P or not t.is_contiguous() or t.device != dev):
            raise RuntimeError("statistics tensors must be contiguous float32 with P elements on the inputs' device")
    with torch.cuda.device(dev):
        _native.check(lib.gs2m_view_stats_update(P, _ptr(radii.contiguous()), _ptr(observe.contiguous()), _ptr(max_radii2D),
                                                 _ptr(observe_cnt), torch.cuda.current_stream(dev).cuda_stream),
                      "gs2m_view_stats_update")


def control_block(raster_settings, state: RasterState):
    """The forward's device-side control block (int32[8]: R, V, flags, R used, V used, ...) as a tensor aliasing the image arena.
    A caller that renders with ``no_wait=True`` never learns on the host whether the instance capacity sufficed; it can OR
    ``control_block(...)[2]`` of its views into a device flag word and read that once per step (flag bits: ``GS2M_BIN_*``)."""
    lib = _native.load()
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    v = _native.StateView()
    _native.check(lib.gs2m_state_view_get(0, W, H, 0, None, None, _ptr(state.img), v), "gs2m_state_view_get")
    off = v.bin_info - state.img.data_ptr()
    return state.img[off:off + 32].view(torch.int32)


def state_view(P, raster_settings, state: RasterState):
    """Typed tensors aliasing the opaque arenas (tests only): depths, rec_a, rec_b, rgb, cov3D, clamped,
    tiles_touched, point_offsets, grad_acc, keys_sorted, point_list, masks, dense_gid, dense_pos, final_T, n_contrib,
    ranges, bin_info, block_ranges, n_contrib_dense."""
    lib = _native.load()
    H, W = int(raster_settings.image_height), int(raster_settings.image_width)
    R = int(state.num_rendered)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    v = _native.StateView()
    _native.check(lib.gs2m_state_view_get(P, W, H, int(state.capacity) or R, _ptr(state.geom) if P else None,
                                          _ptr(state.binning) if state.binning.numel() else None,
                                          _ptr(state.img), v), "gs2m_state_view_get")

    def view(arena, ptr, count, dtype):
        if ptr is None or count == 0:
            return torch.empty(0, dtype=dtype, device=arena.device)
        off = ptr - arena.data_ptr()
        nbytes = count * torch.empty((), dtype=dtype).element_size()
        return arena[off:off + nbytes].view(dtype)

    out = {}
    g, bn, im = state.geom, state.binning, state.img
    if P:
        out["depths"] = view(g, v.depths, P, torch.float32)
        out["rec_a"] = view(g, v.rec_a, 4 * P, torch.float32).view(P, 4)
        out["rec_b"] = view(g, v.rec_b, 4 * P, torch.float32).view(P, 4)
        out["rgb"] = view(g, v.rgb, 4 * P, torch.float32).view(P, 4)
        out["cov3D"] = view(g, v.cov3D, 6 * P, torch.float32).view(P, 6)
        out["clamped"] = view(g, v.clamped, 4 * P, torch.uint8).view(P, 4)
        out["tiles_touched"] = view(g, v.tiles_touched, P, torch.int32)
        out["point_offsets"] = view(g, v.point_offsets, P, torch.int32)
        out["grad_acc"] = view(g, v.grad_acc, _native.ACC_STRIDE * P, torch.float32).view(P, _native.ACC_STRIDE)
    out["keys_sorted"] = view(bn, v.keys_sorted, R, torch.int64)
    out["point_list"] = view(bn, v.point_list, R, torch.int32)
    out["masks"] = view(bn, v.masks, R, torch.uint8)
    cap = int(state.capacity) or R            # the per-warp-block lists are laid out [8][arena capacity]
    out["dense_gid"] = view(bn, v.dense_gid, 8 * cap, torch.int32).view(8, cap)
    out["dense_pos"] = view(bn, v.dense_pos, 8 * cap, torch.int32).view(8, cap)
    out["block_ranges"] = view(im, v.block_ranges, 16 * tiles, torch.int32).view(tiles, 8, 2)
    out["n_contrib_dense"] = view(im, v.n_contrib_dense, H * W, torch.int32).view(H, W)
    out["bin_info"] = view(im, v.bin_info, 8, torch.int32)
    out["final_T"] = view(im, v.final_T, H * W, torch.float32).view(H, W)
    out["n_contrib"] = view(im, v.n_contrib, H * W, torch.int32).view(H, W)
    out["ranges"] = view(im, v.ranges, 2 * tiles, torch.int32).view(tiles, 2)
    return out


# ----------------------------------------------------------------------------------------------------------------
# autograd surface
# ----------------------------------------------------------------------------------------------------------------
class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, features,
                raster_settings):
        needs_grad = any(ctx.needs_input_grad)      # inference calls (torch.no_grad) skip the backward preparation
        color, radii, observe, buffer, state = forward_raw(
            means3D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, features, raster_settings,
            for_backward=needs_grad)
        ctx.raster_settings = raster_settings
        ctx.raster_state = state
        ctx.save_for_backward(means3D, shs, colors_precomp, scales, rotations, cov3Ds_precomp, features, radii,
                              state.geom, state.binning, state.img)
        ctx.mark_non_differentiable(radii, observe)
        return color, radii, observe, buffer

    @staticmethod
    def backward(ctx, grad_color, _grad_radii, _grad_observe, grad_buffer):
        (means3D, shs, colors_precomp, scales, rotations, cov3Ds_precomp, features, radii,
         geom, binning, img) = ctx.saved_tensors
        state = ctx.raster_state      # (the arenas travel through save_for_backward so that autograd tracks their lifetime)
        # gradients nobody receives are not computed: dL_dconic always, dL_dcolor / dL_dcov3D when those inputs are absent
        skip = ("dL_dconic",) + (("dL_dcolor",) if _absent(colors_precomp) else ()) + (("dL_dcov3D",) if _absent(cov3Ds_precomp) else ())
        P_, M_ = int(means3D.shape[0]), (0 if _absent(shs) else int(shs.shape[1]))
        g = backward_raw(grad_color, grad_buffer, means3D, shs, colors_precomp, scales, rotations, cov3Ds_precomp,
                         features, radii, ctx.raster_settings, state, grads=alloc_grads(P_, M_, means3D.device, skip=skip))
        # slots: means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, features, settings
        return (g["dL_dmeans3D"], g["dL_dmeans2D"],
                None if _absent(shs) else g["dL_dsh"],
                None if _absent(colors_precomp) else g["dL_dcolor"],
                g["dL_dopacity"],
                None if _absent(scales) else g["dL_dscale"],
                None if _absent(rotations) else g["dL_drot"],
                None if _absent(cov3Ds_precomp) else g["dL_dcov3D"],
                None if _absent(features) else g["dL_dfeatures"],
                None)


def rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, features,
                        raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, features, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """bool[P]: view-space z > 0.2 (reference ``:162-171`` -> checkFrustum, rasterizer_impl.cu:48-59)."""
        lib = _native.load()
        rs = self.raster_settings
        with torch.no_grad():
            if not positions.is_cuda:
                raise RuntimeError("positions must be a CUDA tensor")
            dev = positions.device
            pos = _dev_f32(positions, dev, "positions")
            vm = _dev_f32(rs.viewmatrix, dev, "viewmatrix")
            pm = _dev_f32(rs.projmatrix, dev, "projmatrix")
            P = int(pos.shape[0])
            present = torch.zeros((P,), dtype=torch.bool, device=dev)
            with torch.cuda.device(dev):
                _native.check(lib.gs2m_mark_visible(P, _ptr(pos), _ptr(vm), _ptr(pm), _ptr(present),
                                                    torch.cuda.current_stream(dev).cuda_stream), "gs2m_mark_visible")
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, features=None):
        have_sh, have_col = shs is not None, colors_precomp is not None
        if have_sh == have_col:
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        have_sr = scales is not None or rotations is not None
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (have_sr and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.Tensor([])  # CPU placeholder for absent optionals, as the reference passes (binding :192-203)
        return rasterize_gaussians(
            means3D, means2D,
            shs if have_sh else empty,
            colors_precomp if have_col else empty,
            opacities,
            scales if scales is not None else empty,
            rotations if rotations is not None else empty,
            cov3D_precomp if cov3D_precomp is not None else empty,
            features if features is not None else empty,
            self.raster_settings)
