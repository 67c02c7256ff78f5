"""Fused parameter activation + feature packing (the stage in front of the rasterizer).

``activate_and_pack(xyz, scaling, rotation, opacity, albedo, roughness, metallic, world_view_transform, camera_center,
z_depth=False, blend_metallic=False) -> (scales, rotations, opacities, features)`` replaces, with one CUDA kernel forward
and one backward, what GS-2M's render facade does with ~15 / ~40 PyTorch kernels per call
(gaussian_renderer/__init__.py:49-96 over the getters of scene/gaussian_model.py:113-172): the four returned tensors
are exactly the ``scales=, rotations=, opacities=, features=`` arguments of ``GaussianRasterizer.forward``; gradients
flow back to the seven RAW parameter tensors.  How render() would use it:

    scales, rotations, opacity, features = activate_and_pack(pc._xyz, pc._scaling, pc._rotation, pc._opacity, pc._albedo,
        pc._roughness, pc._metallic, viewpoint_camera.world_view_transform, viewpoint_camera.camera_center,
        z_depth=pipe.z_depth, blend_metallic=blend_metallic)
"""
import torch

from . import _native


def _f32(t, dev, name):
    if t.dtype != torch.float32 or t.device != dev:
        raise RuntimeError("%s must be a float32 tensor on %s" % (name, dev))
    return t.contiguous()


class _ActivateAndPack(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scaling, rotation, opacity, albedo, roughness, metallic, wvt, campos, z_depth, blend_metallic):
        lib = _native.load()
        if not xyz.is_cuda:
            raise RuntimeError("activate_and_pack has no CPU path: inputs must be CUDA tensors")
        dev = xyz.device
        P = int(xyz.shape[0])
        args = [_f32(t, dev, n) for t, n in ((xyz, "xyz"), (scaling, "scaling"), (rotation, "rotation"), (opacity, "opacity"),
                                             (albedo, "albedo"), (roughness, "roughness"), (metallic, "metallic"),
                                             (wvt, "world_view_transform"), (campos, "camera_center"))]
        shapes = ((P, 3), (P, 3), (P, 4), (P, 1), (P, 3), (P, 1), (P, 1), (4, 4), (3,))
        for t, s in zip(args, shapes):
            if tuple(t.shape) != s:
                raise RuntimeError("activate_and_pack: expected shape %s, got %s" % (s, tuple(t.shape)))
        scales = torch.empty((P, 3), dtype=torch.float32, device=dev)
        rotations = torch.empty((P, 4), dtype=torch.float32, device=dev)
        opacities = torch.empty((P, 1), dtype=torch.float32, device=dev)
        features = torch.empty((P, 10), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_pack_forward(P, *[t.data_ptr() for t in args], int(bool(z_depth)), int(bool(blend_metallic)),
                                                scales.data_ptr(), rotations.data_ptr(), opacities.data_ptr(),
                                                features.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                          "gs2m_pack_forward")
        ctx.save_for_backward(*args)
        ctx.flags = (int(bool(z_depth)), int(bool(blend_metallic)))
        return scales, rotations, opacities, features

    @staticmethod
    def backward(ctx, g_scales, g_rotations, g_opacities, g_features):
        lib = _native.load()
        args = ctx.saved_tensors
        xyz = args[0]
        dev, P = xyz.device, int(xyz.shape[0])
        z = lambda shape: torch.zeros(shape, dtype=torch.float32, device=dev)  # noqa: E731
        ups = [g.contiguous() if g is not None else z(s) for g, s in
               ((g_scales, (P, 3)), (g_rotations, (P, 4)), (g_opacities, (P, 1)), (g_features, (P, 10)))]
        outs = [torch.empty(s, dtype=torch.float32, device=dev) for s in ((P, 3), (P, 3), (P, 4), (P, 1), (P, 3), (P, 1), (P, 1))]
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_pack_backward(P, *[t.data_ptr() for t in args], ctx.flags[0], ctx.flags[1],
                                                 *[t.data_ptr() for t in ups], *[t.data_ptr() for t in outs],
                                                 torch.cuda.current_stream(dev).cuda_stream), "gs2m_pack_backward")
        return (*outs, None, None, None, None)


def activate_and_pack(xyz, scaling, rotation, opacity, albedo, roughness, metallic, world_view_transform, camera_center,
                      z_depth=False, blend_metallic=False):
    return _ActivateAndPack.apply(xyz, scaling, rotation, opacity, albedo, roughness, metallic, world_view_transform,
                                  camera_center, z_depth, blend_metallic)


def pack_backward_accumulate(xyz, scaling, rotation, opacity, albedo, roughness, metallic, world_view_transform, camera_center,
                             g_scales, g_rotations, g_opacities, g_features, d_xyz, d_scaling, d_rotation, d_opacity, d_albedo,
                             d_roughness, d_metallic, radii=None, z_depth=False, blend_metallic=False):
    """Packing-stage backward of one view with ``+=`` into the seven raw-parameter gradient tensors (C-ABI
    ``gs2m_pack_backward_accumulate``); with ``radii`` (int32 ``[P]`` of that view's forward) culled Gaussians are skipped."""
    lib = _native.load()
    if not xyz.is_cuda:
        raise RuntimeError("pack_backward_accumulate has no CPU path: inputs must be CUDA tensors")
    dev, P = xyz.device, int(xyz.shape[0])
    ts = [_f32(t, dev, "tensor %d" % k) for k, t in enumerate(
        (xyz, scaling, rotation, opacity, albedo, roughness, metallic, world_view_transform, camera_center,
         g_scales, g_rotations, g_opacities, g_features))]
    outs = (d_xyz, d_scaling, d_rotation, d_opacity, d_albedo, d_roughness, d_metallic)
    for t, c in zip(outs, (3, 3, 4, 1, 3, 1, 1)):
        if t.dtype != torch.float32 or t.device != dev or not t.is_contiguous() or t.numel() != P * c:
            raise RuntimeError("pack_backward_accumulate: gradient outputs must be contiguous float32 [P,%d] on %s" % (c, dev))
    if radii is not None and (radii.dtype != torch.int32 or radii.numel() != P or not radii.is_contiguous()):
        raise RuntimeError("pack_backward_accumulate: radii must be a contiguous int32 [P] tensor")
    with torch.cuda.device(dev):
        _native.check(lib.gs2m_pack_backward_accumulate(
            P, *[t.data_ptr() for t in ts[:9]], int(bool(z_depth)), int(bool(blend_metallic)), *[t.data_ptr() for t in ts[9:]],
            *[t.data_ptr() for t in outs], None if radii is None else radii.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream), "gs2m_pack_backward_accumulate")


class _DeriveMaps(torch.autograd.Function):
    @staticmethod
    def forward(ctx, buffer, wvt, fx, fy, cx, cy, z_depth):
        lib = _native.load()
        if not buffer.is_cuda:
            raise RuntimeError("derive_maps has no CPU path: buffer must be a CUDA tensor")
        dev = buffer.device
        buffer = _f32(buffer, dev, "buffer")
        wvt = _f32(wvt, dev, "world_view_transform")
        if buffer.dim() != 3 or buffer.shape[0] != 10:
            raise RuntimeError("buffer must be (10, H, W)")
        H, W = int(buffer.shape[1]), int(buffer.shape[2])
        local_normal = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
        mask = torch.empty((1, H, W), dtype=torch.bool, device=dev)
        ctx.geom = (W, H, float(fx), float(fy), float(cx), float(cy), int(bool(z_depth)))
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_postblend_forward(*ctx.geom, wvt.data_ptr(), buffer.data_ptr(), local_normal.data_ptr(),
                                                     depth.data_ptr(), mask.data_ptr(),
                                                     torch.cuda.current_stream(dev).cuda_stream), "gs2m_postblend_forward")
        ctx.save_for_backward(buffer, wvt)
        ctx.mark_non_differentiable(mask)
        return local_normal, depth, mask

    @staticmethod
    def backward(ctx, g_local_normal, g_depth, _g_mask):
        lib = _native.load()
        buffer, wvt = ctx.saved_tensors
        dev = buffer.device
        W, H = ctx.geom[0], ctx.geom[1]
        gl = g_local_normal.contiguous() if g_local_normal is not None else torch.zeros((3, H, W), device=dev)
        gd = g_depth.contiguous() if g_depth is not None else torch.zeros((1, H, W), device=dev)
        g_buffer = torch.empty((10, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_postblend_backward(*ctx.geom, wvt.data_ptr(), buffer.data_ptr(), gl.data_ptr(), gd.data_ptr(),
                                                      g_buffer.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                          "gs2m_postblend_backward")
        return g_buffer, None, None, None, None, None, None


def derive_maps(buffer, world_view_transform, fx, fy, cx, cy, z_depth=False):
    """(local_normal_map[3,H,W], depth_map[1,H,W], normal_mask[1,H,W] bool) from the rasterizer's buffer — the fused form of
    gaussian_renderer/__init__.py:125-141 (with scene/cameras.py:71-81 rays); differentiable w.r.t. ``buffer``."""
    return _DeriveMaps.apply(buffer, world_view_transform, fx, fy, cx, cy, z_depth)


class _SobelNormal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, alpha, bg, wvt, fx, fy, cx, cy):
        lib = _native.load()
        if not depth.is_cuda:
            raise RuntimeError("sobel_normal_map has no CPU path: depth must be a CUDA tensor")
        dev = depth.device
        depth, alpha = _f32(depth, dev, "depth"), _f32(alpha, dev, "alpha")
        bg, wvt = _f32(bg, dev, "bg_color"), _f32(wvt, dev, "world_view_transform")
        if depth.dim() != 2 or depth.shape != alpha.shape or bg.numel() != 3:
            raise RuntimeError("sobel_normal_map: depth and alpha must be (H, W), bg_color (3,)")
        H, W = int(depth.shape[0]), int(depth.shape[1])
        out = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        ctx.geom = (W, H, float(fx), float(fy), float(cx), float(cy))
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_sobel_normal_forward(*ctx.geom, wvt.data_ptr(), bg.data_ptr(), depth.data_ptr(), alpha.data_ptr(),
                                                        out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                          "gs2m_sobel_normal_forward")
        ctx.save_for_backward(depth, alpha, bg, wvt)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _native.load()
        depth, alpha, bg, wvt = ctx.saved_tensors
        dev = depth.device
        d_depth, d_alpha = torch.empty_like(depth), torch.empty_like(alpha)
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_sobel_normal_backward(*ctx.geom, wvt.data_ptr(), bg.data_ptr(), depth.data_ptr(), alpha.data_ptr(),
                                                         g_out.contiguous().data_ptr(), d_depth.data_ptr(), d_alpha.data_ptr(),
                                                         torch.cuda.current_stream(dev).cuda_stream),
                          "gs2m_sobel_normal_backward")
        return d_depth, d_alpha, None, None, None, None, None, None


def sobel_normal_map(depth_map, alpha_map, bg_color, world_view_transform, fx, fy, cx, cy):
    """``render_normal_from_depth_map`` (gaussian_renderer/__init__.py:163-175 over utils/normal_utils.py:30-85) as one kernel
    each way: world-space normals from the ``(H, W)`` depth map, composited over ``bg_color`` with ``alpha_map``; returns
    ``(3, H, W)``, differentiable w.r.t. depth and alpha."""
    return _SobelNormal.apply(depth_map, alpha_map, bg_color, world_view_transform, fx, fy, cx, cy)


class _PhotometricLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, gt, lambda_ssim):
        lib = _native.load()
        if not render.is_cuda:
            raise RuntimeError("photometric_loss has no CPU path: images must be CUDA tensors")
        dev = render.device
        render, gt = _f32(render, dev, "render"), _f32(gt, dev, "gt")
        if render.dim() != 3 or render.shape != gt.shape:
            raise RuntimeError("photometric_loss: render and gt must both be (C, H, W)")
        C_, H, W = (int(d) for d in render.shape)
        maps = torch.empty((3, C_, H, W), dtype=torch.float32, device=dev)
        sums = torch.zeros(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_photometric_loss_forward(C_, H, W, render.data_ptr(), gt.data_ptr(), maps[0].data_ptr(),
                                                            maps[1].data_ptr(), maps[2].data_ptr(), sums.data_ptr(),
                                                            torch.cuda.current_stream(dev).cuda_stream),
                          "gs2m_photometric_loss_forward")
        ctx.save_for_backward(render, gt, maps)
        ctx.lambda_ssim = float(lambda_ssim)
        means = sums / float(C_ * H * W)
        ctx.mark_non_differentiable(means)
        return (1.0 - ctx.lambda_ssim) * means[0] + ctx.lambda_ssim * (1.0 - means[1]), means

    @staticmethod
    def backward(ctx, g_loss, _g_means):
        lib = _native.load()
        render, gt, maps = ctx.saved_tensors
        dev = render.device
        C_, H, W = (int(d) for d in render.shape)
        grad = torch.empty_like(render)
        # the incoming gradient stays on the device (no host read: the step pipeline stays asynchronous and capturable)
        up = g_loss.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_photometric_loss_backward(C_, H, W, render.data_ptr(), gt.data_ptr(), maps[0].data_ptr(),
                                                             maps[1].data_ptr(), maps[2].data_ptr(), ctx.lambda_ssim,
                                                             1.0, up.data_ptr(), grad.data_ptr(),
                                                             torch.cuda.current_stream(dev).cuda_stream),
                          "gs2m_photometric_loss_backward")
        return grad, None, None


def photometric_loss(render, gt, lambda_ssim=0.2, return_terms=False):
    """``(1 - lambda) * l1_loss(render, gt) + lambda * (1 - ssim(render, gt))`` on ``(C,H,W)`` images — train.py:102-107 with
    utils/loss_utils.py:24-70 / the fused-ssim submodule — as one CUDA kernel forward and one backward, differentiable w.r.t.
    ``render``.  With ``return_terms`` also returns the (detached) ``[mean |render - gt|, mean SSIM]``."""
    loss, means = _PhotometricLoss.apply(render, gt, lambda_ssim)
    return (loss, means) if return_terms else loss


class FusedAdam:
    """``torch.optim.Adam(groups, lr=0.0, eps=1e-15)`` of ``GaussianModel.training_setup`` (scene/gaussian_model.py:230-242) as ONE
    kernel launch per step over all parameter groups (C-ABI ``gs2m_adam_step``): default betas, per-group ``lr``, no weight
    decay, no amsgrad.

    ``groups`` is a list of dicts ``{"name", "param": float32 CUDA tensor (updated in place), "lr"}``.  ``step(grads)`` takes a
    dict name -> gradient; a gradient may be a *column slice view* of a wider row-major tensor (``buckets.tensors["sh"][:, :1]``
    for ``f_dc`` and ``[:, 1:]`` for ``f_rest``), which is how the view-sharded step keeps the SH gradient."""

    def __init__(self, groups, betas=(0.9, 0.999), eps=1e-15):
        self.groups = [dict(g) for g in groups]
        if not 0 < len(self.groups) <= 16:
            raise RuntimeError("FusedAdam takes 1..16 parameter groups")
        self.betas, self.eps, self.t = (float(betas[0]), float(betas[1])), float(eps), 0
        for g in self.groups:
            p = g["param"]
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam: parameters must be contiguous float32 CUDA tensors (no CPU path)")
            g["exp_avg"], g["exp_avg_sq"] = torch.zeros_like(p), torch.zeros_like(p)

    def set_lr(self, name, lr):
        for g in self.groups:
            if g["name"] == name:
                g["lr"] = float(lr)

    # ---- optimizer-state surgery of GS-2M's densification (scene/gaussian_model.py:372-403, 437-456): every method returns
    # {name: new parameter tensor}, like the reference's helpers return ``optimizable_tensors`` ----
    def replace_tensor(self, tensor, name):
        """``replace_tensor_to_optimizer`` (reset_opacity): new parameter values, moments back to zero."""
        out = {}
        for g in self.groups:
            if g["name"] == name:
                g["param"] = tensor.detach().contiguous()
                g["exp_avg"], g["exp_avg_sq"] = torch.zeros_like(g["param"]), torch.zeros_like(g["param"])
                out[name] = g["param"]
        return out

    def prune(self, valid_mask):
        """``_prune_optimizer``: keep the rows where ``valid_mask`` is true, in parameters and moments."""
        out = {}
        for g in self.groups:
            g["param"] = g["param"][valid_mask].contiguous()
            g["exp_avg"], g["exp_avg_sq"] = g["exp_avg"][valid_mask].contiguous(), g["exp_avg_sq"][valid_mask].contiguous()
            out[g["name"]] = g["param"]
        return out

    def cat(self, tensors_dict):
        """``cat_tensors_to_optimizer`` (densification_postfix): append new rows; their moments start at zero."""
        out = {}
        for g in self.groups:
            ext = tensors_dict[g["name"]].detach().to(g["param"])
            g["param"] = torch.cat((g["param"], ext), dim=0).contiguous()
            g["exp_avg"] = torch.cat((g["exp_avg"], torch.zeros_like(ext)), dim=0)
            g["exp_avg_sq"] = torch.cat((g["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
            out[g["name"]] = g["param"]
        return out

    def step(self, grads):
        lib = _native.load()
        arr = (_native.AdamGroup * len(self.groups))()
        dev = self.groups[0]["param"].device
        for k, g in enumerate(self.groups):
            p, gr = g["param"], grads[g["name"]]
            rows = int(p.shape[0]) if p.dim() > 0 else 1
            width = p.numel() // max(rows, 1)
            if gr.dtype != torch.float32 or gr.device != p.device or gr.numel() != p.numel() or int(gr.shape[0]) != rows:
                raise RuntimeError("FusedAdam: gradient of %s does not match its parameter" % g["name"])
            if g["exp_avg"].shape != p.shape or g["exp_avg_sq"].shape != p.shape or not p.is_contiguous():
                raise RuntimeError("FusedAdam: state of %s does not match its parameter (resize with prune / cat / replace_tensor)"
                                   % g["name"])
            if gr.is_contiguous():
                stride, off_ptr = width, gr.data_ptr()
            else:   # a column slice of a wider row-major block: rows keep the parent's stride, the inner part is dense
                inner = gr.stride()[1:] if gr.dim() > 1 else ()
                dense, expect = True, 1
                for d, st in zip(reversed(gr.shape[1:]), reversed(inner)):
                    dense &= (st == expect) or d == 1
                    expect *= d
                if not dense:
                    raise RuntimeError("FusedAdam: gradient of %s must be contiguous or a column slice of a row-major tensor" % g["name"])
                stride, off_ptr = int(gr.stride(0)), gr.data_ptr()
            a = arr[k]
            a.param, a.exp_avg, a.exp_avg_sq, a.grad = p.data_ptr(), g["exp_avg"].data_ptr(), g["exp_avg_sq"].data_ptr(), off_ptr
            a.rows, a.width, a.grad_row_stride, a.grad_col_offset, a.lr = rows, width, stride, 0, float(g["lr"])
        self.t += 1          # (after validation: a rejected call does not advance the bias correction)
        with torch.cuda.device(dev):
            _native.check(lib.gs2m_adam_step(arr, len(self.groups), self.t, self.betas[0], self.betas[1], self.eps,
                                             torch.cuda.current_stream(dev).cuda_stream), "gs2m_adam_step")
