#!/usr/bin/env python
"""Fused activation + feature packing vs the PyTorch-eager ops GS-2M runs (oracle/pack_reference.py on the GPU):
forward+backward time at P = 3 M, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import pack_reference as ref  # noqa: E402
import synthetic_scenes as syn  # noqa: E402
from diff_gaussian_rasterization.packing import activate_and_pack  # noqa: E402
from test_feature_pack import _raw_params  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
cam = syn.camera_to(syn.make_cameras(1, 1959, 1090, radius=2.2)[0], "cuda")
raw = {k: v.cuda().requires_grad_(True) for k, v in _raw_params(P).items()}
ups = None


def run(fn):
    global ups
    out = fn(*raw.values(), cam.world_view_transform, cam.camera_center, blend_metallic=True)
    if ups is None:
        ups = [torch.randn_like(t) for t in out]
    torch.autograd.backward(list(out), ups)
    for v in raw.values():
        v.grad = None


for name, fn in (("torch eager (reference ops)", ref.activate_and_pack), ("fused CUDA", activate_and_pack)):
    for _ in range(3):
        run(fn)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(fn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-30s P=%d fwd+bwd  min %.3f ms  median %.3f ms" % (name, P, min(ts), sorted(ts)[len(ts) // 2]))


# ---- post-blend map derivation at the headline resolution ----
from diff_gaussian_rasterization.packing import derive_maps  # noqa: E402

H, W = 1090, 1959
buf = torch.randn(10, H, W, device="cuda").requires_grad_(True)
gl, gd = torch.randn(3, H, W, device="cuda"), torch.randn(1, H, W, device="cuda")


def run_maps(fn):
    ln, d, _ = fn(buf, cam.world_view_transform, 1.1 * W, 1.1 * W, 0.5 * W, 0.5 * H)
    torch.autograd.backward([ln, d], [gl, gd])
    buf.grad = None


for name, fn in (("torch eager (reference ops)", ref.derive_maps), ("fused CUDA", derive_maps)):
    for _ in range(3):
        run_maps(fn)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_maps(fn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-30s derive_maps %dx%d fwd+bwd  min %.3f ms  median %.3f ms" % (name, W, H, min(ts), sorted(ts)[len(ts) // 2]))


# ---- densification statistics: the reference's eager ops vs the fused forward-side kernel ----
# (the gradient-norm half is fused into the backward kernel: +0.025 ms per view, bench.py stage `preprocess_bwd`)
import diff_gaussian_rasterization as dgr  # noqa: E402
import view_parallel as vp  # noqa: E402

st = vp.DensificationStats(P, "cuda")
radii = torch.randint(0, 30, (P,), device="cuda", dtype=torch.int32)
observe = torch.randint(0, 3, (P,), device="cuda", dtype=torch.int32)
g2d = torch.randn(P, 4, device="cuda")


def eager_stats():
    mask = (observe > 0) & (radii > 0)
    st.max_radii2D.copy_(torch.where(mask, torch.max(st.max_radii2D, radii), st.max_radii2D))
    st.observe_cnt[observe > 0] += 1
    st.update_backward_eager(g2d, radii)


def fused_stats():
    dgr.update_view_stats(radii, observe, st.max_radii2D, st.observe_cnt.view(-1))


for name, fn in (("torch eager (reference ops, fwd+bwd halves)", eager_stats), ("fused CUDA (forward half)", fused_stats)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-45s densification stats P=%d  min %.3f ms  median %.3f ms" % (name, P, min(ts), sorted(ts)[len(ts) // 2]))


# ---- photometric loss: the reference's torch ops (11x11 grouped conv2d x5) vs the fused kernels ----
from diff_gaussian_rasterization.packing import photometric_loss  # noqa: E402

render = torch.rand(3, H, W, device="cuda").requires_grad_(True)
gt_img = torch.rand(3, H, W, device="cuda")


def run_loss(fn):
    out = fn(render, gt_img, 0.2)
    (out[0] if isinstance(out, tuple) else out).backward()
    render.grad = None


for name, fn in (("torch eager (utils/loss_utils.py ops)", ref.photometric_loss), ("fused CUDA", photometric_loss)):
    for _ in range(3):
        run_loss(fn)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_loss(fn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-40s L1+SSIM loss %dx%d fwd+bwd  min %.3f ms  median %.3f ms" % (name, W, H, min(ts), sorted(ts)[len(ts) // 2]))


# ---- Adam over GS-2M's nine parameter groups: torch.optim.Adam vs the one-launch kernel ----
from diff_gaussian_rasterization.packing import FusedAdam  # noqa: E402

shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4),
          "albedo": (P, 3), "roughness": (P, 1), "metallic": (P, 1)}
params = {k: torch.nn.Parameter(torch.randn(s, device="cuda")) for k, s in shapes.items()}
sh_grad = torch.randn(P, 16, 3, device="cuda")
grads = {k: torch.randn(s, device="cuda") for k, s in shapes.items() if not k.startswith("f_")}
grads["f_dc"], grads["f_rest"] = sh_grad[:, :1], sh_grad[:, 1:]
for k in shapes:
    params[k].grad = grads[k].contiguous()
opt_default = torch.optim.Adam([{"params": [params[k]], "lr": 1e-3} for k in shapes], lr=0.0, eps=1e-15)
opt_fused = torch.optim.Adam([{"params": [params[k]], "lr": 1e-3} for k in shapes], lr=0.0, eps=1e-15, fused=True)
ours = FusedAdam([{"name": k, "param": params[k].data, "lr": 1e-3} for k in shapes])
for name, fn in (("torch.optim.Adam (reference configuration)", opt_default.step), ("torch.optim.Adam(fused=True)", opt_fused.step),
                 ("gs2m_adam_step (one launch)", lambda: ours.step(grads))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-45s Adam step P=%d (64 floats/Gaussian)  min %.3f ms  median %.3f ms" % (name, P, min(ts), sorted(ts)[len(ts) // 2]))


# ---- normal map from depth (rest of SURVEY.md section 8f rank 2) ----
from diff_gaussian_rasterization.packing import sobel_normal_map  # noqa: E402

depth_img = (2.0 + torch.rand(H, W, device="cuda")).requires_grad_(True)
alpha_img = torch.rand(H, W, device="cuda").requires_grad_(True)
bg3 = torch.zeros(3, device="cuda")
g_sobel = torch.randn(3, H, W, device="cuda")


def run_sobel(fn):
    fn(depth_img, alpha_img, bg3, cam.world_view_transform, 1.1 * W, 1.1 * W, 0.5 * W, 0.5 * H).backward(g_sobel)
    depth_img.grad = alpha_img.grad = None


for name, fn in (("torch eager (utils/normal_utils.py ops)", ref.sobel_normal_map), ("fused CUDA", sobel_normal_map)):
    for _ in range(3):
        run_sobel(fn)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run_sobel(fn); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-40s sobel normal %dx%d fwd+bwd  min %.3f ms  median %.3f ms" % (name, W, H, min(ts), sorted(ts)[len(ts) // 2]))
