"""Shared helpers of the parity tests: run the compiled reference / the CUDA implementation on one synthetic view
and return comparable dictionaries of tensors."""
import numpy as np
import torch

import synthetic_scenes as syn


def align128(x):
    return (x + 127) & ~127


def decode_reference_buffers(P, W, H, R, geom, binning, img):
    """typed views of the reference's opaque byte tensors
    (cuda_rasterizer/rasterizer_impl.cu:145-181, rasterizer_impl.h:18-30)."""
    out = {}
    N = W * H

    def take(buf, off, count, dtype, itemsize):
        off = align128(off)
        t = buf[off:off + count * itemsize].view(dtype)
        return t, off + count * itemsize

    base = geom.data_ptr() % 128
    assert base == 0
    off = 0
    out["depths"], off = take(geom, off, P, torch.float32, 4)
    out["clamped"], off = take(geom, off, 3 * P, torch.uint8, 1)
    _, off = take(geom, off, P, torch.int32, 4)
    m2d, off = take(geom, off, 2 * P, torch.float32, 4)
    out["means2D"] = m2d.view(P, 2)
    c3, off = take(geom, off, 6 * P, torch.float32, 4)
    out["cov3D"] = c3.view(P, 6)
    co, off = take(geom, off, 4 * P, torch.float32, 4)
    out["conic_opacity"] = co.view(P, 4)
    rgb, off = take(geom, off, 3 * P, torch.float32, 4)
    out["rgb"] = rgb.view(P, 3)
    out["tiles_touched"], off = take(geom, off, P, torch.int32, 4)
    start = geom.numel() - 128 - 4 * P
    out["point_offsets"] = geom[start:start + 4 * P].view(torch.int32)

    off = 0
    out["point_list"], off = take(binning, off, R, torch.int32, 4)
    _, off = take(binning, off, R, torch.int32, 4)
    out["keys_sorted"], off = take(binning, off, R, torch.int64, 8)

    off = 0
    ft, off = take(img, off, N, torch.float32, 4)
    out["final_T"] = ft.view(H, W)
    nc, off = take(img, off, N, torch.int32, 4)
    out["n_contrib"] = nc.view(H, W)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    rg, off = take(img, off, 2 * tiles, torch.int32, 4)
    out["ranges"] = rg.view(tiles, 2)
    return out


def make_view(P, W, H, F, shell=0.0, cam_radius=3.0, view=0, n_views=1, device="cuda", scene_seed=syn.SCENE_SEED, cluster=None,
              opacity_cap=None):
    scene = syn.make_scene(P, seed=scene_seed, shell_fraction=shell, cluster=cluster, opacity_cap=opacity_cap)
    cams = syn.make_cameras(max(n_views, view + 1), W, H, radius=cam_radius)
    cam = cams[view]
    feats = syn.pack_features(scene, cam, F)
    gc, gb = syn.make_upstream_grads(W, H, F)
    scene_d = syn.scene_to(scene, device)
    cam_d = syn.camera_to(cam, device)
    return scene_d, cam_d, feats.to(device), gc.to(device), gb.to(device)


def run_reference(ref, scene, cam, feats, F, grad_color=None, grad_buffer=None, sh_degree=3, bg=None):
    """Call the compiled reference's _C entry points directly (binding __init__.py:59-83,99-126)."""
    dev = scene.means3D.device
    P = scene.means3D.shape[0]
    W, H = cam.image_width, cam.image_height
    if bg is None:
        bg = torch.zeros(3, device=dev)
    empty = torch.Tensor([])
    R, color, radii, observe, buffer, geom, binning, img = ref._C.rasterize_gaussians(
        bg, scene.means3D, empty, scene.opacities, scene.scales, scene.rotations, 1.0, empty, feats,
        cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, H, W, scene.shs, sh_degree,
        cam.camera_center, False, F)
    out = dict(R=R, color=color, radii=radii, observe=observe, buffer=buffer)
    out.update(decode_reference_buffers(P, W, H, R, geom, binning, img))
    if grad_color is not None:
        g = ref._C.rasterize_gaussians_backward(
            bg, scene.means3D, radii, buffer, empty, scene.scales, scene.rotations, 1.0, empty, feats,
            cam.world_view_transform, cam.full_proj_transform, cam.tanfovx, cam.tanfovy, grad_color, grad_buffer,
            scene.shs, sh_degree, cam.camera_center, geom, R, binning, img, F)
        names = ["dL_dmeans2D", "dL_dcolor", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscale",
                 "dL_drot", "dL_dfeatures"]
        out.update(dict(zip(names, g)))
    return out


def run_ours(dgr, scene, cam, feats, F, grad_color=None, grad_buffer=None, sh_degree=3, bg=None):
    settings = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings, sh_degree=sh_degree, bg=bg)
    P = scene.means3D.shape[0]
    color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities,
                                                           scene.scales, scene.rotations, None, feats, settings)
    out = dict(R=state.num_rendered, color=color, radii=radii, observe=observe, buffer=buffer)
    sv = dgr.state_view(P, settings, state)
    out.update(sv)
    if P:
        out["means2D"] = sv["rec_a"][:, :2]
        out["conic_opacity"] = torch.cat([sv["rec_a"][:, 2:4], sv["rec_b"][:, 0:2]], dim=1)
        out["rgb"] = sv["rgb"][:, :3]
        out["clamped"] = sv["clamped"][:, :3].reshape(-1)
    if grad_color is not None:
        g = dgr.backward_raw(grad_color, grad_buffer, scene.means3D, scene.shs, None, scene.scales, scene.rotations,
                             None, feats, radii, settings, state)
        out.update(g)
    return out


def bits(t):
    """Bit pattern view for exact float comparison."""
    return t.contiguous().view(torch.int32) if t.dtype == torch.float32 else t


def grad_errors(ours, ref):
    """(max|d| / max|ref|, relative L2) of one gradient tensor."""
    d = (ours.double() - ref.double())
    denom = ref.double().abs().max().clamp_min(1e-30)
    l2 = d.norm() / ref.double().norm().clamp_min(1e-30)
    return float(d.abs().max() / denom), float(l2)


def assert_close_except_flips(actual, expected, rtol, atol, max_flip_frac=2e-3, flip_atol=5e-2):
    """Rendered planes vs the host-precision oracle: all pixels within (rtol, atol) except a tiny fraction where a
    borderline alpha >= 1/255 / T < 1e-4 decision is taken differently in host floating point (those move by up to
    one Gaussian's alpha*colour)."""
    a, e = actual.double(), expected.double()
    bad = (a - e).abs() > (atol + rtol * e.abs())
    frac = bad.double().mean().item()
    assert frac <= max_flip_frac, "%.5f of the values differ beyond rtol=%g atol=%g" % (frac, rtol, atol)
    if bad.any():
        worst = (a - e).abs()[bad].max().item()
        assert worst <= flip_atol, "largest outlier %.4g exceeds the decision-flip bound %.3g" % (worst, flip_atol)
