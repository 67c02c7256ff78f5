"""Oracle for the fused activation + feature-packing stage: the SAME PyTorch eager ops GS-2M runs, written out in order,
differentiable through autograd.  TEST INFRASTRUCTURE ONLY (never imported by the product package).

Follows (paths relative to /root/reference):
  scene/gaussian_model.py:34-44,113-172   activations: exp, torch.nn.functional.normalize, sigmoid
  scene/gaussian_model.py:146-160         get_normals (argmin one-hot, bmm with build_rotation, in-place flip, normalise)
  utils/general_utils.py:72-92            build_rotation
  gaussian_renderer/__init__.py:82-96     cam_normals, cam_points, features columns
Pinning: the reference ships no tests for these stages, so every function here is checked against the reference's OWN code,
imported from /root/reference with stubbed third-party modules: directly on the CPU where the tree is
present (tests/test_caller_oracle_pinned.py) and through committed vectors generated from it
(tests/golden/caller_stage_golden.npz by tests/golden/make_caller_golden.py; tests/golden/facade_*.npz by
tests/golden/make_facade_golden.py, the real render() around the compiled reference rasterizer on a B200).
"""
import torch


def build_rotation(r):
    norm = torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3])
    q = r / norm[:, None]
    R = torch.zeros((q.size(0), 3, 3), dtype=r.dtype, device=r.device)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z)
    R[:, 0, 1] = 2 * (x * y - r_ * z)
    R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z)
    R[:, 1, 1] = 1 - 2 * (x * x + z * z)
    R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y)
    R[:, 2, 1] = 2 * (y * z + r_ * x)
    R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def activate_and_pack(xyz, scaling, rotation, opacity, albedo, roughness, metallic, world_view_transform, camera_center,
                      z_depth=False, blend_metallic=False):
    scales = torch.exp(scaling)
    rotations = torch.nn.functional.normalize(rotation)
    opacities = torch.sigmoid(opacity)
    alb, rough, metal = torch.sigmoid(albedo), torch.sigmoid(roughness), torch.sigmoid(metallic)
    # get_normals
    min_idx = torch.argmin(scales, dim=-1, keepdim=True)
    min_axes = torch.zeros_like(scales).scatter(1, min_idx, 1)
    R = build_rotation(rotations)
    normals = torch.bmm(R, min_axes.unsqueeze(-1)).squeeze(-1)
    view_dirs = camera_center[None] - xyz
    flip = torch.sum(normals * view_dirs, dim=-1) < 0.0
    normals = torch.where(flip[:, None], -normals, normals)
    normals = normals / normals.norm(dim=1, keepdim=True)
    # render(): features
    cam_normals = normals @ world_view_transform[:3, :3]
    cam_points = xyz @ world_view_transform[:3, :3] + world_view_transform[3, :3]
    cols = [torch.ones_like(xyz[:, :1]),
            (cam_points[:, 2:3] if z_depth else (cam_normals * cam_points).sum(dim=-1, keepdim=True).abs()),
            normals, alb, rough, (metal if blend_metallic else torch.zeros_like(metal))]
    return scales, rotations, opacities, torch.cat(cols, dim=1)


def derive_maps(buffer, world_view_transform, fx, fy, cx, cy, z_depth=False):
    """gaussian_renderer/__init__.py:125-141 with the rays of scene/cameras.py:71-81 (scale = 1), eager ops in order."""
    H, W = buffer.shape[1], buffer.shape[2]
    normal_map = buffer[2:5, ...]
    normal_mask = (normal_map != 0).all(0, keepdim=True).detach()
    local_normals = normal_map.permute(1, 2, 0).reshape(-1, 3)
    local_normals = local_normals @ world_view_transform[:3, :3]
    local_normal_map = local_normals.reshape(H, W, 3).permute(2, 0, 1)
    depth_map = buffer[1:2, ...]
    if not z_depth:
        u, v = torch.meshgrid(torch.arange(W, dtype=buffer.dtype, device=buffer.device),
                              torch.arange(H, dtype=buffer.dtype, device=buffer.device), indexing="xy")
        rays = torch.stack(((u - cx) / fx, (v - cy) / fy, torch.ones_like(u)), dim=-1).view(-1, 3)
        denoms = torch.sum(local_normals * rays, dim=-1).view(1, H, W)
        depth_map = buffer[1:2, ...] / -(denoms + 1e-8)
    return local_normal_map, depth_map, normal_mask


def photometric_loss(render, gt, lambda_ssim=0.2):
    """train.py:102-107: (1 - lambda) * l1_loss + lambda * (1 - ssim), with utils/loss_utils.py:24-25 and :30-70 written out
    (11-tap Gaussian window, sigma 1.5, zero padding 5, one group per channel, C1 = 0.01^2, C2 = 0.03^2)."""
    import math
    import torch.nn.functional as F
    ch = render.shape[0]
    g = torch.tensor([math.exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    window = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(ch, 1, 11, 11).contiguous().to(render)
    img1, img2 = render.unsqueeze(0), gt.unsqueeze(0)
    mu1 = F.conv2d(img1, window, padding=5, groups=ch)
    mu2 = F.conv2d(img2, window, padding=5, groups=ch)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=5, groups=ch) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=5, groups=ch) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=5, groups=ch) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    l1 = torch.abs(render - gt).mean()
    return (1.0 - lambda_ssim) * l1 + lambda_ssim * (1.0 - ssim_map.mean()), l1, ssim_map.mean()


def sobel_normal_map(depth, alpha, bg_color, world_view_transform, fx, fy, cx, cy):
    """render_normal_from_depth_map (gaussian_renderer/__init__.py:163-175) with normal_from_depth_image / depth2point /
    depth_pcd2normal (utils/normal_utils.py:11-85, offset=None, view_space=False) and get_calib_matrix_nerf
    (scene/cameras.py:83-89) written out with the same torch ops."""
    H, W = depth.shape
    intrinsic = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=depth.dtype, device=depth.device)
    extrinsic = world_view_transform.transpose(0, 1).contiguous()
    valid_x = torch.arange(W, dtype=torch.float32, device=depth.device) / (W - 1)      # divided in float32 whatever the depth's
    valid_y = torch.arange(H, dtype=torch.float32, device=depth.device) / (H - 1)      # dtype is (normal_utils.py:14-15)
    valid_x, valid_y = torch.meshgrid(valid_x, valid_y, indexing="xy")
    ndc_xyz = torch.stack([valid_x.to(depth.dtype), valid_y.to(depth.dtype), depth], dim=-1)
    inv_scale = torch.tensor([[W - 1, H - 1]], dtype=depth.dtype, device=depth.device)
    cam_z = ndc_xyz[..., 2:3]
    cam_xy = ndc_xyz[..., :2] * inv_scale * cam_z
    cam_xyz = torch.cat([cam_xy, cam_z], dim=-1) @ torch.inverse(intrinsic.t())
    xyz_cam = cam_xyz.reshape(-1, 3)
    xyz_world = torch.cat([xyz_cam, torch.ones_like(xyz_cam[..., 0:1])], axis=-1) @ torch.inverse(extrinsic).transpose(0, 1)
    xyz = xyz_world[..., :3].reshape(H, W, 3)
    bottom_point, top_point = xyz[2:H, 1:W - 1, :], xyz[0:H - 2, 1:W - 1, :]
    right_point, left_point = xyz[1:H - 1, 2:W, :], xyz[1:H - 1, 0:W - 2, :]
    xyz_normal = torch.cross(right_point - left_point, top_point - bottom_point, dim=-1)
    xyz_normal = torch.nn.functional.normalize(xyz_normal, p=2, dim=-1)
    xyz_normal = torch.nn.functional.pad(xyz_normal.permute(2, 0, 1), (1, 1, 1, 1), mode="constant").permute(1, 2, 0)
    normal_ref = xyz_normal * alpha[..., None] + bg_color[None, None, ...] * (1.0 - alpha[..., None])
    return normal_ref.permute(2, 0, 1)


def render_like(rasterizer, raw, shs, wvt, full_proj, campos, tanfovx, tanfovy, W, H, bg, active_sh_degree=3,
                geometry_stage=False, material_stage=False, sobel_normal=False, blend_metallic=False, z_depth=False):
    """The render facade (gaussian_renderer/__init__.py:21-175) restated with the functions of this file around a rasterizer
    module that has the reference's Python surface: activations + packing (:49-96), settings (:98-110), the rasterizer call
    (:113-123), the post-blend maps (:125-141), the result dict (:143-158) and the depth-derived normal (:160-175).
    ``raw`` holds the nine raw parameter tensors (xyz, scaling, rotation, opacity, albedo, roughness, metallic) and ``shs`` the
    (P,16,3) SH block; everything differentiable through autograd.  Pinned against the real facade by tests/golden/facade_*.npz."""
    P = raw["xyz"].shape[0]
    screenspace_points = torch.zeros((P, 4), dtype=raw["xyz"].dtype, requires_grad=True, device=raw["xyz"].device) + 0
    screenspace_points.retain_grad()
    scales, rotations, opacity, features = activate_and_pack(
        raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"], raw["albedo"], raw["roughness"], raw["metallic"], wvt,
        campos, z_depth=z_depth, blend_metallic=blend_metallic)
    feature_count = (9 if material_stage else 5 if geometry_stage else 1) + (1 if blend_metallic else 0)
    settings = rasterizer.GaussianRasterizationSettings(
        image_height=int(H), image_width=int(W), tanfovx=tanfovx, tanfovy=tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=wvt,
        projmatrix=full_proj, sh_degree=active_sh_degree, campos=campos, prefiltered=False, feature_count=feature_count)
    rendered_image, radii, observe, buffer = rasterizer.GaussianRasterizer(raster_settings=settings)(
        means3D=raw["xyz"], means2D=screenspace_points, opacities=opacity, shs=shs, colors_precomp=None, scales=scales,
        rotations=rotations, cov3D_precomp=None, features=features)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    local_normal_map, depth_map, normal_mask = derive_maps(buffer, wvt, fx, fy, 0.5 * W, 0.5 * H, z_depth=z_depth)
    out = {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
           "observe": observe, "alpha_map": buffer[0:1], "distance_map": None if z_depth else buffer[1:2], "depth_map": depth_map,
           "normal_map": buffer[2:5], "albedo_map": buffer[5:8], "roughness_map": buffer[8:9], "metallic_map": buffer[9:10],
           "normal_mask": normal_mask, "local_normal_map": local_normal_map}
    if sobel_normal:
        out["sobel_map"] = sobel_normal_map(out["depth_map"].squeeze(0), out["alpha_map"][0], bg, wvt, fx, fy, 0.5 * W, 0.5 * H)
    return out
