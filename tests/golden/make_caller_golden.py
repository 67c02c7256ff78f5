#!/usr/bin/env python
"""Generates tests/golden/caller_stage_golden.npz on the CPU, in this container, by calling the reference's OWN functions
(imported from /root/reference through oracle/build_ref.load_facade, third-party modules stubbed as SURVEY 8c describes):

    utils/general_utils.py:72-92     build_rotation
    scene/gaussian_model.py:113-172  the activated getters and get_normals of a real GaussianModel
    utils/normal_utils.py:65-85      normal_from_depth_image (view_space=False), as render_normal_from_depth_map calls it
    utils/loss_utils.py:24-70        l1_loss, ssim

on small seeded inputs.  The vectors pin oracle/pack_reference.py (tests/test_caller_oracle_pinned.py) wherever the reference
tree is absent.  Run:  python tests/golden/make_caller_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import build_ref  # noqa: E402
import synthetic_scenes as syn  # noqa: E402


def inputs():
    g = torch.Generator().manual_seed(404)
    P, H, W = 300, 23, 31
    scene = syn.make_scene(P, seed=9, shell_fraction=0.5)
    raw = syn.raw_parameters(scene)
    raw["scaling"][: P // 6, 1] = raw["scaling"][: P // 6, 0]          # exact ties in the thinnest-axis argmin
    cam = syn.make_cameras(1, W, H)[0]
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 2.0 + 0.01 * xx + 0.02 * yy + 0.05 * torch.rand(H, W, generator=g)
    img1 = torch.rand(3, 40, 52, generator=g)
    img2 = (img1 + 0.1 * torch.randn(3, 40, 52, generator=g)).clamp(0, 1)
    return raw, cam, depth, img1, img2


def main():
    import diff_gaussian_rasterization as dgr      # only so that the facade module can be imported; nothing is rasterized
    facade = build_ref.load_facade(dgr, "ours")
    from scene.gaussian_model import GaussianModel
    from utils.general_utils import build_rotation
    from utils.normal_utils import normal_from_depth_image
    lu = build_ref.load_loss_utils(facade)
    raw, cam, depth, img1, img2 = inputs()
    pc = GaussianModel(3)
    pc._xyz, pc._scaling, pc._rotation, pc._opacity = raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"]
    pc._albedo, pc._roughness, pc._metallic = raw["albedo"], raw["roughness"], raw["metallic"]
    W, H = cam.image_width, cam.image_height
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    intrinsic = torch.tensor([[fx, 0, 0.5 * W], [0, fy, 0.5 * H], [0, 0, 1]]).float()
    extrinsic = cam.world_view_transform.transpose(0, 1).contiguous()
    with build_ref.cuda_literals_on_cpu():
        arrays = reference_values(pc, raw, cam, depth, img1, img2, intrinsic, extrinsic, fx, fy, build_rotation,
                                  normal_from_depth_image, lu)
    out = os.path.join(ROOT, "tests", "golden", "caller_stage_golden.npz")
    np.savez_compressed(out, **{k: v.detach().cpu().numpy() for k, v in arrays.items()})
    print("wrote", out, {k: tuple(v.shape) for k, v in arrays.items()})


def reference_values(pc, raw, cam, depth, img1, img2, intrinsic, extrinsic, fx, fy, build_rotation, normal_from_depth_image, lu):
    return {
        "build_rotation": build_rotation(raw["rotation"]),
        "get_scaling": pc.get_scaling, "get_rotation": pc.get_rotation, "get_opacity": pc.get_opacity,
        "get_albedo": pc.get_albedo, "get_roughness": pc.get_roughness, "get_metallic": pc.get_metallic,
        "get_normals": pc.get_normals(cam.camera_center),
        "normal_from_depth": normal_from_depth_image(depth, intrinsic, extrinsic, view_space=False),
        "l1_loss": lu.l1_loss(img1, img2).reshape(1), "ssim": lu.ssim(img1, img2).reshape(1),
        "fx_fy": torch.tensor([fx, fy], dtype=torch.float64),
    }


if __name__ == "__main__":
    main()
