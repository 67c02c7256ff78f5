"""Pins oracle/pack_reference.py (the restated caller-side stages: activations, get_normals, build_rotation, the normal map
from depth, L1 + SSIM) against the reference's OWN functions: through the committed vectors of
tests/golden/caller_stage_golden.npz (made by tests/golden/make_caller_golden.py from /root/reference) and, where the reference
tree is present, by calling those functions live.  CPU only."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import build_ref
import pack_reference as ref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "caller_stage_golden.npz")


def _maker():
    spec = importlib.util.spec_from_file_location("make_caller_golden", os.path.join(HERE, "golden", "make_caller_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _oracle_values(raw, cam, depth, img1, img2):
    W, H = cam.image_width, cam.image_height
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    scales, rotations, opacities, feats = ref.activate_and_pack(
        raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"], raw["albedo"], raw["roughness"], raw["metallic"],
        cam.world_view_transform, cam.camera_center, blend_metallic=True)
    ones, zeros = torch.ones(depth.shape), torch.zeros(3)
    sobel = ref.sobel_normal_map(depth, ones, zeros, cam.world_view_transform, fx, fy, 0.5 * W, 0.5 * H)
    loss, l1, ssim = ref.photometric_loss(img1, img2, 0.2)
    return {"build_rotation": ref.build_rotation(raw["rotation"]), "get_scaling": scales, "get_rotation": rotations,
            "get_opacity": opacities, "get_albedo": feats[:, 5:8], "get_roughness": feats[:, 8:9], "get_metallic": feats[:, 9:10],
            "get_normals": feats[:, 2:5], "normal_from_depth": sobel.permute(1, 2, 0), "l1_loss": l1.reshape(1),
            "ssim": ssim.reshape(1), "loss": loss}


def test_oracle_matches_the_committed_reference_vectors():
    gold = np.load(GOLDEN)
    raw, cam, depth, img1, img2 = _maker().inputs()
    mine = _oracle_values(raw, cam, depth, img1, img2)
    for k in gold.files:
        if k == "fx_fy":
            continue
        torch.testing.assert_close(mine[k], torch.from_numpy(gold[k]), rtol=2e-6, atol=2e-7, msg=k)
    # the combination train.py:102-107 forms from the two pinned terms
    want = 0.8 * float(gold["l1_loss"][0]) + 0.2 * (1.0 - float(gold["ssim"][0]))
    assert abs(float(mine["loss"]) - want) <= 1e-6


@pytest.mark.skipif(not os.path.isdir(build_ref.REF_ROOT), reason="reference tree not present (GPU box)")
def test_oracle_matches_the_live_reference_functions():
    """Same comparison against the reference's functions imported from /root/reference, on a second seed, including autograd
    gradients of the activations / normals chain and of the depth-derived normal."""
    import diff_gaussian_rasterization as dgr
    import synthetic_scenes as syn
    build_ref.build_facade()
    facade = build_ref.load_facade(dgr, "ours")
    from scene.gaussian_model import GaussianModel
    from utils.normal_utils import normal_from_depth_image
    lu = build_ref.load_loss_utils(facade)
    g = torch.Generator().manual_seed(8)
    P, H, W = 400, 19, 27
    scene = syn.make_scene(P, seed=31, shell_fraction=0.3)
    raw = {k: v.clone().requires_grad_(True) for k, v in syn.raw_parameters(scene).items()}     # build_rotation is float32-only
    cam = syn.make_cameras(2, W, H)[1]
    wvt, campos = cam.world_view_transform, cam.camera_center
    pc = GaussianModel(3)
    pc._xyz, pc._scaling, pc._rotation, pc._opacity = raw["xyz"], raw["scaling"], raw["rotation"], raw["opacity"]
    pc._albedo, pc._roughness, pc._metallic = raw["albedo"], raw["roughness"], raw["metallic"]
    up_n = torch.randn(P, 3, generator=g)
    with build_ref.cuda_literals_on_cpu():
        n_ref = pc.get_normals(campos)
        (n_ref * up_n).sum().backward()
    g_ref = {k: v.grad.clone() for k, v in raw.items() if v.grad is not None}
    for v in raw.values():
        v.grad = None
    scales, rotations, opacities, feats = ref.activate_and_pack(*[raw[k] for k in (
        "xyz", "scaling", "rotation", "opacity", "albedo", "roughness", "metallic")], wvt, campos, blend_metallic=True)
    torch.testing.assert_close(feats[:, 2:5], n_ref, rtol=1e-6, atol=1e-7)
    (feats[:, 2:5] * up_n).sum().backward()
    for k, gr in g_ref.items():
        torch.testing.assert_close(raw[k].grad, gr, rtol=2e-4, atol=1e-5 * float(gr.abs().max()), msg=k)
    torch.testing.assert_close(scales, pc.get_scaling) and torch.testing.assert_close(opacities, pc.get_opacity)
    # normal from depth, with gradient
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    wvt = wvt.double()
    depth = (2.0 + torch.rand(H, W, generator=g, dtype=torch.float64)).requires_grad_(True)
    intrinsic = torch.tensor([[fx, 0, 0.5 * W], [0, fy, 0.5 * H], [0, 0, 1]], dtype=torch.float64)
    up = torch.randn(H, W, 3, generator=g, dtype=torch.float64)
    a = normal_from_depth_image(depth, intrinsic, wvt.transpose(0, 1).contiguous(), view_space=False)
    (a * up).sum().backward()
    ga, depth.grad = depth.grad.clone(), None
    b = ref.sobel_normal_map(depth, torch.ones(H, W, dtype=torch.float64), torch.zeros(3, dtype=torch.float64), wvt, fx, fy,
                             0.5 * W, 0.5 * H).permute(1, 2, 0)
    (b * up).sum().backward()
    torch.testing.assert_close(b, a, rtol=1e-9, atol=1e-11)
    torch.testing.assert_close(depth.grad, ga, rtol=1e-8, atol=1e-10)
    # photometric terms
    img1 = torch.rand(3, 33, 47, generator=g)
    img2 = (img1 + 0.2 * torch.randn(3, 33, 47, generator=g)).clamp(0, 1)
    loss, l1, ssim = ref.photometric_loss(img1, img2, 0.2)
    assert abs(float(l1) - float(lu.l1_loss(img1, img2))) <= 1e-7 and abs(float(ssim) - float(lu.ssim(img1, img2))) <= 1e-6
