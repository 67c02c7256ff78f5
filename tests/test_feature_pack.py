"""Fused activation + feature packing: oracle self-consistency on CPU, CUDA kernel vs the
PyTorch-eager oracle (forward and autograd backward) on GPU."""
import pytest
import torch

import pack_reference as ref
import synthetic_scenes as syn


def _raw_params(P, seed=3, device="cpu", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    raw = dict(xyz=r(P, 3), scaling=r(P, 3) * 0.7 - 4.0, rotation=r(P, 4) * 2.0, opacity=r(P, 1) * 2.0, albedo=r(P, 3),
               roughness=r(P, 1), metallic=r(P, 1))
    raw["scaling"][: P // 8, 1] = raw["scaling"][: P // 8, 0]          # exact ties in the thinnest-axis argmin
    return {k: v.to(device=device, dtype=dtype) for k, v in raw.items()}


def test_oracle_matches_the_synthetic_scene_packer():
    """Two independent restatements of gaussian_renderer/__init__.py:82-96 agree (the bench/test scenes use the second)."""
    P = 500
    scene = syn.make_scene(P, shell_fraction=0.5)
    cam = syn.make_cameras(1, 64, 48)[0]
    for F, metal in ((9, False), (10, True)):
        feats = syn.pack_features(scene, cam, F)
        s, q, o, f = ref.activate_and_pack(scene.means3D, torch.log(scene.scales), scene.rotations,
                                           torch.logit(scene.opacities), torch.logit(scene.albedo), torch.logit(scene.roughness),
                                           torch.logit(scene.metallic), cam.world_view_transform, cam.camera_center,
                                           blend_metallic=metal)
        torch.testing.assert_close(f, feats, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(s, scene.scales, rtol=1e-5, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("z_depth,blend_metallic", [(False, False), (False, True), (True, True)])
def test_cuda_pack_matches_oracle(z_depth, blend_metallic):
    from diff_gaussian_rasterization.packing import activate_and_pack
    P = 20_000
    cam = syn.make_cameras(1, 320, 240)[0]
    raw64 = {k: v.requires_grad_(True) for k, v in _raw_params(P, dtype=torch.float64).items()}
    out_ref = ref.activate_and_pack(*raw64.values(), cam.world_view_transform.double(), cam.camera_center.double(),
                                    z_depth=z_depth, blend_metallic=blend_metallic)
    raw = {k: v.detach().float().cuda().requires_grad_(True) for k, v in raw64.items()}
    out = activate_and_pack(*raw.values(), cam.world_view_transform.cuda(), cam.camera_center.cuda(),
                            z_depth=z_depth, blend_metallic=blend_metallic)
    g = torch.Generator().manual_seed(11)
    ups = [torch.randn(t.shape, generator=g, dtype=torch.float64) for t in out_ref]
    for a, b in zip(out, out_ref):
        torch.testing.assert_close(a.cpu().double(), b, rtol=1e-5, atol=1e-6)
    torch.autograd.backward(list(out_ref), ups)
    torch.autograd.backward(list(out), [u.float().cuda() for u in ups])
    for k in raw:
        a, b = raw[k].grad.cpu().double(), raw64[k].grad
        if b is None:            # no path in the eager graph (e.g. metallic when it is not blended): we return zeros
            assert float(a.abs().max()) == 0.0, k
            continue
        err = (a - b).abs().max() / b.abs().max().clamp_min(1e-30)
        assert err <= 2e-5, "%s: %.3e" % (k, err)


@pytest.mark.gpu
def test_packed_features_drive_the_rasterizer_like_the_python_packer():
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization.packing import activate_and_pack
    P, W, H, F = 30_000, 400, 300, 10
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=0.6), "cuda")
    cam = syn.camera_to(syn.make_cameras(1, W, H)[0], "cuda")
    s, q, o, f = activate_and_pack(scene.means3D, torch.log(scene.scales), scene.rotations, torch.logit(scene.opacities),
                                   torch.logit(scene.albedo), torch.logit(scene.roughness), torch.logit(scene.metallic),
                                   cam.world_view_transform, cam.camera_center, blend_metallic=True)
    feats = syn.pack_features(scene, cam, F)
    st = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
    a = dgr.forward_raw(scene.means3D, scene.shs, None, o, s, q, None, f, st)
    b = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, feats, st)
    torch.testing.assert_close(a[0], b[0], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(a[3], b[3], rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("z_depth", [False, True])
def test_cuda_derive_maps_matches_oracle(z_depth):
    """Fused post-blend map derivation vs the eager ops, forward and backward."""
    from diff_gaussian_rasterization.packing import derive_maps
    H, W = 213, 331
    g = torch.Generator().manual_seed(5)
    buf64 = torch.randn(10, H, W, generator=g, dtype=torch.float64)
    buf64[2:5, :40, :50] = 0.0                     # background pixels: mask false
    buf64[1] = buf64[1].abs() + 0.5
    buf64.requires_grad_(True)
    cam = syn.make_cameras(1, W, H)[0]
    fx, fy, cx, cy = 1.1 * W, 1.1 * W, 0.5 * W, 0.5 * H
    ref_out = ref.derive_maps(buf64, cam.world_view_transform.double(), fx, fy, cx, cy, z_depth=z_depth)
    buf = buf64.detach().float().cuda().requires_grad_(True)
    out = derive_maps(buf, cam.world_view_transform.cuda(), fx, fy, cx, cy, z_depth=z_depth)
    torch.testing.assert_close(out[0].cpu().double(), ref_out[0], rtol=1e-5, atol=1e-6)
    # plane depth divides by (n . ray): where that is ~0 the value (and its gradient) is ill-conditioned; compare where it is not
    denom = (buf64[1:2] / ref_out[1]).detach().abs() if not z_depth else torch.ones_like(ref_out[1])
    ok = denom > 0.05
    torch.testing.assert_close(out[1].cpu().double()[ok], ref_out[1][ok], rtol=1e-4, atol=1e-5)
    assert torch.equal(out[2].cpu(), ref_out[2])
    ups = [torch.randn(3, H, W, generator=g, dtype=torch.float64), torch.randn(1, H, W, generator=g, dtype=torch.float64) * 1e-2]
    torch.autograd.backward([ref_out[0], ref_out[1]], ups)
    torch.autograd.backward([out[0], out[1]], [u.float().cuda() for u in ups])
    a, b = buf.grad.cpu().double(), buf64.grad
    ok = ok.expand_as(b)
    err = ((a - b).abs() / (b.abs() + 1e-2 * b.abs().mean()))[ok].max()
    assert err <= 1e-3, "relative error %.3e" % err


def test_photometric_loss_oracle_basics():
    """Identical images: SSIM = 1 and L1 = 0, so the loss vanishes; the oracle runs on the CPU in fp64."""
    g = torch.Generator().manual_seed(2)
    img = torch.rand(3, 40, 52, generator=g, dtype=torch.float64)
    loss, l1, ss = ref.photometric_loss(img, img.clone())
    assert float(l1) == 0.0 and abs(float(ss) - 1.0) < 1e-12 and abs(float(loss)) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 213, 331), (3, 32, 32), (1, 7, 5), (3, 1090, 1959)])
def test_cuda_photometric_loss_matches_oracle(shape):
    """Fused L1 + SSIM loss and its gradient vs the reference's torch ops in fp64."""
    from diff_gaussian_rasterization.packing import photometric_loss
    g = torch.Generator().manual_seed(9)
    render64 = torch.rand(shape, generator=g, dtype=torch.float64)
    gt64 = (render64 + 0.15 * torch.randn(shape, generator=g, dtype=torch.float64)).clamp(0, 1)
    render64[:, : shape[1] // 3] = gt64[:, : shape[1] // 3]          # a region of exact agreement: sign(0) = 0 in the L1 term
    render64.requires_grad_(True)
    loss64, l1_64, ssim64 = ref.photometric_loss(render64, gt64, 0.2)
    loss64.backward()
    render = render64.detach().float().cuda().requires_grad_(True)
    loss, terms = photometric_loss(render, gt64.float().cuda(), 0.2, return_terms=True)
    (loss * 1.7).backward()
    assert abs(float(loss.detach()) - float(loss64.detach())) <= 2e-5 * abs(float(loss64.detach())) + 1e-7
    assert abs(float(terms[0]) - float(l1_64)) <= 2e-5 * float(l1_64) + 1e-7
    assert abs(float(terms[1]) - float(ssim64)) <= 2e-5
    a, b = render.grad.cpu().double() / 1.7, render64.grad
    err = (a - b).abs().max() / b.abs().max()
    assert err <= 2e-4, "gradient: %.3e" % err


@pytest.mark.gpu
@pytest.mark.parametrize("P", [5000, 5001])      # 128-bit state access / element-wise fallback for sizes not divisible by 4
def test_fused_adam_matches_torch_adam_over_the_nine_groups(P):
    """One-launch Adam vs torch.optim.Adam(l, lr=0.0, eps=1e-15) with GS-2M's nine groups
    (scene/gaussian_model.py:230-242); the SH gradient arrives as one (P,16,3) block whose column slices feed f_dc / f_rest."""
    from diff_gaussian_rasterization.packing import FusedAdam
    g = torch.Generator().manual_seed(21)
    shapes = {"xyz": (P, 3), "f_dc": (P, 1, 3), "f_rest": (P, 15, 3), "opacity": (P, 1), "scaling": (P, 3), "rotation": (P, 4),
              "albedo": (P, 3), "roughness": (P, 1), "metallic": (P, 1)}
    lrs = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 5e-3, "rotation": 1e-3,
           "albedo": 0.05, "roughness": 0.05, "metallic": 0.05}
    init = {k: torch.randn(s, generator=g) for k, s in shapes.items()}
    ref_params = {k: torch.nn.Parameter(v.clone().cuda()) for k, v in init.items()}
    opt = torch.optim.Adam([{"params": [ref_params[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
    ours = {k: v.clone().cuda() for k, v in init.items()}
    fused = FusedAdam([{"name": k, "param": ours[k], "lr": lrs[k]} for k in shapes])
    for it in range(4):
        sh_grad = (torch.randn(P, 16, 3, generator=g) * 10.0 ** (-it)).cuda()
        grads = {k: (torch.randn(s, generator=g) * 10.0 ** (-2 * it)).cuda() for k, s in shapes.items() if not k.startswith("f_")}
        grads["f_dc"], grads["f_rest"] = sh_grad[:, :1], sh_grad[:, 1:]          # non-contiguous column slices
        if it == 2:
            lrs["xyz"] = 3e-5                                                        # update_learning_rate
            fused.set_lr("xyz", lrs["xyz"])
            opt.param_groups[0]["lr"] = lrs["xyz"]
        for k in shapes:
            ref_params[k].grad = grads[k].contiguous().clone()
        opt.step()
        fused.step(grads)
    for k in shapes:
        torch.testing.assert_close(ours[k], ref_params[k].detach(), rtol=2e-6, atol=2e-7, msg=k)
        torch.testing.assert_close(fused.groups[list(shapes).index(k)]["exp_avg_sq"], opt.state[ref_params[k]]["exp_avg_sq"],
                                   rtol=1e-6, atol=1e-30)


def test_sobel_normal_oracle_on_a_fronto_parallel_plane():
    """A constant depth map is a plane facing the camera: every interior normal is minus the camera's forward axis in world
    space (a x b = (0, 0, -4 z^2 / (fx fy)) in camera space), the border is background, alpha blends between the two."""
    H, W = 20, 30
    cam = syn.make_cameras(1, W, H)[0]
    wvt = cam.world_view_transform.double()
    depth = torch.full((H, W), 3.0, dtype=torch.float64)
    alpha = torch.full((H, W), 0.25, dtype=torch.float64)
    bg = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float64)
    out = ref.sobel_normal_map(depth, alpha, bg, wvt, 1.1 * W, 1.3 * W, 0.5 * W, 0.5 * H)
    expect = -wvt[:3, 2] * 0.25 + bg * 0.75
    torch.testing.assert_close(out[:, 1:-1, 1:-1], expect[:, None, None].expand(3, H - 2, W - 2), rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(out[:, 0, :], (bg * 0.75)[:, None].expand(3, W), rtol=0, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(97, 131), (2, 5), (3, 3)])
def test_cuda_sobel_normal_matches_oracle(shape):
    """Normal map from the depth map (the rest of SURVEY.md section 8f rank 2) vs the reference's torch ops in fp64."""
    from diff_gaussian_rasterization.packing import sobel_normal_map
    H, W = shape
    g = torch.Generator().manual_seed(13)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    depth64 = (2.0 + 0.01 * xx + 0.02 * yy + 0.05 * torch.rand(H, W, generator=g, dtype=torch.float64)).requires_grad_(True)
    alpha64 = torch.rand(H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    bg64 = torch.tensor([0.2, 0.5, 1.0], dtype=torch.float64)
    cam = syn.make_cameras(1, max(W, 4), max(H, 4))[0]
    fx, fy, cx, cy = 1.1 * W, 1.2 * W, 0.5 * W, 0.5 * H
    ref_out = ref.sobel_normal_map(depth64, alpha64, bg64, cam.world_view_transform.double(), fx, fy, cx, cy)
    up = torch.randn(3, H, W, generator=g, dtype=torch.float64)
    ref_out.backward(up)
    depth = depth64.detach().float().cuda().requires_grad_(True)
    alpha = alpha64.detach().float().cuda().requires_grad_(True)
    out = sobel_normal_map(depth, alpha, bg64.float().cuda(), cam.world_view_transform.cuda(), fx, fy, cx, cy)
    out.backward(up.float().cuda())
    torch.testing.assert_close(out.detach().cpu().double(), ref_out.detach(), rtol=2e-4, atol=2e-4)
    torch.testing.assert_close(alpha.grad.cpu().double(), alpha64.grad, rtol=2e-4, atol=2e-4)
    a, b = depth.grad.cpu().double(), depth64.grad
    assert float((a - b).abs().max()) <= 2e-3 * float(b.abs().max()) + 1e-6


@pytest.mark.gpu
def test_fused_adam_follows_densification_state_surgery():
    """prune / cat / replace_tensor mirror _prune_optimizer, cat_tensors_to_optimizer and replace_tensor_to_optimizer
    (scene/gaussian_model.py:372-403, 437-456): after each, FusedAdam keeps stepping like torch.optim.Adam whose state got the
    reference's surgery."""
    from diff_gaussian_rasterization.packing import FusedAdam
    g = torch.Generator().manual_seed(33)
    P = 4001
    shapes = {"xyz": (3,), "opacity": (1,), "f_rest": (15, 3)}
    lrs = {"xyz": 1.6e-4, "opacity": 0.05, "f_rest": 1.25e-4}
    params = {k: torch.nn.Parameter(torch.randn((P,) + s, generator=g).cuda()) for k, s in shapes.items()}
    opt = torch.optim.Adam([{"params": [params[k]], "lr": lrs[k], "name": k} for k in shapes], lr=0.0, eps=1e-15)
    fused = FusedAdam([{"name": k, "param": params[k].detach().clone(), "lr": lrs[k]} for k in shapes])

    def both_step():
        n = fused.groups[0]["param"].shape[0]
        grads = {k: torch.randn((n,) + s, generator=g).cuda() for k, s in shapes.items()}
        for grp in opt.param_groups:
            grp["params"][0].grad = grads[grp["name"]].clone()
        opt.step()
        fused.step(grads)

    def surgery(fn):
        for grp in opt.param_groups:
            old = grp["params"][0]
            st = opt.state.pop(old)
            new_p, st["exp_avg"], st["exp_avg_sq"] = fn(grp["name"], old.detach(), st["exp_avg"], st["exp_avg_sq"])
            grp["params"][0] = torch.nn.Parameter(new_p.requires_grad_(True))
            opt.state[grp["params"][0]] = st

    both_step(); both_step()
    with pytest.raises(RuntimeError, match="does not match"):
        fused.step({k: torch.zeros((P - 1,) + s).cuda() for k, s in shapes.items()})     # stale size is an error, not a wild write
    keep = (torch.rand(P, generator=g) > 0.3).cuda()
    surgery(lambda n, p, m, v: (p[keep], m[keep], v[keep]))
    out = fused.prune(keep)
    assert out["xyz"].shape[0] == int(keep.sum())
    both_step()
    ext = {k: torch.randn((257,) + s, generator=g).cuda() for k, s in shapes.items()}
    surgery(lambda n, p, m, v: (torch.cat((p, ext[n])), torch.cat((m, torch.zeros_like(ext[n]))), torch.cat((v, torch.zeros_like(ext[n])))))
    fused.cat(ext)
    both_step()
    new_op = torch.randn(fused.groups[1]["param"].shape, generator=g).cuda()
    surgery(lambda n, p, m, v: (new_op.clone(), torch.zeros_like(m), torch.zeros_like(v)) if n == "opacity" else (p, m, v))
    fused.replace_tensor(new_op.clone(), "opacity")
    both_step(); both_step()
    for grp, fg in zip(opt.param_groups, fused.groups):
        torch.testing.assert_close(fg["param"], grp["params"][0].detach(), rtol=2e-6, atol=2e-7, msg=grp["name"])
