"""Densification statistics fused into the library against the reference's eager ops
(train.py:225-241, scene/gaussian_model.py:569-573) on the same per-view outputs."""
import pytest
import torch

import synthetic_scenes as syn
import view_parallel as vp


def test_eager_statistics_follow_the_reference_ops():
    P = 9
    st = vp.DensificationStats(P, "cpu")
    radii = torch.tensor([0, 3, 5, 0, 2, 7, 1, 0, 4], dtype=torch.int32)
    observe = torch.tensor([0, 2, 0, 1, 9, 0, 1, 0, 0], dtype=torch.int32)
    st.max_radii2D.fill_(2.5)
    st.update_forward(radii, observe)
    assert st.max_radii2D.tolist() == [2.5, 3.0, 2.5, 2.5, 2.5, 2.5, 2.5, 2.5, 2.5]
    assert st.observe_cnt.view(-1).tolist() == [0, 1, 0, 1, 1, 0, 1, 0, 0]
    g = torch.arange(4 * P, dtype=torch.float32).view(P, 4)
    st.update_backward_eager(g, radii)
    vis = radii > 0
    assert torch.equal(st.denom.view(-1), vis.float())
    torch.testing.assert_close(st.xyz_gradient_accum.view(-1), torch.where(vis, g[:, :2].norm(dim=1), torch.zeros(P)))
    torch.testing.assert_close(st.xyz_gradient_accum_abs.view(-1), torch.where(vis, g[:, 2:].norm(dim=1), torch.zeros(P)))


@pytest.mark.gpu
def test_fused_statistics_match_eager_ops_over_several_views():
    import diff_gaussian_rasterization as dgr
    P, W, H, F, n_views = 40_000, 400, 300, 10, 3
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=0.6), "cuda")
    cams = [syn.camera_to(c, "cuda") for c in syn.make_cameras(n_views, W, H)]
    gc, gb = (t.cuda() for t in syn.make_upstream_grads(W, H, F))
    fused, eager = vp.DensificationStats(P, "cuda"), vp.DensificationStats(P, "cuda")
    fused.max_radii2D.fill_(6.0)
    eager.max_radii2D.fill_(6.0)
    grads = dgr.alloc_grads(P, 16, "cuda")
    for k, cam in enumerate(cams):
        feats = syn.pack_features(scene, cam, F)
        st = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
        color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                               scene.rotations, None, feats, st)
        # fused: accumulate mode (the bucket holds the running sum, the statistics must still see this view's gradient)
        dgr.update_view_stats(radii, observe, fused.max_radii2D, fused.observe_cnt.view(-1))
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii, st, state,
                         grads=grads, accumulate=k > 0, densify_stats=fused.backward_args())
        # eager: the reference's ops on the view's own gradient
        own = dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii, st, state)
        mask = (observe > 0) & (radii > 0)
        eager.max_radii2D.copy_(torch.where(mask, torch.max(eager.max_radii2D, radii), eager.max_radii2D))
        eager.observe_cnt[observe > 0] += 1
        vis = radii > 0
        g2d = own["dL_dmeans2D"]
        eager.xyz_gradient_accum[vis] += torch.norm(g2d[vis, :2], dim=-1, keepdim=True)
        eager.xyz_gradient_accum_abs[vis] += torch.norm(g2d[vis, 2:], dim=-1, keepdim=True)
        eager.denom[vis] += 1
    assert float(eager.denom.max()) == n_views and float(eager.max_radii2D.max()) > 6.0
    assert torch.equal(fused.max_radii2D, eager.max_radii2D)
    assert torch.equal(fused.observe_cnt, eager.observe_cnt)
    assert torch.equal(fused.denom, eager.denom)
    torch.testing.assert_close(fused.xyz_gradient_accum, eager.xyz_gradient_accum, rtol=1e-5, atol=1e-12)
    torch.testing.assert_close(fused.xyz_gradient_accum_abs, eager.xyz_gradient_accum_abs, rtol=1e-5, atol=1e-12)
