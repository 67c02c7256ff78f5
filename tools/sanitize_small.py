#!/usr/bin/env python
"""Small forward+backward runs for compute-sanitizer (memcheck / racecheck / synccheck): all binning paths, a few F."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import helpers  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

for path in ("depthfirst", "sort64"):
    os.environ["GS2M_BINNING"] = path
    for (P, W, H, F) in ((1500, 100, 70, 10), (800, 64, 48, 5), (300, 33, 17, 0)):
        scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
        sc = scene.scales.clone()
        sc[:10] *= 40.0
        scene = scene._replace(scales=sc.contiguous())
        for rep in range(3):    # depthfirst: the first call is exact, the next speculative; the last one is made to overflow
            if rep == 2 and path == "depthfirst":
                dgr._R_HINT.clear()
                dgr._R_HINT[(0, W, H)] = 1
            o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
            torch.cuda.synchronize()
        print(path, P, W, H, F, "R =", o["R"], "sum|dL_dmeans3D| =", float(o["dL_dmeans3D"].abs().sum()))

# caller-side stages: packing forward/backward (+ accumulate), post-blend maps, densification statistics, accumulate mode 2
import synthetic_scenes as syn  # noqa: E402
import view_parallel as vp  # noqa: E402
from diff_gaussian_rasterization.packing import activate_and_pack, derive_maps  # noqa: E402

os.environ.pop("GS2M_BINNING", None)
P, W, H, F = 1200, 90, 60, 10
scene = syn.scene_to(syn.make_scene(P, shell_fraction=0.6), "cuda")
cam = syn.camera_to(syn.make_cameras(1, W, H)[0], "cuda")
raw = {k: v.cuda().requires_grad_(True) for k, v in syn.raw_parameters(scene).items()}
s, q, o, f = activate_and_pack(*raw.values(), cam.world_view_transform, cam.camera_center, blend_metallic=True)
st = syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, o.detach(), s.detach(), q.detach(), None,
                                                       f.detach(), st)
gc, gb = (t.cuda() for t in syn.make_upstream_grads(W, H, F))
buckets = vp.ParameterBuckets(P, 16, "cuda")
stats = vp.DensificationStats(P, "cuda")
stats.update_forward(radii, observe)
for acc in (False, True):
    dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, s.detach(), q.detach(), None, f.detach(), radii, st, state,
                     grads=buckets.raster, accumulate=buckets.raster_accumulate_mode(acc), densify_stats=stats.backward_args())
    buckets.chain_view({k: v.detach() for k, v in raw.items()}, cam.world_view_transform, cam.camera_center, radii, blend_metallic=True)
(s.sum() + q.sum() + o.sum() + f.sum()).backward()
bufg = buffer.clone().requires_grad_(True)
ln, d, m = derive_maps(bufg, cam.world_view_transform, 1.1 * W, 1.1 * W, 0.5 * W, 0.5 * H)
(ln.sum() + d.sum()).backward()
torch.cuda.synchronize()
print("caller-side stages ok: |xyz grad| =", float(buckets.tensors["xyz"].abs().sum()), "denom max =", float(stats.denom.max()))
