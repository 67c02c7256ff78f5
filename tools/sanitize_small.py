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

for path in ("depthfirst", "sort64", "ranked"):
    os.environ["GS2M_BINNING"] = path
    for (P, W, H, F) in ((1500, 100, 70, 10), (800, 64, 48, 5), (300, 33, 17, 0)):
        scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=0.6)
        sc = scene.scales.clone()
        sc[:10] *= 40.0
        scene = scene._replace(scales=sc.contiguous())
        o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
        torch.cuda.synchronize()
        print(path, P, W, H, F, "R =", o["R"], "sum|dL_dmeans3D| =", float(o["dL_dmeans3D"].abs().sum()))
