#!/usr/bin/env python
"""Runs K forward+backward passes of one implementation on one BASELINE config (to be wrapped by ncu).
Usage: python tools/profile_view.py {ours|reference} <config> [iters] [views]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import helpers  # noqa: E402
import synthetic_scenes as syn  # noqa: E402

impl, cfg_name = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = syn.CONFIGS[cfg_name]
scene, cam, feats, gc, gb = helpers.make_view(cfg["P"], cfg["W"], cfg["H"], cfg["F"], shell=cfg["shell"],
                                              cam_radius=cfg["cam_radius"])
if impl == "ours":
    import diff_gaussian_rasterization as dgr
    run = lambda: helpers.run_ours(dgr, scene, cam, feats, cfg["F"], gc, gb)
else:
    import build_ref
    ref = build_ref.load()
    run = lambda: helpers.run_reference(ref, scene, cam, feats, cfg["F"], gc, gb)
for _ in range(iters):
    out = run()
torch.cuda.synchronize()
print("R =", out["R"], "visible =", int((out["radii"] > 0).sum()))
