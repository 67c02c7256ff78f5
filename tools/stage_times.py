#!/usr/bin/env python
"""Per-stage device times (library cudaEvent hooks) of our rasterizer on one BASELINE config: quick perf iteration."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import helpers  # noqa: E402
import synthetic_scenes as syn  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402
from diff_gaussian_rasterization import _native  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "tnt-3m"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = syn.CONFIGS[cfg_name]
scene, cam, feats, gc, gb = helpers.make_view(cfg["P"], cfg["W"], cfg["H"], cfg["F"], shell=cfg["shell"], cam_radius=cfg["cam_radius"])
lib = _native.load()
for _ in range(3):
    helpers.run_ours(dgr, scene, cam, feats, cfg["F"], gc, gb)
torch.cuda.synchronize()
lib.gs2m_profile_enable(1)
_native.profile_read()
for _ in range(iters):
    helpers.run_ours(dgr, scene, cam, feats, cfg["F"], gc, gb)
torch.cuda.synchronize()
st = _native.profile_read()
tot = 0.0
for k, (ms, n) in st.items():
    print("%-16s %8.4f ms" % (k, ms / max(n, 1)))
    tot += ms / max(n, 1)
print("%-16s %8.4f ms" % ("sum", tot))
