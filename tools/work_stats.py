#!/usr/bin/env python
"""Work statistics of the blend stage on one BASELINE config: how many (instance, pixel-block) pairs have at least one
blending pixel for different block shapes, and how full those blocks are.  Guides the lane mapping of the blend kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import helpers  # noqa: E402
import synthetic_scenes as syn  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

cfg = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "tnt-3m"]
scene, cam, feats, gc, gb = helpers.make_view(cfg["P"], cfg["W"], cfg["H"], cfg["F"], shell=cfg["shell"], cam_radius=cfg["cam_radius"])
o = helpers.run_ours(dgr, scene, cam, feats, cfg["F"])
W, H = cfg["W"], cfg["H"]
tiles_x = (W + 15) // 16
R = o["R"]
keys = o["keys_sorted"]
tile = (keys >> 32).long()
pl = o["point_list"].long()
ranges = o["ranges"].long()
pos = torch.arange(R, device="cuda") - ranges[tile, 0]          # position of the instance in its tile list
ncon = torch.zeros(((H + 15) // 16) * 16, tiles_x * 16, dtype=torch.int64, device="cuda")
ncon[:H, :W] = o["n_contrib"].long().view(H, W)
ra, rb = o["rec_a"], o["rec_b"]
ly, lx = torch.meshgrid(torch.arange(16, device="cuda"), torch.arange(16, device="cuda"), indexing="ij")
shapes = {"16x16": (16, 16), "8x4": (4, 8), "8x2": (2, 8), "4x4": (4, 4), "4x2": (2, 4), "2x2": (2, 2)}
hits = {k: 0 for k in shapes}
valid_px = 0
CH = 200_000
for s in range(0, R, CH):
    e = min(R, s + CH)
    g = pl[s:e]
    t = tile[s:e]
    ty, tx = t // tiles_x, t % tiles_x
    px = (tx * 16)[:, None, None] + lx[None]
    py = (ty * 16)[:, None, None] + ly[None]
    a = ra[g]
    b = rb[g]
    dx = a[:, 0, None, None] - px.float()
    dy = a[:, 1, None, None] - py.float()
    power = -0.5 * (a[:, 2, None, None] * dx * dx + b[:, 0, None, None] * dy * dy) - a[:, 3, None, None] * dx * dy
    alpha = torch.clamp(b[:, 1, None, None] * torch.exp(power), max=0.99)
    reached = pos[s:e, None, None] < ncon[py, px]
    v = (power <= 0) & (alpha >= 1.0 / 255.0) & reached
    valid_px += int(v.sum())
    for k, (bh, bw) in shapes.items():
        blk = v.view(-1, 16 // bh, bh, 16 // bw, bw).any(dim=4).any(dim=2)
        hits[k] += int(blk.sum())
print("R = %d instances, %d blending (pixel, instance) pairs" % (R, valid_px))
for k, (bh, bw) in shapes.items():
    n = hits[k]
    print("%-6s blocks with >=1 blending pixel: %11d   fill %.3f   lane-slots %12d" % (k, n, valid_px / (n * bh * bw), n * bh * bw))
