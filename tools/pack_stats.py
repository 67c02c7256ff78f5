#!/usr/bin/env python
"""What would packing several list entries into one warp iteration buy the blend kernels?  For a sample of tiles: every
(entry, 8x4 warp block) pair the backward evaluates, the set of 4x2 (or 2x2) pixel sub-blocks that hold a blending pixel, and a
simulation of the greedy rule "take consecutive entries while their sub-block sets are disjoint" (per-pixel list order is kept).
Prints iterations now vs packed, and lane utilisation.  Usage: python tools/pack_stats.py [config] [n_tiles]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import helpers  # noqa: E402
import synthetic_scenes as syn  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

cfg = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "tnt-3m"]
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 160
scene, cam, feats, gc, gb = helpers.make_view(cfg["P"], cfg["W"], cfg["H"], cfg["F"], shell=cfg["shell"], cam_radius=cfg["cam_radius"],
                                              cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap"))
o = helpers.run_ours(dgr, scene, cam, feats, cfg["F"])
W, H = cfg["W"], cfg["H"]
tiles_x, tiles_y = (W + 15) // 16, (H + 15) // 16
ranges = o["ranges"].long().cpu().numpy()
pl = o["point_list"].long()
masks = o["masks"].cpu().numpy()
ncon = torch.zeros(tiles_y * 16, tiles_x * 16, dtype=torch.int64, device="cuda")
ncon[:H, :W] = o["n_contrib"].long().view(H, W)
ra, rb = o["rec_a"], o["rec_b"]
ly, lx = torch.meshgrid(torch.arange(16, device="cuda"), torch.arange(16, device="cuda"), indexing="ij")
rng = np.random.RandomState(0)
tiles = rng.choice(tiles_x * tiles_y, size=min(n_sample, tiles_x * tiles_y), replace=False)
tot = dict(hit=0, live=0, lanes=0, it42=0, it22=0, it42_2=0)
for t in tiles:
    b, e = ranges[t]
    if e <= b:
        continue
    g = pl[b:e]
    ty, tx = t // tiles_x, t % tiles_x
    px = (tx * 16 + lx)[None].float()
    py = (ty * 16 + ly)[None].float()
    a, bb = ra[g], rb[g]
    dx = a[:, 0, None, None] - px
    dy = a[:, 1, None, None] - py
    power = -0.5 * (a[:, 2, None, None] * dx * dx + bb[:, 0, None, None] * dy * dy) - a[:, 3, None, None] * dx * dy
    alpha = torch.clamp(bb[:, 1, None, None] * torch.exp(power), max=0.99)
    pos = torch.arange(e - b, device="cuda")[:, None, None]
    v = (power <= 0) & (alpha >= 1.0 / 255.0) & (pos < ncon[ty * 16:(ty + 1) * 16, tx * 16:(tx + 1) * 16][None])
    v = v.view(-1, 4, 4, 2, 8)                       # [entry, block row, y in block, block col, x in block]
    v = v.permute(0, 1, 3, 2, 4)                     # [entry, block row, block col, 4, 8]
    m = masks[b:e]
    for w in range(8):
        blk = v[:, w >> 1, w & 1]                    # [entry, 4, 8]
        hit = torch.from_numpy(((m >> w) & 1).astype(bool)).cuda()
        deep = int(ncon[ty * 16 + (w >> 1) * 4: ty * 16 + (w >> 1) * 4 + 4, tx * 16 + (w & 1) * 8: tx * 16 + (w & 1) * 8 + 8].max())
        hit = hit & (torch.arange(e - b, device="cuda") < deep)          # the backward never walks behind the deepest contributor
        blk = blk[hit]
        tot["hit"] += int(hit.sum())
        live = blk.any(dim=2).any(dim=1)
        blk = blk[live]
        tot["live"] += int(live.sum())
        tot["lanes"] += int(blk.sum())
        s42 = blk.view(-1, 2, 2, 2, 4).any(dim=4).any(dim=2).view(-1, 4).cpu().numpy()     # 4 sub-blocks of 4x2 pixels
        s22 = blk.view(-1, 2, 2, 4, 2).any(dim=4).any(dim=2).view(-1, 8).cpu().numpy()     # 8 sub-blocks of 2x2 pixels
        for key, sm in (("it42", s42), ("it22", s22)):
            bits = (sm * (1 << np.arange(sm.shape[1]))).sum(axis=1)
            used, it = 0, 0
            for x in bits:
                if used & x or used == 0:
                    it += 1
                    used = 0
                used |= int(x)
            tot[key] += it
        # packing at most 2 entries per iteration (cheaper bookkeeping)
        bits = (s42 * (1 << np.arange(4))).sum(axis=1)
        it, k = 0, 0
        while k < len(bits):
            if k + 1 < len(bits) and not (bits[k] & bits[k + 1]):
                k += 2
            else:
                k += 1
            it += 1
        tot["it42_2"] += it
print("sampled %d tiles of %s" % (len(tiles), sys.argv[1] if len(sys.argv) > 1 else "tnt-3m"))
print("mask hits (staged entries per warp block)      %9d" % tot["hit"])
print("entries with >= 1 blending pixel               %9d   lane utilisation %.3f" % (tot["live"], tot["lanes"] / (32.0 * tot["live"])))
for key, name in (("it42", "greedy packing, 4x2 sub-blocks"), ("it42_2", "pairs only,     4x2 sub-blocks"), ("it22", "greedy packing, 2x2 sub-blocks")):
    print("%-46s %9d   %.3f of the iterations, lane utilisation %.3f" % (name, tot[key], tot[key] / tot["live"], tot["lanes"] / (32.0 * tot[key])))
