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
scene, cam, feats, gc, gb = helpers.make_view(cfg["P"], cfg["W"], cfg["H"], cfg["F"], shell=cfg["shell"], cam_radius=cfg["cam_radius"],
                                              cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap"))
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
for k, (ms, n) in st.items():      # per VIEW (a stage may bracket several kernels / be entered twice per view, e.g. the two sorts)
    print("%-16s %8.4f ms   (%d brackets per view)" % (k, ms / iters, n // iters))
    tot += ms / iters
print("%-16s %8.4f ms" % ("sum", tot))
o = helpers.run_ours(dgr, scene, cam, feats, cfg["F"])
ln = (o["ranges"][:, 1] - o["ranges"][:, 0]).float()
print("R = %d, visible = %d, tile lists: mean %.0f, max %d (%.1fx mean); mean n_contrib %.1f" % (
    o["R"], int((o["radii"] > 0).sum()), float(ln.mean()), int(ln.max()), float(ln.max() / ln.mean().clamp_min(1)), float(o["n_contrib"].float().mean())))

# ---- host-side view: wall time per fwd+bwd (one sync at the end) and host time spent inside each call ----
import time
settings = syn.raster_settings_for(cam, cfg["F"], dgr.GaussianRasterizationSettings)
lib.gs2m_profile_enable(0)
torch.cuda.synchronize()
t_f = t_b = 0.0
t0 = time.perf_counter()
for _ in range(iters):
    a = time.perf_counter()
    color, radii, observe, buffer, state = dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                           scene.rotations, None, feats, settings)
    b = time.perf_counter()
    dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats, radii, settings, state)
    c = time.perf_counter()
    t_f += b - a
    t_b += c - b
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / iters * 1e3
print("wall per view %.3f ms | host in forward_raw %.3f ms (includes the R read-back sync) | host in backward_raw %.3f ms"
      % (wall, t_f / iters * 1e3, t_b / iters * 1e3))
print("torch allocator: num_alloc_retries", torch.cuda.memory_stats().get("num_alloc_retries"), "cudaMalloc segments",
      torch.cuda.memory_stats().get("segment.all.allocated"))
