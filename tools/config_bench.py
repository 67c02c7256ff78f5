#!/usr/bin/env python
"""Drop-in-API throughput on every BASELINE config: ours vs the compiled reference, both driven exactly the way an
unmodified GS-2M iteration drives them (gaussian_renderer/__init__.py:111-123 + loss.backward()):

    GaussianRasterizer(settings)(means3D=, means2D=, opacities=, shs=, scales=, rotations=, features=)
    torch.autograd.backward([color, buffer], [grad_color, grad_buffer])

on ONE stream (torch's current stream), `views_per_iter` views per iteration with the gradients accumulating in the
leaves' .grad (train.py:95 + utils/loss_utils.py:253 render two views per iteration).  Timed with CUDA events around a
run of iterations and with the host clock next to it (a single-stream caller pays the host-side bubbles).

Usage: python tools/config_bench.py [--configs a,b,...] [--iters N] [--out gpurun_out/r2_configs.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import synthetic_scenes as syn  # noqa: E402


def make_leaves(scene):
    return dict(means3D=scene.means3D.clone().requires_grad_(True), opacities=scene.opacities.clone().requires_grad_(True),
                shs=scene.shs.clone().requires_grad_(True), scales=scene.scales.clone().requires_grad_(True),
                rotations=scene.rotations.clone().requires_grad_(True))


def time_impl(mod, cfg, F, n_views, iters, warmup, device):
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=cfg["shell"], cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap")), device)
    cams = [syn.camera_to(c, device) for c in syn.make_cameras(n_views, W, H, radius=cfg["cam_radius"])]
    feats = [syn.pack_features(scene, c, F).requires_grad_(True) for c in cams]
    gc, gb = [t.to(device) for t in syn.make_upstream_grads(W, H, F)]
    leaves = make_leaves(scene)
    settings = [syn.raster_settings_for(c, F, mod.GaussianRasterizationSettings) for c in cams]
    m2d = torch.zeros(P, 4, device=device, requires_grad=True)
    info = {}

    def fwd(v):
        return mod.GaussianRasterizer(settings[v])(
            means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"], colors_precomp=None,
            scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None, features=feats[v])

    def iteration():
        for t in list(leaves.values()) + feats + [m2d]:
            t.grad = None                                       # optimizer.zero_grad(set_to_none=True), train.py:259
        for v in range(n_views):
            color, radii, observe, buffer = fwd(v)
            torch.autograd.backward([color, buffer], [gc, gb])
        info["visible"] = radii

    for _ in range(warmup):
        iteration()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        iteration()
    e1.record()
    torch.cuda.synchronize(device)
    wall = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    # forward and backward apart (events around each call; includes the host bubbles inside each)
    f_ms = b_ms = 0.0
    for _ in range(iters):
        for v in range(n_views):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            color, radii, observe, buffer = fwd(v)
            b.record()
            torch.autograd.backward([color, buffer], [gc, gb])
            c.record()
            torch.cuda.synchronize(device)
            f_ms += a.elapsed_time(b)
            b_ms += b.elapsed_time(c)
    n = iters * n_views
    return {"ms_per_view": dev_ms / n, "wall_ms_per_view": wall / n, "fwd_ms": f_ms / n, "bwd_ms": b_ms / n,
            "visible": int((info["visible"] > 0).sum())}


def time_graphed(dgr, cfg, F, n_views, iters, device):
    """The same views with forward (no_wait: nothing touches the host) + backward captured ONCE into a CUDA graph per view and
    replayed: what the kernels cost without Python, ctypes and per-kernel launch overhead (fixed shapes; the camera matrices
    live in static buffers a caller would refresh with three tiny copies)."""
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=cfg["shell"], cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap")), device)
    cams = [syn.camera_to(c, device) for c in syn.make_cameras(n_views, W, H, radius=cfg["cam_radius"])]
    feats = [syn.pack_features(scene, c, F) for c in cams]
    gc, gb = [t.to(device) for t in syn.make_upstream_grads(W, H, F)]
    settings = [syn.raster_settings_for(c, F, dgr.GaussianRasterizationSettings) for c in cams]
    graphs = []
    side = torch.cuda.Stream()
    for v in range(n_views):
        args = (scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, feats[v], settings[v])
        cap = int(1.25 * dgr.forward_raw(*args, capacity=0)[4].num_rendered) + 65536
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            out = dgr.forward_raw(*args, capacity=cap, no_wait=True)
            dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v], out[1],
                             settings[v], out[4])
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = dgr.forward_raw(*args, capacity=cap, no_wait=True)
            grads = dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v],
                                     out[1], settings[v], out[4])
        graphs.append((g, out, grads))
    for g, _, _ in graphs:
        g.replay()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for g, _, _ in graphs:
            g.replay()
    e1.record()
    torch.cuda.synchronize(device)
    return {"ms_per_view": e0.elapsed_time(e1) / (iters * n_views)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="plumbing-100k,dtu-300k,shiny-500k,shiny-500k:10,tnt-3m,clustered-1m")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--views-per-iter", type=int, default=2)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r2_configs.json"))
    ap.add_argument("--no-reference", action="store_true")
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    import diff_gaussian_rasterization as dgr
    import build_ref
    ref = None if args.no_reference or not build_ref.available() else build_ref.load()
    rows = []
    for name in args.configs.split(","):
        cname, _, f_over = name.partition(":")
        cfg = syn.CONFIGS[cname]
        F = int(f_over) if f_over else cfg["F"]
        iters = args.iters if cfg["P"] < 2_000_000 else max(3, args.iters // 3)
        row = {"config": cname, "P": cfg["P"], "W": cfg["W"], "H": cfg["H"], "F": F, "views_per_iter": args.views_per_iter}
        row["ours"] = time_impl(dgr, cfg, F, args.views_per_iter, iters, args.warmup, device)
        torch.cuda.empty_cache()
        row["ours_graph"] = time_graphed(dgr, cfg, F, args.views_per_iter, iters, device)
        torch.cuda.empty_cache()
        if ref is not None:
            row["reference"] = time_impl(ref, cfg, F, args.views_per_iter, iters, args.warmup, device)
            row["speedup"] = row["reference"]["ms_per_view"] / row["ours"]["ms_per_view"]
            row["speedup_wall"] = row["reference"]["wall_ms_per_view"] / row["ours"]["wall_ms_per_view"]
        rows.append(row)
        print(json.dumps(row))
        sys.stdout.flush()
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)
    print("| config | P | image | F | ours ms/view (fwd / bwd) | ours, CUDA-graph replay | reference ms/view (fwd / bwd) | speed-up (device / wall) |")
    print("|---|---|---|---|---|---|---|---|")
    for r in rows:
        o = r["ours"]
        rr = r.get("reference")
        print("| %s | %d | %dx%d | %d | %.3f (%.3f / %.3f) | %.3f | %s | %s |" % (
            r["config"], r["P"], r["W"], r["H"], r["F"], o["ms_per_view"], o["fwd_ms"], o["bwd_ms"], r["ours_graph"]["ms_per_view"],
            "%.3f (%.3f / %.3f)" % (rr["ms_per_view"], rr["fwd_ms"], rr["bwd_ms"]) if rr else "-",
            "%.2fx / %.2fx" % (r["speedup"], r["speedup_wall"]) if rr else "-"))


if __name__ == "__main__":
    main()
