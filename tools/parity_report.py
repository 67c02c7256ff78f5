#!/usr/bin/env python
"""Diagnostic (not a test): our CUDA rasterizer vs the compiled reference on seeded synthetic views.
Prints mismatch counts for the bit-exact fields, error norms for the float fields, and CUDA-event timings.
Usage: python tools/parity_report.py [P W H F [cam_radius [shell]]] ...  (defaults: a small and a medium case)"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import build_ref  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402
import helpers  # noqa: E402


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def report(P, W, H, F, cam_radius=3.0, shell=0.0, timing=True):
    ref = build_ref.load()
    scene, cam, feats, gc, gb = helpers.make_view(P, W, H, F, shell=shell, cam_radius=cam_radius)
    r = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    o = helpers.run_ours(dgr, scene, cam, feats, F, gc, gb)
    torch.cuda.synchronize()
    res = {"P": P, "W": W, "H": H, "F": F, "R_ref": int(r["R"]), "R_ours": int(o["R"])}
    vis = r["radii"] > 0
    res["visible"] = int(vis.sum())
    res["radii_mismatch"] = int((r["radii"] != o["radii"]).sum())
    res["tiles_touched_mismatch"] = int((r["tiles_touched"] != o["tiles_touched"]).sum())
    for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
        a, b = r[k][vis], o[k][vis]
        res[k + "_bit_mismatch"] = int((helpers.bits(a) != helpers.bits(b)).sum())
        res[k + "_maxabs"] = float((a - b).abs().max()) if a.numel() else 0.0
    cl_r = r["clamped"].view(P, 3)[vis]
    cl_o = o["clamped"].view(P, 3)[vis]
    res["clamped_mismatch"] = int((cl_r != cl_o).sum())
    if r["R"] == o["R"]:
        res["keys_mismatch"] = int((r["keys_sorted"] != o["keys_sorted"]).sum())
        res["point_list_mismatch"] = int((r["point_list"] != o["point_list"]).sum())
    res["ranges_mismatch"] = int((r["ranges"] != o["ranges"]).sum())
    res["n_contrib_mismatch"] = int((r["n_contrib"] != o["n_contrib"]).sum())
    res["final_T_bit_mismatch"] = int((helpers.bits(r["final_T"]) != helpers.bits(o["final_T"])).sum())
    res["observe_mismatch"] = int((r["observe"] != o["observe"]).sum())
    for k in ("color", "buffer"):
        d = (r[k] - o[k]).abs()
        res[k + "_maxabs"] = float(d.max())
        res[k + "_maxrel"] = float((d / r[k].abs().clamp_min(1e-6)).max())
    for k in ("dL_dmeans2D", "dL_dcolor", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscale", "dL_drot",
              "dL_dfeatures"):
        e, l2 = helpers.grad_errors(o[k], r[k])
        res[k] = "max/max=%.2e l2=%.2e" % (e, l2)
    # reference self-noise (float atomics): run the reference backward twice
    r2 = helpers.run_reference(ref, scene, cam, feats, F, gc, gb)
    res["ref_self_noise"] = {k: "%.2e" % helpers.grad_errors(r2[k], r[k])[0] for k in (
        "dL_dmeans2D", "dL_dcolor", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscale", "dL_drot", "dL_dfeatures")}
    res["list_len_mean"] = float((r["ranges"][:, 1] - r["ranges"][:, 0]).float().mean())
    res["list_len_max"] = int((r["ranges"][:, 1] - r["ranges"][:, 0]).max())
    res["n_contrib_mean"] = float(r["n_contrib"].float().mean())
    if timing:
        settings = helpers.syn.raster_settings_for(cam, F, dgr.GaussianRasterizationSettings)
        empty = torch.Tensor([])
        bg = torch.zeros(3, device="cuda")

        def ref_fwd():
            return ref._C.rasterize_gaussians(bg, scene.means3D, empty, scene.opacities, scene.scales, scene.rotations, 1.0,
                                              empty, feats, cam.world_view_transform, cam.full_proj_transform, cam.tanfovx,
                                              cam.tanfovy, H, W, scene.shs, 3, cam.camera_center, False, F)
        Rr, color, radii, observe, buffer, geom, binning, img = ref_fwd()

        def ref_bwd():
            return ref._C.rasterize_gaussians_backward(bg, scene.means3D, radii, buffer, empty, scene.scales, scene.rotations,
                                                       1.0, empty, feats, cam.world_view_transform, cam.full_proj_transform,
                                                       cam.tanfovx, cam.tanfovy, gc, gb, scene.shs, 3, cam.camera_center, geom,
                                                       Rr, binning, img, F)

        def our_fwd():
            return dgr.forward_raw(scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None,
                                   feats, settings)
        c2, radii2, obs2, buf2, state = our_fwd()

        def our_bwd():
            return dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats,
                                    radii2, settings, state)
        res["ms_ref_fwd(min,med)"] = timed(ref_fwd)
        res["ms_ref_bwd(min,med)"] = timed(ref_bwd)
        res["ms_our_fwd(min,med)"] = timed(our_fwd)
        res["ms_our_bwd(min,med)"] = timed(our_bwd)
    return res


if __name__ == "__main__":
    args = [float(a) for a in sys.argv[1:]]
    cases = []
    while args:
        chunk, args = args[:6], args[6:]
        cases.append(chunk)
    if not cases:
        cases = [[20000, 320, 240, 10, 3.0, 0.7], [100000, 800, 800, 5, 3.0, 0.7]]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    allres = []
    for c in cases:
        P, W, H, F = int(c[0]), int(c[1]), int(c[2]), int(c[3])
        rad = c[4] if len(c) > 4 else 3.0
        shell = c[5] if len(c) > 5 else 0.0
        t0 = time.time()
        res = report(P, W, H, F, rad, shell)
        res["wall_s"] = round(time.time() - t0, 1)
        allres.append(res)
        print(json.dumps(res, indent=1))
        sys.stdout.flush()
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(allres, f, indent=1)
