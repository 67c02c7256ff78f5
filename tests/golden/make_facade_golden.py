#!/usr/bin/env python
"""Generates tests/golden/facade_*.npz: the reference's UNMODIFIED render() (gaussian_renderer/__init__.py:21-175, from the
bytecode oracle/build_ref.py makes of it) around the COMPILED REFERENCE rasterizer, on small seeded scenes on a GPU:

    gpurun -- python tests/golden/make_facade_golden.py      # writes gpurun_out/golden/facade_*.npz; copy to tests/golden/

Each file holds every entry of the dict render() returns, the seeded loss weights, and the gradients of the nine raw parameter
groups and of viewspace_points after ``scalar_loss(out).backward()``.  The inputs are regenerated from the seeds in `in_meta`.
The GPU box that runs the tests has no /root/reference: these vectors are what pins oracle/pack_reference.render_like (the
restated facade) and the fused caller-side kernels against the real one there.
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import build_ref  # noqa: E402
import facade_harness as fh  # noqa: E402
import synthetic_scenes as syn  # noqa: E402

CASES = {
    # name: variant of facade_harness.VARIANTS, P, W, H, bg, shell, scene seed, active SH degree
    "facade_material_metallic_sobel": dict(variant="material_metallic_sobel", P=1500, W=96, H=64, bg=(0.1, 0.2, 0.3), shell=0.7,
                                           seed=1234, D=3),
    "facade_geometry_zdepth": dict(variant="geometry_zdepth", P=900, W=75, H=50, bg=(0.0, 0.0, 0.0), shell=0.5, seed=77, D=2),
    "facade_plain": dict(variant="plain", P=600, W=40, H=56, bg=(1.0, 1.0, 1.0), shell=0.0, seed=5, D=0),
}


def case_inputs(c, device="cuda"):
    scene = syn.make_scene(c["P"], seed=c["seed"], shell_fraction=c["shell"])
    cam = syn.make_cameras(1, c["W"], c["H"])[0]
    bg = torch.tensor(c["bg"], dtype=torch.float32, device=device)
    return scene, cam, bg


def main():
    ref = build_ref.load()
    module = fh.facade(ref, "ref")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        scene, cam, bg = case_inputs(c)
        pc = fh.make_model(scene, active_sh_degree=c["D"])
        camera = fh.make_camera(cam)
        out, grads, vs_grad, weights = fh.run_variant(module, pc, camera, bg, c["variant"])
        torch.cuda.synchronize()
        arrays = {"in_meta": np.array([c["P"], c["W"], c["H"], c["seed"], c["D"]], dtype=np.int64),
                  "in_shell": np.array([c["shell"]]), "in_bg": np.array(c["bg"], dtype=np.float32),
                  "in_tanfov": np.array([math.tan(camera.FoVx * 0.5), math.tan(camera.FoVy * 0.5)], dtype=np.float64)}
        for k, v in out.items():
            if v is not None and k != "viewspace_points":
                arrays["out_" + k] = v.detach().cpu().numpy()
        for k, v in weights.items():
            arrays["w_" + k] = v.cpu().numpy()
        for k, v in grads.items():
            if v is not None:
                arrays["grad" + k] = v.cpu().numpy()
        arrays["grad_viewspace_points"] = vs_grad.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
