#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the COMPILED REFERENCE rasterizer (oracle/_ref, built by
oracle/build_ref.py from /root/reference) on small seeded scenes on a GPU:

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz; copy them to tests/golden/

Each file holds the inputs and every output of the reference's forward and backward for one view, so the CPU oracle
(oracle/cpu_rasterizer.py) can be pinned against the reference without a GPU, and the CUDA implementation can be
checked against the same vectors on a GPU box that has no /root/reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))
import build_ref  # noqa: E402
import helpers  # noqa: E402
import synthetic_scenes as syn  # noqa: E402

CASES = {
    # name: P, W, H, F, sh_degree, bg, shell, scene_seed
    "full_f10": dict(P=1200, W=96, H=64, F=10, D=3, bg=(0.0, 0.0, 0.0), shell=0.7, seed=1234),
    "ragged_f5_bg": dict(P=800, W=75, H=50, F=5, D=1, bg=(0.2, 0.5, 0.7), shell=0.5, seed=77),
}


def main():
    ref = build_ref.load()
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        scene, cam, feats, gc, gb = helpers.make_view(c["P"], c["W"], c["H"], c["F"], shell=c["shell"],
                                                      scene_seed=c["seed"])
        bg = torch.tensor(c["bg"], dtype=torch.float32, device="cuda")
        r = helpers.run_reference(ref, scene, cam, feats, c["F"], gc, gb, sh_degree=c["D"], bg=bg)
        torch.cuda.synchronize()
        arrays = {"in_" + k: getattr(scene, k).cpu().numpy() for k in scene._fields}
        arrays.update(in_features=feats.cpu().numpy(), in_grad_color=gc.cpu().numpy(), in_grad_buffer=gb.cpu().numpy(),
                      in_viewmatrix=cam.world_view_transform.cpu().numpy(),
                      in_projmatrix=cam.full_proj_transform.cpu().numpy(), in_campos=cam.camera_center.cpu().numpy(),
                      in_bg=bg.cpu().numpy(),
                      in_meta=np.array([c["P"], c["W"], c["H"], c["F"], c["D"]], dtype=np.int64),
                      in_tanfov=np.array([cam.tanfovx, cam.tanfovy], dtype=np.float64))
        for k, v in r.items():
            if k == "R":
                arrays["out_R"] = np.array([v], dtype=np.int64)
            elif k in ("clamped", "tiles_touched", "point_offsets", "depths", "means2D", "conic_opacity", "rgb", "cov3D"):
                vis = (r["radii"] > 0)
                t = v.view(c["P"], -1) if k == "clamped" else v
                t = t.clone()
                t[~vis] = 0   # culled entries are uninitialised memory in the reference
                arrays["out_" + k] = t.cpu().numpy()
            else:
                arrays["out_" + k] = v.cpu().numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
        print(name, "R =", r["R"], {k: v.shape for k, v in arrays.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
