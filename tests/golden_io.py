"""Loads tests/golden/*.npz (outputs of the compiled reference, see tests/golden/make_golden.py)."""
import glob
import os
from typing import NamedTuple

import numpy as np
import torch

import synthetic_scenes as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Settings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    feature_count: int


def golden_files():
    """Rasterizer vectors (make_golden.py); the facade_* / caller_* files next to them have their own loaders."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("facade_", "caller_")))


def facade_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "facade_*.npz")))


def load(path, device="cpu", settings_cls=Settings):
    z = np.load(path)
    P, W, H, F, D = [int(v) for v in z["in_meta"]]
    t = lambda k: torch.from_numpy(z[k]).to(device)  # noqa: E731
    scene = syn.Scene(*[t("in_" + k) for k in syn.Scene._fields])
    settings = settings_cls(image_height=H, image_width=W, tanfovx=float(z["in_tanfov"][0]),
                            tanfovy=float(z["in_tanfov"][1]), bg=t("in_bg"), scale_modifier=1.0,
                            viewmatrix=t("in_viewmatrix"), projmatrix=t("in_projmatrix"), sh_degree=D,
                            campos=t("in_campos"), prefiltered=False, feature_count=F)
    inputs = dict(scene=scene, features=t("in_features"), grad_color=t("in_grad_color"),
                  grad_buffer=t("in_grad_buffer"), settings=settings, P=P, W=W, H=H, F=F, D=D)
    outputs = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out_")}
    return inputs, outputs
