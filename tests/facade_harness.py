"""Drives the reference's UNMODIFIED render facade (gaussian_renderer/__init__.py:21-175, loaded from the bytecode that
oracle/build_ref.py makes of it) with real ``GaussianModel`` / ``Camera`` objects filled from the seeded synthetic scenes,
once per rasterizer module.  Test infrastructure ("the drop-in's acceptance harness")."""
import math
from types import SimpleNamespace

import torch

import build_ref
import synthetic_scenes as syn

PARAM_NAMES = ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "_albedo", "_roughness", "_metallic")

# name -> render() keyword arguments and pipe flags (gaussian_renderer/__init__.py:25-29,61,71,83,90,143)
VARIANTS = {
    "plain": dict(),
    "geometry": dict(geometry_stage=True),
    "geometry_sobel": dict(geometry_stage=True, sobel_normal=True),
    "material": dict(material_stage=True),
    "material_metallic_sobel": dict(material_stage=True, blend_metallic=True, sobel_normal=True),
    "geometry_zdepth": dict(geometry_stage=True, pipe=dict(z_depth=True)),
    "material_python_sh": dict(material_stage=True, pipe=dict(convert_SHs_python=True)),
    "geometry_python_cov": dict(geometry_stage=True, pipe=dict(compute_cov3D_python=True)),
}
FLOAT_MAPS = ("render", "alpha_map", "distance_map", "depth_map", "normal_map", "albedo_map", "roughness_map", "metallic_map",
              "local_normal_map", "sobel_map")
EXACT = ("visibility_filter", "radii", "observe", "normal_mask")


def facade(rasterizer_module, tag):
    return build_ref.load_facade(rasterizer_module, tag)


def make_model(scene: syn.Scene, device="cuda", sh_degree=3, active_sh_degree=3):
    """A reference ``GaussianModel`` whose raw parameters reproduce ``scene`` (inverse activations, synthetic_scenes.raw_parameters)."""
    from scene.gaussian_model import GaussianModel
    raw = syn.raw_parameters(scene)
    pc = GaussianModel(sh_degree)
    pc.active_sh_degree = active_sh_degree
    leaf = lambda t: torch.nn.Parameter(t.detach().clone().to(device).contiguous().requires_grad_(True))  # noqa: E731
    pc._xyz = leaf(raw["xyz"])
    pc._features_dc = leaf(scene.shs[:, :1, :])
    pc._features_rest = leaf(scene.shs[:, 1:, :])
    pc._scaling, pc._rotation, pc._opacity = leaf(raw["scaling"]), leaf(raw["rotation"]), leaf(raw["opacity"])
    pc._albedo, pc._roughness, pc._metallic = leaf(raw["albedo"]), leaf(raw["roughness"]), leaf(raw["metallic"])
    return pc


def make_camera(cam: syn.Camera, device="cuda"):
    """A reference ``Camera`` (scene/cameras.py:19-97) without its image-loading constructor: the attributes render() and the
    camera's own get_rays / get_calib_matrix_nerf read."""
    from scene.cameras import Camera
    c = Camera.__new__(Camera)
    torch.nn.Module.__init__(c)
    c.image_width, c.image_height = int(cam.image_width), int(cam.image_height)
    c.FoVx, c.FoVy = 2.0 * math.atan(cam.tanfovx), 2.0 * math.atan(cam.tanfovy)
    c.Fx = c.image_width / (2.0 * math.tan(c.FoVx / 2.0))        # utils/graphics_utils.py fov2focal
    c.Fy = c.image_height / (2.0 * math.tan(c.FoVy / 2.0))
    c.Cx, c.Cy = 0.5 * c.image_width, 0.5 * c.image_height
    c.world_view_transform = cam.world_view_transform.to(device)
    c.full_proj_transform = cam.full_proj_transform.to(device)
    c.camera_center = cam.camera_center.to(device)
    return c


def make_pipe(**flags):
    p = dict(compute_cov3D_python=False, convert_SHs_python=False, z_depth=False)
    p.update(flags)
    return SimpleNamespace(**p)


def loss_weights(out, seed=77):
    """Seeded random weights, one per differentiable map of the render() dict (the scalar the backward pass starts from)."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for k in FLOAT_MAPS:
        if out.get(k) is not None:
            w[k] = (torch.randn(out[k].shape, generator=g) / out[k].numel()).to(out[k].device)
    return w


def scalar_loss(out, weights):
    total = 0.0
    for k, w in weights.items():
        m = out[k]
        if k == "depth_map":
            m = m.clamp(-20.0, 20.0)      # plane depth divides by (n . ray): keep the near-singular pixels from dominating
        total = total + (m * w).sum()
    return total


def run_variant(module, pc, camera, bg, name, weights=None):
    """render() + backward for one variant.  Returns (out dict, {param name: grad}, viewspace_points.grad, weights)."""
    kw = dict(VARIANTS[name])
    pipe = make_pipe(**kw.pop("pipe", {}))
    for n in PARAM_NAMES:
        getattr(pc, n).grad = None
    out = module.render(camera, pc, pipe, bg, **kw)
    if weights is None:
        weights = loss_weights(out)
    scalar_loss(out, weights).backward()
    grads = {n: (getattr(pc, n).grad.detach().clone() if getattr(pc, n).grad is not None else None) for n in PARAM_NAMES}
    vs = out["viewspace_points"].grad
    return out, grads, (None if vs is None else vs.detach().clone()), weights
