"""Seeded synthetic scenes, cameras and upstream gradients for the rasterizer's tests and benchmark
(generator spec: SURVEY.md section 8(d)), plus the caller-side packing of the 10-column ``features`` tensor the way
GS-2M's render facade builds it (gaussian_renderer/__init__.py:82-111, scene/gaussian_model.py:146-160,
scene/cameras.py:64-67, utils/graphics_utils.py:51-71).

Everything is generated on the CPU in fp32 from ``torch.Generator`` seeds, so the reference rasterizer, the CPU
oracle and the CUDA implementation all see identical bits.  Pure torch; no dependency on the CUDA library.
"""
import math
from typing import NamedTuple

import torch

SCENE_SEED, CAMERA_SEED, GRAD_SEED = 1234, 4321, 999

# BASELINE.json configs: name -> (P, W, H, feature_count, n_views, camera radius, surface fraction)
CONFIGS = {
    "plumbing-100k": dict(P=100_000, W=800, H=800, F=5, views=1, cam_radius=3.0, shell=0.7),
    "dtu-300k": dict(P=300_000, W=800, H=600, F=5, views=49, cam_radius=3.0, shell=0.7),
    "shiny-500k": dict(P=500_000, W=800, H=800, F=9, views=8, cam_radius=3.0, shell=0.7),
    "tnt-3m": dict(P=3_000_000, W=1959, H=1090, F=10, views=8, cam_radius=2.2, shell=0.0),
    "dp-6m": dict(P=6_000_000, W=1959, H=1090, F=10, views=64, cam_radius=2.2, shell=0.0),
    # load-imbalance probe (not a BASELINE config): 40 % of the Gaussians in a tight blob at the origin, so that a few dozen
    # tiles carry lists more than ten times the mean length
    "clustered-1m": dict(P=1_000_000, W=1959, H=1090, F=10, views=8, cam_radius=2.2, shell=0.0, cluster=(0.4, 0.02)),
    # the same scene right after GS-2M's periodic opacity reset (scene/gaussian_model.py reset_opacity: min(opacity, 0.01)):
    # no pixel saturates, every pixel walks its whole tile list, so the longest lists decide when the blend kernels finish
    "clustered-1m-reset": dict(P=1_000_000, W=1959, H=1090, F=10, views=8, cam_radius=2.2, shell=0.0, cluster=(0.4, 0.02),
                               opacity_cap=0.01),
}


class Scene(NamedTuple):
    means3D: torch.Tensor    # (P,3)
    scales: torch.Tensor     # (P,3)  activated (exp)
    rotations: torch.Tensor  # (P,4)  unit quaternions (w,x,y,z)
    opacities: torch.Tensor  # (P,1)  activated (sigmoid)
    shs: torch.Tensor        # (P,16,3)
    albedo: torch.Tensor     # (P,3)
    roughness: torch.Tensor  # (P,1)
    metallic: torch.Tensor   # (P,1)


class Camera(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    world_view_transform: torch.Tensor  # (4,4) row-major tensor of W2V^T
    full_proj_transform: torch.Tensor   # (4,4) row-major tensor of (P W2V)^T
    camera_center: torch.Tensor         # (3,)


def make_scene(P, seed=SCENE_SEED, shell_fraction=0.0, extent=1.0, cluster=None, opacity_cap=None):
    """``cluster = (fraction, sigma)``: that fraction of the Gaussians (taken from the end of the arrays) is moved into an
    isotropic normal blob of that standard deviation around the origin (a dense object in a sparse scene)."""
    g = torch.Generator().manual_seed(seed)
    means = (torch.rand(P, 3, generator=g) * 2.0 - 1.0) * extent
    if shell_fraction > 0:
        n_shell = int(P * shell_fraction)
        d = torch.randn(n_shell, 3, generator=g)
        d = d / d.norm(dim=1, keepdim=True)
        r = 0.6 + 0.02 * torch.randn(n_shell, 1, generator=g)
        means[:n_shell] = d * r
    mu = math.log(0.35 * P ** (-1.0 / 3.0))
    log_scales = mu + 0.5 * torch.randn(P, 3, generator=g)
    flat_axis = torch.randint(0, 3, (P,), generator=g)
    log_scales[torch.arange(P), flat_axis] += math.log(0.1)
    scales = torch.exp(log_scales)
    q = torch.randn(P, 4, generator=g)
    rotations = q / q.norm(dim=1, keepdim=True)
    opacities = torch.sigmoid(2.0 * torch.randn(P, 1, generator=g))
    shs = torch.empty(P, 16, 3)
    shs[:, 0, :] = torch.rand(P, 3, generator=g) * 3.0 - 1.5
    shs[:, 1:, :] = 0.1 * torch.randn(P, 15, 3, generator=g)
    albedo = torch.sigmoid(torch.randn(P, 3, generator=g))
    roughness = torch.sigmoid(torch.randn(P, 1, generator=g))
    metallic = torch.sigmoid(torch.randn(P, 1, generator=g))
    if cluster is not None:
        n_cl = int(P * cluster[0])
        if n_cl > 0:
            means[P - n_cl:] = cluster[1] * torch.randn(n_cl, 3, generator=g)
    if opacity_cap is not None:
        opacities = opacities.clamp(max=opacity_cap)
    return Scene(means.float(), scales.float(), rotations.float(), opacities.float(), shs.float(), albedo.float(),
                 roughness.float(), metallic.float())


def _projection(znear, zfar, tan_half_x, tan_half_y):
    # utils/graphics_utils.py:51-71 with symmetric frustum
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 1.0 / tan_half_x
    Pm[1, 1] = 1.0 / tan_half_y
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_cameras(n, W, H, radius=3.0, seed=CAMERA_SEED, focal_scale=1.1):
    g = torch.Generator().manual_seed(seed)
    cams = []
    fx = focal_scale * W
    tanx, tany = W / (2.0 * fx), H / (2.0 * fx)
    for _ in range(n):
        d = torch.randn(3, generator=g)
        d = d / d.norm()
        centre = d * radius * (1.0 + 0.05 * (torch.rand(1, generator=g).item() - 0.5))
        fwd = -centre / centre.norm()
        down0 = torch.tensor([0.0, -1.0, 0.0])
        right = torch.linalg.cross(down0, fwd)
        if right.norm() < 1e-4:
            right = torch.tensor([1.0, 0.0, 0.0])
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        w2v = torch.eye(4)
        w2v[:3, :3] = torch.stack([right, down, fwd])
        w2v[:3, 3] = -w2v[:3, :3] @ centre
        wvt = w2v.t().contiguous().float()                       # scene/cameras.py:64
        proj_t = _projection(0.01, 100.0, tanx, tany).t().float()
        full = (wvt.unsqueeze(0).bmm(proj_t.unsqueeze(0))).squeeze(0).contiguous()  # scene/cameras.py:66
        campos = wvt.inverse()[3, :3].contiguous()               # scene/cameras.py:67
        cams.append(Camera(H, W, tanx, tany, wvt, full, campos))
    return cams


def make_upstream_grads(W, H, F, seed=GRAD_SEED):
    g = torch.Generator().manual_seed(seed + 1)
    n = float(W * H)
    grad_color = torch.randn(3, H, W, generator=g) / n
    grad_buffer = torch.randn(10, H, W, generator=g) / n
    grad_buffer[F:] = 0.0
    return grad_color.float(), grad_buffer.float()


def rotation_matrices(q):
    """(P,3,3) rotation matrices of unit quaternions (w,x,y,z) (utils/general_utils.py:72-104)."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.empty(q.shape[0], 3, 3, dtype=q.dtype, device=q.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def raw_parameters(scene: Scene):
    """The scene as GS-2M's raw (pre-activation) parameter tensors: inverse of the getters of scene/gaussian_model.py:113-172
    (log scale, un-normalised quaternion, logit opacity / albedo / roughness / metallic)."""
    eps = 1e-6

    def logit(p):
        p = p.clamp(eps, 1 - eps)
        return torch.log(p / (1 - p))
    return {"xyz": scene.means3D.clone(), "scaling": torch.log(scene.scales), "rotation": scene.rotations * 1.7,
            "opacity": logit(scene.opacities), "albedo": logit(scene.albedo), "roughness": logit(scene.roughness),
            "metallic": logit(scene.metallic)}


def pack_features(scene: Scene, cam: Camera, feature_count, z_depth=False):
    """The (P,10) side-channel tensor: [1, distance, normal(3), albedo(3), roughness, metallic]
    (gaussian_renderer/__init__.py:82-96). Column 9 is only filled when feature_count is even (blend_metallic)."""
    P = scene.means3D.shape[0]
    R = rotation_matrices(scene.rotations)
    axis = torch.argmin(scene.scales, dim=1)
    normals = R[torch.arange(P, device=R.device), :, axis]                 # column of R for the thinnest axis
    flip = ((cam.camera_center[None] - scene.means3D) * normals).sum(-1) < 0
    normals = torch.where(flip[:, None], -normals, normals)
    normals = normals / normals.norm(dim=1, keepdim=True)
    wvt = cam.world_view_transform
    cam_normals = normals @ wvt[:3, :3]
    cam_points = scene.means3D @ wvt[:3, :3] + wvt[3, :3]
    feats = torch.zeros(P, 10, dtype=torch.float32, device=scene.means3D.device)
    feats[:, 0] = 1.0
    feats[:, 1] = cam_points[:, 2] if z_depth else (cam_normals * cam_points).sum(-1).abs()
    feats[:, 2:5] = normals
    feats[:, 5:8] = scene.albedo
    feats[:, 8:9] = scene.roughness
    if feature_count in (2, 6, 10):
        feats[:, 9:10] = scene.metallic
    return feats


def scene_to(scene: Scene, device):
    return Scene(*[t.to(device) for t in scene])


def camera_to(cam: Camera, device):
    return Camera(cam.image_height, cam.image_width, cam.tanfovx, cam.tanfovy, cam.world_view_transform.to(device),
                  cam.full_proj_transform.to(device), cam.camera_center.to(device))


def raster_settings_for(cam: Camera, feature_count, settings_cls, sh_degree=3, bg=None):
    """Build the 12-field settings tuple the way render() does (gaussian_renderer/__init__.py:98-110)."""
    dev = cam.world_view_transform.device
    if bg is None:
        bg = torch.zeros(3, dtype=torch.float32, device=dev)
    return settings_cls(image_height=int(cam.image_height), image_width=int(cam.image_width), tanfovx=cam.tanfovx,
                        tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0, viewmatrix=cam.world_view_transform,
                        projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center,
                        prefiltered=False, feature_count=feature_count)
