"""CPU oracle: a pure-PyTorch restatement of GS-2M's differentiable Gaussian rasterizer (forward + backward).

TEST INFRASTRUCTURE ONLY — never imported by the product package (``gs-2m_b200/``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` fallback legs use it.

PARITY PINNING: the reference ships no golden vectors or tests for this path, so this
restatement is pinned against outputs of the *compiled reference itself* (``oracle/_ref``, built by
``oracle/build_ref.py``) run on a B200: ``tests/golden/*.npz`` hold those outputs for small seeded scenes together
with the generating script ``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks this file against
them on the CPU.  Floating point here is IEEE fp32/fp64 on the host, so integer outputs (radii, n_contrib) can differ
from the GPU on measure-zero borderline cases; bit-exact integer parity is asserted GPU-vs-GPU against the compiled
reference, not against this file.

What follows the reference (file:line under submodules/diff-gaussian-rasterization/cuda_rasterizer/):
  preprocess_forward  forward.cu:145-241 (+ computeCov3D :109-142, computeCov2D :70-104 without the 0.3 dilation,
                      computeColorFromSH :20-67, in_frustum auxiliary.h:140-162, ndc2Pix :40-42, getRect :44-53)
  build_lists         rasterizer_impl.cu:63-129,265-305 (duplicateWithKeys, 64-bit key sort, identifyTileRanges)
  blend_forward       forward.cu:246-372
  blend_backward      backward.cu:413-598
  preprocess backward backward.cu:153-410 through autograd of ``preprocess_forward`` with the reference's two
                      deliberate inconsistencies re-created: the conic's gradient is taken at cov2D + 0.3*I
                      (backward.cu:205-207) and the frustum-clamped t.x / t.y are constants w.r.t. t.z (:183-184,268-270).
                      (The reference's 1/(denom^2 + 1e-7) regulariser, :211, is not reproduced: relative effect
                      1e-7/denom^2 <= ~1e-5 since denom >= 0.09.)
The blend is vectorised per tile as a [pixels x list] alpha matrix with an exclusive cumulative product for T and
first-index termination, as BASELINE.md section 3 prescribes for the CPU baseline.
"""
import math

import numpy as np
import torch

TILE = 16
SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435)


def _sh_to_rgb(deg, shs, dirs):
    """forward.cu:20-67 (same basis as utils/sh_utils.py:57-115). shs (P,M,3), dirs (P,3) unit."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * shs[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * shs[:, 1] + SH_C1 * z * shs[:, 2] - SH_C1 * x * shs[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + SH_C2[0] * xy * shs[:, 4] + SH_C2[1] * yz * shs[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * shs[:, 6]
               + SH_C2[3] * xz * shs[:, 7] + SH_C2[4] * (xx - yy) * shs[:, 8])
        if deg > 2:
            res = (res + SH_C3[0] * y * (3 * xx - yy) * shs[:, 9] + SH_C3[1] * xy * z * shs[:, 10]
                   + SH_C3[2] * y * (4 * zz - xx - yy) * shs[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * shs[:, 12]
                   + SH_C3[4] * x * (4 * zz - xx - yy) * shs[:, 13] + SH_C3[5] * z * (xx - yy) * shs[:, 14]
                   + SH_C3[6] * x * (xx - 3 * yy) * shs[:, 15])
    return res + 0.5


def _cov3d(scales, mod, rot):
    """forward.cu:109-142: Sigma = R diag(mod*s)^2 R^T, quaternion (r,x,y,z) used as given."""
    r, x, y, z = rot[:, 0], rot[:, 1], rot[:, 2], rot[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)
    S = mod * scales
    M = R * S[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    return torch.stack([Sigma[:, 0, 0], Sigma[:, 0, 1], Sigma[:, 0, 2], Sigma[:, 1, 1], Sigma[:, 1, 2], Sigma[:, 2, 2]], dim=1)


def preprocess_forward(means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, sh_degree=0,
                       shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                       scale_modifier=1.0, backward_quirks=False):
    """Per-Gaussian stage. Returns a dict of (P,...) tensors; entries of culled Gaussians are zero / radius 0.

    With ``backward_quirks`` the returned conic / mean carry the *reference's backward* dependency structure (see the
    module docstring) while their values stay those of the forward; used only to drive autograd in ``backward``.
    """
    dt = means3D.dtype
    P = means3D.shape[0]
    vm, pm = viewmatrix.to(dt), projmatrix.to(dt)
    focal_x, focal_y = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    ones = torch.ones(P, 1, dtype=dt)
    p_hom4 = torch.cat([means3D, ones], dim=1)
    p_view = p_hom4 @ vm                                     # transformPoint4x3 (auxiliary.h:67-74)
    depth = p_view[:, 2]
    in_front = depth > 0.2                                   # auxiliary.h:150
    p_clip = p_hom4 @ pm
    p_w = 1.0 / (p_clip[:, 3] + 0.0000001)
    ndc = p_clip[:, :2] * p_w[:, None]

    cov3D = cov3D_precomp if cov3D_precomp is not None else _cov3d(scales, scale_modifier, rotations)

    # EWA projection (forward.cu:70-104)
    tx, ty, tz = p_view[:, 0], p_view[:, 1], p_view[:, 2]
    tz_safe = torch.where(in_front, tz, torch.ones_like(tz))
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    txtz, tytz = tx / tz_safe, ty / tz_safe
    cx = torch.clamp(txtz, -limx, limx) * tz_safe
    cy = torch.clamp(tytz, -limy, limy) * tz_safe
    if backward_quirks:  # clamped t.x/t.y are constants in the reference's backward (backward.cu:183-184,268-270)
        cx = torch.where((txtz < -limx) | (txtz > limx), cx.detach(), tx)
        cy = torch.where((tytz < -limy) | (tytz > limy), cy.detach(), ty)
    zeros = torch.zeros_like(tz)
    J = torch.stack([focal_x / tz_safe, zeros, -(focal_x * cx) / (tz_safe * tz_safe),
                     zeros, focal_y / tz_safe, -(focal_y * cy) / (tz_safe * tz_safe)], dim=1).view(P, 2, 3)
    Rw = vm[:3, :3].t()                                      # W2V rotation (math matrix)
    T = J @ Rw                                               # (P,2,3)
    Sig = torch.stack([cov3D[:, 0], cov3D[:, 1], cov3D[:, 2], cov3D[:, 1], cov3D[:, 3], cov3D[:, 4],
                       cov3D[:, 2], cov3D[:, 4], cov3D[:, 5]], dim=1).view(P, 3, 3)
    cov2 = T @ Sig @ T.transpose(1, 2)
    a, b, c = cov2[:, 0, 0], cov2[:, 0, 1], cov2[:, 1, 1]
    det = a * c - b * b
    ok = in_front & (det != 0)
    det_safe = torch.where(ok, det, torch.ones_like(det))
    conic = torch.stack([c / det_safe, -b / det_safe, a / det_safe], dim=1)
    if backward_quirks:  # gradient of the conic as if cov2D had 0.3 on its diagonal (backward.cu:205-219)
        a3, c3 = a + 0.3, c + 0.3
        det3 = a3 * c3 - b * b
        conic_q = torch.stack([c3 / det3, -b / det3, a3 / det3], dim=1)
        conic = conic_q + (conic - conic_q).detach()
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.clamp(lam, min=0.0))).to(torch.int64)
    # ndc2Pix in double (auxiliary.h:40-42)
    pix = (((ndc.double() + 1.0) * torch.tensor([W, H], dtype=torch.float64) - 1.0) * 0.5).to(dt)
    # getRect (auxiliary.h:44-53)
    tiles_x, tiles_y = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    rf = radius.to(dt)
    pd = pix.detach()

    def tile_lo(p, n):
        return torch.clamp(torch.trunc((p - rf) / TILE), 0, n).to(torch.int64)

    def tile_hi(p, n):
        return torch.clamp(torch.trunc((p + rf + TILE - 1) / TILE), 0, n).to(torch.int64)
    x0, x1 = tile_lo(pd[:, 0], tiles_x), tile_hi(pd[:, 0], tiles_x)
    y0, y1 = tile_lo(pd[:, 1], tiles_y), tile_hi(pd[:, 1], tiles_y)
    tiles_touched = (x1 - x0) * (y1 - y0)
    visible = ok & (tiles_touched > 0)

    if colors_precomp is None:
        d = means3D - campos.to(dt)[None]
        d = d / d.norm(dim=1, keepdim=True)
        raw = _sh_to_rgb(sh_degree, shs, d)
        clamped = raw < 0
        rgb = torch.clamp_min(raw, 0.0)
    else:
        rgb, clamped = colors_precomp, torch.zeros(P, 3, dtype=torch.bool)

    vi = visible.to(torch.int64)
    return dict(depth=depth, visible=visible, radii=(radius * vi).to(torch.int32), means2D=pix, conic=conic,
                opacity=opacities.reshape(-1), rgb=rgb, clamped=clamped, cov3D=cov3D,
                rect=torch.stack([x0, y0, x1, y1], dim=1), tiles_touched=(tiles_touched * vi).to(torch.int32),
                tiles_x=tiles_x, tiles_y=tiles_y)


def build_lists(pre):
    """Keys ((tile << 32) | depth bits), stable sort, tile ranges (rasterizer_impl.cu:63-129,288-305)."""
    vis = torch.nonzero(pre["visible"]).reshape(-1)
    rect = pre["rect"][vis].numpy()
    depth_bits = pre["depth"].detach().to(torch.float32)[vis].numpy().view(np.uint32).astype(np.uint64)
    tiles_x, tiles_y = pre["tiles_x"], pre["tiles_y"]
    w = rect[:, 2] - rect[:, 0]
    h = rect[:, 3] - rect[:, 1]
    cnt = (w * h).astype(np.int64)
    R = int(cnt.sum())
    owner = np.repeat(np.arange(len(vis)), cnt)
    start = np.cumsum(cnt) - cnt
    local = np.arange(R) - np.repeat(start, cnt)
    ww = np.repeat(w, cnt)
    ty = np.repeat(rect[:, 1], cnt) + local // np.maximum(ww, 1)
    tx = np.repeat(rect[:, 0], cnt) + local % np.maximum(ww, 1)
    tile = (ty * tiles_x + tx).astype(np.uint64)
    keys = (tile << np.uint64(32)) | depth_bits[owner]
    vals = vis.numpy()[owner].astype(np.uint32)
    order = np.argsort(keys, kind="stable")
    keys_sorted, point_list = keys[order], vals[order]
    n_tiles = tiles_x * tiles_y
    ranges = np.zeros((n_tiles, 2), dtype=np.uint32)
    if R > 0:
        t_sorted = (keys_sorted >> np.uint64(32)).astype(np.int64)
        bounds = np.nonzero(np.diff(t_sorted))[0] + 1
        starts = np.concatenate([[0], bounds])
        ends = np.concatenate([bounds, [R]])
        ranges[t_sorted[starts], 0] = starts
        ranges[t_sorted[starts], 1] = ends
    return dict(keys_sorted=keys_sorted, point_list=point_list, ranges=ranges, R=R)


def _tile_pixels(t, tiles_x, W, H, dt):
    ty, tx = divmod(t, tiles_x)
    ys = torch.arange(ty * TILE, min((ty + 1) * TILE, H))
    xs = torch.arange(tx * TILE, min((tx + 1) * TILE, W))
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return xx.reshape(-1), yy.reshape(-1)


def _tile_alpha(pre, ids, px, py, dt):
    """forward.cu:325-339 for all (pixel, list entry) pairs of a tile. Returns alpha (0 where skipped), G, d."""
    xy = pre["means2D"].detach()[ids]
    con = pre["conic"].detach()[ids]
    op = pre["opacity"].detach()[ids]
    dx = xy[None, :, 0] - px[:, None].to(dt)
    dy = xy[None, :, 1] - py[:, None].to(dt)
    power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
    G = torch.exp(power)
    alpha = torch.clamp_max(op[None, :] * G, 0.99)
    valid = (power <= 0) & (alpha >= 1.0 / 255.0)
    return torch.where(valid, alpha, torch.zeros_like(alpha)), valid, G, dx, dy


def blend_forward(pre, lists, features, bg, W, H, F, tiles=None):
    """forward.cu:246-372. Returns color (3,H,W), buffer (10,H,W), final_T (H,W), n_contrib (H,W), observe (P)."""
    dt = pre["means2D"].dtype
    P = pre["means2D"].shape[0]
    color = torch.zeros(3, H, W, dtype=dt)
    buffer = torch.zeros(10, H, W, dtype=dt)
    final_T = torch.ones(H, W, dtype=dt)
    n_contrib = torch.zeros(H, W, dtype=torch.int32)
    observe = torch.zeros(P, dtype=torch.int64)
    rgb = pre["rgb"].detach()
    pl = torch.from_numpy(lists["point_list"].astype(np.int64))
    tiles_x = pre["tiles_x"]
    tile_iter = range(tiles_x * pre["tiles_y"]) if tiles is None else tiles
    for t in tile_iter:
        s, e = int(lists["ranges"][t, 0]), int(lists["ranges"][t, 1])
        px, py = _tile_pixels(t, tiles_x, W, H, dt)
        if e <= s:
            color[:, py, px] = bg.to(dt)[:, None].expand(3, px.numel()).clone()
            continue
        ids = pl[s:e]
        alpha, valid, G, dx, dy = _tile_alpha(pre, ids, px, py, dt)
        one_m = 1.0 - alpha
        T_incl = torch.cumprod(one_m, dim=1)                       # test_T after each entry
        T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
        stop = valid & (T_incl < 0.0001)
        L = alpha.shape[1]
        idx = torch.arange(L)[None, :].expand_as(stop)
        first_stop = torch.where(stop, idx, torch.full_like(idx, L)).min(dim=1).values   # termination index
        contrib = valid & (idx < first_stop[:, None])
        w = torch.where(contrib, alpha * T_excl, torch.zeros_like(alpha))
        color[:, py, px] = (w @ rgb[ids]).t()
        if F > 0:
            buffer[:F, py, px] = (w @ features[ids][:, :F].to(dt)).t()
        T_fin = torch.where(first_stop < L, T_excl.gather(1, first_stop.clamp_max(L - 1)[:, None])[:, 0], T_incl[:, -1])
        final_T[py, px] = T_fin
        color[:, py, px] += T_fin[None, :] * bg.to(dt)[:, None]
        last = torch.where(contrib, idx + 1, torch.zeros_like(idx)).max(dim=1).values
        n_contrib[py, px] = last.to(torch.int32)
        obs = (contrib & (T_excl > 0.5)).sum(dim=0)
        observe.index_add_(0, ids, obs)
    return color, buffer, final_T, n_contrib, observe.to(torch.int32)


def blend_backward(pre, lists, features, bg, W, H, F, final_T, n_contrib, grad_color, grad_buffer, tiles=None):
    """backward.cu:413-598 vectorised per tile. Returns dL_dmeans2D (P,4), dL_dconic (P,3: xx,xy,yy), dL_dopacity (P),
    dL_dcolor (P,3), dL_dfeatures (P,10)."""
    dt = pre["means2D"].dtype
    P = pre["means2D"].shape[0]
    dmean = torch.zeros(P, 4, dtype=dt)
    dconic = torch.zeros(P, 3, dtype=dt)
    dopac = torch.zeros(P, dtype=dt)
    dcol = torch.zeros(P, 3, dtype=dt)
    dfeat = torch.zeros(P, 10, dtype=dt)
    rgb = pre["rgb"].detach()
    pl = torch.from_numpy(lists["point_list"].astype(np.int64))
    tiles_x = pre["tiles_x"]
    bgd = bg.to(dt)
    tile_iter = range(tiles_x * pre["tiles_y"]) if tiles is None else tiles
    for t in tile_iter:
        s, e = int(lists["ranges"][t, 0]), int(lists["ranges"][t, 1])
        if e <= s:
            continue
        px, py = _tile_pixels(t, tiles_x, W, H, dt)
        ids = pl[s:e]
        alpha, valid, G, dx, dy = _tile_alpha(pre, ids, px, py, dt)
        L = alpha.shape[1]
        idx = torch.arange(L)[None, :].expand_as(valid)
        contrib = valid & (idx < n_contrib[py, px].to(torch.int64)[:, None])      # backward.cu:517-519
        a = torch.where(contrib, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a
        T_excl = torch.cumprod(one_m, dim=1)
        T_excl = torch.cat([torch.ones_like(T_excl[:, :1]), T_excl[:, :-1]], dim=1)  # T before each entry
        w = a * T_excl                                                              # dchannel_dcolor (:533)
        gpix = torch.cat([grad_color[:, py, px], grad_buffer[:F, py, px]], dim=0).t().to(dt)   # (pix, 3+F)
        cvec = torch.cat([rgb[ids], features[ids][:, :F].to(dt)], dim=1)            # (L, 3+F)
        # colour / feature gradients (:551,560)
        gc = w.t() @ gpix
        dcol.index_add_(0, ids, gc[:, :3])
        if F > 0:
            dfeat[:, :F] = dfeat[:, :F].index_add(0, ids, gc[:, 3:])
        # dL/dalpha: (c - accum) . dL * T, accum = colour blended behind the entry (:546-562)
        cd = gpix @ cvec.t()                                                        # (pix, L)  <c_j, dL_p>
        wcd = w * cd
        behind = torch.flip(torch.cumsum(torch.flip(wcd, dims=[1]), dim=1), dims=[1]) - wcd   # sum_{k>j} w_k cd_k
        T_after = T_excl * one_m
        dL_dalpha = (cd * T_excl - behind / one_m.clamp_min(1e-30))
        T_fin = final_T[py, px].to(dt)
        bg_dot = (gpix[:, :3] * bgd[None, :]).sum(dim=1)
        dL_dalpha = dL_dalpha + (-T_fin[:, None] / one_m) * bg_dot[:, None]         # :566-572
        dL_dalpha = torch.where(contrib, dL_dalpha, torch.zeros_like(dL_dalpha))
        del T_after
        con = pre["conic"].detach()[ids]
        op = pre["opacity"].detach()[ids]
        dL_dG = op[None, :] * dL_dalpha
        gdx, gdy = G * dx, G * dy
        dG_ddelx = -gdx * con[None, :, 0] - gdy * con[None, :, 1]
        dG_ddely = -gdy * con[None, :, 2] - gdx * con[None, :, 1]
        mx = dL_dG * dG_ddelx * (0.5 * W)
        my = dL_dG * dG_ddely * (0.5 * H)
        dmean.index_add_(0, ids, torch.stack([mx.sum(0), my.sum(0), mx.abs().sum(0), my.abs().sum(0)], dim=1))
        dconic.index_add_(0, ids, torch.stack([(-0.5 * gdx * dx * dL_dG).sum(0), (-0.5 * gdx * dy * dL_dG).sum(0),
                                               (-0.5 * gdy * dy * dL_dG).sum(0)], dim=1))
        dopac.index_add_(0, ids, (G * dL_dalpha).sum(0))
    return dmean, dconic, dopac, dcol, dfeat


class CpuRasterizer:
    """Forward + backward of one view on the CPU with the reference binding's argument meaning
    (diff_gaussian_rasterization/__init__.py:42-141)."""

    def __init__(self, dtype=torch.float32):
        self.dtype = dtype

    def forward(self, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, features, settings,
                tiles=None):
        dt = self.dtype
        cv = lambda t: None if t is None or t.numel() == 0 else t.detach().to("cpu", dt)  # noqa: E731
        self.inp = dict(means3D=cv(means3D), shs=cv(shs), colors_precomp=cv(colors_precomp), opacities=cv(opacities),
                        scales=cv(scales), rotations=cv(rotations), cov3D_precomp=cv(cov3D_precomp),
                        features=cv(features))
        self.s = settings
        self.W, self.H, self.F = int(settings.image_width), int(settings.image_height), int(settings.feature_count)
        i = self.inp
        self.pre = preprocess_forward(
            i["means3D"], i["opacities"], settings.viewmatrix.cpu(), settings.projmatrix.cpu(), settings.campos.cpu(),
            self.W, self.H, settings.tanfovx, settings.tanfovy, sh_degree=settings.sh_degree, shs=i["shs"],
            colors_precomp=i["colors_precomp"], scales=i["scales"], rotations=i["rotations"],
            cov3D_precomp=i["cov3D_precomp"], scale_modifier=settings.scale_modifier)
        self.lists = build_lists(self.pre)
        feats = i["features"] if i["features"] is not None else torch.zeros(i["means3D"].shape[0], 10, dtype=dt)
        self.bg = settings.bg.detach().cpu().to(dt)
        color, buffer, final_T, n_contrib, observe = blend_forward(self.pre, self.lists, feats, self.bg, self.W, self.H,
                                                                   self.F, tiles=tiles)
        self.final_T, self.n_contrib = final_T, n_contrib
        return color, self.pre["radii"], observe, buffer

    def backward(self, grad_color, grad_buffer, tiles=None, final_T=None, n_contrib=None):
        """``final_T`` / ``n_contrib`` default to this object's own forward; tests may inject the reference's saved
        per-pixel state instead (the reference's backward reads exactly these two arrays, backward.cu:460-468) so that
        a borderline termination decision taken differently in host floating point does not pollute the comparison."""
        dt = self.dtype
        i, s = self.inp, self.s
        if final_T is not None:
            self.final_T = final_T.detach().cpu().to(dt)
        if n_contrib is not None:
            self.n_contrib = n_contrib.detach().cpu().to(torch.int32)
        feats = i["features"] if i["features"] is not None else torch.zeros(i["means3D"].shape[0], 10, dtype=dt)
        gc, gb = grad_color.detach().cpu().to(dt), grad_buffer.detach().cpu().to(dt)
        dmean2D, dconic, dopac, dcol, dfeat = blend_backward(self.pre, self.lists, feats, self.bg, self.W, self.H,
                                                             self.F, self.final_T, self.n_contrib, gc, gb, tiles=tiles)
        # per-Gaussian stage through autograd with the reference's backward-only dependency structure
        leaves = {k: (v.clone().requires_grad_(True) if v is not None else None)
                  for k, v in i.items() if k in ("means3D", "shs", "scales", "rotations", "cov3D_precomp")}
        pre = preprocess_forward(
            leaves["means3D"], i["opacities"], s.viewmatrix.cpu(), s.projmatrix.cpu(), s.campos.cpu(), self.W, self.H,
            s.tanfovx, s.tanfovy, sh_degree=s.sh_degree, shs=leaves["shs"], colors_precomp=i["colors_precomp"],
            scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=leaves["cov3D_precomp"],
            scale_modifier=s.scale_modifier, backward_quirks=True)
        vis = self.pre["visible"].to(dt)[:, None]
        # dL_dmeans2D.xy is already scaled by (W/2, H/2): it is the gradient w.r.t. NDC (backward.cu:490-491,378-392)
        ndc_like = torch.stack([pre["means2D"][:, 0] * (2.0 / self.W), pre["means2D"][:, 1] * (2.0 / self.H)], dim=1)
        # K5 stores HALF of dL/d(conic.y) (backward.cu:591: the off-diagonal appears twice in the symmetric matrix) and
        # K6 doubles it again (:216-218), so the true gradient w.r.t. the packed (xx, xy, yy) conic is (g0, 2*g1, g2).
        dconic_true = dconic * torch.tensor([1.0, 2.0, 1.0], dtype=dt)
        loss = (ndc_like * (dmean2D[:, :2] * vis)).sum() + (pre["conic"] * (dconic_true * vis)).sum()
        if i["colors_precomp"] is None:
            loss = loss + (pre["rgb"] * (dcol * vis)).sum()
        wanted = [(k, v) for k, v in leaves.items() if v is not None]
        if leaves["cov3D_precomp"] is None:
            wanted.append(("cov3D_precomp", pre["cov3D"]))   # the reference also returns dL/dcov3D when it is derived
        grads = torch.autograd.grad(loss, [v for _, v in wanted], allow_unused=True)
        out = {k: (g if g is not None else torch.zeros_like(v)) for (k, v), g in zip(wanted, grads)}
        P = i["means3D"].shape[0]
        z = lambda *shape: torch.zeros(*shape, dtype=dt)  # noqa: E731
        return dict(dL_dmeans2D=dmean2D, dL_dconic=dconic, dL_dopacity=dopac[:, None], dL_dcolor=dcol,
                    dL_dfeatures=dfeat, dL_dmeans3D=out.get("means3D", z(P, 3)),
                    dL_dsh=out.get("shs", z(P, 0, 3)), dL_dscale=out.get("scales", z(P, 3)),
                    dL_drot=out.get("rotations", z(P, 4)), dL_dcov3D=out.get("cov3D_precomp", z(P, 6)))
