// Forward per-Gaussian stage: near-plane cull, projection, 3-D covariance, EWA 2-D covariance, conic, screen radius,
// tile rectangle, SH -> RGB.  One thread per Gaussian.
//
// Behavioural reference: preprocessCUDA (cuda_rasterizer/forward.cu:145-241) with computeCov3D (:109-142),
// computeCov2D (:70-104, NO +0.3 dilation), computeColorFromSH (:20-67), in_frustum (auxiliary.h:140-162),
// ndc2Pix in double (auxiliary.h:40-42) and getRect (auxiliary.h:44-53).
//
// Bit-exactness: sort keys, tile ranges and per-pixel contributor counts downstream are integer functions of the
// floats produced here, so every float on the geometry path is computed with explicit round-to-nearest intrinsics
// in the association order the reference has *as compiled by nvcc 12.9 for sm_100 with FMA contraction* (read from
// its SASS): 3-term dot products are fma(z, fma(x, mul(y))) — the middle product is the plain multiply — and
// correctly-rounded div / rcp / sqrt are implementation independent.  Nothing here may be compiled with fast-math.
#include "common.cuh"

namespace gs2m {

namespace {

// a.x*b.x + a.y*b.y + a.z*b.z in the reference's contraction order: fma(z, fma(x, mul(y)))
__device__ __forceinline__ float dot3m(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by)));
}

struct Cov3 { float c0, c1, c2, c3, c4, c5; };

// Sigma = R diag(s)^2 R^T with the quaternion used as given (no normalisation; forward.cu:117).
__device__ __forceinline__ Cov3 covariance_from_scale_rotation(float sx, float sy, float sz, float mod,
                                                               float r, float x, float y, float z) {
    // doubled rotation-matrix entries, each a single fma over one plain product
    const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float xz_p_ry = __fmaf_rn(r, y, xz);
    const float xz_m_ry = __fmaf_rn(-r, y, xz);
    const float yz_m_rx = __fmaf_rn(y, z, -rx);
    const float yz_p_rx = __fmaf_rn(y, z, rx);
    const float xy_m_rz = __fmaf_rn(x, y, -rz);
    const float xy_p_rz = __fmaf_rn(x, y, rz);
    const float xx_p_yy = __fmaf_rn(x, x, yy);
    const float yy_p_zz = __fadd_rn(yy, zz);
    const float xx_p_zz = __fmaf_rn(x, x, zz);
    const float r00 = __fadd_rn(-__fadd_rn(yy_p_zz, yy_p_zz), 1.0f);
    const float r11 = __fadd_rn(-__fadd_rn(xx_p_zz, xx_p_zz), 1.0f);
    const float r22 = __fadd_rn(-__fadd_rn(xx_p_yy, xx_p_yy), 1.0f);
    const float s0 = __fmul_rn(sx, mod), s1 = __fmul_rn(sy, mod), s2 = __fmul_rn(sz, mod);
    // columns of R scaled by the matching axis length
    const float a0 = __fmul_rn(s0, r00);
    const float a1 = __fmul_rn(s0, __fadd_rn(xy_p_rz, xy_p_rz));
    const float a2 = __fmul_rn(s0, __fadd_rn(xz_m_ry, xz_m_ry));
    const float b0 = __fmul_rn(s1, __fadd_rn(xy_m_rz, xy_m_rz));
    const float b1 = __fmul_rn(s1, r11);
    const float b2 = __fmul_rn(s1, __fadd_rn(yz_p_rx, yz_p_rx));
    const float c0 = __fmul_rn(s2, __fadd_rn(xz_p_ry, xz_p_ry));
    const float c1 = __fmul_rn(s2, __fadd_rn(yz_m_rx, yz_m_rx));
    const float c2 = __fmul_rn(s2, r22);
    Cov3 o;
    o.c0 = __fmaf_rn(c0, c0, __fmaf_rn(a0, a0, __fmul_rn(b0, b0)));
    o.c1 = __fmaf_rn(c0, c1, __fmaf_rn(a0, a1, __fmul_rn(b0, b1)));
    o.c2 = __fmaf_rn(c0, c2, __fmaf_rn(a0, a2, __fmul_rn(b0, b2)));
    o.c3 = __fmaf_rn(c1, c1, __fmaf_rn(a1, a1, __fmul_rn(b1, b1)));
    o.c4 = __fmaf_rn(c1, c2, __fmaf_rn(a1, a2, __fmul_rn(b1, b2)));
    o.c5 = __fmaf_rn(c2, c2, __fmaf_rn(a2, a2, __fmul_rn(b2, b2)));
    return o;
}

// view-dependent colour from degree-D real SH (basis/sign convention of auxiliary.h:21-38, forward.cu:33-58)
// `sh` points at this Gaussian's [M][3] block.  Returns res (before +0.5); accumulation order as the reference.
template <typename LoadSH>
__device__ __forceinline__ void sh_to_rgb(int D, float x, float y, float z, LoadSH sh, float res[3]) {
    const float C0 = 0.28209479177387814f;
#pragma unroll
    for (int c = 0; c < 3; ++c) res[c] = __fmul_rn(sh(0, c), C0);
    if (D < 1) return;
    {
        const float C1 = 0.4886025119029199f;
        const float ty = __fmul_rn(y, C1), tz = __fmul_rn(z, C1), tx = __fmul_rn(x, C1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fmaf_rn(-ty, sh(1, c), res[c]);
            v = __fmaf_rn(tz, sh(2, c), v);
            res[c] = __fmaf_rn(-tx, sh(3, c), v);
        }
    }
    if (D < 2) return;
    const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float xy = __fmul_rn(y, x), zz2 = __fadd_rn(zz, zz);
    const float xx_m_yy = __fadd_rn(xx, -yy);
    {
        const float k4 = __fmul_rn(xy, 1.0925484305920792f);
        const float k5 = __fmul_rn(__fmul_rn(z, y), -1.0925484305920792f);
        const float k6 = __fmul_rn(__fadd_rn(-yy, __fadd_rn(-xx, zz2)), 0.31539156525252005f);
        const float k7 = __fmul_rn(__fmul_rn(z, x), -1.0925484305920792f);
        const float k8 = __fmul_rn(xx_m_yy, 0.5462742152960396f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fmaf_rn(k4, sh(4, c), res[c]);
            v = __fmaf_rn(k5, sh(5, c), v);
            v = __fmaf_rn(k6, sh(6, c), v);
            v = __fmaf_rn(k7, sh(7, c), v);
            res[c] = __fmaf_rn(k8, sh(8, c), v);
        }
    }
    if (D < 3) return;
    {
        const float q = __fadd_rn(-yy, __fmaf_rn(zz, 4.0f, -xx));                 // 4zz - xx - yy
        const float k9 = __fmul_rn(__fmul_rn(y, -0.5900435899266435f), __fmaf_rn(xx, 3.0f, -yy));
        const float k10 = __fmul_rn(__fmul_rn(xy, 2.890611442640554f), z);
        const float k11 = __fmul_rn(__fmul_rn(y, -0.4570457994644658f), q);
        const float k12 = __fmul_rn(__fmul_rn(z, 0.3731763325901154f), __fmaf_rn(yy, -3.0f, __fmaf_rn(xx, -3.0f, zz2)));
        const float k13 = __fmul_rn(q, __fmul_rn(x, -0.4570457994644658f));
        const float k14 = __fmul_rn(xx_m_yy, __fmul_rn(z, 1.445305721320277f));
        const float k15 = __fmul_rn(__fmul_rn(x, -0.5900435899266435f), __fmaf_rn(yy, -3.0f, xx));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fmaf_rn(k9, sh(9, c), res[c]);
            v = __fmaf_rn(k10, sh(10, c), v);
            v = __fmaf_rn(k11, sh(11, c), v);
            v = __fmaf_rn(k12, sh(12, c), v);
            v = __fmaf_rn(k13, sh(13, c), v);
            v = __fmaf_rn(k14, sh(14, c), v);
            res[c] = __fmaf_rn(k15, sh(15, c), v);
        }
    }
}

// ZERO_ACC: also zero the backward accumulator row of every visible Gaussian (the backward blend adds into it with vector
// reductions; rows of culled Gaussians are never read) — replaces a memset of the whole [P,24] array in the backward.
// `bin_flags`: when `prefiltered` is set the caller has promised that no Gaussian fails the near-plane test; one that does
// raises GS2M_BIN_PREFILTERED (the reference printf()s and traps the device there, auxiliary.h:154-160).
template <bool ZERO_ACC>
__global__ void __launch_bounds__(256) preprocess_forward_kernel(FwdParams p, GeomState g, int* __restrict__ radii,
                                                                 int* __restrict__ observe, uint32_t* __restrict__ bin_flags,
                                                                 bool prefiltered) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;

    // every load that does not depend on the culling decision is issued up front, next to the position's
    const float px = p.means3D[3 * idx + 0], py = p.means3D[3 * idx + 1], pz = p.means3D[3 * idx + 2];
    const float opacity = p.opacities[idx];
    float sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.cov3D_precomp == nullptr) {
        const float* sc = p.scales + 3 * (size_t)idx;
        sc0 = sc[0]; sc1 = sc[1]; sc2 = sc[2];
        q = *reinterpret_cast<const float4*>(p.rotations + 4 * (size_t)idx);
    }

    // outputs every Gaussian gets, visible or not
    radii[idx] = 0;
    observe[idx] = 0;
    g.tiles_touched[idx] = 0;

    const float* __restrict__ vm = p.viewmatrix;
    const float* __restrict__ pm = p.projmatrix;

    // near-plane cull on view-space depth (auxiliary.h:150-152; the x/y frustum test is disabled in the reference)
    const float depth = __fadd_rn(dot3m(px, py, pz, vm[2], vm[6], vm[10]), vm[14]);
    if (depth <= 0.2f) {
        if (prefiltered) atomicOr(bin_flags, (uint32_t)GS2M_BIN_PREFILTERED);
        return;
    }

    // clip-space position and perspective divide
    const float hx = __fadd_rn(dot3m(px, py, pz, pm[0], pm[4], pm[8]), pm[12]);
    const float hy = __fadd_rn(dot3m(px, py, pz, pm[1], pm[5], pm[9]), pm[13]);
    const float hw = __fadd_rn(dot3m(px, py, pz, pm[3], pm[7], pm[11]), pm[15]);
    const float inv_w = __frcp_rn(__fadd_rn(hw, 0.0000001f));
    const float ndc_x = __fmul_rn(hx, inv_w), ndc_y = __fmul_rn(hy, inv_w);

    Cov3 S;
    if (p.cov3D_precomp != nullptr) {
        const float* c = p.cov3D_precomp + 6 * (size_t)idx;
        S.c0 = c[0]; S.c1 = c[1]; S.c2 = c[2]; S.c3 = c[3]; S.c4 = c[4]; S.c5 = c[5];
    } else {
        S = covariance_from_scale_rotation(sc0, sc1, sc2, p.scale_modifier, q.x, q.y, q.z, q.w);
        float* o = g.cov3D + 6 * (size_t)idx;
        o[0] = S.c0; o[1] = S.c1; o[2] = S.c2; o[3] = S.c3; o[4] = S.c4; o[5] = S.c5;
    }

    // EWA projection of the covariance (Zwicker et al. 2002, eq. 29/31)
    const float tz = depth;  // same dot product
    const float tx = __fadd_rn(dot3m(px, py, pz, vm[0], vm[4], vm[8]), vm[12]);
    const float ty = __fadd_rn(dot3m(px, py, pz, vm[1], vm[5], vm[9]), vm[13]);
    const float limx = __fmul_rn(p.tan_fovx, 1.3f), limy = __fmul_rn(p.tan_fovy, 1.3f);
    const float cx = fminf(fmaxf(__fdiv_rn(tx, tz), -limx), limx);
    const float cy = fminf(fmaxf(__fdiv_rn(ty, tz), -limy), limy);
    const float tz2 = __fmul_rn(tz, tz);
    const float j00 = __fdiv_rn(p.focal_x, tz);
    const float j02 = __fdiv_rn(__fmul_rn(__fmul_rn(tz, -cx), p.focal_x), tz2);
    const float j11 = __fdiv_rn(p.focal_y, tz);
    const float j12 = __fdiv_rn(__fmul_rn(__fmul_rn(tz, -cy), p.focal_y), tz2);
    // rows of J * R_w2v
    const float t0x = __fmaf_rn(vm[2], j02, __fmul_rn(vm[0], j00));
    const float t0y = __fmaf_rn(vm[6], j02, __fmul_rn(vm[4], j00));
    const float t0z = __fmaf_rn(vm[10], j02, __fmul_rn(vm[8], j00));
    const float t1x = __fmaf_rn(vm[2], j12, __fmul_rn(vm[1], j11));
    const float t1y = __fmaf_rn(vm[6], j12, __fmul_rn(vm[5], j11));
    const float t1z = __fmaf_rn(vm[10], j12, __fmul_rn(vm[9], j11));
    // Sigma * rows
    const float v00 = dot3m(t0x, t0y, t0z, S.c0, S.c1, S.c2);
    const float v01 = dot3m(t0x, t0y, t0z, S.c1, S.c3, S.c4);
    const float v02 = dot3m(t0x, t0y, t0z, S.c2, S.c4, S.c5);
    const float v10 = dot3m(t1x, t1y, t1z, S.c0, S.c1, S.c2);
    const float v11 = dot3m(t1x, t1y, t1z, S.c1, S.c3, S.c4);
    const float v12 = dot3m(t1x, t1y, t1z, S.c2, S.c4, S.c5);
    const float cov_a = dot3m(t0x, t0y, t0z, v00, v01, v02);
    const float cov_b = dot3m(t0x, t0y, t0z, v10, v11, v12);
    const float cov_c = dot3m(t1x, t1y, t1z, v10, v11, v12);

    const float det = __fmaf_rn(cov_a, cov_c, -__fmul_rn(cov_b, cov_b));
    if (det == 0.0f) return;
    const float det_inv = __frcp_rn(det);
    const float conic_x = __fmul_rn(cov_c, det_inv);
    const float conic_y = __fmul_rn(cov_b, -det_inv);
    const float conic_z = __fmul_rn(cov_a, det_inv);

    // screen-space extent from the larger eigenvalue
    const float mid = __fmul_rn(__fadd_rn(cov_a, cov_c), 0.5f);
    const float root = __fsqrt_rn(fmaxf(__fmaf_rn(mid, mid, -det), 0.1f));
    const float lam = fmaxf(__fadd_rn(mid, root), __fadd_rn(mid, -root));
    const int radius = (int)ceilf(__fmul_rn(__fsqrt_rn(lam), 3.0f));

    // pixel-centre coordinates, evaluated in double like ndc2Pix
    const float pix_x = __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)ndc_x, 1.0), (double)p.W, -1.0), 0.5));
    const float pix_y = __double2float_rn(__dmul_rn(__fma_rn(__dadd_rn((double)ndc_y, 1.0), (double)p.H, -1.0), 0.5));

    int x0, y0, x1, y1;
    tile_rect(pix_x, pix_y, radius, p.tiles_x, p.tiles_y, x0, y0, x1, y1);
    const unsigned n_tiles = (unsigned)(x1 - x0) * (unsigned)(y1 - y0);
    if (n_tiles == 0) return;

    float4 rgb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.colors_precomp == nullptr) {
        const float dx = __fadd_rn(px, -p.cam_pos[0]);
        const float dy = __fadd_rn(py, -p.cam_pos[1]);
        const float dz = __fadd_rn(pz, -p.cam_pos[2]);
        const float len = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
        const float ux = __fdiv_rn(dx, len), uy = __fdiv_rn(dy, len), uz = __fdiv_rn(dz, len);
        const float* __restrict__ shp = p.shs + (size_t)idx * p.M * 3;
        float res[3];
        if (((3 * p.M) & 3) == 0) {
            // rows are 16-byte aligned: fetch the (D+1)^2 x 3 coefficients with 128-bit loads, all in flight at once
            float c[48];
            const int n4 = (3 * (p.D + 1) * (p.D + 1) + 3) >> 2;
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < n4) t = __ldg(reinterpret_cast<const float4*>(shp) + i);
                c[4 * i] = t.x; c[4 * i + 1] = t.y; c[4 * i + 2] = t.z; c[4 * i + 3] = t.w;
            }
            sh_to_rgb(p.D, ux, uy, uz, [&](int k, int ch) { return c[3 * k + ch]; }, res);
        } else {
            sh_to_rgb(p.D, ux, uy, uz, [&](int k, int ch) { return __ldg(shp + 3 * k + ch); }, res);
        }
        uchar4 cl;
        cl.x = !(res[0] >= -0.5f); cl.y = !(res[1] >= -0.5f); cl.z = !(res[2] >= -0.5f); cl.w = 0;
        rgb.x = cl.x ? 0.f : __fadd_rn(res[0], 0.5f);
        rgb.y = cl.y ? 0.f : __fadd_rn(res[1], 0.5f);
        rgb.z = cl.z ? 0.f : __fadd_rn(res[2], 0.5f);
        reinterpret_cast<uchar4*>(g.clamped)[idx] = cl;
    } else {
        const float* c = p.colors_precomp + 3 * (size_t)idx;
        rgb.x = c[0]; rgb.y = c[1]; rgb.z = c[2];
    }

    // Footprint threshold for the blend kernels' conservative culling: a pixel can only receive
    // alpha >= 1/255 where  conic-quadratic q(d) <= 2*ln(255*opacity).  Negative => can never contribute.
    const float thr = (opacity >= 0.00392156885936856f) ? 2.0f * logf(255.0f * opacity) : -1.0f;

    g.depths[idx] = depth;
    radii[idx] = radius;
    g.xy_conic_ab[idx] = make_float4(pix_x, pix_y, conic_x, conic_y);
    g.conic_c_opac[idx] = make_float4(conic_z, opacity, thr, 0.0f);
    g.rgb[idx] = rgb;
    g.tiles_touched[idx] = n_tiles;
    if (ZERO_ACC) {
        float4* acc = reinterpret_cast<float4*>(g.grad_acc + (size_t)idx * GS2M_ACC_STRIDE);
#pragma unroll
        for (int i = 0; i < GS2M_ACC_STRIDE / 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__global__ void __launch_bounds__(256) mark_visible_kernel(int P, const float* __restrict__ means3D,
                                                           const float* __restrict__ vm, uint8_t* __restrict__ present) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float px = means3D[3 * idx + 0], py = means3D[3 * idx + 1], pz = means3D[3 * idx + 2];
    const float depth = __fadd_rn(dot3m(px, py, pz, vm[2], vm[6], vm[10]), vm[14]);
    present[idx] = !(depth <= 0.2f);
}

}  // namespace

int launch_preprocess_forward(const FwdParams& p, const GeomState& g, int* radii, int* observe, uint32_t* bin_flags,
                              bool prefiltered, bool zero_grad_acc, cudaStream_t s) {
    if (p.P == 0) return GS2M_OK;
    count_launches(1);
    if (zero_grad_acc) preprocess_forward_kernel<true><<<(p.P + 255) / 256, 256, 0, s>>>(p, g, radii, observe, bin_flags, prefiltered);
    else preprocess_forward_kernel<false><<<(p.P + 255) / 256, 256, 0, s>>>(p, g, radii, observe, bin_flags, prefiltered);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s) {
    if (P == 0) return GS2M_OK;
    count_launches(1);
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, viewmatrix, present);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace gs2m
