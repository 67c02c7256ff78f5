// Backward of the per-Gaussian stage, fused into one kernel: conic -> 2-D covariance -> 3-D covariance and mean,
// projection of the 2-D mean gradient, SH colour backward, covariance -> scale / rotation.
//
// Behavioural reference: computeCov2DCUDA (cuda_rasterizer/backward.cu:153-281, which recomputes the 2-D
// covariance WITH a +0.3 dilation the forward does not apply — reproduced on purpose), preprocessCUDA backward
// (:352-410), computeColorFromSH backward (:23-148), computeCov3D backward (:285-347, no quaternion-normalisation
// Jacobian), dnormvdv (auxiliary.h:111-120).
//
// Input is the packed per-Gaussian accumulator written by the backward blend ([P][24] floats):
//   0,1 dL/dmean2D (already scaled by W/2,H/2)   2,3 sum of |.|   4,5,6 dL/dconic (xx, xy, yy)   7 dL/dopacity
//   8..10 dL/drgb   11..20 dL/dfeatures
// Every element of every output tensor is written (zeros for culled Gaussians), so the caller never has to
// zero-fill 344 B/Gaussian the way rasterize_points.cu:150-159 does.  With `accumulate` the nine user-visible
// tensors are updated with += instead (several views summed before one all-reduce).
#include "common.cuh"

namespace gs2m {
namespace {

template <bool ACC>
__device__ __forceinline__ void put(float* p, float v) {
    if (ACC) *p += v; else *p = v;
}

__device__ __forceinline__ float3 operator*(float s, float3 v) { return make_float3(s * v.x, s * v.y, s * v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

template <bool ACC>
__global__ void __launch_bounds__(256) preprocess_backward_kernel(BwdParams p, GeomState g) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.P) return;
    const size_t i = (size_t)idx;
    const bool visible = p.radii[idx] > 0;
    const int M = p.M;

    float acc[GS2M_ACC_STRIDE];
    if (visible) {
        const float4* a4 = reinterpret_cast<const float4*>(g.grad_acc + i * GS2M_ACC_STRIDE);
#pragma unroll
        for (int k = 0; k < GS2M_ACC_STRIDE / 4; ++k) {
            const float4 t = a4[k];
            acc[4 * k] = t.x; acc[4 * k + 1] = t.y; acc[4 * k + 2] = t.z; acc[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < GS2M_ACC_STRIDE; ++k) acc[k] = 0.f;
    }

    // ---- pass-through gradients ----
    {
        float4* o = reinterpret_cast<float4*>(p.dL_dmeans2D) + i;
        if (ACC) { float4 t = *o; t.x += acc[0]; t.y += acc[1]; t.z += acc[2]; t.w += acc[3]; *o = t; }
        else *o = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (p.dL_dconic) reinterpret_cast<float4*>(p.dL_dconic)[i] = make_float4(acc[4], acc[5], 0.f, acc[6]);
        put<ACC>(p.dL_dopacity + i, acc[7]);
        put<ACC>(p.dL_dcolor + 3 * i + 0, acc[8]);
        put<ACC>(p.dL_dcolor + 3 * i + 1, acc[9]);
        put<ACC>(p.dL_dcolor + 3 * i + 2, acc[10]);
#pragma unroll
        for (int k = 0; k < GS2M_NUM_FEATURES; ++k) put<ACC>(p.dL_dfeatures + GS2M_NUM_FEATURES * i + k, (k < p.F) ? acc[11 + k] : 0.f);
    }

    if (!visible) {
        if (!ACC) {
            for (int k = 0; k < 3; ++k) p.dL_dmeans3D[3 * i + k] = 0.f;
            for (int k = 0; k < 6; ++k) p.dL_dcov3D[6 * i + k] = 0.f;
            for (int k = 0; k < 3; ++k) p.dL_dscale[3 * i + k] = 0.f;
            for (int k = 0; k < 4; ++k) p.dL_drot[4 * i + k] = 0.f;
            if (p.dL_dsh) for (int k = 0; k < 3 * M; ++k) p.dL_dsh[3 * M * i + k] = 0.f;
        }
        return;
    }

    const float* __restrict__ vm = p.viewmatrix;
    const float* __restrict__ pm = p.projmatrix;
    const float3 m = make_float3(p.means3D[3 * i], p.means3D[3 * i + 1], p.means3D[3 * i + 2]);

    // =================== conic -> cov2D -> cov3D, mean (via the Jacobian) ===================
    const float* cov3D = (p.cov3D_precomp ? p.cov3D_precomp : g.cov3D) + 6 * i;
    const float c0 = cov3D[0], c1 = cov3D[1], c2 = cov3D[2], c3 = cov3D[3], c4 = cov3D[4], c5 = cov3D[5];
    float3 t = make_float3(vm[0] * m.x + vm[4] * m.y + vm[8] * m.z + vm[12],
                           vm[1] * m.x + vm[5] * m.y + vm[9] * m.z + vm[13],
                           vm[2] * m.x + vm[6] * m.y + vm[10] * m.z + vm[14]);
    const float limx = 1.3f * p.tan_fovx, limy = 1.3f * p.tan_fovy;
    const float txtz = t.x / t.z, tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    const float x_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float fx = p.focal_x, fy = p.focal_y;
    const float j00 = fx / t.z, j02 = -(fx * t.x) / (t.z * t.z);
    const float j11 = fy / t.z, j12 = -(fy * t.y) / (t.z * t.z);
    // rows of R_w2v
    const float3 w0 = make_float3(vm[0], vm[4], vm[8]);
    const float3 w1 = make_float3(vm[1], vm[5], vm[9]);
    const float3 w2 = make_float3(vm[2], vm[6], vm[10]);
    const float3 T0 = j00 * w0 + j02 * w2;   // rows of J * R_w2v
    const float3 T1 = j11 * w1 + j12 * w2;
    const float3 V0 = make_float3(c0 * T0.x + c1 * T0.y + c2 * T0.z, c1 * T0.x + c3 * T0.y + c4 * T0.z,
                                  c2 * T0.x + c4 * T0.y + c5 * T0.z);   // Sigma * T0
    const float3 V1 = make_float3(c0 * T1.x + c1 * T1.y + c2 * T1.z, c1 * T1.x + c3 * T1.y + c4 * T1.z,
                                  c2 * T1.x + c4 * T1.y + c5 * T1.z);   // Sigma * T1
    const float a = dot(T0, V0) + 0.3f;      // the backward-only dilation (backward.cu:205-207)
    const float b = dot(T0, V1);
    const float c = dot(T1, V1) + 0.3f;
    const float gx = acc[4], gy = acc[5], gz = acc[6];   // dL/dconic (xx, xy, yy)
    const float denom = a * c - b * b;
    const float denom2inv = 1.0f / (denom * denom + 0.0000001f);
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (denom2inv != 0.f) {
        dL_da = denom2inv * (-c * c * gx + 2.f * b * c * gy + (denom - a * c) * gz);
        dL_dc = denom2inv * (-a * a * gz + 2.f * a * b * gy + (denom - a * c) * gx);
        dL_db = denom2inv * 2.f * (b * c * gx - (denom + 2.f * b * b) * gy + a * b * gz);
        dcov[0] = T0.x * T0.x * dL_da + T0.x * T1.x * dL_db + T1.x * T1.x * dL_dc;
        dcov[3] = T0.y * T0.y * dL_da + T0.y * T1.y * dL_db + T1.y * T1.y * dL_dc;
        dcov[5] = T0.z * T0.z * dL_da + T0.z * T1.z * dL_db + T1.z * T1.z * dL_dc;
        dcov[1] = 2.f * T0.x * T0.y * dL_da + (T0.x * T1.y + T0.y * T1.x) * dL_db + 2.f * T1.x * T1.y * dL_dc;
        dcov[2] = 2.f * T0.x * T0.z * dL_da + (T0.x * T1.z + T0.z * T1.x) * dL_db + 2.f * T1.x * T1.z * dL_dc;
        dcov[4] = 2.f * T0.z * T0.y * dL_da + (T0.y * T1.z + T0.z * T1.y) * dL_db + 2.f * T1.y * T1.z * dL_dc;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) put<ACC>(p.dL_dcov3D + 6 * i + k, dcov[k]);

    // gradient w.r.t. the rows of J*R, then J, then the view-space mean
    const float3 dT0 = (2.f * dL_da) * V0 + dL_db * V1;
    const float3 dT1 = (2.f * dL_dc) * V1 + dL_db * V0;
    const float dJ00 = dot(w0, dT0), dJ02 = dot(w2, dT0);
    const float dJ11 = dot(w1, dT1), dJ12 = dot(w2, dT1);
    const float tz = 1.f / t.z, tz2 = tz * tz, tz3 = tz2 * tz;
    const float dtx = x_mul * -fx * tz2 * dJ02;
    const float dty = y_mul * -fy * tz2 * dJ12;
    const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.f * fx * t.x) * tz3 * dJ02 + (2.f * fy * t.y) * tz3 * dJ12;
    float3 dmean = make_float3(vm[0] * dtx + vm[1] * dty + vm[2] * dtz, vm[4] * dtx + vm[5] * dty + vm[6] * dtz,
                               vm[8] * dtx + vm[9] * dty + vm[10] * dtz);

    // =================== 2-D mean -> 3-D mean through the projection ===================
    {
        const float hw = pm[3] * m.x + pm[7] * m.y + pm[11] * m.z + pm[15];
        const float mw = 1.0f / (hw + 0.0000001f);
        const float mul1 = (pm[0] * m.x + pm[4] * m.y + pm[8] * m.z + pm[12]) * mw * mw;
        const float mul2 = (pm[1] * m.x + pm[5] * m.y + pm[9] * m.z + pm[13]) * mw * mw;
        const float g2x = acc[0], g2y = acc[1];
        dmean.x += (pm[0] * mw - pm[3] * mul1) * g2x + (pm[1] * mw - pm[3] * mul2) * g2y;
        dmean.y += (pm[4] * mw - pm[7] * mul1) * g2x + (pm[5] * mw - pm[7] * mul2) * g2y;
        dmean.z += (pm[8] * mw - pm[11] * mul1) * g2x + (pm[9] * mw - pm[11] * mul2) * g2y;
    }

    // =================== colour -> SH coefficients and view direction ===================
    if (p.shs != nullptr) {
        const float3 d0 = make_float3(m.x - p.cam_pos[0], m.y - p.cam_pos[1], m.z - p.cam_pos[2]);
        const float inv_len = 1.0f / sqrtf(dot(d0, d0));
        const float x = d0.x * inv_len, y = d0.y * inv_len, z = d0.z * inv_len;
        const uchar4 cl = reinterpret_cast<const uchar4*>(g.clamped)[idx];
        const float3 dRGB = make_float3(cl.x ? 0.f : acc[8], cl.y ? 0.f : acc[9], cl.z ? 0.f : acc[10]);
        const float* __restrict__ sh = p.shs + 3 * (size_t)M * i;
        float* __restrict__ dsh = p.dL_dsh + 3 * (size_t)M * i;

        // basis value B[k] and its gradient (Bx,By,Bz) w.r.t. the (unnormalised-treated) direction, per coefficient
        const int D = p.D;
        const int n_active = (D + 1) * (D + 1);
        float B[16], Bx[16], By[16], Bz[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) { B[k] = 0.f; Bx[k] = 0.f; By[k] = 0.f; Bz[k] = 0.f; }
        const float C0 = 0.28209479177387814f, C1 = 0.4886025119029199f;
        const float C20 = 1.0925484305920792f, C21 = -1.0925484305920792f, C22 = 0.31539156525252005f,
                    C23 = -1.0925484305920792f, C24 = 0.5462742152960396f;
        const float C30 = -0.5900435899266435f, C31 = 2.890611442640554f, C32 = -0.4570457994644658f,
                    C33 = 0.3731763325901154f, C34 = -0.4570457994644658f, C35 = 1.445305721320277f,
                    C36 = -0.5900435899266435f;
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        B[0] = C0;
        if (D > 0) {
            B[1] = -C1 * y; By[1] = -C1;
            B[2] = C1 * z;  Bz[2] = C1;
            B[3] = -C1 * x; Bx[3] = -C1;
        }
        if (D > 1) {
            B[4] = C20 * xy;                   Bx[4] = C20 * y;        By[4] = C20 * x;
            B[5] = C21 * yz;                   By[5] = C21 * z;        Bz[5] = C21 * y;
            B[6] = C22 * (2.f * zz - xx - yy); Bx[6] = -2.f * C22 * x; By[6] = -2.f * C22 * y; Bz[6] = 4.f * C22 * z;
            B[7] = C23 * xz;                   Bx[7] = C23 * z;        Bz[7] = C23 * x;
            B[8] = C24 * (xx - yy);            Bx[8] = 2.f * C24 * x;  By[8] = -2.f * C24 * y;
        }
        if (D > 2) {
            B[9] = C30 * y * (3.f * xx - yy);
            Bx[9] = C30 * 6.f * xy;            By[9] = C30 * 3.f * (xx - yy);
            B[10] = C31 * xy * z;
            Bx[10] = C31 * yz;                 By[10] = C31 * xz;                  Bz[10] = C31 * xy;
            B[11] = C32 * y * (4.f * zz - xx - yy);
            Bx[11] = C32 * -2.f * xy;          By[11] = C32 * (4.f * zz - xx - 3.f * yy); Bz[11] = C32 * 8.f * yz;
            B[12] = C33 * z * (2.f * zz - 3.f * xx - 3.f * yy);
            Bx[12] = C33 * -6.f * xz;          By[12] = C33 * -6.f * yz;           Bz[12] = C33 * 3.f * (2.f * zz - xx - yy);
            B[13] = C34 * x * (4.f * zz - xx - yy);
            Bx[13] = C34 * (4.f * zz - 3.f * xx - yy); By[13] = C34 * -2.f * xy;   Bz[13] = C34 * 8.f * xz;
            B[14] = C35 * z * (xx - yy);
            Bx[14] = C35 * 2.f * xz;           By[14] = C35 * -2.f * yz;           Bz[14] = C35 * (xx - yy);
            B[15] = C36 * x * (xx - 3.f * yy);
            Bx[15] = C36 * 3.f * (xx - yy);    By[15] = C36 * -6.f * xy;
        }
        float3 ddir = make_float3(0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < M) {
                const bool active = k < n_active;
                const float bk = active ? B[k] : 0.f;
                put<ACC>(dsh + 3 * k + 0, bk * dRGB.x);
                put<ACC>(dsh + 3 * k + 1, bk * dRGB.y);
                put<ACC>(dsh + 3 * k + 2, bk * dRGB.z);
                if (active && k > 0) {
                    const float s = sh[3 * k] * dRGB.x + sh[3 * k + 1] * dRGB.y + sh[3 * k + 2] * dRGB.z;
                    ddir.x = fmaf(Bx[k], s, ddir.x);
                    ddir.y = fmaf(By[k], s, ddir.y);
                    ddir.z = fmaf(Bz[k], s, ddir.z);
                }
            }
        }
        if (!ACC) for (int k = 16; k < M; ++k) { dsh[3 * k] = 0.f; dsh[3 * k + 1] = 0.f; dsh[3 * k + 2] = 0.f; }
        // through the normalisation dir = d0/|d0|
        const float s2 = dot(d0, d0);
        const float inv32 = 1.0f / sqrtf(s2 * s2 * s2);
        dmean.x += ((s2 - d0.x * d0.x) * ddir.x - d0.y * d0.x * ddir.y - d0.z * d0.x * ddir.z) * inv32;
        dmean.y += (-d0.x * d0.y * ddir.x + (s2 - d0.y * d0.y) * ddir.y - d0.z * d0.y * ddir.z) * inv32;
        dmean.z += (-d0.x * d0.z * ddir.x - d0.y * d0.z * ddir.y + (s2 - d0.z * d0.z) * ddir.z) * inv32;
    }
    put<ACC>(p.dL_dmeans3D + 3 * i + 0, dmean.x);
    put<ACC>(p.dL_dmeans3D + 3 * i + 1, dmean.y);
    put<ACC>(p.dL_dmeans3D + 3 * i + 2, dmean.z);

    // =================== cov3D -> scale, rotation ===================
    if (p.scales != nullptr) {
        const float4 q = *reinterpret_cast<const float4*>(p.rotations + 4 * i);
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float s0 = p.scale_modifier * p.scales[3 * i], s1 = p.scale_modifier * p.scales[3 * i + 1],
                    s2 = p.scale_modifier * p.scales[3 * i + 2];
        // columns of the rotation matrix
        const float3 r0 = make_float3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + r * z), 2.f * (x * z - r * y));
        const float3 r1 = make_float3(2.f * (x * y - r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + r * x));
        const float3 r2 = make_float3(2.f * (x * z + r * y), 2.f * (y * z - r * x), 1.f - 2.f * (x * x + y * y));
        // symmetric dL/dSigma (off-diagonals were doubled above)
        const float g00 = dcov[0], g01 = 0.5f * dcov[1], g02 = 0.5f * dcov[2], g11 = dcov[3], g12 = 0.5f * dcov[4],
                    g22 = dcov[5];
        auto Gmul = [&](float3 v) {
            return make_float3(g00 * v.x + g01 * v.y + g02 * v.z, g01 * v.x + g11 * v.y + g12 * v.z,
                               g02 * v.x + g12 * v.y + g22 * v.z);
        };
        const float3 Gr0 = Gmul(r0), Gr1 = Gmul(r1), Gr2 = Gmul(r2);
        // Sigma = sum_k s_k^2 r_k r_k^T  =>  dL/ds_k = 2 s_k r_k^T G r_k  (w.r.t. the modified scale, as the reference)
        put<ACC>(p.dL_dscale + 3 * i + 0, 2.f * s0 * dot(r0, Gr0));
        put<ACC>(p.dL_dscale + 3 * i + 1, 2.f * s1 * dot(r1, Gr1));
        put<ACC>(p.dL_dscale + 3 * i + 2, 2.f * s2 * dot(r2, Gr2));
        // D_k = dL/dr_k = 2 s_k^2 G r_k ; chain through r_k(q)
        const float3 D0 = (2.f * s0 * s0) * Gr0, D1 = (2.f * s1 * s1) * Gr1, D2 = (2.f * s2 * s2) * Gr2;
        const float dq_r = 2.f * z * (D0.y - D1.x) + 2.f * y * (D2.x - D0.z) + 2.f * x * (D1.z - D2.y);
        const float dq_x = 2.f * y * (D1.x + D0.y) + 2.f * z * (D2.x + D0.z) + 2.f * r * (D1.z - D2.y) - 4.f * x * (D2.z + D1.y);
        const float dq_y = 2.f * x * (D1.x + D0.y) + 2.f * r * (D2.x - D0.z) + 2.f * z * (D1.z + D2.y) - 4.f * y * (D2.z + D0.x);
        const float dq_z = 2.f * r * (D0.y - D1.x) + 2.f * x * (D2.x + D0.z) + 2.f * y * (D1.z + D2.y) - 4.f * z * (D1.y + D0.x);
        put<ACC>(p.dL_drot + 4 * i + 0, dq_r);
        put<ACC>(p.dL_drot + 4 * i + 1, dq_x);
        put<ACC>(p.dL_drot + 4 * i + 2, dq_y);
        put<ACC>(p.dL_drot + 4 * i + 3, dq_z);
    } else if (!ACC) {
        for (int k = 0; k < 3; ++k) p.dL_dscale[3 * i + k] = 0.f;
        for (int k = 0; k < 4; ++k) p.dL_drot[4 * i + k] = 0.f;
    }
    if (p.shs == nullptr && !ACC && p.dL_dsh) {
        for (int k = 0; k < 3 * M; ++k) p.dL_dsh[3 * M * i + k] = 0.f;
    }
}

}  // namespace

int launch_preprocess_backward(const BwdParams& p, const GeomState& g, cudaStream_t s) {
    if (p.P == 0) return GS2M_OK;
    const int blocks = (p.P + 255) / 256;
    count_launches(1);
    if (p.accumulate) preprocess_backward_kernel<true><<<blocks, 256, 0, s>>>(p, g);
    else preprocess_backward_kernel<false><<<blocks, 256, 0, s>>>(p, g);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace gs2m
