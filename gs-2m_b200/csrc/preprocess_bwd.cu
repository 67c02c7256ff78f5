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
#include <algorithm>
#include <atomic>
#include "common.cuh"
#include "pack_math.cuh"

// Gaussians (= threads) per block of the staged kernel.  The API's row ranges start at multiples of 256, which every value
// that divides 256 satisfies.  Measured on B200, config 4: 256 threads 0.445 ms, 128 0.464 ms, 64 0.424 ms (a block walks
// through load / compute / store phases separated by barriers; 19 KB blocks give 11 of them per SM to overlap those phases).
#ifndef GS2M_PB_THREADS
#define GS2M_PB_THREADS 64
#endif

// register cap of the staged kernel (64-thread blocks, 19 KB of shared memory: 11 blocks per SM need <= 93)
#ifndef GS2M_PB_MAXNREG
#define GS2M_PB_MAXNREG 88
#endif

namespace gs2m {
namespace {

constexpr int PBT = GS2M_PB_THREADS;
static_assert(256 % PBT == 0 && PBT % 32 == 0, "block size must divide 256");

// accumulate mode adds with fire-and-forget vector reductions (red.global.add.v4.f32): the read-modify-write happens in
// L2, so the store phase of a block has no load latency to wait for
__device__ __forceinline__ void red_add_f4(float4* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool ACC>
__device__ __forceinline__ void put(float* p, float v) {
    if (ACC) *p += v; else *p = v;
}

__device__ __forceinline__ float3 operator*(float s, float3 v) { return make_float3(s * v.x, s * v.y, s * v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// Per-Gaussian backward math, written once; `io` decides where the results go (straight to global memory, or into
// the block's shared staging rows that are then copied out with coalesced 128-bit stores).
// the Gaussian's packed accumulator row written by the backward blend (zeros for a culled Gaussian)
__device__ __forceinline__ void load_acc_row(const GeomState& g, size_t i, bool visible, float (&acc)[GS2M_ACC_STRIDE]) {
    if (visible) {
        const float4* a4 = reinterpret_cast<const float4*>(g.grad_acc + i * GS2M_ACC_STRIDE);
#pragma unroll
        for (int k = 0; k < GS2M_ACC_STRIDE / 4; ++k) {
            const float4 t = __ldg(a4 + k);
            acc[4 * k] = t.x; acc[4 * k + 1] = t.y; acc[4 * k + 2] = t.z; acc[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < GS2M_ACC_STRIDE; ++k) acc[k] = 0.f;
    }
}

// Everything the per-Gaussian math reads from global memory for one (Gaussian, view), fetched in one burst through the
// non-coherent path before any of it is used: the reductions and stores in the math (asm volatile with a memory clobber) would
// otherwise pin each load right in front of its first use, and a thread would sit through four or five memory latencies per
// Gaussian instead of one (ncu: long_scoreboard 56 % of the stall samples before this change).
struct GaussianInputs {
    float acc[GS2M_ACC_STRIDE];
    float cov[6];
    float3 mean, scale;
    float4 rot;
    uchar4 clamped;
};
__device__ __forceinline__ void load_gaussian_inputs(const BwdParams& p, const GeomState& g, size_t i, bool visible, GaussianInputs& in) {
    load_acc_row(g, i, visible, in.acc);
    if (!visible) return;
    in.mean = make_float3(__ldg(p.means3D + 3 * i), __ldg(p.means3D + 3 * i + 1), __ldg(p.means3D + 3 * i + 2));
    const float* cov3D = (p.cov3D_precomp ? p.cov3D_precomp : g.cov3D) + 6 * i;
#pragma unroll
    for (int k = 0; k < 6; ++k) in.cov[k] = __ldg(cov3D + k);
    in.clamped = (p.shs != nullptr) ? __ldg(reinterpret_cast<const uchar4*>(g.clamped) + i) : make_uchar4(0, 0, 0, 0);
    if (p.scales != nullptr) {
        in.rot = __ldg(reinterpret_cast<const float4*>(p.rotations + 4 * i));
        in.scale = make_float3(__ldg(p.scales + 3 * i), __ldg(p.scales + 3 * i + 1), __ldg(p.scales + 3 * i + 2));
    }
}

template <class IO>
__device__ __forceinline__ void gaussian_backward(const BwdParams& p, const int idx, const bool visible,
                                                  const GaussianInputs& gin, IO& io) {
    const float (&acc)[GS2M_ACC_STRIDE] = gin.acc;
    const size_t i = (size_t)idx;
    const int M = p.M;

    // a culled Gaussian contributes exactly zero: tensors that are accumulated (+=) need nothing for it, tensors that are
    // overwritten get zeros
    if (!visible) {
        if (!IO::kAccOther) {
            io.mean2d(make_float4(0.f, 0.f, 0.f, 0.f));
            io.conic(make_float4(0.f, 0.f, 0.f, 0.f));
            io.opacity(0.f);
            for (int k = 0; k < 3; ++k) io.color(k, 0.f);
            for (int k = 0; k < GS2M_NUM_FEATURES; ++k) io.feature(k, 0.f);
            for (int k = 0; k < 6; ++k) io.cov(k, 0.f);
            for (int k = 0; k < 3; ++k) io.scale(k, 0.f);
            io.rot(make_float4(0.f, 0.f, 0.f, 0.f));
        }
        if (!IO::kAccParams) {
            for (int k = 0; k < 3; ++k) io.mean3d(k, 0.f);
#pragma unroll
            for (int k = 0; k < 48; ++k) { if (k < 3 * M) io.sh_out(k, 0.f); }
            if (IO::kAnyM) { for (int k = 48; k < 3 * M; ++k) io.sh_out(k, 0.f); }
        }
        return;
    }

    // ---- pass-through gradients ----
    io.mean2d(make_float4(acc[0], acc[1], acc[2], acc[3]));
    if (visible) {   // add_densification_stats (scene/gaussian_model.py:569-573), on this view's gradient
        if (p.densify_grad_accum) atomicAdd(p.densify_grad_accum + idx, sqrtf(acc[0] * acc[0] + acc[1] * acc[1]));
        if (p.densify_grad_accum_abs) atomicAdd(p.densify_grad_accum_abs + idx, sqrtf(acc[2] * acc[2] + acc[3] * acc[3]));
        if (p.densify_denom) atomicAdd(p.densify_denom + idx, 1.0f);
    }
    io.conic(make_float4(acc[4], acc[5], 0.f, acc[6]));
    io.opacity(acc[7]);
    io.color(0, acc[8]); io.color(1, acc[9]); io.color(2, acc[10]);
#pragma unroll
    for (int k = 0; k < GS2M_NUM_FEATURES; ++k) io.feature(k, (k < p.F) ? acc[11 + k] : 0.f);

    const float* __restrict__ vm = p.viewmatrix;
    const float* __restrict__ pm = p.projmatrix;
    const float3 m = gin.mean;

    // =================== conic -> cov2D -> cov3D, mean (via the Jacobian) ===================
    // This block is numerically ill-conditioned (differences of large products), so the reference's own outputs
    // move by 1e-5..1e-4 (relative to the tensor max) between runs from atomic-order noise in dL/dconic alone.
    // To add nothing on top of that, every operation below uses explicit round-to-nearest intrinsics in exactly the
    // association order of the reference's computeCov2DCUDA as compiled for sm_100 (read from its SASS), so that
    // identical inputs give bit-identical outputs.
    const float c0 = gin.cov[0], c1 = gin.cov[1], c2 = gin.cov[2], c3 = gin.cov[3], c4 = gin.cov[4], c5 = gin.cov[5];
    const float tz_v = __fadd_rn(__fmaf_rn(m.z, vm[10], __fmaf_rn(m.x, vm[2], __fmul_rn(m.y, vm[6]))), vm[14]);
    const float tx_v = __fadd_rn(__fmaf_rn(m.z, vm[8], __fmaf_rn(m.x, vm[0], __fmul_rn(m.y, vm[4]))), vm[12]);
    const float ty_v = __fadd_rn(__fmaf_rn(m.z, vm[9], __fmaf_rn(m.x, vm[1], __fmul_rn(m.y, vm[5]))), vm[13]);
    const float limx = __fmul_rn(p.tan_fovx, 1.3f), limy = __fmul_rn(p.tan_fovy, 1.3f);
    const float txtz = __fdiv_rn(tx_v, tz_v), tytz = __fdiv_rn(ty_v, tz_v);
    const float t_x = __fmul_rn(tz_v, fminf(fmaxf(txtz, -limx), limx));
    const float t_y = __fmul_rn(tz_v, fminf(fmaxf(tytz, -limy), limy));
    const float x_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float fx = p.focal_x, fy = p.focal_y;
    const float tzsq = __fmul_rn(tz_v, tz_v);
    const float j00 = __fdiv_rn(fx, tz_v), j02 = __fdiv_rn(__fmul_rn(t_x, -fx), tzsq);
    const float j11 = __fdiv_rn(fy, tz_v), j12 = __fdiv_rn(__fmul_rn(t_y, -fy), tzsq);
    // rows of J * R_w2v
    const float X0 = __fmaf_rn(vm[2], j02, __fmul_rn(vm[0], j00));
    const float Y0 = __fmaf_rn(vm[6], j02, __fmul_rn(vm[4], j00));
    const float Z0 = __fmaf_rn(vm[10], j02, __fmul_rn(vm[8], j00));
    const float X1 = __fmaf_rn(vm[2], j12, __fmul_rn(vm[1], j11));
    const float Y1 = __fmaf_rn(vm[6], j12, __fmul_rn(vm[5], j11));
    const float Z1 = __fmaf_rn(vm[10], j12, __fmul_rn(vm[9], j11));
    // Sigma * rows
    const float V00 = __fmaf_rn(c2, Z0, __fmaf_rn(c0, X0, __fmul_rn(c1, Y0)));
    const float V01 = __fmaf_rn(c4, Z0, __fmaf_rn(c1, X0, __fmul_rn(c3, Y0)));
    const float V02 = __fmaf_rn(c5, Z0, __fmaf_rn(c2, X0, __fmul_rn(c4, Y0)));
    const float V10 = __fmaf_rn(c2, Z1, __fmaf_rn(c0, X1, __fmul_rn(c1, Y1)));
    const float V11 = __fmaf_rn(c4, Z1, __fmaf_rn(c1, X1, __fmul_rn(c3, Y1)));
    const float V12 = __fmaf_rn(c5, Z1, __fmaf_rn(c2, X1, __fmul_rn(c4, Y1)));
    // cov2D with the backward-only 0.3 dilation (backward.cu:205-207)
    const float a = __fadd_rn(__fmaf_rn(Z0, V02, __fmaf_rn(Y0, V01, __fmul_rn(X0, V00))), 0.3f);
    const float c = __fadd_rn(__fmaf_rn(Z1, V12, __fmaf_rn(Y1, V11, __fmul_rn(X1, V10))), 0.3f);
    const float b = __fmaf_rn(Z0, V12, __fmaf_rn(Y0, V11, __fmul_rn(X0, V10)));
    const float gx = acc[4], gy = acc[5], gz = acc[6];   // dL/dconic (xx, xy, yy)
    const float ac = __fmul_rn(a, c);
    const float denom = __fmaf_rn(-b, b, ac);
    const float k = __frcp_rn(__fmaf_rn(denom, denom, 0.0000001f));
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (k != 0.f) {
        const float b2 = __fadd_rn(b, b);
        const float d_m_ac = __fadd_rn(-ac, denom);
        dL_da = __fmul_rn(__fmaf_rn(gz, d_m_ac, __fmaf_rn(gy, __fmul_rn(c, b2), -__fmul_rn(gx, __fmul_rn(c, c)))), k);
        dL_dc = __fmul_rn(__fmaf_rn(gx, d_m_ac, __fmaf_rn(gy, __fmul_rn(b, __fadd_rn(a, a)), -__fmul_rn(gz, __fmul_rn(a, a)))), k);
        dL_db = __fmul_rn(__fadd_rn(k, k),
                          __fmaf_rn(gz, __fmul_rn(b, a), __fmaf_rn(gx, __fmul_rn(b, c), -__fmul_rn(gy, __fmaf_rn(b, b2, denom)))));
        dcov[0] = __fmaf_rn(dL_dc, __fmul_rn(X1, X1), __fmaf_rn(dL_da, __fmul_rn(X0, X0), __fmul_rn(dL_db, __fmul_rn(X0, X1))));
        dcov[3] = __fmaf_rn(dL_dc, __fmul_rn(Y1, Y1), __fmaf_rn(dL_da, __fmul_rn(Y0, Y0), __fmul_rn(dL_db, __fmul_rn(Y0, Y1))));
        dcov[5] = __fmaf_rn(dL_dc, __fmul_rn(Z1, Z1), __fmaf_rn(dL_da, __fmul_rn(Z0, Z0), __fmul_rn(dL_db, __fmul_rn(Z0, Z1))));
        dcov[1] = __fmaf_rn(dL_dc, __fmul_rn(Y1, __fadd_rn(X1, X1)),
                            __fmaf_rn(dL_da, __fmul_rn(Y0, __fadd_rn(X0, X0)), __fmul_rn(dL_db, __fmaf_rn(X0, Y1, __fmul_rn(Y0, X1)))));
        dcov[2] = __fmaf_rn(dL_dc, __fmul_rn(Z1, __fadd_rn(X1, X1)),
                            __fmaf_rn(dL_da, __fmul_rn(Z0, __fadd_rn(X0, X0)), __fmul_rn(dL_db, __fmaf_rn(X0, Z1, __fmul_rn(Z0, X1)))));
        dcov[4] = __fmaf_rn(dL_dc, __fmul_rn(Z1, __fadd_rn(Y1, Y1)),
                            __fmaf_rn(dL_da, __fmul_rn(Y0, __fadd_rn(Z0, Z0)), __fmul_rn(dL_db, __fmaf_rn(Y0, Z1, __fmul_rn(Z0, Y1)))));
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) io.cov(q, dcov[q]);

    // gradient w.r.t. the rows of J*R, then J, then the view-space mean
    const float dT00 = __fmaf_rn(__fadd_rn(V00, V00), dL_da, __fmul_rn(V10, dL_db));
    const float dT01 = __fmaf_rn(__fadd_rn(V01, V01), dL_da, __fmul_rn(V11, dL_db));
    const float dT02 = __fmaf_rn(__fadd_rn(V02, V02), dL_da, __fmul_rn(V12, dL_db));
    const float dT10 = __fmaf_rn(V00, dL_db, __fmul_rn(__fadd_rn(V10, V10), dL_dc));
    const float dT11 = __fmaf_rn(V01, dL_db, __fmul_rn(__fadd_rn(V11, V11), dL_dc));
    const float dT12 = __fmaf_rn(V02, dL_db, __fmul_rn(__fadd_rn(V12, V12), dL_dc));
    const float dJ00 = __fmaf_rn(vm[8], dT02, __fmaf_rn(vm[0], dT00, __fmul_rn(vm[4], dT01)));
    const float dJ02 = __fmaf_rn(vm[10], dT02, __fmaf_rn(vm[2], dT00, __fmul_rn(vm[6], dT01)));
    const float dJ11 = __fmaf_rn(vm[9], dT12, __fmaf_rn(vm[1], dT10, __fmul_rn(vm[5], dT11)));
    const float dJ12 = __fmaf_rn(vm[10], dT12, __fmaf_rn(vm[2], dT10, __fmul_rn(vm[6], dT11)));
    const float tz = __frcp_rn(tz_v);
    const float tz2 = __fmul_rn(tz, tz), tz3 = __fmul_rn(tz2, tz);
    const float dtx = __fmul_rn(dJ02, __fmul_rn(tz2, __fmul_rn(x_mul, -fx)));
    const float dty = __fmul_rn(dJ12, __fmul_rn(tz2, __fmul_rn(y_mul, -fy)));
    float dtz = __fmaf_rn(dJ00, __fmul_rn(tz2, -fx), -__fmul_rn(dJ11, __fmul_rn(tz2, fy)));
    dtz = __fmaf_rn(dJ02, __fmul_rn(tz3, __fmul_rn(t_x, __fadd_rn(fx, fx))), dtz);
    dtz = __fmaf_rn(dJ12, __fmul_rn(tz3, __fmul_rn(t_y, __fadd_rn(fy, fy))), dtz);
    float3 dmean = make_float3(__fmaf_rn(dtz, vm[2], __fmaf_rn(dtx, vm[0], __fmul_rn(dty, vm[1]))),
                               __fmaf_rn(dtz, vm[6], __fmaf_rn(dtx, vm[4], __fmul_rn(dty, vm[5]))),
                               __fmaf_rn(dtz, vm[10], __fmaf_rn(dtx, vm[8], __fmul_rn(dty, vm[9]))));

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
        const float3 d0 = make_float3(m.x - __ldg(p.cam_pos), m.y - __ldg(p.cam_pos + 1), m.z - __ldg(p.cam_pos + 2));
        const float inv_len = 1.0f / sqrtf(dot(d0, d0));
        const float x = d0.x * inv_len, y = d0.y * inv_len, z = d0.z * inv_len;
        const uchar4 cl = gin.clamped;
        const float3 dRGB = make_float3(cl.x ? 0.f : acc[8], cl.y ? 0.f : acc[9], cl.z ? 0.f : acc[10]);

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
                // read this coefficient before its slot is overwritten with the gradient (the staged sink aliases them)
                const float s = (active && k > 0) ? io.sh_in(3 * k) * dRGB.x + io.sh_in(3 * k + 1) * dRGB.y + io.sh_in(3 * k + 2) * dRGB.z : 0.f;
                io.sh_out(3 * k + 0, bk * dRGB.x);
                io.sh_out(3 * k + 1, bk * dRGB.y);
                io.sh_out(3 * k + 2, bk * dRGB.z);
                if (active && k > 0) {
                    ddir.x = fmaf(Bx[k], s, ddir.x);
                    ddir.y = fmaf(By[k], s, ddir.y);
                    ddir.z = fmaf(Bz[k], s, ddir.z);
                }
            }
        }
        if (IO::kAnyM) { for (int k = 48; k < 3 * M; ++k) io.sh_out(k, 0.f); }      // coefficients beyond degree 3 (M > 16)
        // through the normalisation dir = d0/|d0|
        const float s2 = dot(d0, d0);
        const float inv32 = 1.0f / sqrtf(s2 * s2 * s2);
        dmean.x += ((s2 - d0.x * d0.x) * ddir.x - d0.y * d0.x * ddir.y - d0.z * d0.x * ddir.z) * inv32;
        dmean.y += (-d0.x * d0.y * ddir.x + (s2 - d0.y * d0.y) * ddir.y - d0.z * d0.y * ddir.z) * inv32;
        dmean.z += (-d0.x * d0.z * ddir.x - d0.y * d0.z * ddir.y + (s2 - d0.z * d0.z) * ddir.z) * inv32;
    }
    io.mean3d(0, dmean.x); io.mean3d(1, dmean.y); io.mean3d(2, dmean.z);

    // =================== cov3D -> scale, rotation ===================
    if (p.scales != nullptr) {
        const float4 q = gin.rot;
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        const float s0 = p.scale_modifier * gin.scale.x, s1 = p.scale_modifier * gin.scale.y, s2 = p.scale_modifier * gin.scale.z;
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
        io.scale(0, 2.f * s0 * dot(r0, Gr0));
        io.scale(1, 2.f * s1 * dot(r1, Gr1));
        io.scale(2, 2.f * s2 * dot(r2, Gr2));
        // D_k = dL/dr_k = 2 s_k^2 G r_k ; chain through r_k(q)
        const float3 D0 = (2.f * s0 * s0) * Gr0, D1 = (2.f * s1 * s1) * Gr1, D2 = (2.f * s2 * s2) * Gr2;
        const float dq_r = 2.f * z * (D0.y - D1.x) + 2.f * y * (D2.x - D0.z) + 2.f * x * (D1.z - D2.y);
        const float dq_x = 2.f * y * (D1.x + D0.y) + 2.f * z * (D2.x + D0.z) + 2.f * r * (D1.z - D2.y) - 4.f * x * (D2.z + D1.y);
        const float dq_y = 2.f * x * (D1.x + D0.y) + 2.f * r * (D2.x - D0.z) + 2.f * z * (D1.z + D2.y) - 4.f * y * (D2.z + D0.x);
        const float dq_z = 2.f * r * (D0.y - D1.x) + 2.f * x * (D2.x + D0.z) + 2.f * y * (D1.z + D2.y) - 4.f * z * (D1.y + D0.x);
        io.rot(make_float4(dq_r, dq_x, dq_y, dq_z));
    } else {
        for (int k = 0; k < 3; ++k) io.scale(k, 0.f);
        io.rot(make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (p.shs == nullptr) {
#pragma unroll
        for (int k = 0; k < 48; ++k) { if (k < 3 * M) io.sh_out(k, 0.f); }
        if (IO::kAnyM) { for (int k = 48; k < 3 * M; ++k) io.sh_out(k, 0.f); }
    }
}


// Accumulate modes (gs2m_backward_args::accumulate): 0 = every tensor is overwritten; 1 = every caller-visible tensor is
// updated with +=; 2 = only the gradients of GS-2M's view-independent raw parameters (dL_dmeans3D, dL_dsh) are updated with
// +=, the view-dependent rest is overwritten (its chain rule through the caller's packing stage has to run per view).
template <int MODE>
struct AccPolicy {
    static constexpr bool kAccParams = MODE != 0;   // dL_dmeans3D, dL_dsh
    static constexpr bool kAccOther = MODE == 1;    // dL_dmeans2D, dL_dopacity, dL_dcolor, dL_dfeatures, dL_dcov3D, dL_dscale, dL_drot
};

// ---- sink 1: straight to global memory (any M) ----
template <int MODE>
struct GlobalIO : AccPolicy<MODE> {
    using AccPolicy<MODE>::kAccParams;
    using AccPolicy<MODE>::kAccOther;
    static constexpr bool kAnyM = true;      // handles rows of more than 48 floats (M > 16)
    const BwdParams& p; size_t i;
    __device__ GlobalIO(const BwdParams& p_, size_t i_) : p(p_), i(i_) {}
    __device__ void mean2d(float4 v) {
        float4* o = reinterpret_cast<float4*>(p.dL_dmeans2D) + i;
        if (kAccOther) { float4 t = *o; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        *o = v;
    }
    __device__ void conic(float4 v) { if (p.dL_dconic) reinterpret_cast<float4*>(p.dL_dconic)[i] = v; }
    __device__ void opacity(float v) { put<kAccOther>(p.dL_dopacity + i, v); }
    __device__ void color(int k, float v) { if (p.dL_dcolor) put<kAccOther>(p.dL_dcolor + 3 * i + k, v); }
    __device__ void feature(int k, float v) { put<kAccOther>(p.dL_dfeatures + GS2M_NUM_FEATURES * i + k, v); }
    __device__ void mean3d(int k, float v) { put<kAccParams>(p.dL_dmeans3D + 3 * i + k, v); }
    __device__ void cov(int k, float v) { if (p.dL_dcov3D) put<kAccOther>(p.dL_dcov3D + 6 * i + k, v); }
    __device__ void scale(int k, float v) { put<kAccOther>(p.dL_dscale + 3 * i + k, v); }
    __device__ void rot(float4 v) {
        float4* o = reinterpret_cast<float4*>(p.dL_drot) + i;
        if (kAccOther) { float4 t = *o; v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
        *o = v;
    }
    __device__ float sh_in(int k) const { return __ldg(p.shs + 3 * (size_t)p.M * i + k); }
    __device__ void sh_out(int k, float v) { if (p.dL_dsh) put<kAccParams>(p.dL_dsh + 3 * (size_t)p.M * i + k, v); }
};

// ---- sink 2: rows of a shared staging area (M <= 16); the block copies them out coalesced ----
constexpr int ST_SH = 48, ST_FEAT = 10, ST_COV = 6, ST_V3 = 3;
struct StageSmem {
    float sh[PBT * (ST_SH + 1)];    // SH coefficients on the way in, dL/dsh on the way out; row stride (3*M)|1 (odd:
                                    // conflict-free when every thread walks its own row)
    float feat[PBT * ST_FEAT];
    float cov[PBT * ST_COV];
    float mean3d[PBT * ST_V3];
    float scale[PBT * ST_V3];
    float color[PBT * ST_V3];
    unsigned char vis[PBT];         // row visibility: accumulate mode skips the rows of culled Gaussians
};
template <int MODE>
struct StagedIO : AccPolicy<MODE> {
    using AccPolicy<MODE>::kAccOther;
    static constexpr bool kAnyM = false;
    const BwdParams& p; size_t i; StageSmem& sm; int t; int sh_row;
    __device__ StagedIO(const BwdParams& p_, size_t i_, StageSmem& sm_, int t_) : p(p_), i(i_), sm(sm_), t(t_), sh_row((3 * p_.M) | 1) {}
    __device__ void mean2d(float4 v) {
        if (p.dL_dmeans2D == nullptr) return;       // (optional in chain mode)
        float4* o = reinterpret_cast<float4*>(p.dL_dmeans2D) + i;
        if (kAccOther) red_add_f4(o, v);
        else *o = v;
    }
    __device__ void conic(float4 v) { if (p.dL_dconic) reinterpret_cast<float4*>(p.dL_dconic)[i] = v; }
    __device__ void opacity(float v) { if (kAccOther) atomicAdd(p.dL_dopacity + i, v); else p.dL_dopacity[i] = v; }
    __device__ void color(int k, float v) { sm.color[t * ST_V3 + k] = v; }
    __device__ void feature(int k, float v) { sm.feat[t * ST_FEAT + k] = v; }
    __device__ void mean3d(int k, float v) { sm.mean3d[t * ST_V3 + k] = v; }
    __device__ void cov(int k, float v) { sm.cov[t * ST_COV + k] = v; }
    __device__ void scale(int k, float v) { sm.scale[t * ST_V3 + k] = v; }
    __device__ void rot(float4 v) {
        float4* o = reinterpret_cast<float4*>(p.dL_drot) + i;
        if (kAccOther) red_add_f4(o, v);
        else *o = v;
    }
    __device__ float sh_in(int k) const { return sm.sh[t * sh_row + k]; }
    __device__ void sh_out(int k, float v) { sm.sh[t * sh_row + k] = v; }
};

// ---- sink 3: sink 2 for dL/dsh, dL/dmeans2D and the optional precomputed-input gradients; the gradients w.r.t. the activated
// scale / rotation / opacity, the feature columns and the 3-D mean stay in registers and are chained through the caller-side
// packing stage (pack_math.cuh) before anything is written: what leaves the kernel are raw-parameter gradients
// (gs2m_backward_args::chain).  Saves the 72 B/Gaussian scratch round trip and a separate chain kernel per view.
template <int MODE>
struct ChainIO : StagedIO<MODE> {
    float g_scale[3], g_feat[GS2M_NUM_FEATURES], g_mean[3], g_opac;
    float4 g_rot;
    __device__ ChainIO(const BwdParams& p_, size_t i_, StageSmem& sm_, int t_) : StagedIO<MODE>(p_, i_, sm_, t_) {}
    __device__ void opacity(float v) { g_opac = v; }
    __device__ void feature(int k, float v) { g_feat[k] = v; }
    __device__ void mean3d(int k, float v) { g_mean[k] = v; }
    __device__ void scale(int k, float v) { g_scale[k] = v; }
    __device__ void rot(float4 v) { g_rot = v; }
};

// coalesced copy-out of `n` floats of the block's contiguous output region (16-byte aligned start)
template <bool ACC, int ROW>
__device__ __forceinline__ void block_store(float* __restrict__ dst, const float* __restrict__ src, int n,
                                            const unsigned char* __restrict__ vis) {
    const int n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int e = threadIdx.x; e < n4; e += PBT) {
        if (ACC && !(vis[(4 * e) / ROW] | vis[(4 * e + 1) / ROW] | vis[(4 * e + 2) / ROW] | vis[(4 * e + 3) / ROW]))
            continue;                                                           // += 0 for culled rows: skip the RMW
        float4 v = s4[e];
        if (ACC) {
            // rows of culled Gaussians hold no staged values in accumulate mode: mask them out element-wise
            v.x = vis[(4 * e) / ROW] ? v.x : 0.f;
            v.y = vis[(4 * e + 1) / ROW] ? v.y : 0.f;
            v.z = vis[(4 * e + 2) / ROW] ? v.z : 0.f;
            v.w = vis[(4 * e + 3) / ROW] ? v.w : 0.f;
            red_add_f4(d4 + e, v);
        } else {
            d4[e] = v;
        }
    }
    for (int e = (n4 << 2) + threadIdx.x; e < n; e += PBT) {
        if (ACC) { if (vis[e / ROW]) atomicAdd(dst + e, src[e]); }
        else dst[e] = src[e];
    }
}

template <int MODE>
__global__ void __launch_bounds__(256) preprocess_backward_generic_kernel(BwdParams p, GeomState g) {
    const int idx = p.row_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.row_end) return;
    GlobalIO<MODE> io(p, (size_t)idx);
    const bool visible = p.radii[idx] > 0;
    GaussianInputs gin;
    load_gaussian_inputs(p, g, (size_t)idx, visible, gin);
    gaussian_backward(p, idx, visible, gin, io);
}

template <int MODE, bool CHAIN>
__global__ void __maxnreg__(GS2M_PB_MAXNREG) preprocess_backward_staged_kernel(BwdParams p, GeomState g) {
    constexpr bool kAccParams = AccPolicy<MODE>::kAccParams, kAccOther = AccPolicy<MODE>::kAccOther;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StageSmem& sm = *reinterpret_cast<StageSmem*>(smem_raw);
    const int t = threadIdx.x;
    const size_t row0 = (size_t)p.row_begin + (size_t)blockIdx.x * PBT;      // row_begin is a multiple of 256 (checked by the API)
    const int rows = (int)min((size_t)PBT, (size_t)p.row_end - row0);
    const int idx = (int)row0 + t;
    const bool inside = t < rows;
    const bool visible = inside && p.radii[idx] > 0;
    const int sh_row = 3 * p.M;
    // the thread's own inputs are requested first: their latency passes under the cooperative SH staging below
    GaussianInputs gin;
    load_gaussian_inputs(p, g, (size_t)(inside ? idx : 0), visible, gin);
    // coalesced load of the SH rows of the block's visible Gaussians
    const int sh_pad = sh_row | 1;
    if (p.shs != nullptr) {
        const float* __restrict__ src = p.shs + row0 * sh_row;
        const int n = rows * sh_row;
        if ((sh_row & 3) == 0) {      // 128-bit global loads; a float4 never straddles two rows
            for (int e4 = t; e4 < (n >> 2); e4 += PBT) {
                const int r = (4 * e4) / sh_row, c = (4 * e4) - r * sh_row;
                if (p.radii[row0 + r] > 0) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + e4);
                    float* d = sm.sh + r * sh_pad + c;
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        } else {
            for (int e = t; e < n; e += PBT) {
                const int r = e / sh_row, c = e - r * sh_row;
                if (p.radii[row0 + r] > 0) sm.sh[r * sh_pad + c] = __ldg(src + e);
            }
        }
        __syncthreads();
    }
    sm.vis[t] = visible ? 1 : 0;
    float* const s_scalar = sm.feat;      // chain mode: the feature staging is unused; it carries three [PBT] scalar columns
    if (inside) {
        if (CHAIN) {
            ChainIO<MODE> io(p, (size_t)idx, sm, t);
            gaussian_backward(p, idx, visible, gin, io);
            RawGrads r;
            float mean[3] = {0.f, 0.f, 0.f};
            if (visible) {
                const PackIn in{p.P, p.means3D, p.chain.scaling_raw, p.chain.rotation_raw, p.chain.opacity_raw, p.chain.albedo_raw,
                                p.chain.roughness_raw, p.chain.metallic_raw, p.viewmatrix, p.cam_pos, p.chain.z_depth,
                                p.chain.blend_metallic};
                const Derived d = derive(in, idx);
                r = pack_chain(in, idx, d, io.g_scale, io.g_rot, io.g_opac, io.g_feat);
#pragma unroll
                for (int k = 0; k < 3; ++k) mean[k] = io.g_mean[k] + r.dp[k];
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) { r.dscaling[k] = 0.f; r.dalbedo[k] = 0.f; }
                r.drot = make_float4(0.f, 0.f, 0.f, 0.f);
                r.dopacity = r.droughness = r.dmetallic = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                sm.mean3d[t * ST_V3 + k] = mean[k];
                sm.scale[t * ST_V3 + k] = r.dscaling[k];
                sm.color[t * ST_V3 + k] = r.dalbedo[k];
            }
            s_scalar[t] = r.dopacity; s_scalar[PBT + t] = r.droughness; s_scalar[2 * PBT + t] = r.dmetallic;
            float4* o = reinterpret_cast<float4*>(p.chain.d_rotation_raw) + idx;
            if (kAccParams) { if (visible) red_add_f4(o, r.drot); }
            else *o = r.drot;
        } else {
            StagedIO<MODE> io(p, (size_t)idx, sm, t);
            gaussian_backward(p, idx, visible, gin, io);
        }
    }
    __syncthreads();
    if (p.dL_dsh && sh_row > 0) {
        float* __restrict__ dst = p.dL_dsh + row0 * sh_row;
        const int n = rows * sh_row;
        if ((sh_row & 3) == 0) {
            for (int e4 = t; e4 < (n >> 2); e4 += PBT) {
                const int r = (4 * e4) / sh_row, c = (4 * e4) - r * sh_row;
                if (kAccParams && !sm.vis[r]) continue;
                const float* q = sm.sh + r * sh_pad + c;
                float4 v = make_float4(q[0], q[1], q[2], q[3]);
                float4* d4 = reinterpret_cast<float4*>(dst) + e4;
                if (kAccParams) red_add_f4(d4, v);
                else *d4 = v;
            }
        } else {
            for (int e = t; e < n; e += PBT) {
                const int r = e / sh_row, c = e - r * sh_row;
                if (kAccParams && !sm.vis[r]) continue;
                const float v = sm.sh[r * sh_pad + c];
                dst[e] = kAccParams ? dst[e] + v : v;
            }
        }
    }
    if (CHAIN) {      // raw-parameter gradients: overwritten (MODE 0) or added for the visible rows (MODE 2)
        block_store<kAccParams, ST_V3>(p.chain.d_xyz + row0 * ST_V3, sm.mean3d, rows * ST_V3, sm.vis);
        block_store<kAccParams, ST_V3>(p.chain.d_scaling_raw + row0 * ST_V3, sm.scale, rows * ST_V3, sm.vis);
        block_store<kAccParams, ST_V3>(p.chain.d_albedo_raw + row0 * ST_V3, sm.color, rows * ST_V3, sm.vis);
        block_store<kAccParams, 1>(p.chain.d_opacity_raw + row0, s_scalar, rows, sm.vis);
        block_store<kAccParams, 1>(p.chain.d_roughness_raw + row0, s_scalar + PBT, rows, sm.vis);
        block_store<kAccParams, 1>(p.chain.d_metallic_raw + row0, s_scalar + 2 * PBT, rows, sm.vis);
        if (p.dL_dcov3D) block_store<kAccOther, ST_COV>(p.dL_dcov3D + row0 * ST_COV, sm.cov, rows * ST_COV, sm.vis);
        return;
    }
    block_store<kAccOther, ST_FEAT>(p.dL_dfeatures + row0 * ST_FEAT, sm.feat, rows * ST_FEAT, sm.vis);
    if (p.dL_dcov3D) block_store<kAccOther, ST_COV>(p.dL_dcov3D + row0 * ST_COV, sm.cov, rows * ST_COV, sm.vis);
    block_store<kAccParams, ST_V3>(p.dL_dmeans3D + row0 * ST_V3, sm.mean3d, rows * ST_V3, sm.vis);
    block_store<kAccOther, ST_V3>(p.dL_dscale + row0 * ST_V3, sm.scale, rows * ST_V3, sm.vis);
    if (p.dL_dcolor) block_store<kAccOther, ST_V3>(p.dL_dcolor + row0 * ST_V3, sm.color, rows * ST_V3, sm.vis);
}


// ================================ several views in one pass (chain mode) ================================
// A view-sharded step owes every raw parameter the SUM of its gradients over the rank's views.  Launched per view, this stage
// re-reads the Gaussian's parameters and SH row and read-modify-writes all 64 output floats once per view that sees it; here a
// thread owns one Gaussian for ALL the views of the launch: it walks the views whose radii say "visible", runs the same
// per-Gaussian math + packing chain on each view's accumulator row, sums the 16 chained gradients in registers and the 48 SH
// gradients in its shared-memory row, and the block writes every output element exactly once, coalesced.
constexpr int MAX_VIEWS = 8;      // views per launch (kernel-parameter space: 8 x ~0.5 KB); longer lists run in groups
struct ViewSet {
    int n;
    BwdParams p[MAX_VIEWS];
    GeomState g[MAX_VIEWS];
};
struct ViewsSmem {
    float sh_in[PBT * (ST_SH + 1)];    // SH coefficients of the block's Gaussians (loaded once, read by every view)
    float sh_acc[PBT * (ST_SH + 1)];   // sum over the views of dL/dsh
    float mean3d[PBT * ST_V3];
    float scale[PBT * ST_V3];
    float color[PBT * ST_V3];
    float scalar[3 * PBT];
    unsigned char vis[PBT];            // visible in at least one view of the launch
};
struct ViewsIO {
    static constexpr bool kAccParams = false, kAccOther = false, kAnyM = false;     // (only read on the culled path, unused here)
    const BwdParams& p; size_t i; ViewsSmem& sm; int t; int sh_row;
    float g_scale[3], g_feat[GS2M_NUM_FEATURES], g_mean[3], g_opac;
    float4 g_rot;
    __device__ ViewsIO(const BwdParams& p_, size_t i_, ViewsSmem& sm_, int t_) : p(p_), i(i_), sm(sm_), t(t_), sh_row((3 * p_.M) | 1) {}
    __device__ void mean2d(float4 v) { if (p.dL_dmeans2D) reinterpret_cast<float4*>(p.dL_dmeans2D)[i] = v; }     // per view
    __device__ void conic(float4 v) { if (p.dL_dconic) reinterpret_cast<float4*>(p.dL_dconic)[i] = v; }
    __device__ void opacity(float v) { g_opac = v; }
    __device__ void color(int, float) {}
    __device__ void feature(int k, float v) { g_feat[k] = v; }
    __device__ void mean3d(int k, float v) { g_mean[k] = v; }
    __device__ void cov(int, float) {}
    __device__ void scale(int k, float v) { g_scale[k] = v; }
    __device__ void rot(float4 v) { g_rot = v; }
    __device__ float sh_in(int k) const { return sm.sh_in[t * sh_row + k]; }
    __device__ void sh_out(int k, float v) { sm.sh_acc[t * sh_row + k] += v; }
};

template <bool ACC>
__global__ void __maxnreg__(144) preprocess_backward_views_kernel(const __grid_constant__ ViewSet vs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ViewsSmem& sm = *reinterpret_cast<ViewsSmem*>(smem_raw);
    const BwdParams& p0 = vs.p[0];
    const int t = threadIdx.x;
    const size_t row0 = (size_t)p0.row_begin + (size_t)blockIdx.x * PBT;
    const int rows = (int)min((size_t)PBT, (size_t)p0.row_end - row0);
    const int idx = (int)row0 + t;
    const bool inside = t < rows;
    const int sh_row = 3 * p0.M, sh_pad = sh_row | 1;
    // which views see this Gaussian
    uint32_t vmask = 0;
    if (inside) {
#pragma unroll
        for (int v = 0; v < MAX_VIEWS; ++v)
            if (v < vs.n && __ldg(vs.p[v].radii + idx) > 0) vmask |= 1u << v;
    }
    sm.vis[t] = vmask != 0;
    for (int k = 0; k < sh_pad; ++k) sm.sh_acc[t * sh_pad + k] = 0.f;
    __syncthreads();
    // coalesced load of the SH rows of the Gaussians some view sees
    if (p0.shs != nullptr) {
        const float* __restrict__ src = p0.shs + row0 * sh_row;
        const int n = rows * sh_row;
        if ((sh_row & 3) == 0) {
            for (int e4 = t; e4 < (n >> 2); e4 += PBT) {
                const int r = (4 * e4) / sh_row, c = (4 * e4) - r * sh_row;
                if (sm.vis[r]) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + e4);
                    float* d = sm.sh_in + r * sh_pad + c;
                    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
                }
            }
        } else {
            for (int e = t; e < n; e += PBT) {
                const int r = e / sh_row, c = e - r * sh_row;
                if (sm.vis[r]) sm.sh_in[r * sh_pad + c] = __ldg(src + e);
            }
        }
        __syncthreads();
    }
    RawGrads sum;
    float mean[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) { sum.dscaling[k] = 0.f; sum.dalbedo[k] = 0.f; }
    sum.drot = make_float4(0.f, 0.f, 0.f, 0.f);
    sum.dopacity = sum.droughness = sum.dmetallic = 0.f;
    // the camera-independent half of the packing stage, once per Gaussian
    Derived d;
    SigmoidSlopes slopes;
    float pos[3];
    if (vmask) {
        const PackIn in0{p0.P, p0.means3D, p0.chain.scaling_raw, p0.chain.rotation_raw, p0.chain.opacity_raw, p0.chain.albedo_raw,
                         p0.chain.roughness_raw, p0.chain.metallic_raw, p0.viewmatrix, p0.cam_pos, p0.chain.z_depth, p0.chain.blend_metallic};
        derive_gaussian(in0, idx, d, pos);
        slopes = sigmoid_slopes(in0, idx);
    }
#pragma unroll 1
    for (int v = 0; v < vs.n; ++v) {
        if (!((vmask >> v) & 1u)) continue;
        const BwdParams& p = vs.p[v];
        const GeomState& g = vs.g[v];
        GaussianInputs gin;
        load_gaussian_inputs(p, g, (size_t)idx, true, gin);
        ViewsIO io(p, (size_t)idx, sm, t);
        gaussian_backward(p, idx, true, gin, io);
        const PackIn in{p.P, p.means3D, p.chain.scaling_raw, p.chain.rotation_raw, p.chain.opacity_raw, p.chain.albedo_raw,
                        p.chain.roughness_raw, p.chain.metallic_raw, p.viewmatrix, p.cam_pos, p.chain.z_depth, p.chain.blend_metallic};
        derive_view(in, pos, d);
        const RawGrads r = pack_chain(in, d, slopes, io.g_scale, io.g_rot, io.g_opac, io.g_feat);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mean[k] += io.g_mean[k] + r.dp[k];
            sum.dscaling[k] += r.dscaling[k];
            sum.dalbedo[k] += r.dalbedo[k];
        }
        sum.drot.x += r.drot.x; sum.drot.y += r.drot.y; sum.drot.z += r.drot.z; sum.drot.w += r.drot.w;
        sum.dopacity += r.dopacity; sum.droughness += r.droughness; sum.dmetallic += r.dmetallic;
    }
    if (inside) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            sm.mean3d[t * ST_V3 + k] = mean[k];
            sm.scale[t * ST_V3 + k] = sum.dscaling[k];
            sm.color[t * ST_V3 + k] = sum.dalbedo[k];
        }
        sm.scalar[t] = sum.dopacity; sm.scalar[PBT + t] = sum.droughness; sm.scalar[2 * PBT + t] = sum.dmetallic;
        float4* o = reinterpret_cast<float4*>(p0.chain.d_rotation_raw) + idx;
        if (ACC) { if (vmask) red_add_f4(o, sum.drot); }
        else *o = sum.drot;
    }
    __syncthreads();
    if (p0.dL_dsh && sh_row > 0) {
        float* __restrict__ dst = p0.dL_dsh + row0 * sh_row;
        const int n = rows * sh_row;
        if ((sh_row & 3) == 0) {
            for (int e4 = t; e4 < (n >> 2); e4 += PBT) {
                const int r = (4 * e4) / sh_row, c = (4 * e4) - r * sh_row;
                if (ACC && !sm.vis[r]) continue;
                const float* q = sm.sh_acc + r * sh_pad + c;
                const float4 v = make_float4(q[0], q[1], q[2], q[3]);
                float4* d4 = reinterpret_cast<float4*>(dst) + e4;
                if (ACC) red_add_f4(d4, v);
                else *d4 = v;
            }
        } else {
            for (int e = t; e < n; e += PBT) {
                const int r = e / sh_row, c = e - r * sh_row;
                if (ACC && !sm.vis[r]) continue;
                const float v = sm.sh_acc[r * sh_pad + c];
                dst[e] = ACC ? dst[e] + v : v;
            }
        }
    }
    block_store<ACC, ST_V3>(p0.chain.d_xyz + row0 * ST_V3, sm.mean3d, rows * ST_V3, sm.vis);
    block_store<ACC, ST_V3>(p0.chain.d_scaling_raw + row0 * ST_V3, sm.scale, rows * ST_V3, sm.vis);
    block_store<ACC, ST_V3>(p0.chain.d_albedo_raw + row0 * ST_V3, sm.color, rows * ST_V3, sm.vis);
    block_store<ACC, 1>(p0.chain.d_opacity_raw + row0, sm.scalar, rows, sm.vis);
    block_store<ACC, 1>(p0.chain.d_roughness_raw + row0, sm.scalar + PBT, rows, sm.vis);
    block_store<ACC, 1>(p0.chain.d_metallic_raw + row0, sm.scalar + 2 * PBT, rows, sm.vis);
}

}  // namespace

int launch_preprocess_backward(const BwdParams& p, const GeomState& g, cudaStream_t s) {
    if (p.P == 0 || p.row_end <= p.row_begin) return GS2M_OK;
    const int blocks = (p.row_end - p.row_begin + 255) / 256;             // generic kernel: 256 threads
    const int blocks_staged = (p.row_end - p.row_begin + PBT - 1) / PBT;
    count_launches(1);
    if (p.accumulate < 0 || p.accumulate > 2) { set_error("accumulate mode %d outside 0..2", p.accumulate); return GS2M_ERR_INVALID_ARGUMENT; }
    if (p.M <= 16) {
        static PerDeviceOnce configured;
        int dev;
        if (configured.need(dev)) {
#define GS2M_PB_ATTR(MODE, CHAIN) GS2M_CUDA(cudaFuncSetAttribute(preprocess_backward_staged_kernel<MODE, CHAIN>, \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StageSmem)))
            GS2M_PB_ATTR(0, false); GS2M_PB_ATTR(1, false); GS2M_PB_ATTR(2, false); GS2M_PB_ATTR(0, true); GS2M_PB_ATTR(2, true);
#undef GS2M_PB_ATTR
            configured.done(dev);
        }
        const size_t smem = sizeof(StageSmem);
        if (p.has_chain) {
            if (p.accumulate == 2) preprocess_backward_staged_kernel<2, true><<<blocks_staged, PBT, smem, s>>>(p, g);
            else preprocess_backward_staged_kernel<0, true><<<blocks_staged, PBT, smem, s>>>(p, g);
        } else if (p.accumulate == 1) preprocess_backward_staged_kernel<1, false><<<blocks_staged, PBT, smem, s>>>(p, g);
        else if (p.accumulate == 2) preprocess_backward_staged_kernel<2, false><<<blocks_staged, PBT, smem, s>>>(p, g);
        else preprocess_backward_staged_kernel<0, false><<<blocks_staged, PBT, smem, s>>>(p, g);
    } else {
        if (p.has_chain) { set_error("chain mode needs M <= 16"); return GS2M_ERR_INVALID_ARGUMENT; }
        if (p.accumulate == 1) preprocess_backward_generic_kernel<1><<<blocks, 256, 0, s>>>(p, g);
        else if (p.accumulate == 2) preprocess_backward_generic_kernel<2><<<blocks, 256, 0, s>>>(p, g);
        else preprocess_backward_generic_kernel<0><<<blocks, 256, 0, s>>>(p, g);
    }
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int launch_preprocess_backward_views(const BwdParams* ps, const GeomState* gs, int n_views, bool accumulate, cudaStream_t s) {
    const BwdParams& p0 = ps[0];
    if (p0.P == 0 || p0.row_end <= p0.row_begin || n_views <= 0) return GS2M_OK;
    if (p0.M > 16 || !p0.has_chain) { set_error("backward_views needs a chain and M <= 16"); return GS2M_ERR_INVALID_ARGUMENT; }
    static PerDeviceOnce configured;
    int dev;
    if (configured.need(dev)) {
        GS2M_CUDA(cudaFuncSetAttribute(preprocess_backward_views_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ViewsSmem)));
        GS2M_CUDA(cudaFuncSetAttribute(preprocess_backward_views_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ViewsSmem)));
        configured.done(dev);
    }
    const int blocks = (p0.row_end - p0.row_begin + PBT - 1) / PBT;
    for (int v0 = 0; v0 < n_views; v0 += MAX_VIEWS) {       // groups of MAX_VIEWS: the first overwrites (unless told to add)
        ViewSet vs;
        vs.n = std::min(MAX_VIEWS, n_views - v0);
        for (int v = 0; v < vs.n; ++v) { vs.p[v] = ps[v0 + v]; vs.g[v] = gs[v0 + v]; }
        for (int v = vs.n; v < MAX_VIEWS; ++v) { vs.p[v] = ps[v0]; vs.g[v] = gs[v0]; }
        count_launches(1);
        if (accumulate || v0 > 0) preprocess_backward_views_kernel<true><<<blocks, PBT, sizeof(ViewsSmem), s>>>(vs);
        else preprocess_backward_views_kernel<false><<<blocks, PBT, sizeof(ViewsSmem), s>>>(vs);
        GS2M_CUDA(cudaGetLastError());
    }
    return GS2M_OK;
}

}  // namespace gs2m
