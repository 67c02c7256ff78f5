// Per-Gaussian math of the caller-side packing stage (parameter activations, thinnest-axis normal, camera-space distance),
// shared by the stand-alone packing kernels (feature_pack.cu) and by the rasterizer's per-Gaussian backward when it chains its
// gradients straight through this stage (preprocess_bwd.cu, gs2m_backward_args::chain).
//
// Behavioural reference: scene/gaussian_model.py:113-172 (getters, get_normals), utils/general_utils.py:72-92 (build_rotation),
// gaussian_renderer/__init__.py:82-96 (cam_normals / cam_points / features).
#pragma once
#include "common.cuh"

namespace gs2m {

struct PackIn {
    int P;
    const float *xyz, *scaling, *rotation, *opacity, *albedo, *roughness, *metallic;   // raw (pre-activation) parameters
    const float *wvt, *campos;                                                          // world_view_transform (4x4 row-major), camera centre
    int z_depth, blend_metallic;
};

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

struct Derived {
    float s[3], q[4], nq, qh[4], nqh, col[3], n[3], nm, camn[3], camp[3], u;
    int axis;
    bool flip;
};

__device__ __forceinline__ void column_of_R(const float* q, int axis, float* c) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    if (axis == 0) { c[0] = 1.f - 2.f * (y * y + z * z); c[1] = 2.f * (x * y + r * z); c[2] = 2.f * (x * z - r * y); }
    else if (axis == 1) { c[0] = 2.f * (x * y - r * z); c[1] = 1.f - 2.f * (x * x + z * z); c[2] = 2.f * (y * z + r * x); }
    else { c[0] = 2.f * (x * z + r * y); c[1] = 2.f * (y * z - r * x); c[2] = 1.f - 2.f * (x * x + y * y); }
}

// derive() in two halves: what depends only on the Gaussian (activations, normalised quaternion, thinnest axis and its
// rotation-matrix column) and what depends on the camera (flip towards the camera, camera-space normal / position, distance).
// A caller that visits one Gaussian for several views (preprocess_bwd.cu, backward_views) runs the first half once.
__device__ __forceinline__ void derive_gaussian(const PackIn& in, int i, Derived& d, float (&p)[3]) {
    p[0] = __ldg(in.xyz + 3 * (size_t)i); p[1] = __ldg(in.xyz + 3 * (size_t)i + 1); p[2] = __ldg(in.xyz + 3 * (size_t)i + 2);
#pragma unroll
    for (int k = 0; k < 3; ++k) d.s[k] = expf(__ldg(in.scaling + 3 * (size_t)i + k));
    const float4 rq = __ldg(reinterpret_cast<const float4*>(in.rotation + 4 * (size_t)i));
    d.nq = fmaxf(sqrtf(rq.x * rq.x + rq.y * rq.y + rq.z * rq.z + rq.w * rq.w), 1e-12f);   // F.normalize eps
    d.q[0] = rq.x / d.nq; d.q[1] = rq.y / d.nq; d.q[2] = rq.z / d.nq; d.q[3] = rq.w / d.nq;
    d.nqh = sqrtf(d.q[0] * d.q[0] + d.q[1] * d.q[1] + d.q[2] * d.q[2] + d.q[3] * d.q[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) d.qh[k] = d.q[k] / d.nqh;
    d.axis = 0;
    if (d.s[1] < d.s[d.axis]) d.axis = 1;       // first minimum, like torch.argmin
    if (d.s[2] < d.s[d.axis]) d.axis = 2;
    column_of_R(d.qh, d.axis, d.col);
}

__device__ __forceinline__ void derive_view(const PackIn& in, const float (&p)[3], Derived& d) {
    const float* W = in.wvt;
    const float vd = d.col[0] * (in.campos[0] - p[0]) + d.col[1] * (in.campos[1] - p[1]) + d.col[2] * (in.campos[2] - p[2]);
    d.flip = vd < 0.0f;
    float m[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) m[k] = d.flip ? -d.col[k] : d.col[k];
    d.nm = sqrtf(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) d.n[k] = m[k] / d.nm;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        d.camn[k] = d.n[0] * W[k] + d.n[1] * W[4 + k] + d.n[2] * W[8 + k];
        d.camp[k] = p[0] * W[k] + p[1] * W[4 + k] + p[2] * W[8 + k] + W[12 + k];
    }
    d.u = d.camn[0] * d.camp[0] + d.camn[1] * d.camp[1] + d.camn[2] * d.camp[2];
}

__device__ __forceinline__ Derived derive(const PackIn& in, int i) {
    Derived d;
    float p[3];
    derive_gaussian(in, i, d, p);
    derive_view(in, p, d);
    return d;
}

// Adjoint of derive(): upstream gradients w.r.t. the activated scales (gs), rotations (gq), opacity (go) and the 10 feature columns
// (gf) -> gradients w.r.t. the raw parameters.  `dp` is the extra position gradient through the distance / depth column.
struct RawGrads { float dp[3], dscaling[3], dalbedo[3]; float4 drot; float dopacity, droughness, dmetallic; };

// d sigmoid / d raw of the six sigmoid-activated parameters: per Gaussian, the same for every view
struct SigmoidSlopes { float albedo[3], opacity, roughness, metallic; };
__device__ __forceinline__ SigmoidSlopes sigmoid_slopes(const PackIn& in, int i) {
    SigmoidSlopes k;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = sigmoidf(__ldg(in.albedo + 3 * (size_t)i + c));
        k.albedo[c] = a * (1.0f - a);
    }
    const float o = sigmoidf(__ldg(in.opacity + i));
    k.opacity = o * (1.0f - o);
    const float ro = sigmoidf(__ldg(in.roughness + i));
    k.roughness = ro * (1.0f - ro);
    const float me = sigmoidf(__ldg(in.metallic + i));
    k.metallic = me * (1.0f - me);
    return k;
}

// The gradients that arrive here are often tiny (a Gaussian that barely touches a view), and an IEEE division with a tiny or
// subnormal numerator leaves the hardware's fast path for a ~100-instruction routine (measured: 40 % of the per-Gaussian
// backward's instructions).  The three norms are O(1), so the adjoints multiply by their reciprocals instead.
__device__ __forceinline__ RawGrads pack_chain(const PackIn& in, const Derived& d, const SigmoidSlopes& sl, const float* gs,
                                               float4 gq, float go, const float* gf) {
    const float* W = in.wvt;
    RawGrads out;
    // distance / depth column
    const float du = in.z_depth ? 0.0f : ((d.u > 0.f) ? gf[1] : ((d.u < 0.f) ? -gf[1] : 0.0f));
    float dn[3], dp[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float Wcp = W[4 * j] * d.camp[0] + W[4 * j + 1] * d.camp[1] + W[4 * j + 2] * d.camp[2];
        const float Wcn = W[4 * j] * d.camn[0] + W[4 * j + 1] * d.camn[1] + W[4 * j + 2] * d.camn[2];
        dn[j] = gf[2 + j] + du * Wcp;
        dp[j] = du * Wcn + (in.z_depth ? gf[1] * W[4 * j + 2] : 0.0f);
    }
    // normalisation n = m / |m|, flip, column of R(qh)
    const float inv_nm = 1.0f / d.nm, inv_nqh = 1.0f / d.nqh, inv_nq = 1.0f / d.nq;
    const float ndn = d.n[0] * dn[0] + d.n[1] * dn[1] + d.n[2] * dn[2];
    float dc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dm = (dn[k] - d.n[k] * ndn) * inv_nm;
        dc[k] = d.flip ? -dm : dm;
    }
    const float r = d.qh[0], x = d.qh[1], y = d.qh[2], z = d.qh[3];
    float dqh[4];
    if (d.axis == 0) {
        dqh[0] = 2.f * z * dc[1] - 2.f * y * dc[2];
        dqh[1] = 2.f * y * dc[1] + 2.f * z * dc[2];
        dqh[2] = -4.f * y * dc[0] + 2.f * x * dc[1] - 2.f * r * dc[2];
        dqh[3] = -4.f * z * dc[0] + 2.f * r * dc[1] + 2.f * x * dc[2];
    } else if (d.axis == 1) {
        dqh[0] = -2.f * z * dc[0] + 2.f * x * dc[2];
        dqh[1] = 2.f * y * dc[0] - 4.f * x * dc[1] + 2.f * r * dc[2];
        dqh[2] = 2.f * x * dc[0] + 2.f * z * dc[2];
        dqh[3] = -2.f * r * dc[0] - 4.f * z * dc[1] + 2.f * y * dc[2];
    } else {
        dqh[0] = 2.f * y * dc[0] - 2.f * x * dc[1];
        dqh[1] = 2.f * z * dc[0] - 2.f * r * dc[1] - 4.f * x * dc[2];
        dqh[2] = 2.f * r * dc[0] + 2.f * z * dc[1] - 4.f * y * dc[2];
        dqh[3] = 2.f * x * dc[0] + 2.f * y * dc[1];
    }
    // qh = q / |q|  (build_rotation), then add the rasterizer's gradient w.r.t. q, then q = raw / max(|raw|, eps)
    const float qd = d.qh[0] * dqh[0] + d.qh[1] * dqh[1] + d.qh[2] * dqh[2] + d.qh[3] * dqh[3];
    const float dq[4] = {(dqh[0] - d.qh[0] * qd) * inv_nqh + gq.x, (dqh[1] - d.qh[1] * qd) * inv_nqh + gq.y,
                         (dqh[2] - d.qh[2] * qd) * inv_nqh + gq.z, (dqh[3] - d.qh[3] * qd) * inv_nqh + gq.w};
    const float qq = d.q[0] * dq[0] + d.q[1] * dq[1] + d.q[2] * dq[2] + d.q[3] * dq[3];
    const bool clamped = d.nq <= 1e-12f;   // F.normalize divides by the clamped norm: no projection term then
    out.drot = make_float4((dq[0] - (clamped ? 0.f : d.q[0] * qq)) * inv_nq, (dq[1] - (clamped ? 0.f : d.q[1] * qq)) * inv_nq,
                           (dq[2] - (clamped ? 0.f : d.q[2] * qq)) * inv_nq, (dq[3] - (clamped ? 0.f : d.q[3] * qq)) * inv_nq);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        out.dp[k] = dp[k];
        out.dscaling[k] = gs[k] * d.s[k];
        out.dalbedo[k] = gf[5 + k] * sl.albedo[k];
    }
    out.dopacity = go * sl.opacity;
    out.droughness = gf[8] * sl.roughness;
    out.dmetallic = in.blend_metallic ? gf[9] * sl.metallic : 0.0f;
    return out;
}

__device__ __forceinline__ RawGrads pack_chain(const PackIn& in, int i, const Derived& d, const float* gs, float4 gq, float go,
                                               const float* gf) {
    return pack_chain(in, d, sigmoid_slopes(in, i), gs, gq, go, gf);
}

}  // namespace gs2m
