// Caller-side stage in front of the rasterizer: parameter activations and the packing of the
// 10-column `features` tensor, forward and backward, one thread per Gaussian, one kernel each.
//
// Behavioural reference (all PyTorch eager ops + autograd in GS-2M, ~15 kernels forward and ~40 backward per render call):
//   scene/gaussian_model.py:113-172  getters: exp(_scaling), normalize(_rotation), sigmoid(_opacity/_albedo/_roughness/_metallic)
//   scene/gaussian_model.py:146-160  get_normals: thinnest axis of R(q) (first minimum), flipped to face the camera, normalised
//   utils/general_utils.py:72-92     build_rotation (normalises the quaternion once more)
//   gaussian_renderer/__init__.py:82-96  cam_normals / cam_points, features = [1, |n_cam . p_cam| or z, n(3), albedo(3), rough, metal]
// The backward is the analytic adjoint of exactly that graph (argmin and the flip test carry no gradient, |.| uses sign()).
#include "common.cuh"

namespace gs2m {
namespace {

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

__device__ __forceinline__ Derived derive(const PackIn& in, int i) {
    Derived d;
    const float* W = in.wvt;
    const float p[3] = {in.xyz[3 * i], in.xyz[3 * i + 1], in.xyz[3 * i + 2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) d.s[k] = expf(in.scaling[3 * i + k]);
    const float4 rq = *reinterpret_cast<const float4*>(in.rotation + 4 * (size_t)i);
    d.nq = fmaxf(sqrtf(rq.x * rq.x + rq.y * rq.y + rq.z * rq.z + rq.w * rq.w), 1e-12f);   // F.normalize eps
    d.q[0] = rq.x / d.nq; d.q[1] = rq.y / d.nq; d.q[2] = rq.z / d.nq; d.q[3] = rq.w / d.nq;
    d.nqh = sqrtf(d.q[0] * d.q[0] + d.q[1] * d.q[1] + d.q[2] * d.q[2] + d.q[3] * d.q[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) d.qh[k] = d.q[k] / d.nqh;
    d.axis = 0;
    if (d.s[1] < d.s[d.axis]) d.axis = 1;       // first minimum, like torch.argmin
    if (d.s[2] < d.s[d.axis]) d.axis = 2;
    column_of_R(d.qh, d.axis, d.col);
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
    return d;
}

__global__ void __launch_bounds__(256) pack_forward_kernel(PackIn in, float* __restrict__ scales, float* __restrict__ rotations,
                                                           float* __restrict__ opacities, float* __restrict__ features) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= in.P) return;
    const Derived d = derive(in, i);
#pragma unroll
    for (int k = 0; k < 3; ++k) scales[3 * (size_t)i + k] = d.s[k];
    *reinterpret_cast<float4*>(rotations + 4 * (size_t)i) = make_float4(d.q[0], d.q[1], d.q[2], d.q[3]);
    opacities[i] = sigmoidf(in.opacity[i]);
    float2* f = reinterpret_cast<float2*>(features + GS2M_NUM_FEATURES * (size_t)i);
    f[0] = make_float2(1.0f, in.z_depth ? d.camp[2] : fabsf(d.u));
    f[1] = make_float2(d.n[0], d.n[1]);
    f[2] = make_float2(d.n[2], sigmoidf(in.albedo[3 * (size_t)i]));
    f[3] = make_float2(sigmoidf(in.albedo[3 * (size_t)i + 1]), sigmoidf(in.albedo[3 * (size_t)i + 2]));
    f[4] = make_float2(sigmoidf(in.roughness[i]), in.blend_metallic ? sigmoidf(in.metallic[i]) : 0.0f);
}

struct PackGrads {
    const float *g_scales, *g_rotations, *g_opacities, *g_features;   // upstream (w.r.t. the activated tensors / features)
    float *d_xyz, *d_scaling, *d_rotation, *d_opacity, *d_albedo, *d_roughness, *d_metallic;
};

// ACC: the raw-parameter gradients are updated with += (several views of a data-parallel step, each chained through its own
// camera); Gaussians the view culled (radii == 0: all upstream gradients are zero) are skipped then.
template <bool ACC>
__device__ __forceinline__ void emit(float* p, float v) { if (ACC) *p += v; else *p = v; }

template <bool ACC>
__global__ void __launch_bounds__(256) pack_backward_kernel(PackIn in, PackGrads g, const int* __restrict__ radii) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= in.P) return;
    if (ACC && radii != nullptr && radii[i] <= 0) return;
    const Derived d = derive(in, i);
    const float* W = in.wvt;
    float gf[GS2M_NUM_FEATURES];
    {
        const float2* f = reinterpret_cast<const float2*>(g.g_features + GS2M_NUM_FEATURES * (size_t)i);
#pragma unroll
        for (int k = 0; k < 5; ++k) { const float2 t = f[k]; gf[2 * k] = t.x; gf[2 * k + 1] = t.y; }
    }
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
    const float ndn = d.n[0] * dn[0] + d.n[1] * dn[1] + d.n[2] * dn[2];
    float dc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dm = (dn[k] - d.n[k] * ndn) / d.nm;
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
    const float4 gq = *reinterpret_cast<const float4*>(g.g_rotations + 4 * (size_t)i);
    float dq[4] = {(dqh[0] - d.qh[0] * qd) / d.nqh + gq.x, (dqh[1] - d.qh[1] * qd) / d.nqh + gq.y,
                   (dqh[2] - d.qh[2] * qd) / d.nqh + gq.z, (dqh[3] - d.qh[3] * qd) / d.nqh + gq.w};
    const float qq = d.q[0] * dq[0] + d.q[1] * dq[1] + d.q[2] * dq[2] + d.q[3] * dq[3];
    const bool clamped = d.nq <= 1e-12f;   // F.normalize divides by the clamped norm: no projection term then
    float4 dr = make_float4((dq[0] - (clamped ? 0.f : d.q[0] * qq)) / d.nq, (dq[1] - (clamped ? 0.f : d.q[1] * qq)) / d.nq,
                            (dq[2] - (clamped ? 0.f : d.q[2] * qq)) / d.nq, (dq[3] - (clamped ? 0.f : d.q[3] * qq)) / d.nq);
    float4* o4 = reinterpret_cast<float4*>(g.d_rotation + 4 * (size_t)i);
    if (ACC) { const float4 t = *o4; dr.x += t.x; dr.y += t.y; dr.z += t.z; dr.w += t.w; }
    *o4 = dr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        emit<ACC>(g.d_xyz + 3 * (size_t)i + k, dp[k]);
        emit<ACC>(g.d_scaling + 3 * (size_t)i + k, g.g_scales[3 * (size_t)i + k] * d.s[k]);
        const float a = sigmoidf(in.albedo[3 * (size_t)i + k]);
        emit<ACC>(g.d_albedo + 3 * (size_t)i + k, gf[5 + k] * a * (1.0f - a));
    }
    const float o = sigmoidf(in.opacity[i]);
    emit<ACC>(g.d_opacity + i, g.g_opacities[i] * o * (1.0f - o));
    const float ro = sigmoidf(in.roughness[i]);
    emit<ACC>(g.d_roughness + i, gf[8] * ro * (1.0f - ro));
    const float me = sigmoidf(in.metallic[i]);
    emit<ACC>(g.d_metallic + i, in.blend_metallic ? gf[9] * me * (1.0f - me) : 0.0f);
}

}  // namespace
}  // namespace gs2m

using namespace gs2m;

extern "C" {

int gs2m_pack_forward(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                      const float* albedo_raw, const float* roughness_raw, const float* metallic_raw,
                      const float* world_view_transform, const float* campos, int z_depth, int blend_metallic,
                      float* scales, float* rotations, float* opacities, float* features, void* stream) {
    if (P < 0) { set_error("pack_forward: negative P"); return GS2M_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GS2M_OK;
    if (!xyz || !scaling_raw || !rotation_raw || !opacity_raw || !albedo_raw || !roughness_raw || !metallic_raw ||
        !world_view_transform || !campos || !scales || !rotations || !opacities || !features) {
        set_error("pack_forward: NULL pointer"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    PackIn in{P, xyz, scaling_raw, rotation_raw, opacity_raw, albedo_raw, roughness_raw, metallic_raw, world_view_transform,
              campos, z_depth, blend_metallic};
    count_launches(1);
    pack_forward_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, scales, rotations, opacities, features);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

static int pack_backward_impl(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                       const float* albedo_raw, const float* roughness_raw, const float* metallic_raw,
                       const float* world_view_transform, const float* campos, int z_depth, int blend_metallic,
                       const float* dL_dscales, const float* dL_drotations, const float* dL_dopacities, const float* dL_dfeatures,
                       float* d_xyz, float* d_scaling_raw, float* d_rotation_raw, float* d_opacity_raw, float* d_albedo_raw,
                       float* d_roughness_raw, float* d_metallic_raw, bool accumulate, const int* radii, void* stream) {
    if (P < 0) { set_error("pack_backward: negative P"); return GS2M_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GS2M_OK;
    if (!xyz || !scaling_raw || !rotation_raw || !opacity_raw || !albedo_raw || !roughness_raw || !metallic_raw ||
        !world_view_transform || !campos || !dL_dscales || !dL_drotations || !dL_dopacities || !dL_dfeatures || !d_xyz ||
        !d_scaling_raw || !d_rotation_raw || !d_opacity_raw || !d_albedo_raw || !d_roughness_raw || !d_metallic_raw) {
        set_error("pack_backward: NULL pointer"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    PackIn in{P, xyz, scaling_raw, rotation_raw, opacity_raw, albedo_raw, roughness_raw, metallic_raw, world_view_transform,
              campos, z_depth, blend_metallic};
    PackGrads g{dL_dscales, dL_drotations, dL_dopacities, dL_dfeatures, d_xyz, d_scaling_raw, d_rotation_raw, d_opacity_raw,
                d_albedo_raw, d_roughness_raw, d_metallic_raw};
    count_launches(1);
    if (accumulate) pack_backward_kernel<true><<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, g, radii);
    else pack_backward_kernel<false><<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, g, nullptr);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int gs2m_pack_backward(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                       const float* albedo_raw, const float* roughness_raw, const float* metallic_raw,
                       const float* world_view_transform, const float* campos, int z_depth, int blend_metallic,
                       const float* dL_dscales, const float* dL_drotations, const float* dL_dopacities, const float* dL_dfeatures,
                       float* d_xyz, float* d_scaling_raw, float* d_rotation_raw, float* d_opacity_raw, float* d_albedo_raw,
                       float* d_roughness_raw, float* d_metallic_raw, void* stream) {
    return pack_backward_impl(P, xyz, scaling_raw, rotation_raw, opacity_raw, albedo_raw, roughness_raw, metallic_raw,
                              world_view_transform, campos, z_depth, blend_metallic, dL_dscales, dL_drotations, dL_dopacities,
                              dL_dfeatures, d_xyz, d_scaling_raw, d_rotation_raw, d_opacity_raw, d_albedo_raw, d_roughness_raw,
                              d_metallic_raw, false, nullptr, stream);
}

int gs2m_pack_backward_accumulate(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw,
                                  const float* opacity_raw, const float* albedo_raw, const float* roughness_raw,
                                  const float* metallic_raw, const float* world_view_transform, const float* campos, int z_depth,
                                  int blend_metallic, const float* dL_dscales, const float* dL_drotations,
                                  const float* dL_dopacities, const float* dL_dfeatures, float* d_xyz, float* d_scaling_raw,
                                  float* d_rotation_raw, float* d_opacity_raw, float* d_albedo_raw, float* d_roughness_raw,
                                  float* d_metallic_raw, const int* radii, void* stream) {
    return pack_backward_impl(P, xyz, scaling_raw, rotation_raw, opacity_raw, albedo_raw, roughness_raw, metallic_raw,
                              world_view_transform, campos, z_depth, blend_metallic, dL_dscales, dL_drotations, dL_dopacities,
                              dL_dfeatures, d_xyz, d_scaling_raw, d_rotation_raw, d_opacity_raw, d_albedo_raw, d_roughness_raw,
                              d_metallic_raw, true, radii, stream);
}

}  // extern "C"
