// Caller-side stage in front of the rasterizer: parameter activations and the packing of the
// 10-column `features` tensor, forward and backward, one thread per Gaussian, one kernel each.
//
// Behavioural reference (all PyTorch eager ops + autograd in GS-2M, ~15 kernels forward and ~40 backward per render call):
//   scene/gaussian_model.py:113-172  getters: exp(_scaling), normalize(_rotation), sigmoid(_opacity/_albedo/_roughness/_metallic)
//   scene/gaussian_model.py:146-160  get_normals: thinnest axis of R(q) (first minimum), flipped to face the camera, normalised
//   utils/general_utils.py:72-92     build_rotation (normalises the quaternion once more)
//   gaussian_renderer/__init__.py:82-96  cam_normals / cam_points, features = [1, |n_cam . p_cam| or z, n(3), albedo(3), rough, metal]
// The backward is the analytic adjoint of exactly that graph (argmin and the flip test carry no gradient, |.| uses sign()).
#include "pack_math.cuh"

namespace gs2m {
namespace {

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
    float gf[GS2M_NUM_FEATURES];
    {
        const float2* f = reinterpret_cast<const float2*>(g.g_features + GS2M_NUM_FEATURES * (size_t)i);
#pragma unroll
        for (int k = 0; k < 5; ++k) { const float2 t = f[k]; gf[2 * k] = t.x; gf[2 * k + 1] = t.y; }
    }
    const float gs[3] = {g.g_scales[3 * (size_t)i], g.g_scales[3 * (size_t)i + 1], g.g_scales[3 * (size_t)i + 2]};
    const float4 gq = *reinterpret_cast<const float4*>(g.g_rotations + 4 * (size_t)i);
    const RawGrads r = pack_chain(in, i, d, gs, gq, g.g_opacities[i], gf);
    float4 dr = r.drot;
    float4* o4 = reinterpret_cast<float4*>(g.d_rotation + 4 * (size_t)i);
    if (ACC) { const float4 t = *o4; dr.x += t.x; dr.y += t.y; dr.z += t.z; dr.w += t.w; }
    *o4 = dr;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        emit<ACC>(g.d_xyz + 3 * (size_t)i + k, r.dp[k]);
        emit<ACC>(g.d_scaling + 3 * (size_t)i + k, r.dscaling[k]);
        emit<ACC>(g.d_albedo + 3 * (size_t)i + k, r.dalbedo[k]);
    }
    emit<ACC>(g.d_opacity + i, r.dopacity);
    emit<ACC>(g.d_roughness + i, r.droughness);
    emit<ACC>(g.d_metallic + i, r.dmetallic);
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
