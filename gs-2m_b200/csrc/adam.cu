// Caller-side stage behind the gradient all-reduce: one Adam step over all of GS-2M's parameter
// groups in a single launch.
//
// Behavioural reference: torch.optim.Adam(l, lr=0.0, eps=1e-15) over nine groups with their own learning rates
// (scene/gaussian_model.py:230-242), default betas, no weight decay, no amsgrad — per element:
//   exp_avg += (g - exp_avg) * (1 - beta1);  exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) * g * g;
//   param -= (lr / (1 - beta1^t)) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps)
// A group's gradient may be a column slice of a wider row-major matrix (`grad_row_stride`, `grad_col_offset`): the view-sharded
// step keeps dL/d(features_dc) and dL/d(features_rest) side by side in one (P, M, 3) block, like `get_features` concatenates them.
#include "common.cuh"

namespace gs2m {
namespace {

constexpr int ADAM_MAX_GROUPS = 16;

struct AdamGroups {
    gs2m_adam_group g[ADAM_MAX_GROUPS];
    long long end[ADAM_MAX_GROUPS];   // exclusive prefix end of each group's work units
    unsigned char vec[ADAM_MAX_GROUPS];       // units of 4 elements, 128-bit state access
    unsigned char grad_vec[ADAM_MAX_GROUPS];  // the (contiguous) gradient is 16-byte aligned too
    int n;
};

__device__ __forceinline__ void adam_update(float g, float& m, float& v, float& p, float beta2, float one_minus_beta1,
                                            float one_minus_beta2, float eps, float step_size, float inv_sqrt_bias2) {
    m = fmaf(g - m, one_minus_beta1, m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(one_minus_beta2 * g, g, v * beta2);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    p -= step_size * (m / (sqrtf(v) * inv_sqrt_bias2 + eps));
}

// Work is counted in units: a group whose element count is a multiple of 4 and whose state is 16-byte aligned is walked four
// elements at a time with 128-bit loads/stores of param / exp_avg / exp_avg_sq (its gradient too when it is contiguous; a
// column-slice gradient is gathered element by element); other groups fall back to one element per unit.
// IndexT = uint32_t when all groups together have fewer than 2^31 elements (no 64-bit divisions in the loop).
template <typename IndexT>
__global__ void __launch_bounds__(256) adam_step_kernel(AdamGroups gs, long long total_units, float beta2, float one_minus_beta1,
                                                        float one_minus_beta2, float eps, float inv_bias1, float inv_sqrt_bias2) {
    const IndexT stride = (IndexT)gridDim.x * 256, n = (IndexT)total_units;
    for (IndexT u = (IndexT)blockIdx.x * 256 + threadIdx.x; u < n; u += stride) {
        int k = 0;
        while (k + 1 < gs.n && (long long)u >= gs.end[k]) ++k;
        const gs2m_adam_group& G = gs.g[k];
        const IndexT lu = u - (IndexT)(k ? gs.end[k - 1] : 0);
        const float step_size = G.lr * inv_bias1;
        const bool dense = G.grad_row_stride == G.width && G.grad_col_offset == 0;
        if (gs.vec[k]) {
            const IndexT local = lu * 4;
            float4 p = reinterpret_cast<float4*>(G.param)[lu];
            float4 m = reinterpret_cast<float4*>(G.exp_avg)[lu];
            float4 v = reinterpret_cast<float4*>(G.exp_avg_sq)[lu];
            float g[4];
            if (dense && gs.grad_vec[k]) {
                const float4 t = reinterpret_cast<const float4*>(G.grad)[lu];
                g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
            } else {
                IndexT row = local / (IndexT)G.width;
                int col = (int)(local - row * (IndexT)G.width);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    g[i] = G.grad[(size_t)row * G.grad_row_stride + G.grad_col_offset + col];
                    if (++col == G.width) { col = 0; ++row; }
                }
            }
            adam_update(g[0], m.x, v.x, p.x, beta2, one_minus_beta1, one_minus_beta2, eps, step_size, inv_sqrt_bias2);
            adam_update(g[1], m.y, v.y, p.y, beta2, one_minus_beta1, one_minus_beta2, eps, step_size, inv_sqrt_bias2);
            adam_update(g[2], m.z, v.z, p.z, beta2, one_minus_beta1, one_minus_beta2, eps, step_size, inv_sqrt_bias2);
            adam_update(g[3], m.w, v.w, p.w, beta2, one_minus_beta1, one_minus_beta2, eps, step_size, inv_sqrt_bias2);
            reinterpret_cast<float4*>(G.param)[lu] = p;
            reinterpret_cast<float4*>(G.exp_avg)[lu] = m;
            reinterpret_cast<float4*>(G.exp_avg_sq)[lu] = v;
        } else {
            size_t gi = lu;
            if (!dense) {
                const IndexT row = lu / (IndexT)G.width;
                gi = (size_t)row * G.grad_row_stride + G.grad_col_offset + (lu - row * (IndexT)G.width);
            }
            float m = G.exp_avg[lu], v = G.exp_avg_sq[lu], p = G.param[lu];
            adam_update(G.grad[gi], m, v, p, beta2, one_minus_beta1, one_minus_beta2, eps, step_size, inv_sqrt_bias2);
            G.exp_avg[lu] = m; G.exp_avg_sq[lu] = v; G.param[lu] = p;
        }
    }
}

}  // namespace
}  // namespace gs2m

using namespace gs2m;

extern "C" int gs2m_adam_step(const gs2m_adam_group* groups, int n_groups, int step, double beta1, double beta2, double eps,
                              void* stream) {
    if (!groups || n_groups <= 0 || n_groups > ADAM_MAX_GROUPS || step < 1) {
        set_error("adam_step: need 1..%d groups and step >= 1", ADAM_MAX_GROUPS); return GS2M_ERR_INVALID_ARGUMENT;
    }
    AdamGroups gs;
    long long total = 0, total_elems = 0;
    for (int k = 0; k < n_groups; ++k) {
        const gs2m_adam_group& G = groups[k];
        if (!G.param || !G.exp_avg || !G.exp_avg_sq || !G.grad || G.rows < 0 || G.width <= 0 || G.grad_col_offset < 0 ||
            G.grad_row_stride < G.grad_col_offset + G.width) {
            set_error("adam_step: group %d is malformed", k); return GS2M_ERR_INVALID_ARGUMENT;
        }
        gs.g[k] = G;
        const long long numel = G.rows * G.width;
        auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        gs.vec[k] = (numel % 4 == 0) && aligned(G.param) && aligned(G.exp_avg) && aligned(G.exp_avg_sq);
        gs.grad_vec[k] = aligned(G.grad);
        total_elems += numel;
        total += gs.vec[k] ? numel / 4 : numel;
        gs.end[k] = total;
    }
    gs.n = n_groups;
    if (total == 0) return GS2M_OK;
    // scalars are formed in double like the Python floats of torch.optim.Adam and rounded to fp32 once
    const double bias1 = 1.0 - pow(beta1, (double)step), bias2 = 1.0 - pow(beta2, (double)step);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    count_launches(1);
    if (total_elems < (1ll << 31))
        adam_step_kernel<uint32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            gs, total, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)(1.0 / bias1),
            (float)(1.0 / sqrt(bias2)));
    else
        adam_step_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
            gs, total, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)(1.0 / bias1),
            (float)(1.0 / sqrt(bias2)));
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}
