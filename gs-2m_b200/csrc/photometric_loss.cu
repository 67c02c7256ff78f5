// Caller-side stage behind the rasterizer: GS-2M's photometric loss on the rendered image,
//   Lrgb = (1 - lambda) * mean|render - gt|  +  lambda * (1 - mean SSIM(render, gt))          (train.py:102-107)
// and its gradient with respect to the render, which is the `grad_color` the rasterizer backward consumes.
//
// Behavioural reference: utils/loss_utils.py:24-25 (l1_loss) and :30-70 (_ssim: 11x11 Gaussian window, sigma 1.5, zero
// "same" padding, per channel, C1 = 0.01^2, C2 = 0.03^2; the fused-ssim submodule train.py calls computes the same map).
// Two kernels: (1) separable window sums of (a, b, a^2, b^2, ab) per 32x32 tile in shared memory -> SSIM value, the three
// partial derivatives d m / d E[a], d m / d E[a^2], d m / d E[ab], and the two loss sums; (2) the same separable window over
// the three derivative maps and the pointwise combination with the L1 sign term -> dL/d render.
#include "common.cuh"

namespace gs2m {
namespace {

constexpr int LT = 32;            // tile edge
constexpr int LH = LT + 10;       // tile + window halo
constexpr int LTHREADS = 256;

struct Window { float w[11]; };

__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < LTHREADS / 32) t = s_red[threadIdx.x];
    if (warp == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;   // valid in thread 0
}

__global__ void __launch_bounds__(LTHREADS) ssim_l1_forward_kernel(int H, int W, const float* __restrict__ render,
                                                                   const float* __restrict__ gt, Window win, float C1, float C2,
                                                                   float* __restrict__ dm_dE1, float* __restrict__ dm_dE11,
                                                                   float* __restrict__ dm_dE12, float* __restrict__ sums) {
    __shared__ float sa[LH][LH + 1], sb[LH][LH + 1];
    __shared__ float hz[5][LH][LT];
    __shared__ float s_red[LTHREADS / 32];
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const float* a_img = render + (size_t)ch * H * W;
    const float* b_img = gt + (size_t)ch * H * W;
    for (int i = threadIdx.x; i < LH * LH; i += LTHREADS) {
        const int ly = i / LH, lx = i - ly * LH;
        const int y = y0 + ly - 5, x = x0 + lx - 5;
        const bool in = (x >= 0) && (x < W) && (y >= 0) && (y < H);
        sa[ly][lx] = in ? a_img[(size_t)y * W + x] : 0.f;
        sb[ly][lx] = in ? b_img[(size_t)y * W + x] : 0.f;
    }
    __syncthreads();
    // horizontal window over the LH rows of the halo
    for (int i = threadIdx.x; i < LH * LT; i += LTHREADS) {
        const int ly = i / LT, lx = i - ly * LT;
        float e1 = 0.f, e2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float a = sa[ly][lx + k], b = sb[ly][lx + k], w = win.w[k];
            e1 = fmaf(w, a, e1); e2 = fmaf(w, b, e2);
            e11 = fmaf(w, a * a, e11); e22 = fmaf(w, b * b, e22); e12 = fmaf(w, a * b, e12);
        }
        hz[0][ly][lx] = e1; hz[1][ly][lx] = e2; hz[2][ly][lx] = e11; hz[3][ly][lx] = e22; hz[4][ly][lx] = e12;
    }
    __syncthreads();
    float ssim_sum = 0.f, l1_sum = 0.f;
    const int lx = threadIdx.x & 31;
    for (int ly = threadIdx.x >> 5; ly < LT; ly += LTHREADS / 32) {
        const int x = x0 + lx, y = y0 + ly;
        if (x >= W || y >= H) continue;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = win.w[k];
            mu1 = fmaf(w, hz[0][ly + k][lx], mu1); mu2 = fmaf(w, hz[1][ly + k][lx], mu2);
            e11 = fmaf(w, hz[2][ly + k][lx], e11); e22 = fmaf(w, hz[3][ly + k][lx], e22);
            e12 = fmaf(w, hz[4][ly + k][lx], e12);
        }
        const float s1 = e11 - mu1 * mu1, s2 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        const float A = 2.f * mu1 * mu2 + C1, B = 2.f * s12 + C2;
        const float Cc = mu1 * mu1 + mu2 * mu2 + C1, D = s1 + s2 + C2;
        const float inv_cd = 1.f / (Cc * D);
        const float m = A * B * inv_cd;
        const float dm_ds1 = -m / D;                 // = d m / d E[a^2]
        const float dm_ds12 = 2.f * A * inv_cd;      // = d m / d E[ab]
        // d m / d E[a] with E[a^2], E[ab] held fixed: direct terms through A and Cc, plus s1 = E11 - E1^2, s12 = E12 - E1*mu2
        const float dm_dmu1 = 2.f * mu2 * B * inv_cd - 2.f * mu1 * m / Cc - 2.f * mu1 * dm_ds1 - mu2 * dm_ds12;
        const size_t o = (size_t)ch * H * W + (size_t)y * W + x;
        dm_dE1[o] = dm_dmu1; dm_dE11[o] = dm_ds1; dm_dE12[o] = dm_ds12;
        ssim_sum += m;
        l1_sum += fabsf(sa[ly + 5][lx + 5] - sb[ly + 5][lx + 5]);
    }
    const float t0 = block_sum(l1_sum, s_red);
    const float t1 = block_sum(ssim_sum, s_red);
    if (threadIdx.x == 0) { atomicAdd(sums, t0); atomicAdd(sums + 1, t1); }
}

// dL/d render = g_map * [ win(dm_dE1) + 2 a win(dm_dE11) + b win(dm_dE12) ] + g_l1 * sign(a - b),
// g_map = -lambda * upstream / N (Lssim = 1 - mean m), g_l1 = (1 - lambda) * upstream / N
__global__ void __launch_bounds__(LTHREADS) ssim_l1_backward_kernel(int H, int W, const float* __restrict__ render,
                                                                    const float* __restrict__ gt, Window win,
                                                                    const float* __restrict__ dm_dE1,
                                                                    const float* __restrict__ dm_dE11,
                                                                    const float* __restrict__ dm_dE12, float g_map, float g_l1,
                                                                    const float* __restrict__ upstream_dev,
                                                                    float* __restrict__ dL_drender) {
    if (upstream_dev != nullptr) {      // upstream gradient that lives on the device: no host read in the caller's backward
        const float u = __ldg(upstream_dev);
        g_map *= u; g_l1 *= u;
    }
    __shared__ float sd[3][LH][LH + 1];
    __shared__ float hz[3][LH][LT];
    const int ch = blockIdx.z;
    const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
    const size_t plane = (size_t)ch * H * W;
    for (int i = threadIdx.x; i < LH * LH; i += LTHREADS) {
        const int ly = i / LH, lx = i - ly * LH;
        const int y = y0 + ly - 5, x = x0 + lx - 5;
        const bool in = (x >= 0) && (x < W) && (y >= 0) && (y < H);
        const size_t o = plane + (size_t)y * W + x;
        sd[0][ly][lx] = in ? dm_dE1[o] : 0.f;
        sd[1][ly][lx] = in ? dm_dE11[o] : 0.f;
        sd[2][ly][lx] = in ? dm_dE12[o] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LH * LT; i += LTHREADS) {
        const int ly = i / LT, lx = i - ly * LT;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = win.w[k];
            v0 = fmaf(w, sd[0][ly][lx + k], v0); v1 = fmaf(w, sd[1][ly][lx + k], v1); v2 = fmaf(w, sd[2][ly][lx + k], v2);
        }
        hz[0][ly][lx] = v0; hz[1][ly][lx] = v1; hz[2][ly][lx] = v2;
    }
    __syncthreads();
    const int lx = threadIdx.x & 31;
    for (int ly = threadIdx.x >> 5; ly < LT; ly += LTHREADS / 32) {
        const int x = x0 + lx, y = y0 + ly;
        if (x >= W || y >= H) continue;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = win.w[k];
            v0 = fmaf(w, hz[0][ly + k][lx], v0); v1 = fmaf(w, hz[1][ly + k][lx], v1); v2 = fmaf(w, hz[2][ly + k][lx], v2);
        }
        const size_t o = plane + (size_t)y * W + x;
        const float a = render[o], b = gt[o];
        const float d = a - b;
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);   // torch.abs backward: sign(0) = 0
        dL_drender[o] = g_map * (v0 + 2.f * a * v1 + b * v2) + g_l1 * sgn;
    }
}

Window gaussian_window() {   // utils/loss_utils.py:40-42 (_gaussian(11, 1.5))
    Window win;
    double g[11], s = 0.0;
    for (int i = 0; i < 11; ++i) { g[i] = exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); s += g[i]; }
    for (int i = 0; i < 11; ++i) win.w[i] = (float)(g[i] / s);
    return win;
}

}  // namespace
}  // namespace gs2m

using namespace gs2m;

extern "C" {

int gs2m_photometric_loss_forward(int channels, int height, int width, const float* render, const float* gt, float* dm_dE1,
                                  float* dm_dE11, float* dm_dE12, float* sums /* [2], zeroed by the caller */, void* stream) {
    if (channels <= 0 || height <= 0 || width <= 0 || !render || !gt || !dm_dE1 || !dm_dE11 || !dm_dE12 || !sums) {
        set_error("photometric_loss_forward: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, channels);
    count_launches(1);
    ssim_l1_forward_kernel<<<grid, LTHREADS, 0, (cudaStream_t)stream>>>(height, width, render, gt, gaussian_window(), 0.01f * 0.01f,
                                                                        0.03f * 0.03f, dm_dE1, dm_dE11, dm_dE12, sums);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int gs2m_photometric_loss_backward(int channels, int height, int width, const float* render, const float* gt, const float* dm_dE1,
                                   const float* dm_dE11, const float* dm_dE12, float lambda_ssim, float upstream,
                                   const float* upstream_device, float* dL_drender, void* stream) {
    if (channels <= 0 || height <= 0 || width <= 0 || !render || !gt || !dm_dE1 || !dm_dE11 || !dm_dE12 || !dL_drender) {
        set_error("photometric_loss_backward: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    const dim3 grid((width + LT - 1) / LT, (height + LT - 1) / LT, channels);
    const float inv_n = 1.0f / ((float)channels * (float)height * (float)width);
    count_launches(1);
    ssim_l1_backward_kernel<<<grid, LTHREADS, 0, (cudaStream_t)stream>>>(height, width, render, gt, gaussian_window(), dm_dE1, dm_dE11,
                                                                         dm_dE12, -lambda_ssim * upstream * inv_n,
                                                                         (1.0f - lambda_ssim) * upstream * inv_n, upstream_device,
                                                                         dL_drender);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // extern "C"
