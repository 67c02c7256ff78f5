// Caller-side stage behind the rasterizer: the per-pixel maps GS-2M derives from the blended
// 10-channel buffer, forward and backward, one thread per pixel, one kernel each.
//
// Behavioural reference (PyTorch eager ops + autograd, gaussian_renderer/__init__.py:125-141 and scene/cameras.py:71-81):
//   normal_map  = buffer[2:5];  normal_mask = (normal_map != 0).all(0)
//   local_normal_map = normal_map (as rows) @ world_view_transform[:3,:3]
//   rays = ((x - Cx)/Fx, (y - Cy)/Fy, 1);  depth_map = distance_map / -(sum(local_normals * rays) + 1e-8)   (plane depth)
//   (with pipe.z_depth the depth map is buffer[1] itself)
#include "common.cuh"

namespace gs2m {
namespace {

struct PostIn {
    int W, H;
    float fx, fy, cx, cy;
    int z_depth;
    const float* wvt;       // world_view_transform, 4x4 row-major
    const float* buffer;    // [10,H,W]
};

__global__ void __launch_bounds__(256) postblend_forward_kernel(PostIn in, float* __restrict__ local_normal,
                                                                float* __restrict__ depth, uint8_t* __restrict__ mask) {
    const size_t N = (size_t)in.W * in.H;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / in.W), x = (int)(i - (size_t)y * in.W);
    const float n0 = in.buffer[2 * N + i], n1 = in.buffer[3 * N + i], n2 = in.buffer[4 * N + i];
    const float* W = in.wvt;
    float ln[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ln[k] = n0 * W[k] + n1 * W[4 + k] + n2 * W[8 + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) local_normal[k * N + i] = ln[k];
    mask[i] = (n0 != 0.f) && (n1 != 0.f) && (n2 != 0.f);
    const float dist = in.buffer[N + i];
    if (in.z_depth) {
        depth[i] = dist;
    } else {
        const float rx = ((float)x - in.cx) / in.fx, ry = ((float)y - in.cy) / in.fy;
        const float denom = ln[0] * rx + ln[1] * ry + ln[2];
        depth[i] = dist / -(denom + 1e-8f);
    }
}

__global__ void __launch_bounds__(256) postblend_backward_kernel(PostIn in, const float* __restrict__ g_local_normal,
                                                                 const float* __restrict__ g_depth, float* __restrict__ g_buffer) {
    const size_t N = (size_t)in.W * in.H;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / in.W), x = (int)(i - (size_t)y * in.W);
    const float* W = in.wvt;
    float dln[3] = {g_local_normal[i], g_local_normal[N + i], g_local_normal[2 * N + i]};
    float d_dist;
    const float gd = g_depth[i];
    if (in.z_depth) {
        d_dist = gd;
    } else {
        const float n0 = in.buffer[2 * N + i], n1 = in.buffer[3 * N + i], n2 = in.buffer[4 * N + i];
        float ln[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) ln[k] = n0 * W[k] + n1 * W[4 + k] + n2 * W[8 + k];
        const float rx = ((float)x - in.cx) / in.fx, ry = ((float)y - in.cy) / in.fy;
        const float s = ln[0] * rx + ln[1] * ry + ln[2] + 1e-8f;
        const float dist = in.buffer[N + i];
        d_dist = -gd / s;
        const float d_denom = gd * dist / (s * s);
        dln[0] += d_denom * rx; dln[1] += d_denom * ry; dln[2] += d_denom;
    }
    // buffer channels: 0 alpha (no path), 1 distance, 2..4 normal, 5..9 material (no path)
    g_buffer[i] = 0.f;
    g_buffer[N + i] = d_dist;
#pragma unroll
    for (int j = 0; j < 3; ++j) g_buffer[(2 + j) * N + i] = W[4 * j] * dln[0] + W[4 * j + 1] * dln[1] + W[4 * j + 2] * dln[2];
#pragma unroll
    for (int c = 5; c < GS2M_NUM_FEATURES; ++c) g_buffer[c * N + i] = 0.f;
}

// train.py:225-228 and :238-241 on one view's (radii, observe)
__global__ void __launch_bounds__(256) view_stats_kernel(int P, const int* __restrict__ radii, const int* __restrict__ observe,
                                                         float* __restrict__ max_radii2D, float* __restrict__ observe_cnt) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (observe[i] <= 0) return;
    if (observe_cnt) atomicAdd(observe_cnt + i, 1.0f);
    // non-negative floats order like their bit patterns
    if (max_radii2D && r > 0) atomicMax(reinterpret_cast<int*>(max_radii2D) + i, __float_as_int((float)r));
}

}  // namespace
}  // namespace gs2m

using namespace gs2m;

extern "C" {

int gs2m_postblend_forward(int width, int height, float fx, float fy, float cx, float cy, int z_depth,
                           const float* world_view_transform, const float* buffer, float* local_normal_map, float* depth_map,
                           uint8_t* normal_mask, void* stream) {
    if (width <= 0 || height <= 0 || !world_view_transform || !buffer || !local_normal_map || !depth_map || !normal_mask) {
        set_error("postblend_forward: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    PostIn in{width, height, fx, fy, cx, cy, z_depth, world_view_transform, buffer};
    const size_t N = (size_t)width * height;
    count_launches(1);
    postblend_forward_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, local_normal_map, depth_map, normal_mask);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int gs2m_postblend_backward(int width, int height, float fx, float fy, float cx, float cy, int z_depth,
                            const float* world_view_transform, const float* buffer, const float* dL_dlocal_normal_map,
                            const float* dL_ddepth_map, float* dL_dbuffer, void* stream) {
    if (width <= 0 || height <= 0 || !world_view_transform || !buffer || !dL_dlocal_normal_map || !dL_ddepth_map || !dL_dbuffer) {
        set_error("postblend_backward: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    PostIn in{width, height, fx, fy, cx, cy, z_depth, world_view_transform, buffer};
    const size_t N = (size_t)width * height;
    count_launches(1);
    postblend_backward_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, dL_dlocal_normal_map, dL_ddepth_map, dL_dbuffer);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int gs2m_view_stats_update(int P, const int* radii, const int* observe, float* max_radii2D, float* observe_cnt, void* stream) {
    if (P < 0 || (P > 0 && (!radii || !observe))) { set_error("view_stats_update: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT; }
    if (P == 0 || (!max_radii2D && !observe_cnt)) return GS2M_OK;
    count_launches(1);
    view_stats_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, radii, observe, max_radii2D, observe_cnt);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // extern "C"
