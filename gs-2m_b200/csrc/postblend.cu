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

// ---- normal map from the depth map ("sobel" normal; gaussian_renderer/__init__.py:163-175, utils/normal_utils.py:30-85) ----
// Every pixel is back-projected with the pinhole intrinsics, cam(x,y) = depth * ((x-cx)/fx, (y-cy)/fy, 1); an interior pixel's
// normal is normalize(cross(right - left, top - bottom)) of its four neighbours' points, rotated to world space (the camera
// translation cancels in the differences and a rotation commutes with the cross product); border pixels get 0; the result is
// composited over the background with alpha.
struct SobelIn {
    int W, H;
    float fx, fy, cx, cy;
    const float* bg;      // [3] device
    const float* wvt;     // world_view_transform (4x4 row-major); its upper-left 3x3 maps camera -> world for column vectors
    const float* depth;   // [H,W]
    const float* alpha;   // [H,W]
};

struct SobelPoint { float n[3], a[3], b[3], len; };   // camera-space unit normal, the two difference vectors, |a x b|

__device__ __forceinline__ void cam_point(const SobelIn& in, int x, int y, float p[3]) {
    const float z = in.depth[(size_t)y * in.W + x];
    // depth2point_cam: ndc = x / (W-1) scaled back by (W-1), times z, then K^-1
    p[0] = (((float)x / (float)(in.W - 1)) * (float)(in.W - 1) * z - in.cx * z) / in.fx;
    p[1] = (((float)y / (float)(in.H - 1)) * (float)(in.H - 1) * z - in.cy * z) / in.fy;
    p[2] = z;
}

__device__ __forceinline__ void sobel_point(const SobelIn& in, int x, int y, SobelPoint& s) {
    float l[3], r[3], t[3], bo[3];
    cam_point(in, x - 1, y, l); cam_point(in, x + 1, y, r); cam_point(in, x, y - 1, t); cam_point(in, x, y + 1, bo);
#pragma unroll
    for (int k = 0; k < 3; ++k) { s.a[k] = r[k] - l[k]; s.b[k] = t[k] - bo[k]; }
    const float v[3] = {s.a[1] * s.b[2] - s.a[2] * s.b[1], s.a[2] * s.b[0] - s.a[0] * s.b[2], s.a[0] * s.b[1] - s.a[1] * s.b[0]};
    s.len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const float inv = 1.0f / fmaxf(s.len, 1e-12f);     // F.normalize(p=2, eps=1e-12)
#pragma unroll
    for (int k = 0; k < 3; ++k) s.n[k] = v[k] * inv;
}

__global__ void __launch_bounds__(256) sobel_forward_kernel(SobelIn in, float* __restrict__ out /*[3,H,W]*/) {
    const size_t N = (size_t)in.W * in.H;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / in.W), x = (int)(i - (size_t)y * in.W);
    float nw[3] = {0.f, 0.f, 0.f};
    if (x >= 1 && x < in.W - 1 && y >= 1 && y < in.H - 1) {
        SobelPoint s;
        sobel_point(in, x, y, s);
#pragma unroll
        for (int r = 0; r < 3; ++r) nw[r] = in.wvt[4 * r] * s.n[0] + in.wvt[4 * r + 1] * s.n[1] + in.wvt[4 * r + 2] * s.n[2];
    }
    const float al = in.alpha[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) out[r * N + i] = nw[r] * al + in.bg[r] * (1.0f - al);
}

// d_depth must be zeroed by the launcher: every interior pixel scatters to its four neighbours
__global__ void __launch_bounds__(256) sobel_backward_kernel(SobelIn in, const float* __restrict__ g_out /*[3,H,W]*/,
                                                             float* __restrict__ d_depth, float* __restrict__ d_alpha) {
    const size_t N = (size_t)in.W * in.H;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int y = (int)(i / in.W), x = (int)(i - (size_t)y * in.W);
    const float g[3] = {g_out[i], g_out[N + i], g_out[2 * N + i]};
    float nw[3] = {0.f, 0.f, 0.f};
    if (x >= 1 && x < in.W - 1 && y >= 1 && y < in.H - 1) {
        SobelPoint s;
        sobel_point(in, x, y, s);
#pragma unroll
        for (int r = 0; r < 3; ++r) nw[r] = in.wvt[4 * r] * s.n[0] + in.wvt[4 * r + 1] * s.n[1] + in.wvt[4 * r + 2] * s.n[2];
        const float al = in.alpha[i];
        // world -> camera (transpose of the rotation), then through the normalisation and the cross product
        float dn[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) dn[c] = al * (in.wvt[c] * g[0] + in.wvt[4 + c] * g[1] + in.wvt[8 + c] * g[2]);
        float dv[3] = {0.f, 0.f, 0.f};
        if (s.len > 1e-12f) {
            const float nd = s.n[0] * dn[0] + s.n[1] * dn[1] + s.n[2] * dn[2];
#pragma unroll
            for (int c = 0; c < 3; ++c) dv[c] = (dn[c] - s.n[c] * nd) / s.len;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) dv[c] = dn[c] * 1e12f;     // clamped denominator: v / eps
        }
        // v = a x b:  d_a = b x dv,  d_b = dv x a
        const float da[3] = {s.b[1] * dv[2] - s.b[2] * dv[1], s.b[2] * dv[0] - s.b[0] * dv[2], s.b[0] * dv[1] - s.b[1] * dv[0]};
        const float db[3] = {dv[1] * s.a[2] - dv[2] * s.a[1], dv[2] * s.a[0] - dv[0] * s.a[2], dv[0] * s.a[1] - dv[1] * s.a[0]};
        // cam(q) = depth(q) * ray(q): a = cam(x+1,y) - cam(x-1,y), b = cam(x,y-1) - cam(x,y+1)
        auto scatter = [&](int qx, int qy, const float d[3], float sign) {
            const float rx = (((float)qx / (float)(in.W - 1)) * (float)(in.W - 1) - in.cx) / in.fx;
            const float ry = (((float)qy / (float)(in.H - 1)) * (float)(in.H - 1) - in.cy) / in.fy;
            atomicAdd(d_depth + (size_t)qy * in.W + qx, sign * (rx * d[0] + ry * d[1] + d[2]));
        };
        scatter(x + 1, y, da, 1.f); scatter(x - 1, y, da, -1.f); scatter(x, y - 1, db, 1.f); scatter(x, y + 1, db, -1.f);
    }
    d_alpha[i] = g[0] * (nw[0] - in.bg[0]) + g[1] * (nw[1] - in.bg[1]) + g[2] * (nw[2] - in.bg[2]);
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

int gs2m_sobel_normal_forward(int width, int height, float fx, float fy, float cx, float cy, const float* world_view_transform,
                              const float* bg, const float* depth_map, const float* alpha_map, float* sobel_map, void* stream) {
    if (width < 2 || height < 2 || !world_view_transform || !bg || !depth_map || !alpha_map || !sobel_map) {
        set_error("sobel_normal_forward: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    SobelIn in{width, height, fx, fy, cx, cy, bg, world_view_transform, depth_map, alpha_map};
    const size_t N = (size_t)width * height;
    count_launches(1);
    sobel_forward_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, sobel_map);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int gs2m_sobel_normal_backward(int width, int height, float fx, float fy, float cx, float cy, const float* world_view_transform,
                               const float* bg, const float* depth_map, const float* alpha_map, const float* dL_dsobel_map,
                               float* dL_ddepth_map, float* dL_dalpha_map, void* stream) {
    if (width < 2 || height < 2 || !world_view_transform || !bg || !depth_map || !alpha_map || !dL_dsobel_map || !dL_ddepth_map ||
        !dL_dalpha_map) {
        set_error("sobel_normal_backward: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    SobelIn in{width, height, fx, fy, cx, cy, bg, world_view_transform, depth_map, alpha_map};
    const size_t N = (size_t)width * height;
    GS2M_CUDA(cudaMemsetAsync(dL_ddepth_map, 0, N * sizeof(float), (cudaStream_t)stream));
    count_launches(1);
    sobel_backward_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, dL_dsobel_map, dL_ddepth_map, dL_dalpha_map);
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
