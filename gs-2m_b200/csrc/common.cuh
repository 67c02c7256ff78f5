// Shared device/host helpers of the sm_100a Gaussian rasterizer.
//
// Arena layouts (private to this library, exposed for tests through gs2m_state_view_get):
//   geometry arena  (P)   : GeomState   — per-Gaussian projected records + scan scratch + backward accumulator
//   binning  arena  (R)   : BinState    — unsorted/sorted (key,value) lists + radix-sort scratch
//   image    arena  (W*H) : ImageState  — final transmittance, last-contributor index, tile ranges
// They play the role of GeometryState/BinningState/ImageState of the reference
// (cuda_rasterizer/rasterizer_impl.h:32-63) but use 16-byte records so the blend kernels can stage them with
// 128-bit loads.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stddef.h>
#include "../../include/gs2m_rasterizer.h"

#define GS2M_BLOCK (GS2M_TILE_X * GS2M_TILE_Y)   // 256 pixels per tile
#define GS2M_ACC_STRIDE 24                       // floats per Gaussian in the backward accumulator (21 used)

namespace gs2m {

struct GeomState {
    float*    depths;         // [P]
    float4*   xy_conic_ab;    // [P] (mean2D.x, mean2D.y, conic.x, conic.y)      } the 32-byte "blend record"
    float4*   conic_c_opac;   // [P] (conic.z, opacity, footprint threshold 2 ln(255 o), unused) }
    float4*   rgb;            // [P] (r,g,b,0) SH colour
    float*    cov3D;          // [P,6]
    uint8_t*  clamped;        // [P,4]
    uint32_t* tiles_touched;  // [P]
    uint32_t* point_offsets;  // [P]
    char*     scan_temp;      // gs2m_scan_temp_bytes(P)
    uint32_t* depth_keys;     // [P] } (depth bits, Gaussian index) of the V visible Gaussians: compaction output and the
    uint32_t* depth_keys_alt; // [P] } ping-pong buffers of the depth sort
    uint32_t* order_a;        // [P] }
    uint32_t* order_b;        // [P] }
    char*     rank_temp;      // sort_temp_bytes(P)
    float*    grad_acc;       // [P,GS2M_ACC_STRIDE] backward blend accumulator
    static size_t carve(char* base, int P, GeomState* out);
};

struct BinState {
    uint64_t* keys_unsorted;  // [R]
    uint64_t* keys_sorted;    // [R]
    uint32_t* vals_unsorted;  // [R]
    uint32_t* point_list;     // [R]
    uint8_t*  masks;          // [R] footprint mask of every list entry (bit w: warp block w of the tile may be reached)
    uint32_t* dense_gid;      // [8][R] per warp-block position w: the instance list compacted by mask bit w (Gaussian indices)
    uint32_t* dense_pos;      // [8][R] ... and each entry's position inside its tile's list
    uint32_t* dense_block_totals;   // [8][ceil(R/512)] scan scratch of the compaction
    char*     sort_temp;
    static size_t carve(char* base, int R, BinState* out);
};

struct ImageState {
    float*    final_T;        // [N]
    uint32_t* n_contrib;      // [N]
    uint2*    ranges;         // [tiles]
    uint32_t* tile_order;     // [tiles] tile ids, longest lists first: the order in which the blend kernels' CTAs take the tiles
    uint2*    block_ranges;   // [tiles][8] slice of dense_gid[w] that is the list of (tile, warp block w)
    uint32_t* n_contrib_dense;   // [N] last contributor + 1 in the coordinates of the pixel's warp-block list (for the backward)
    uint32_t* bin_info;       // [8] control block written on the device: BIN_* indices below
    static size_t carve(char* base, int W, int H, ImageState* out);
};

// bin_info words.  R / V are the true counts; R_USED / V_USED are what the count-dependent kernels process: equal to R / V,
// or 0 when the result has to be discarded anyway (R above the arena's capacity or above the 30-bit sort bookkeeping).
enum { BIN_R = 0, BIN_V = 1, BIN_FLAGS = 2, BIN_R_USED = 3, BIN_V_USED = 4, BIN_WORDS = 8 };

template <typename T>
static inline void carve_array(char*& p, T*& ptr, size_t count) {
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 127) & ~uintptr_t(127);
    ptr = reinterpret_cast<T*>(a);
    p = reinterpret_cast<char*>(ptr + count);
}

void set_error(const char* fmt, ...);
void count_launches(int n);                 // bench bookkeeping: kernels launched by this library
struct StageTimer {                         // RAII: cudaEvent pair around one stage when profiling is enabled
    StageTimer(int stage, cudaStream_t s);
    ~StageTimer();
    int stage_; cudaStream_t s_; cudaEvent_t e0_, e1_; bool on_;
};
bool check_cuda(cudaError_t e, const char* what);
#define GS2M_CUDA(call) do { if (!::gs2m::check_cuda((call), #call)) return GS2M_ERR_CUDA; } while (0)

// One-time setup per DEVICE (function attributes such as the dynamic shared-memory opt-in belong to the device's context,
// not to the process): `if (once.need(dev)) { cudaFuncSetAttribute(...); once.done(dev); }` with dev = cudaGetDevice().
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};
    bool need(int& dev) {
        if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        return dev >= 64 || !((mask.load(std::memory_order_acquire) >> dev) & 1ull);
    }
    void done(int dev) { if (dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release); }
};

// ---- stage launchers (each defined in its own .cu) ----
struct FwdParams {
    int P, D, M, W, H, F;
    int tiles_x, tiles_y;
    float tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
    const float *means3D, *shs, *colors_precomp, *opacities, *scales, *rotations, *cov3D_precomp, *features;
    const float *viewmatrix, *projmatrix, *cam_pos, *background;
};

int launch_preprocess_forward(const FwdParams& p, const GeomState& g, int* radii, int* observe, uint32_t* bin_flags,
                              bool prefiltered, bool zero_grad_acc, cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s);
int launch_duplicate_with_keys(int P, const GeomState& g, const int* radii, int tiles_x, int tiles_y,
                               uint64_t* keys, uint32_t* vals, cudaStream_t s);
int launch_identify_tile_ranges(int R, const uint64_t* keys_sorted, uint2* ranges, int n_tiles, cudaStream_t s);
int launch_ranges_and_masks(int R, int tiles_x, int tiles_y, const BinState& b, const GeomState& g, const ImageState& im, cudaStream_t s);
// Count-dependent stages take a capacity (`*_cap`: sizes the grid) and a device pointer to the real count (`n_ptr`, may be
// nullptr = the capacity is the count); the kernels process min(*n_ptr, cap) items.
int launch_ranges_masks_keys(int R_cap, const uint32_t* n_ptr, int tiles_x, int tiles_y, const uint32_t* tile_keys_sorted,
                             const BinState& b, const GeomState& g, const ImageState& im, cudaStream_t s);
// depth-first binning (binning_depthfirst.cu)
size_t compact_temp_bytes(int n);
int binning_df_compact(int P, const GeomState& g, uint32_t* keys, uint32_t* vals, uint32_t* bin_info, uint32_t R_capacity,
                       cudaStream_t s);
int binning_df_emit(int V_cap, const uint32_t* n_ptr, const GeomState& g, const uint32_t* order, const int* radii, int tiles_x,
                    int tiles_y, uint32_t* tile_keys, uint32_t* vals, cudaStream_t s);
int launch_tile_order(int n_tiles, const uint2* ranges, uint32_t* tile_order, cudaStream_t s);
int launch_blend_forward(const FwdParams& p, const GeomState& g, const BinState& b, int R_cap,
                         const ImageState& im, float* out_color, int* out_observe, float* out_buffer, cudaStream_t s);

struct BwdParams {
    int P, D, M, W, H, F, R;
    int tiles_x, tiles_y;
    float tan_fovx, tan_fovy, focal_x, focal_y, scale_modifier;
    const float *means3D, *shs, *colors_precomp, *scales, *rotations, *cov3D_precomp, *features;
    const float *viewmatrix, *projmatrix, *cam_pos, *background;
    const int* radii;
    const float *grad_color, *grad_buffer;
    float *dL_dmeans2D, *dL_dconic, *dL_dopacity, *dL_dcolor, *dL_dmeans3D, *dL_dcov3D, *dL_dsh, *dL_dscale, *dL_drot,
          *dL_dfeatures;
    int accumulate;
    int row_begin, row_end;                                                // Gaussians the per-Gaussian stage covers
    float *densify_grad_accum, *densify_grad_accum_abs, *densify_denom;   // optional (nullptr = off)
    bool has_chain;                                                        // chain the gradients through the packing stage
    gs2m_param_chain chain;
};
int launch_blend_backward(const BwdParams& p, const GeomState& g, const BinState& b, int R_cap, const ImageState& im, cudaStream_t s);
int launch_preprocess_backward(const BwdParams& p, const GeomState& g, cudaStream_t s);
// per-Gaussian stage of n views in one pass (chain mode): outputs written once with the sum over the views (`accumulate`: added)
int launch_preprocess_backward_views(const BwdParams* ps, const GeomState* gs, int n_views, bool accumulate, cudaStream_t s);

size_t sort_temp_bytes(int n);
int sort_pairs_u64(uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n, int end_bit,
                   char* temp, cudaStream_t s);
int sort_pairs_u64_pingpong(uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n,
                            int end_bit, char* temp, cudaStream_t s, int* result_in_input);
int sort_pairs_u32_pingpong(uint32_t* keys_in, uint32_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n_cap,
                            const uint32_t* n_ptr, int end_bit, char* temp, cudaStream_t s, int* result_in_input);
size_t scan_temp_bytes(int n);
int inclusive_sum_u32(const uint32_t* in, uint32_t* out, int n, char* temp, cudaStream_t s);

// ---- small device helpers ----
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// Tile rectangle of a splat (half-open, clamped to the grid). Arithmetic mirrors the reference's getRect
// (cuda_rasterizer/auxiliary.h:44-53) as compiled for sm_100: (p - r) * (1/16) and ((p + r) + 16) - 1) * (1/16),
// truncation toward zero, clamp to [0, grid].
__device__ __forceinline__ void tile_rect(float px, float py, int radius, int tiles_x, int tiles_y,
                                          int& x0, int& y0, int& x1, int& y1) {
    const float r = (float)radius;
    x0 = min(tiles_x, max(0, (int)(__fmul_rn(__fadd_rn(px, -r), 0.0625f))));
    y0 = min(tiles_y, max(0, (int)(__fmul_rn(__fadd_rn(py, -r), 0.0625f))));
    x1 = min(tiles_x, max(0, (int)(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), -1.0f), 0.0625f))));
    y1 = min(tiles_y, max(0, (int)(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), -1.0f), 0.0625f))));
}
#endif

}  // namespace gs2m
