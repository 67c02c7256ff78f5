/*
 * gs2m_rasterizer.h — C-ABI of the B200-native (sm_100a) tile-based differentiable Gaussian rasterizer.
 *
 * Drop-in boundary for ndming/GS-2M's `submodules/diff-gaussian-rasterization`: each entry point replaces one
 * member of the reference's native interface `CudaRasterizer::Rasterizer` (cuda_rasterizer/rasterizer.h:18-91),
 * which the reference binds through pybind11 (ext.cpp:15-18 -> rasterize_points.cu:30-219).  Signatures use plain
 * pointers, sizes and an opaque `cudaStream_t`; no torch / C++ types cross this boundary.  All pointers are DEVICE
 * pointers to contiguous fp32 / int32 data unless noted; an absent optional input is NULL (the reference passes a
 * null data_ptr for the empty CPU tensors its Python layer substitutes, diff_gaussian_rasterization/__init__.py:192-203).
 *
 * The three growable scratch arenas (geometry / binning / image state) are obtained through caller-supplied resize
 * callbacks, exactly as the reference's `std::function<char*(size_t)>` arguments
 * (cuda_rasterizer/rasterizer.h:29-31, rasterize_points.cu:22-28): the caller owns the memory, the library keeps
 * no state between calls, and backward receives the same three blobs verbatim.  Their internal layout is private to
 * this library (it is NOT the reference's GeometryState/BinningState/ImageState layout); `gs2m_state_view` exposes
 * typed pointers into them for parity tests (sorted keys, tile ranges, n_contrib ...).
 *
 * Return value: >= 0 on success (forward: number of rendered Gaussian/tile instances R, like
 * `Rasterizer::forward`'s return, rasterizer_impl.cu:329), negative `GS2M_ERR_*` on failure;
 * `gs2m_last_error()` gives a message.
 */
#ifndef GS2M_RASTERIZER_H_
#define GS2M_RASTERIZER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS2M_ABI_VERSION 3

/* compile-time constants of the path (reference: cuda_rasterizer/config.h:15-18) */
#define GS2M_NUM_CHANNELS 3
#define GS2M_NUM_FEATURES 10
#define GS2M_TILE_X 16
#define GS2M_TILE_Y 16

enum {
    GS2M_OK = 0,
    GS2M_ERR_INVALID_ARGUMENT = -1, /* bad sizes / missing required pointer / feature_count outside 0..10 */
    GS2M_ERR_ALLOC = -2,            /* a resize callback returned NULL */
    GS2M_ERR_CUDA = -3,             /* CUDA runtime error (message in gs2m_last_error) */
    GS2M_ERR_TOO_LARGE = -4,        /* instance count does not fit the 30-bit sort bookkeeping (counted in 64 bits) */
    GS2M_ERR_CAPACITY = -5,         /* speculative forward: more instances than R_capacity; outputs are invalid, call again
                                       with a larger capacity (gs2m_last_instance_count() gives the count) or in exact mode */
    GS2M_ERR_PREFILTERED = -6       /* `prefiltered` was set but a Gaussian failed the near-plane test; the reference traps the
                                       device here (auxiliary.h:154-160), this library reports a checked error instead */
};

/* Resize callback: make the arena at least `bytes` long and return its (>=128-B aligned) device base pointer.
 * Mirrors `std::function<char*(size_t N)>` of rasterizer.h:29-31. */
typedef char* (*gs2m_resize_fn)(void* user, size_t bytes);

/* ---- forward: replaces CudaRasterizer::Rasterizer::forward (rasterizer.h:28-56, rasterizer_impl.cu:185-330) ---- */
typedef struct gs2m_forward_args {
    gs2m_resize_fn geometry_buffer; void* geometry_user;
    gs2m_resize_fn binning_buffer;  void* binning_user;
    gs2m_resize_fn image_buffer;    void* image_user;
    int P;                    /* number of Gaussians */
    int D;                    /* active SH degree (0..3) */
    int M;                    /* SH coefficients per colour stored in `shs` (row stride = 3*M floats) */
    const float* background;  /* [3] */
    int width, height;
    const float* means3D;       /* [P,3] */
    const float* shs;           /* [P,M,3] or NULL */
    const float* colors_precomp;/* [P,3]  or NULL (exactly one of shs / colors_precomp) */
    const float* opacities;     /* [P,1] */
    const float* scales;        /* [P,3] or NULL */
    float scale_modifier;
    const float* rotations;     /* [P,4] (w,x,y,z), NOT normalised by the kernels, or NULL */
    const float* cov3D_precomp; /* [P,6] or NULL (exactly one of scales+rotations / cov3D_precomp) */
    const float* features;      /* [P,10] side channels or NULL when feature_count == 0 */
    const float* viewmatrix;    /* [4,4] row-major tensor of W2V^T (element (r,c) at 4*c+r) */
    const float* projmatrix;    /* [4,4] row-major tensor of (P*W2V)^T */
    const float* cam_pos;       /* [3] */
    float tan_fovx, tan_fovy;
    int prefiltered;            /* caller promises that no Gaussian is behind the near plane: a culled one is an error
                                   (GS2M_ERR_PREFILTERED; the reference printf()s and __trap()s, auxiliary.h:154-160) */
    int feature_count;          /* 0..10: prefix length of `features` columns that are blended */
    float* out_color;           /* [3,H,W]   fully written (no pre-zeroing needed) */
    int*   out_radii;           /* [P]       fully written */
    int*   out_observe;         /* [P]       fully written */
    float* out_buffer;          /* [10,H,W]  fully written (channels >= feature_count are zero) */
    void*  stream;              /* cudaStream_t */
    /* How the instance count R (known only on the device after the per-Gaussian stage) sizes the binning arena.
     *   R_capacity == 0  exact mode: R is read back (one host sync before the R-dependent launches, like
     *                    rasterizer_impl.cu:269-270) and binning_buffer is asked for gs2m_binning_bytes(R).  Returns R.
     *   R_capacity  > 0  speculative mode: binning_buffer is asked for gs2m_binning_bytes(R_capacity) up front, every
     *                    R-dependent kernel is launched on a grid sized for the capacity and reads the real count from device
     *                    memory, so no launch waits for the host.  With no_wait == 0 the call then waits for the count (which
     *                    finished long before the blend it has just queued) and returns R, or GS2M_ERR_CAPACITY /
     *                    GS2M_ERR_TOO_LARGE / GS2M_ERR_PREFILTERED when the result must be discarded.  With no_wait == 1 the
     *                    host is never touched (the call can be captured into a CUDA graph); it returns R_capacity and the
     *                    caller checks gs2m_state_view.bin_info later.  Backward must be given the same R_capacity. */
    int R_capacity;
    int no_wait;
    int no_backward;            /* 1: inference only — skip preparing (zeroing) the backward accumulator rows of the visible Gaussians */
} gs2m_forward_args;

/* Instance count seen by the calling thread's most recent forward that waited for it (exact mode or no_wait == 0). */
long long gs2m_last_instance_count(void);

int gs2m_rasterize_forward(const gs2m_forward_args* args);

/* Optional tail of the backward for GS-2M's training loop: chain the rasterizer's gradients w.r.t. the ACTIVATED scales /
 * rotations / opacities and the 10 `features` columns straight through the caller-side packing stage of THIS view's camera
 * (gs2m_pack_forward: scene/gaussian_model.py:113-172, gaussian_renderer/__init__.py:82-96) inside the per-Gaussian kernel, and
 * add the result to the gradients of the RAW parameters.  Replaces writing dL_dscale / dL_drot / dL_dopacity / dL_dfeatures /
 * dL_dmeans3D (which may then be NULL, like dL_dmeans2D) followed by gs2m_pack_backward_accumulate.  With accumulate = 0 every
 * element of the seven raw-gradient tensors is overwritten (zeros for culled Gaussians); with accumulate = 2 they are updated
 * with += for the visible Gaussians.  d_xyz receives dL_dmeans3D plus the position term of the distance / depth column. */
typedef struct gs2m_param_chain {
    const float *scaling_raw, *rotation_raw, *opacity_raw, *albedo_raw, *roughness_raw, *metallic_raw;   /* [P,3] [P,4] [P,1] [P,3] [P,1] [P,1] */
    int z_depth, blend_metallic;
    float *d_xyz, *d_scaling_raw, *d_rotation_raw, *d_opacity_raw, *d_albedo_raw, *d_roughness_raw, *d_metallic_raw;
} gs2m_param_chain;

/* ---- backward: replaces CudaRasterizer::Rasterizer::backward (rasterizer.h:58-90, rasterizer_impl.cu:334-438) ---- */
typedef struct gs2m_backward_args {
    int P, D, M, R;             /* R = value returned by forward */
    int R_capacity;             /* the forward's R_capacity (0 = exact mode: the binning arena was sized for R itself) */
    const float* background;
    int width, height;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* features;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    float tan_fovx, tan_fovy;
    const int* radii;           /* [P] as returned by forward */
    char* geometry_buffer;      /* the three arenas filled by forward */
    char* binning_buffer;
    char* image_buffer;
    size_t geometry_bytes, binning_bytes, image_bytes;
    int feature_count;
    const float* grad_color;    /* dL/d out_color  [3,H,W] */
    const float* grad_buffer;   /* dL/d out_buffer [10,H,W] (only the first feature_count planes are read) */
    /* outputs.  accumulate = 0: every element is written (culled Gaussians get zeros).  accumulate = 1: the nine
     * caller-visible gradient tensors are updated with += (view-sharded data parallel step: gradients of several views
     * are summed in place before one all-reduce).  accumulate = 2: only dL_dmeans3D and dL_dsh — gradients of GS-2M's raw,
     * view-independent parameters — are updated with +=; the others are overwritten, because they chain through the
     * caller's view-dependent packing stage (gs2m_pack_backward_accumulate) once per view. */
    float* dL_dmeans2D;         /* [P,4]  (.xy signed, .zw sum of |.|)  rasterize_points.cu:151 */
    float* dL_dconic;           /* [P,4]  scratch-like output (x,y,-,w) rasterize_points.cu:154 */
    float* dL_dopacity;         /* [P,1] */
    float* dL_dcolor;           /* [P,3]  gradient of colors_precomp; may be NULL (not written) when shs are used */
    float* dL_dmeans3D;         /* [P,3] */
    float* dL_dcov3D;           /* [P,6]  gradient of cov3D_precomp; may be NULL (not written) when scales+rotations are used */
    float* dL_dsh;              /* [P,M,3] or NULL when M == 0 */
    float* dL_dscale;           /* [P,3] */
    float* dL_drot;             /* [P,4] */
    float* dL_dfeatures;        /* [P,10] */
    int accumulate;
    void* stream;
    /* The backward has two stages: the reverse blend over the tile lists (fills an internal per-Gaussian accumulator in the
     * geometry arena) and the per-Gaussian stage (accumulator -> the ten gradient tensors).  phase 0 runs both; phase 1 only the
     * blend; phase 2 only the per-Gaussian stage, restricted to Gaussians [row_begin, row_end) when row_end > 0 (row_begin must
     * be a multiple of 256).  A view-sharded step runs phase 1 per view as soon as the view is rendered and defers phase 2,
     * range by range over all of its views, so that the all-reduce of a finished range overlaps the next range's work. */
    int phase;
    int row_begin, row_end;
    int grad_acc_dirty;         /* 0 for the first backward after a forward (which zeroed the internal accumulator rows of the
                                   visible Gaussians); 1 when backward runs again over the same forward state, or when the
                                   forward ran with no_backward: the library then clears the accumulator first */
    /* optional (NULL = off): GS-2M's densification statistics, fused into the per-Gaussian backward so that they see the
     * gradient of THIS view even in accumulate mode (scene/gaussian_model.py:569-573 add_densification_stats with
     * update_filter = radii > 0): accum += |dL_dmeans2D.xy|, accum_abs += |dL_dmeans2D.zw|, denom += 1.  float[P] each,
     * updated atomically (several views may be in flight on different streams). */
    float* densify_grad_accum;
    float* densify_grad_accum_abs;
    float* densify_denom;
    const gs2m_param_chain* chain;   /* optional, see above (needs scales + rotations inputs and M <= 16) */
} gs2m_backward_args;

int gs2m_rasterize_backward(const gs2m_backward_args* args);

/* The per-Gaussian stage (phase 2) of `n_views` backward calls in ONE pass over the Gaussians — an extension for view-sharded
 * callers; the reference has no counterpart (its backward is per view, rasterizer_impl.cu:334-438, and autograd sums the views).
 * `views` is an array of n_views argument blocks, each exactly what gs2m_rasterize_backward would take for that view with
 * phase = 2 and a chain: they must share P / D / M, means3D, shs, dL_dsh, the row range, the stream and the chain block (same raw
 * parameters and the same seven raw-gradient outputs), and use SH colours and scale + rotation inputs.  Every thread owns one
 * Gaussian, walks the views that see it and sums their raw-parameter gradients on chip, so each element of dL_dsh and of the
 * seven chain outputs is written once with the sum over the views (views[0].accumulate = 0: overwritten, zeros for Gaussians no
 * view sees; 2: added to what is there) instead of being read-modify-written once per view.  The per-view optional outputs
 * (dL_dmeans2D, dL_dconic, the densify_* statistics) behave as in the per-view call.  Result = calling gs2m_rasterize_backward
 * for views[0] with that accumulate mode and for the others with accumulate = 2, up to fp32 summation order. */
int gs2m_rasterize_backward_views(const gs2m_backward_args* views, int n_views);

/* Arena sizes the resize callbacks will be asked for (the `required<T>()` of rasterizer_impl.h:25-30). */
size_t gs2m_geometry_bytes(int P);
size_t gs2m_image_bytes(int width, int height);
size_t gs2m_binning_bytes(int R);

/* ---- markVisible: replaces CudaRasterizer::Rasterizer::markVisible (rasterizer.h:21-26, rasterizer_impl.cu:132-143) ---- */
int gs2m_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                      uint8_t* present /* [P] bool */, void* stream);

/* ---- typed view into the opaque arenas (parity tests / debugging only) ---- */
typedef struct gs2m_state_view {
    /* geometry arena, all [P] unless noted; entries of culled Gaussians (radii == 0) are undefined */
    const float*    depths;         /* view-space z */
    const float*    rec_a;          /* [P,4] mean2D.x, mean2D.y, conic.x, conic.y */
    const float*    rec_b;          /* [P,4] conic.z, opacity, 2*ln(255*opacity) footprint threshold, unused */
    const float*    rgb;            /* [P,4] (r,g,b,unused) */
    const float*    cov3D;          /* [P,6] */
    const uint8_t*  clamped;        /* [P,4] (r,g,b,unused) */
    const uint32_t* tiles_touched;
    const uint32_t* point_offsets;  /* inclusive scan of tiles_touched */
    const float*    grad_acc;       /* [P,24] packed backward-blend accumulator (valid after backward) */
    /* binning arena, [R] (pass the arena's capacity as `R` when forward ran in speculative mode) */
    const uint64_t* keys_sorted;    /* (tile << 32) | float_bits(depth) */
    const uint32_t* point_list;     /* sorted Gaussian indices */
    const uint8_t*  masks;          /* footprint mask of every list entry: bit w = warp block w of the tile may blend it */
    const uint32_t* dense_gid;      /* [8,R] for each warp-block position w: point_list compacted (stably) by mask bit w */
    const uint32_t* dense_pos;      /* [8,R] ... and each entry's position inside its tile's list */
    /* image arena */
    const float*    final_T;        /* [H*W] */
    const uint32_t* n_contrib;      /* [H*W] */
    const uint32_t* ranges;         /* [tiles,2] */
    const uint32_t* bin_info;       /* [8] {R, V, flags, R used by the kernels, V used, -, -, -}; flags: GS2M_BIN_* */
    const uint32_t* block_ranges;   /* [tiles,8,2] the slice of dense_gid[w] that is the list of (tile, warp block w) */
    const uint32_t* n_contrib_dense;/* [H*W] n_contrib in the coordinates of the pixel's warp-block list */
} gs2m_state_view;
enum { GS2M_BIN_OVERFLOW = 1, GS2M_BIN_PREFILTERED = 2, GS2M_BIN_TOO_LARGE = 4 };

int gs2m_state_view_get(int P, int width, int height, int R,
                        char* geometry_buffer, char* binning_buffer, char* image_buffer, gs2m_state_view* out);

/* ---- building blocks exported on their own (parity tests of the binning stage; same device code the forward uses) ---- */
/* Stable LSD radix sort of (u64 key, u32 value) pairs on key bits [0, end_bit) — the role cub::DeviceRadixSort::SortPairs
 * plays at rasterizer_impl.cu:291-296.  `temp` must hold gs2m_sort_temp_bytes(n) bytes. keys_in/vals_in are clobbered. */
size_t gs2m_sort_temp_bytes(int n);
int gs2m_sort_pairs_u64(uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out,
                        int n, int end_bit, char* temp, void* stream);
/* Inclusive prefix sum of u32 (cub::DeviceScan::InclusiveSum at rasterizer_impl.cu:265). temp: gs2m_scan_temp_bytes(n). */
size_t gs2m_scan_temp_bytes(int n);
int gs2m_inclusive_sum_u32(const uint32_t* in, uint32_t* out, int n, char* temp, void* stream);

/* ---- caller-side stage in front of the rasterizer ----
 * Fuses what GS-2M does in ~15 (forward) / ~40 (backward) PyTorch kernels per render call: the parameter activations
 * (scene/gaussian_model.py:113-172: exp, normalize, sigmoid), get_normals (:146-160, utils/general_utils.py:72-92) and the
 * packing of the 10-column `features` tensor (gaussian_renderer/__init__.py:82-96).  Inputs are the RAW (pre-activation)
 * parameters [P,3] [P,3] [P,4] [P,1] [P,3] [P,1] [P,1], `world_view_transform` (the 4x4 row-major tensor of W2V^T) and the
 * camera centre.  Outputs: scales [P,3], rotations [P,4], opacities [P,1], features [P,10].  The backward takes the
 * rasterizer's gradients w.r.t. those four tensors and returns gradients w.r.t. the raw parameters; `d_xyz` is the extra
 * term through the distance/depth column, to be added to the rasterizer's dL/dmeans3D. */
int gs2m_pack_forward(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                      const float* albedo_raw, const float* roughness_raw, const float* metallic_raw,
                      const float* world_view_transform, const float* campos, int z_depth, int blend_metallic,
                      float* scales, float* rotations, float* opacities, float* features, void* stream);
int gs2m_pack_backward(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw, const float* opacity_raw,
                       const float* albedo_raw, const float* roughness_raw, const float* metallic_raw,
                       const float* world_view_transform, const float* campos, int z_depth, int blend_metallic,
                       const float* dL_dscales, const float* dL_drotations, const float* dL_dopacities, const float* dL_dfeatures,
                       float* d_xyz, float* d_scaling_raw, float* d_rotation_raw, float* d_opacity_raw, float* d_albedo_raw,
                       float* d_roughness_raw, float* d_metallic_raw, void* stream);

/* Same, but the seven raw-parameter gradients are updated with += (the view-sharded step chains every view through its own
 * camera and sums raw-parameter gradients); `radii` (optional, as returned by the forward) lets the
 * kernel skip the Gaussians the view culled, whose upstream gradients are all zero. */
int gs2m_pack_backward_accumulate(int P, const float* xyz, const float* scaling_raw, const float* rotation_raw,
                                  const float* opacity_raw, const float* albedo_raw, const float* roughness_raw,
                                  const float* metallic_raw, const float* world_view_transform, const float* campos, int z_depth,
                                  int blend_metallic, const float* dL_dscales, const float* dL_drotations,
                                  const float* dL_dopacities, const float* dL_dfeatures, float* d_xyz, float* d_scaling_raw,
                                  float* d_rotation_raw, float* d_opacity_raw, float* d_albedo_raw, float* d_roughness_raw,
                                  float* d_metallic_raw, const int* radii, void* stream);

/* ---- caller-side stage behind the rasterizer ----
 * Per-pixel maps GS-2M derives from the blended buffer with ~10 PyTorch kernels (gaussian_renderer/__init__.py:125-141,
 * scene/cameras.py:71-81): local_normal_map[3,H,W] = buffer[2:5] rotated into camera space, normal_mask[H,W] (all three
 * normal channels non-zero), depth_map[1,H,W] = distance / -(local_normal . ray + 1e-8) with ray = ((x-cx)/fx, (y-cy)/fy, 1)
 * (or buffer[1] itself when z_depth).  The backward returns dL/dbuffer [10,H,W] (zero on the channels without a path). */
int gs2m_postblend_forward(int width, int height, float fx, float fy, float cx, float cy, int z_depth,
                           const float* world_view_transform, const float* buffer, float* local_normal_map, float* depth_map,
                           uint8_t* normal_mask, void* stream);
int gs2m_postblend_backward(int width, int height, float fx, float fy, float cx, float cy, int z_depth,
                            const float* world_view_transform, const float* buffer, const float* dL_dlocal_normal_map,
                            const float* dL_ddepth_map, float* dL_dbuffer, void* stream);

/* Normal map from the depth map (the "sobel" normal of the geometry stage; gaussian_renderer/__init__.py:163-175,
 * utils/normal_utils.py:30-85): back-projection with the pinhole intrinsics, normalize(cross(right - left, top - bottom)) of
 * the four neighbours in world space, zero on the one-pixel border, composited over `bg` ([3], device) with alpha_map.
 * depth_map / alpha_map are [H,W], sobel_map is [3,H,W].  The backward returns dL/ddepth_map and dL/dalpha_map. */
int gs2m_sobel_normal_forward(int width, int height, float fx, float fy, float cx, float cy, const float* world_view_transform,
                              const float* bg, const float* depth_map, const float* alpha_map, float* sobel_map, void* stream);
int gs2m_sobel_normal_backward(int width, int height, float fx, float fy, float cx, float cy, const float* world_view_transform,
                               const float* bg, const float* depth_map, const float* alpha_map, const float* dL_dsobel_map,
                               float* dL_ddepth_map, float* dL_dalpha_map, void* stream);

/* ---- photometric loss on the rendered image and its gradient ----
 * Lrgb = (1 - lambda) * mean|render - gt| + lambda * (1 - mean SSIM(render, gt))  (train.py:102-107; utils/loss_utils.py:24-25
 * and :30-70: 11x11 Gaussian window, sigma 1.5, zero "same" padding, per channel, C1 = 0.01^2, C2 = 0.03^2 — what the
 * fused-ssim submodule computes).  Images are [channels,H,W].  forward: sums[0] += sum|render - gt|, sums[1] += sum SSIM (the
 * caller zeroes `sums` and forms the means), and the three [channels,H,W] derivative maps the backward needs.  backward:
 * dL_drender = upstream * dLrgb/drender — directly the `grad_color` of gs2m_rasterize_backward.  The upstream gradient is
 * `upstream` times, when `upstream_device` is not NULL, the float it points to in device memory (an autograd caller passes
 * its incoming gradient tensor there and never reads it on the host). */
int gs2m_photometric_loss_forward(int channels, int height, int width, const float* render, const float* gt, float* dm_dE1,
                                  float* dm_dE11, float* dm_dE12, float* sums, void* stream);
int gs2m_photometric_loss_backward(int channels, int height, int width, const float* render, const float* gt, const float* dm_dE1,
                                   const float* dm_dE11, const float* dm_dE12, float lambda_ssim, float upstream,
                                   const float* upstream_device, float* dL_drender, void* stream);

/* ---- one Adam step over all parameter groups in a single launch ----
 * torch.optim.Adam(groups, eps=1e-15) of scene/gaussian_model.py:230-242: default betas, per-group learning rate, no weight decay,
 * no amsgrad.  param / exp_avg / exp_avg_sq are contiguous [rows, width]; the gradient of a group may be a column slice of a wider
 * row-major matrix (element (r, c) at grad[r * grad_row_stride + grad_col_offset + c]).  `step` counts from 1.  At most 16 groups. */
typedef struct gs2m_adam_group {
    float* param;
    float* exp_avg;
    float* exp_avg_sq;
    const float* grad;
    long long rows;
    int width;
    int grad_row_stride;
    int grad_col_offset;
    float lr;
} gs2m_adam_group;
int gs2m_adam_step(const gs2m_adam_group* groups, int n_groups, int step, double beta1, double beta2, double eps, void* stream);

/* ---- per-view densification statistics of the forward outputs ----
 * train.py:225-228: mask = (observe > 0) & (radii > 0); max_radii2D = where(mask, max(max_radii2D, radii), max_radii2D);
 * train.py:238-241 (multi-view trim): observe_cnt[observe > 0] += 1.  float[P] each (the reference keeps both as float
 * tensors); either may be NULL; updated atomically. */
int gs2m_view_stats_update(int P, const int* radii, const int* observe, float* max_radii2D, float* observe_cnt, void* stream);

/* ---- per-stage device timing (bench.py's roofline leg) ----
 * When enabled, forward/backward bracket every stage with cudaEvents on the launching stream.  gs2m_profile_read
 * synchronises on the recorded events, ADDS the elapsed milliseconds and launch counts of all completed calls since
 * the last read into the caller's arrays (indexed by the enum below) and clears the pending list.  Disabled by
 * default (zero overhead). */
enum {
    GS2M_STAGE_PREPROCESS_FWD = 0, GS2M_STAGE_SCAN, GS2M_STAGE_DUPLICATE, GS2M_STAGE_SORT, GS2M_STAGE_RANGES,
    GS2M_STAGE_BLEND_FWD, GS2M_STAGE_BLEND_BWD, GS2M_STAGE_PREPROCESS_BWD, GS2M_NUM_STAGES
};
void gs2m_profile_enable(int enable);
int gs2m_profile_read(float* stage_ms /* [GS2M_NUM_STAGES] */, int* stage_calls /* [GS2M_NUM_STAGES] */);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
long long gs2m_launch_count(void);

const char* gs2m_last_error(void);
int gs2m_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GS2M_RASTERIZER_H_ */
