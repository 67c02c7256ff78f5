// C-ABI entry points (include/gs2m_rasterizer.h) and host orchestration of the stages.
//
// Behavioural reference: CudaRasterizer::Rasterizer::{forward,backward,markVisible}
// (cuda_rasterizer/rasterizer_impl.cu:185-330, 334-438, 132-143) and the torch glue's argument checks
// (rasterize_points.cu:52-54,78,161).  Differences that matter to callers: kernels run on the caller's stream
// (the reference uses the legacy default stream), every CUDA call is checked, and outputs need no pre-zeroing.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace gs2m {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

bool check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return false;
}

// ---- profiling / bookkeeping ----
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_profile{0};
struct PendingStage { int stage; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<PendingStage> g_pending;

void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

StageTimer::StageTimer(int stage, cudaStream_t s) : stage_(stage), s_(s), e0_(nullptr), e1_(nullptr), on_(false) {
    if (!g_profile.load(std::memory_order_relaxed)) return;
    if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) return;
    on_ = true;
    cudaEventRecord(e0_, s_);
}
StageTimer::~StageTimer() {
    if (!on_) return;
    cudaEventRecord(e1_, s_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_pending.push_back({stage_, e0_, e1_});
}

size_t GeomState::carve(char* base, int P, GeomState* out) {
    GeomState g;
    char* p = base;
    const size_t n = (size_t)P;
    carve_array(p, g.depths, n);
    carve_array(p, g.xy_conic_ab, n);
    carve_array(p, g.conic_c_opac, n);
    carve_array(p, g.rgb, n);
    carve_array(p, g.cov3D, n * 6);
    carve_array(p, g.clamped, n * 4);
    carve_array(p, g.tiles_touched, n);
    carve_array(p, g.point_offsets, n);
    carve_array(p, g.scan_temp, std::max(scan_temp_bytes(P), compact_temp_bytes(P)));
    carve_array(p, g.depth_keys, n);
    carve_array(p, g.depth_keys_alt, n);
    carve_array(p, g.order_a, n);
    carve_array(p, g.order_b, n);
    carve_array(p, g.rank_temp, sort_temp_bytes(P));
    carve_array(p, g.grad_acc, n * GS2M_ACC_STRIDE);
    if (out) *out = g;
    return (size_t)(p - base) + 128;
}

size_t BinState::carve(char* base, int R, BinState* out) {
    BinState b;
    char* p = base;
    const size_t n = (size_t)(R > 0 ? R : 0);
    carve_array(p, b.keys_unsorted, n);
    carve_array(p, b.keys_sorted, n);
    carve_array(p, b.vals_unsorted, n);
    carve_array(p, b.point_list, n);
    carve_array(p, b.masks, n);
    carve_array(p, b.dense_gid, n * 8);
    carve_array(p, b.dense_pos, n * 8);
    carve_array(p, b.dense_block_totals, ((n + 511) / 512) * 8);
    carve_array(p, b.sort_temp, sort_temp_bytes(R));
    if (out) *out = b;
    return (size_t)(p - base) + 128;
}

size_t ImageState::carve(char* base, int W, int H, ImageState* out) {
    ImageState im;
    char* p = base;
    const size_t n = (size_t)W * H;
    const size_t tiles = (size_t)((W + GS2M_TILE_X - 1) / GS2M_TILE_X) * ((H + GS2M_TILE_Y - 1) / GS2M_TILE_Y);
    carve_array(p, im.final_T, n);
    carve_array(p, im.n_contrib, n);
    carve_array(p, im.ranges, tiles);
    carve_array(p, im.tile_order, tiles);
    carve_array(p, im.block_ranges, tiles * 8);
    carve_array(p, im.n_contrib_dense, n);
    carve_array(p, im.bin_info, BIN_WORDS);
    if (out) *out = im;
    return (size_t)(p - base) + 128;
}

// bit_length(n): number of key bits the tile id needs (getHigherMsb, rasterizer_impl.cu:31-44)
static int tile_bits(uint32_t n_tiles) {
    int b = 0;
    while (n_tiles >> b) ++b;
    return b;
}

// Every binning path leaves the sorted lists in BinState::keys_sorted / point_list: the ping-pong sorts are started in
// whichever buffer makes their last digit pass land there (no final copy).
static int digit_passes(int key_bits) { return (key_bits + 7) / 8; }

enum BinningPath { BIN_DEPTHFIRST = 0, BIN_SORT64 = 1 };
static BinningPath binning_path_from_env() {
    const char* e = getenv("GS2M_BINNING");
    if (e && strcmp(e, "sort64") == 0) return BIN_SORT64;
    return BIN_DEPTHFIRST;
}

// Host side of the instance-count read-back: a pinned landing buffer and an event per (host thread, device).
struct HostSlot { uint32_t* pinned; cudaEvent_t ev; };
static thread_local HostSlot g_slots[64] = {};
static thread_local long long g_last_R = 0;
static int host_slot(HostSlot** out) {
    int dev = 0;
    GS2M_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device index %d outside 0..63", dev); return GS2M_ERR_INVALID_ARGUMENT; }
    HostSlot& h = g_slots[dev];
    if (!h.pinned) {
        GS2M_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h.pinned), BIN_WORDS * sizeof(uint32_t), cudaHostAllocDefault));
        GS2M_CUDA(cudaEventCreateWithFlags(&h.ev, cudaEventDisableTiming));
    }
    *out = &h;
    return GS2M_OK;
}

// discard flags of the control block -> error code (0 when the result is valid)
static int check_bin_flags(const uint32_t* info, int R_capacity) {
    const uint32_t flags = info[BIN_FLAGS];
    g_last_R = (long long)info[BIN_R];
    if (flags & GS2M_BIN_TOO_LARGE) { set_error("%u or more Gaussian/tile instances exceed the supported 2^30", info[BIN_R]); return GS2M_ERR_TOO_LARGE; }
    if (flags & GS2M_BIN_PREFILTERED) { set_error("prefiltered was set but a Gaussian is behind the near plane (view-space z <= 0.2)"); return GS2M_ERR_PREFILTERED; }
    if (flags & GS2M_BIN_OVERFLOW) { set_error("%u instances do not fit R_capacity %d", info[BIN_R], R_capacity); return GS2M_ERR_CAPACITY; }
    return GS2M_OK;
}

static int validate_common(int P, int D, int M, int W, int H, int F, const void* means3D, const void* shs,
                           const void* colors_precomp, const void* scales, const void* rotations, const void* cov3D,
                           const void* features, const void* vm, const void* pm, const void* cam) {
    if (P < 0 || W <= 0 || H <= 0) { set_error("invalid sizes P=%d W=%d H=%d", P, W, H); return GS2M_ERR_INVALID_ARGUMENT; }
    if (F < 0 || F > GS2M_NUM_FEATURES) { set_error("feature_count %d outside 0..10", F); return GS2M_ERR_INVALID_ARGUMENT; }
    if (P == 0) return GS2M_OK;
    if (!means3D || !vm || !pm) { set_error("means3D / viewmatrix / projmatrix must not be NULL"); return GS2M_ERR_INVALID_ARGUMENT; }
    if ((shs == nullptr) == (colors_precomp == nullptr)) { set_error("provide exactly one of shs / colors_precomp"); return GS2M_ERR_INVALID_ARGUMENT; }
    if (shs && (M <= 0 || D < 0 || D > 3 || (D + 1) * (D + 1) > M)) { set_error("SH degree %d does not fit M=%d coefficients", D, M); return GS2M_ERR_INVALID_ARGUMENT; }
    if (shs && !cam) { set_error("cam_pos must not be NULL when shs are used"); return GS2M_ERR_INVALID_ARGUMENT; }
    const bool have_sr = scales != nullptr && rotations != nullptr;
    if (have_sr == (cov3D != nullptr) || ((scales != nullptr) != (rotations != nullptr))) {
        set_error("provide exactly one of scales+rotations / cov3D_precomp"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    if (F > 0 && !features) { set_error("features must not be NULL when feature_count > 0"); return GS2M_ERR_INVALID_ARGUMENT; }
    return GS2M_OK;
}

}  // namespace gs2m

using namespace gs2m;

extern "C" {

int gs2m_abi_version(void) { return GS2M_ABI_VERSION; }
long long gs2m_launch_count(void) { return g_launches.load(); }
void gs2m_profile_enable(int enable) { g_profile.store(enable ? 1 : 0); }
int gs2m_profile_read(float* stage_ms, int* stage_calls) {
    std::vector<PendingStage> todo;
    { std::lock_guard<std::mutex> lk(g_prof_mu); todo.swap(g_pending); }
    for (auto& p : todo) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            if (stage_ms) stage_ms[p.stage] += ms;
            if (stage_calls) stage_calls[p.stage] += 1;
        }
        cudaEventDestroy(p.e0);
        cudaEventDestroy(p.e1);
    }
    return GS2M_OK;
}
const char* gs2m_last_error(void) { return g_error; }

size_t gs2m_geometry_bytes(int P) { return GeomState::carve(nullptr, P, nullptr); }
size_t gs2m_image_bytes(int width, int height) { return ImageState::carve(nullptr, width, height, nullptr); }
size_t gs2m_binning_bytes(int R) { return BinState::carve(nullptr, R, nullptr); }
size_t gs2m_sort_temp_bytes(int n) { return sort_temp_bytes(n); }
size_t gs2m_scan_temp_bytes(int n) { return scan_temp_bytes(n); }

int gs2m_sort_pairs_u64(uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n, int end_bit,
                        char* temp, void* stream) {
    return sort_pairs_u64(keys_in, keys_out, vals_in, vals_out, n, end_bit, temp, (cudaStream_t)stream);
}

int gs2m_inclusive_sum_u32(const uint32_t* in, uint32_t* out, int n, char* temp, void* stream) {
    return inclusive_sum_u32(in, out, n, temp, (cudaStream_t)stream);
}

int gs2m_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                      void* stream) {
    (void)projmatrix;  // the reference's frustum test only uses the view matrix (auxiliary.h:148-152)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error("mark_visible: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT; }
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

int gs2m_rasterize_forward(const gs2m_forward_args* a) {
    if (!a) { set_error("null args"); return GS2M_ERR_INVALID_ARGUMENT; }
    int rc = validate_common(a->P, a->D, a->M, a->width, a->height, a->feature_count, a->means3D, a->shs, a->colors_precomp,
                             a->scales, a->rotations, a->cov3D_precomp, a->features, a->viewmatrix, a->projmatrix, a->cam_pos);
    if (rc != GS2M_OK) return rc;
    if (!a->out_color || !a->out_buffer || !a->background || (a->P > 0 && (!a->out_radii || !a->out_observe || !a->opacities))) {
        set_error("forward: missing output / background / opacity pointer"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    if (!a->geometry_buffer || !a->binning_buffer || !a->image_buffer) { set_error("forward: missing resize callbacks"); return GS2M_ERR_INVALID_ARGUMENT; }
    if (a->R_capacity < 0 || a->R_capacity >= (1 << 30)) { set_error("forward: R_capacity %d outside 0..2^30-1", a->R_capacity); return GS2M_ERR_INVALID_ARGUMENT; }
    cudaStream_t s = (cudaStream_t)a->stream;

    FwdParams p;
    p.P = a->P; p.D = a->D; p.M = a->M; p.W = a->width; p.H = a->height; p.F = a->feature_count;
    p.tiles_x = (a->width + GS2M_TILE_X - 1) / GS2M_TILE_X;
    p.tiles_y = (a->height + GS2M_TILE_Y - 1) / GS2M_TILE_Y;
    p.tan_fovx = a->tan_fovx; p.tan_fovy = a->tan_fovy;
    p.focal_y = a->height / (2.0f * a->tan_fovy);   // rasterizer_impl.cu:212-213
    p.focal_x = a->width / (2.0f * a->tan_fovx);
    p.scale_modifier = a->scale_modifier;
    p.means3D = a->means3D; p.shs = a->shs; p.colors_precomp = a->colors_precomp; p.opacities = a->opacities;
    p.scales = a->scales; p.rotations = a->rotations; p.cov3D_precomp = a->cov3D_precomp; p.features = a->features;
    p.viewmatrix = a->viewmatrix; p.projmatrix = a->projmatrix; p.cam_pos = a->cam_pos; p.background = a->background;
    const int n_tiles = p.tiles_x * p.tiles_y;

    char* img_base = a->image_buffer(a->image_user, ImageState::carve(nullptr, p.W, p.H, nullptr));
    if (!img_base) { set_error("image_buffer callback returned NULL"); return GS2M_ERR_ALLOC; }
    ImageState im;
    ImageState::carve(img_base, p.W, p.H, &im);
    GS2M_CUDA(cudaMemsetAsync(im.bin_info, 0, BIN_WORDS * sizeof(uint32_t), s));

    // Binning path (both give bit-identical keys / lists / ranges, tests/test_gpu_parity.py):
    //   "depthfirst" (default)  depth-sort the visible Gaussians, emit in depth order, 2-pass sort on the tile id
    //   "sort64"                duplicate + 64-bit onesweep radix sort, the reference's structure (exact mode only)
    // B200, config 4, binning total: 0.47 / 0.77 ms (profiles/r1_binning_paths.md); GS2M_BINNING=sort64 selects the second.
    BinningPath path = binning_path_from_env();
    // exact mode: R is read back before the binning arena is sized (the reference's one host sync).  speculative mode: the
    // arena is sized for R_capacity and the count-dependent kernels read R / V from the control block on the device.
    if (a->R_capacity > 0 && path != BIN_DEPTHFIRST) { set_error("forward: speculative mode (R_capacity > 0) needs the default binning path"); return GS2M_ERR_INVALID_ARGUMENT; }
    const bool speculative = a->R_capacity > 0;
    int R = 0, R_cap = 0, V_cap = 0;
    GeomState g;
    memset(&g, 0, sizeof(g));
    HostSlot* slot = nullptr;
    if (p.P > 0) {
        char* geom_base = a->geometry_buffer(a->geometry_user, GeomState::carve(nullptr, p.P, nullptr));
        if (!geom_base) { set_error("geometry_buffer callback returned NULL"); return GS2M_ERR_ALLOC; }
        GeomState::carve(geom_base, p.P, &g);

        { StageTimer t(GS2M_STAGE_PREPROCESS_FWD, s);
          rc = launch_preprocess_forward(p, g, a->out_radii, a->out_observe, im.bin_info + BIN_FLAGS, a->prefiltered != 0,
                                         a->no_backward == 0, s); }
        if (rc != GS2M_OK) return rc;
        const bool need_host = !(speculative && a->no_wait);
        if (need_host && (rc = host_slot(&slot)) != GS2M_OK) return rc;
        if (path == BIN_DEPTHFIRST) {
            // point_offsets + compaction of the visible Gaussians to (depth bits, index) + the control block in one scan
            { StageTimer t(GS2M_STAGE_SCAN, s);
              rc = binning_df_compact(p.P, g, g.depth_keys_alt, g.order_b, im.bin_info,
                                      speculative ? (uint32_t)a->R_capacity : 0x3FFFFFFFu, s); }
            if (rc != GS2M_OK) return rc;
            if (need_host) {
                GS2M_CUDA(cudaMemcpyAsync(slot->pinned, im.bin_info, BIN_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
                GS2M_CUDA(cudaEventRecord(slot->ev, s));
            }
            if (speculative) {
                R_cap = a->R_capacity;
                V_cap = p.P;
            } else {
                GS2M_CUDA(cudaEventSynchronize(slot->ev));
                if ((rc = check_bin_flags(slot->pinned, 0)) != GS2M_OK) return rc;
                R = R_cap = (int)slot->pinned[BIN_R];
                V_cap = (int)slot->pinned[BIN_V];
            }
        } else {
            { StageTimer t(GS2M_STAGE_SCAN, s); rc = inclusive_sum_u32(g.tiles_touched, g.point_offsets, p.P, g.scan_temp, s); }
            if (rc != GS2M_OK) return rc;
            GS2M_CUDA(cudaMemcpyAsync(slot->pinned, im.bin_info, BIN_WORDS * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            GS2M_CUDA(cudaMemcpyAsync(slot->pinned + BIN_R, g.point_offsets + (p.P - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            GS2M_CUDA(cudaStreamSynchronize(s));
            if (slot->pinned[BIN_R] >= (1u << 30)) slot->pinned[BIN_FLAGS] |= GS2M_BIN_TOO_LARGE;   // 32-bit sum on this path
            if ((rc = check_bin_flags(slot->pinned, 0)) != GS2M_OK) return rc;
            R = R_cap = (int)slot->pinned[BIN_R];
        }
    }

    char* bin_base = a->binning_buffer(a->binning_user, BinState::carve(nullptr, R_cap, nullptr));
    if (!bin_base) { set_error("binning_buffer callback returned NULL"); return GS2M_ERR_ALLOC; }
    BinState b;
    BinState::carve(bin_base, R_cap, &b);
    if (R_cap > 0 && path == BIN_DEPTHFIRST) {
        const uint32_t* v_used = im.bin_info + BIN_V_USED;
        const uint32_t* r_used = im.bin_info + BIN_R_USED;
        // depth order of the V visible Gaussians (ties keep ascending index: the compaction is stable and so is the sort)
        int in_input = 0;
        { StageTimer t(GS2M_STAGE_SORT, s);
          rc = sort_pairs_u32_pingpong(g.depth_keys_alt, g.depth_keys, g.order_b, g.order_a, V_cap, v_used, 32, g.rank_temp, s, &in_input); }
        if (rc != GS2M_OK) return rc;
        const uint32_t* order = in_input ? g.order_b : g.order_a;
        // 32-bit tile ids ping-pong between the two halves of keys_unsorted; start where the last pass lands in point_list
        int key_bits = 1;
        while (((long long)n_tiles - 1) >> key_bits) ++key_bits;
        uint32_t* tk[2] = {reinterpret_cast<uint32_t*>(b.keys_unsorted), reinterpret_cast<uint32_t*>(b.keys_unsorted) + R_cap};
        uint32_t* tv[2] = {b.vals_unsorted, b.point_list};
        const int start = digit_passes(key_bits) & 1 ? 0 : 1;
        { StageTimer t(GS2M_STAGE_DUPLICATE, s);
          rc = binning_df_emit(V_cap, v_used, g, order, a->out_radii, p.tiles_x, p.tiles_y, tk[start], tv[start], s); }
        if (rc != GS2M_OK) return rc;
        { StageTimer t(GS2M_STAGE_SORT, s);
          rc = sort_pairs_u32_pingpong(tk[start], tk[start ^ 1], tv[start], tv[start ^ 1], R_cap, r_used, key_bits, b.sort_temp, s, &in_input); }
        if (rc != GS2M_OK) return rc;
        if ((in_input != 0) != (start == 1)) { set_error("internal: sort buffer parity mismatch"); return GS2M_ERR_CUDA; }
        { StageTimer t(GS2M_STAGE_RANGES, s);
          rc = launch_ranges_masks_keys(R_cap, r_used, p.tiles_x, p.tiles_y, tk[1], b, g, im, s); }
        if (rc != GS2M_OK) return rc;
    } else if (R_cap > 0) {
        const int key_bits = 32 + tile_bits((uint32_t)n_tiles);
        uint64_t* kb[2] = {b.keys_unsorted, b.keys_sorted};
        uint32_t* vb[2] = {b.vals_unsorted, b.point_list};
        const int start = digit_passes(key_bits) & 1 ? 0 : 1;
        { StageTimer t(GS2M_STAGE_DUPLICATE, s);
          rc = launch_duplicate_with_keys(p.P, g, a->out_radii, p.tiles_x, p.tiles_y, kb[start], vb[start], s); }
        if (rc != GS2M_OK) return rc;
        int in_input = 0;
        { StageTimer t(GS2M_STAGE_SORT, s);
          rc = sort_pairs_u64_pingpong(kb[start], kb[start ^ 1], vb[start], vb[start ^ 1], R, key_bits, b.sort_temp, s, &in_input); }
        if (rc != GS2M_OK) return rc;
        if ((in_input != 0) != (start == 1)) { set_error("internal: sort buffer parity mismatch"); return GS2M_ERR_CUDA; }
        // tile ranges + footprint masks in one pass over the sorted list
        { StageTimer t(GS2M_STAGE_RANGES, s);
          rc = launch_ranges_and_masks(R, p.tiles_x, p.tiles_y, b, g, im, s); }
        if (rc != GS2M_OK) return rc;
    } else {
        StageTimer t(GS2M_STAGE_RANGES, s);
        rc = launch_identify_tile_ranges(0, b.keys_sorted, im.ranges, n_tiles, s);
        if (rc != GS2M_OK) return rc;
        GS2M_CUDA(cudaMemsetAsync(im.block_ranges, 0, (size_t)n_tiles * 8 * sizeof(uint2), s));
    }
    { StageTimer t(GS2M_STAGE_RANGES, s); rc = launch_tile_order(n_tiles, im.ranges, im.tile_order, s); }
    if (rc != GS2M_OK) return rc;
    { StageTimer t(GS2M_STAGE_BLEND_FWD, s);
      rc = launch_blend_forward(p, g, b, R_cap, im, a->out_color, a->out_observe, a->out_buffer, s); }
    if (rc != GS2M_OK) return rc;
    if (speculative && p.P > 0) {
        if (a->no_wait) return a->R_capacity;
        // everything is queued; the count was ready long before the blend that is now running
        GS2M_CUDA(cudaEventSynchronize(slot->ev));
        if ((rc = check_bin_flags(slot->pinned, a->R_capacity)) != GS2M_OK) return rc;
        R = (int)slot->pinned[BIN_R];
    }
    return R;
}

long long gs2m_last_instance_count(void) { return g_last_R; }

// Validates one backward call and assembles the kernels' view of it.  Returns GS2M_OK with `empty` set when there is nothing to do.
struct BackwardCall { BwdParams p; GeomState g; BinState b; ImageState im; cudaStream_t s; bool empty; int R_carve; };

static int assemble_backward(const gs2m_backward_args* a, BackwardCall& c) {
    c.empty = false;
    if (!a) { set_error("null args"); return GS2M_ERR_INVALID_ARGUMENT; }
    int rc = validate_common(a->P, a->D, a->M, a->width, a->height, a->feature_count, a->means3D, a->shs, a->colors_precomp,
                             a->scales, a->rotations, a->cov3D_precomp, a->features, a->viewmatrix, a->projmatrix, a->cam_pos);
    if (rc != GS2M_OK) return rc;
    if (a->P == 0) { c.empty = true; return GS2M_OK; }
    if (!a->radii || !a->geometry_buffer || !a->binning_buffer || !a->image_buffer || !a->grad_color ||
        (a->feature_count > 0 && !a->grad_buffer) || !a->background) {
        set_error("backward: missing saved state / upstream gradient pointer"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    // dL_dcolor / dL_dcov3D are the gradients of the PRECOMPUTED colour / covariance inputs: required when those inputs are
    // used, optional (NULL = not written) when SHs / scale+rotation are
    const gs2m_param_chain* ch = a->chain;
    if (ch) {
        if (!a->scales || !a->rotations || a->M > 16 || a->accumulate == 1) {
            set_error("backward: chain needs scales + rotations inputs, M <= 16 and accumulate 0 or 2"); return GS2M_ERR_INVALID_ARGUMENT;
        }
        if (!ch->scaling_raw || !ch->rotation_raw || !ch->opacity_raw || !ch->albedo_raw || !ch->roughness_raw || !ch->metallic_raw ||
            !ch->d_xyz || !ch->d_scaling_raw || !ch->d_rotation_raw || !ch->d_opacity_raw || !ch->d_albedo_raw ||
            !ch->d_roughness_raw || !ch->d_metallic_raw || !a->cam_pos) {
            set_error("backward: chain with a NULL pointer"); return GS2M_ERR_INVALID_ARGUMENT;
        }
    } else if (!a->dL_dmeans2D || !a->dL_dopacity || !a->dL_dmeans3D || !a->dL_dscale || !a->dL_drot || !a->dL_dfeatures) {
        set_error("backward: missing gradient output pointer"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    if ((a->colors_precomp && !a->dL_dcolor) || (a->cov3D_precomp && !a->dL_dcov3D) || (a->M > 0 && a->shs && !a->dL_dsh)) {
        set_error("backward: missing gradient output pointer"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    if (a->R < 0) { set_error("backward: negative R"); return GS2M_ERR_INVALID_ARGUMENT; }
    if (a->accumulate < 0 || a->accumulate > 2) { set_error("backward: accumulate mode %d outside 0..2", a->accumulate); return GS2M_ERR_INVALID_ARGUMENT; }
    if (a->phase < 0 || a->phase > 2) { set_error("backward: phase %d outside 0..2", a->phase); return GS2M_ERR_INVALID_ARGUMENT; }
    const bool ranged = a->row_end > 0 || a->row_begin != 0;
    if (ranged && (a->phase != 2 || a->row_begin < 0 || a->row_end > a->P || a->row_begin > a->row_end || (a->row_begin & 255))) {
        set_error("backward: rows [%d, %d) need phase 2, 0 <= begin <= end <= P and begin %% 256 == 0", a->row_begin, a->row_end);
        return GS2M_ERR_INVALID_ARGUMENT;
    }
    if (a->R_capacity < 0 || (a->R_capacity > 0 && a->R > a->R_capacity)) { set_error("backward: R %d does not fit R_capacity %d", a->R, a->R_capacity); return GS2M_ERR_INVALID_ARGUMENT; }
    const int R_carve = a->R_capacity > 0 ? a->R_capacity : a->R;    // what forward sized the binning arena for
    if (a->geometry_bytes < GeomState::carve(nullptr, a->P, nullptr) ||
        a->binning_bytes < BinState::carve(nullptr, R_carve, nullptr) ||
        a->image_bytes < ImageState::carve(nullptr, a->width, a->height, nullptr)) {
        set_error("backward: a saved arena is smaller than forward allocated it"); return GS2M_ERR_INVALID_ARGUMENT;
    }
    c.s = (cudaStream_t)a->stream;

    BwdParams& p = c.p;
    p.P = a->P; p.D = a->D; p.M = a->M; p.W = a->width; p.H = a->height; p.F = a->feature_count; p.R = a->R;
    p.tiles_x = (a->width + GS2M_TILE_X - 1) / GS2M_TILE_X;
    p.tiles_y = (a->height + GS2M_TILE_Y - 1) / GS2M_TILE_Y;
    p.tan_fovx = a->tan_fovx; p.tan_fovy = a->tan_fovy;
    p.focal_y = a->height / (2.0f * a->tan_fovy);
    p.focal_x = a->width / (2.0f * a->tan_fovx);
    p.scale_modifier = a->scale_modifier;
    p.means3D = a->means3D; p.shs = a->shs; p.colors_precomp = a->colors_precomp; p.scales = a->scales;
    p.rotations = a->rotations; p.cov3D_precomp = a->cov3D_precomp; p.features = a->features;
    p.viewmatrix = a->viewmatrix; p.projmatrix = a->projmatrix; p.cam_pos = a->cam_pos; p.background = a->background;
    p.radii = a->radii; p.grad_color = a->grad_color; p.grad_buffer = a->grad_buffer;
    p.dL_dmeans2D = a->dL_dmeans2D; p.dL_dconic = a->dL_dconic; p.dL_dopacity = a->dL_dopacity; p.dL_dcolor = a->dL_dcolor;
    p.dL_dmeans3D = a->dL_dmeans3D; p.dL_dcov3D = a->dL_dcov3D; p.dL_dsh = a->dL_dsh; p.dL_dscale = a->dL_dscale;
    p.dL_drot = a->dL_drot; p.dL_dfeatures = a->dL_dfeatures; p.accumulate = a->accumulate;
    p.row_begin = ranged ? a->row_begin : 0;
    p.row_end = ranged ? a->row_end : a->P;
    p.densify_grad_accum = a->densify_grad_accum; p.densify_grad_accum_abs = a->densify_grad_accum_abs;
    p.densify_denom = a->densify_denom;
    p.has_chain = ch != nullptr;
    if (ch) p.chain = *ch; else memset(&p.chain, 0, sizeof(p.chain));

    c.R_carve = R_carve;
    GeomState::carve(a->geometry_buffer, p.P, &c.g);
    BinState::carve(a->binning_buffer, R_carve, &c.b);
    ImageState::carve(a->image_buffer, p.W, p.H, &c.im);
    return GS2M_OK;
}

int gs2m_rasterize_backward(const gs2m_backward_args* a) {
    BackwardCall c;
    int rc = assemble_backward(a, c);
    if (rc != GS2M_OK || c.empty) return rc;
    const BwdParams& p = c.p;
    const GeomState& g = c.g;
    const BinState& b = c.b;
    const ImageState& im = c.im;
    cudaStream_t s = c.s;

    // The accumulator rows of the visible Gaussians were zeroed by the forward (preprocess); a second backward over the same
    // forward state has to start from zero again.
    if (a->phase != 2) {
        if (a->grad_acc_dirty) GS2M_CUDA(cudaMemsetAsync(g.grad_acc, 0, (size_t)p.P * GS2M_ACC_STRIDE * sizeof(float), s));
        { StageTimer t(GS2M_STAGE_BLEND_BWD, s); rc = launch_blend_backward(p, g, b, c.R_carve, im, s); }
        if (rc != GS2M_OK) return rc;
    }
    if (a->phase != 1 && p.row_end > p.row_begin) {
        StageTimer t(GS2M_STAGE_PREPROCESS_BWD, s);
        rc = launch_preprocess_backward(p, g, s);
    }
    return rc;
}

// The per-Gaussian stage (phase 2) of several views in one pass: every thread owns one Gaussian, walks the views that see it,
// sums their chained raw-parameter gradients in registers / shared memory and writes each output element once.
int gs2m_rasterize_backward_views(const gs2m_backward_args* views, int n_views) {
    if (!views || n_views <= 0) { set_error("backward_views: no views"); return GS2M_ERR_INVALID_ARGUMENT; }
    std::vector<BackwardCall> calls((size_t)n_views);
    for (int v = 0; v < n_views; ++v) {
        const gs2m_backward_args* a = views + v;
        int rc = assemble_backward(a, calls[v]);
        if (rc != GS2M_OK) return rc;
        if (calls[v].empty) return GS2M_OK;        // P == 0 (the same for every view, checked below for the others)
        const gs2m_backward_args* f = views;
        if (a->phase != 2 || !a->chain) { set_error("backward_views: every view needs phase 2 and a chain"); return GS2M_ERR_INVALID_ARGUMENT; }
        if (a->dL_dcolor || a->dL_dcov3D || a->colors_precomp || a->cov3D_precomp) {
            set_error("backward_views: precomputed colours / covariances are not supported"); return GS2M_ERR_INVALID_ARGUMENT;
        }
        if (a->P != f->P || a->M != f->M || a->D != f->D || a->means3D != f->means3D || a->shs != f->shs || a->dL_dsh != f->dL_dsh ||
            a->row_begin != f->row_begin || a->row_end != f->row_end || a->stream != f->stream ||
            a->scale_modifier != f->scale_modifier || memcmp(a->chain, f->chain, sizeof(gs2m_param_chain)) != 0) {
            set_error("backward_views: the views must share the Gaussians, the row range, the stream and the chain"); return GS2M_ERR_INVALID_ARGUMENT;
        }
    }
    if (calls[0].p.row_end <= calls[0].p.row_begin) return GS2M_OK;
    StageTimer t(GS2M_STAGE_PREPROCESS_BWD, calls[0].s);
    std::vector<BwdParams> ps((size_t)n_views);
    std::vector<GeomState> gs((size_t)n_views);
    for (int v = 0; v < n_views; ++v) { ps[v] = calls[v].p; gs[v] = calls[v].g; }
    return launch_preprocess_backward_views(ps.data(), gs.data(), n_views, views->accumulate != 0, calls[0].s);
}

int gs2m_state_view_get(int P, int width, int height, int R, char* geometry_buffer, char* binning_buffer,
                        char* image_buffer, gs2m_state_view* out) {
    if (!out || P < 0 || R < 0 || width <= 0 || height <= 0) { set_error("state_view: bad arguments"); return GS2M_ERR_INVALID_ARGUMENT; }
    memset(out, 0, sizeof(*out));
    if (geometry_buffer) {
        GeomState g;
        GeomState::carve(geometry_buffer, P, &g);
        out->depths = g.depths;
        out->rec_a = reinterpret_cast<const float*>(g.xy_conic_ab);
        out->rec_b = reinterpret_cast<const float*>(g.conic_c_opac);
        out->rgb = reinterpret_cast<const float*>(g.rgb);
        out->cov3D = g.cov3D;
        out->clamped = g.clamped;
        out->tiles_touched = g.tiles_touched;
        out->point_offsets = g.point_offsets;
        out->grad_acc = g.grad_acc;
    }
    if (binning_buffer) {
        BinState b;
        BinState::carve(binning_buffer, R, &b);
        out->keys_sorted = b.keys_sorted;
        out->point_list = b.point_list;
        out->masks = b.masks;
        out->dense_gid = b.dense_gid;
        out->dense_pos = b.dense_pos;
    }
    if (image_buffer) {
        ImageState im;
        ImageState::carve(image_buffer, width, height, &im);
        out->final_T = im.final_T;
        out->n_contrib = im.n_contrib;
        out->ranges = reinterpret_cast<const uint32_t*>(im.ranges);
        out->bin_info = im.bin_info;
        out->block_ranges = reinterpret_cast<const uint32_t*>(im.block_ranges);
        out->n_contrib_dense = im.n_contrib_dense;
    }
    return GS2M_OK;
}

}  // extern "C"
