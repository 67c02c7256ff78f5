// Default binning path ("depthfirst"): the (tile | depth) order of the reference's 64-bit sort, produced by two short sorts
// instead of one long one.
//
// Behavioural reference: InclusiveSum + duplicateWithKeys + SortPairs on 32+bit_length(n_tiles) key bits
// (rasterizer_impl.cu:265-296, :63-103).  A stable sort by (tile, depth) equals a stable sort by depth followed by a
// stable sort by tile, and the depth of an instance is the depth of its Gaussian, so:
//
//   1. compact    one scan pass over tiles_touched yields both the reference's point_offsets (kept for parity) and the
//                 stable compaction of the visible Gaussians to (depth bits, index) pairs                       [P]
//   2. depth sort 4 digit passes of the 32-bit onesweep sort over the V visible Gaussians (ties keep index order) [V]
//   3. emit       scan of tiles_touched in depth order, then every Gaussian writes (tile id, index) for its tile
//                 rectangle, row-major like duplicateWithKeys                                                    [R]
//   4. tile sort  ceil(bit_length(n_tiles-1) / 8) = 2 digit passes of the same sort on 32-bit tile ids           [R]
//
// The instance list is bit-identical to the 64-bit path's (same ties: ascending Gaussian index), and the 64-bit keys
// themselves are re-materialised next to it by the range/mask kernel.  Traffic per instance drops from 6 passes x 24 B
// to 2 passes x 16 B (config 4: duplicate + sort 0.60 ms -> emit + both sorts 0.27 ms).
#include "common.cuh"

namespace gs2m {
namespace {

constexpr int CS_THREADS = 256;
constexpr int CS_ITEMS = 8;
constexpr int CS_TILE = CS_THREADS * CS_ITEMS;

__device__ __forceinline__ uint2 warp_inclusive_scan2(uint2 v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t nx = __shfl_up_sync(0xffffffffu, v.x, d);
        const uint32_t ny = __shfl_up_sync(0xffffffffu, v.y, d);
        if (lane >= d) { v.x += nx; v.y += ny; }
    }
    return v;
}

// block-wide exclusive scan of a pair of counters per thread
__device__ __forceinline__ uint2 block_exclusive_scan2(uint2 v, uint2* smem_warp /*[8]*/, uint2& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint2 inc = warp_inclusive_scan2(v, lane);
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    uint2 base = make_uint2(0, 0), tot = make_uint2(0, 0);
#pragma unroll
    for (int w = 0; w < CS_THREADS / 32; ++w) {
        const uint2 s = smem_warp[w];
        if (w < warp) { base.x += s.x; base.y += s.y; }
        tot.x += s.x; tot.y += s.y;
    }
    __syncthreads();
    total = tot;
    return make_uint2(base.x + inc.x - v.x, base.y + inc.y - v.y);
}

// x = tiles touched, y = visible (tiles touched > 0)
__global__ void __launch_bounds__(CS_THREADS) compact_tile_sums_kernel(const uint32_t* __restrict__ tiles_touched, int n,
                                                                       uint2* __restrict__ tile_sums) {
    __shared__ uint2 sw[8];
    const int base = blockIdx.x * CS_TILE + threadIdx.x * CS_ITEMS;
    uint2 s = make_uint2(0, 0);
#pragma unroll
    for (int i = 0; i < CS_ITEMS; ++i)
        if (base + i < n) { const uint32_t t = tiles_touched[base + i]; s.x += t; s.y += (t != 0u); }
    uint2 total;
    block_exclusive_scan2(s, sw, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place.  The number of tile sums is ceil(n / items_per_tile) with n read from
// device memory when n_ptr is given (the emission's scan runs over the V visible Gaussians, a device-side count).
// CONTROL: also writes the forward's control block `info` (common.cuh BIN_*): the true instance count R (summed in 64 bits: the
// 32-bit running offsets may wrap when screen-filling splats push the sum past 2^32) and visible count V, the
// discard flags, and the counts the later kernels use (zero when the result must be discarded, so that nothing is ever
// written past the arena).
template <bool CONTROL>
__global__ void __launch_bounds__(CS_THREADS) compact_spine_kernel(uint2* __restrict__ tile_sums, int n_cap,
                                                                   const uint32_t* __restrict__ n_ptr, int items_per_tile,
                                                                   uint32_t* __restrict__ info, uint32_t R_capacity) {
    __shared__ uint2 sw[8];
    __shared__ unsigned long long s_total;
    if (threadIdx.x == 0) s_total = 0ull;
    const int n = n_ptr ? (int)min(*n_ptr, (uint32_t)n_cap) : n_cap;
    const int n_tiles = (n + items_per_tile - 1) / items_per_tile;
    uint2 carry = make_uint2(0, 0);
    unsigned long long wide = 0ull;
    constexpr int SP_ITEMS = 4;                 // consecutive tile sums per thread: a quarter of the block-scan rounds
    for (int start = 0; start < n_tiles; start += CS_THREADS * SP_ITEMS) {
        const int i0 = start + threadIdx.x * SP_ITEMS;
        uint2 v[SP_ITEMS];
        uint2 mine = make_uint2(0, 0);
#pragma unroll
        for (int k = 0; k < SP_ITEMS; ++k) {
            v[k] = (i0 + k < n_tiles) ? tile_sums[i0 + k] : make_uint2(0, 0);
            mine.x += v[k].x; mine.y += v[k].y;
            wide += v[k].x;
        }
        uint2 total;
        uint2 run = block_exclusive_scan2(mine, sw, total);
        run.x += carry.x; run.y += carry.y;
#pragma unroll
        for (int k = 0; k < SP_ITEMS; ++k) {
            if (i0 + k < n_tiles) tile_sums[i0 + k] = run;
            run.x += v[k].x; run.y += v[k].y;
        }
        carry.x += total.x; carry.y += total.y;
    }
    if (!CONTROL) return;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wide += __shfl_down_sync(0xffffffffu, wide, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_total, wide);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long R64 = s_total;
        uint32_t flags = info[BIN_FLAGS];                    // GS2M_BIN_PREFILTERED may have been raised by the preprocess
        if (R64 >= (1ull << 30)) flags |= GS2M_BIN_TOO_LARGE;
        else if (R64 > (unsigned long long)R_capacity) flags |= GS2M_BIN_OVERFLOW;
        const bool discard = (flags & (GS2M_BIN_TOO_LARGE | GS2M_BIN_OVERFLOW)) != 0;
        info[BIN_R] = (uint32_t)min(R64, 0xFFFFFFFFull);
        info[BIN_V] = carry.y;
        info[BIN_FLAGS] = flags;
        info[BIN_R_USED] = discard ? 0u : (uint32_t)R64;
        info[BIN_V_USED] = discard ? 0u : carry.y;
    }
}

__global__ void __launch_bounds__(CS_THREADS) compact_apply_kernel(const uint32_t* __restrict__ tiles_touched,
                                                                   const float* __restrict__ depths, int n,
                                                                   const uint2* __restrict__ tile_offsets,
                                                                   uint32_t* __restrict__ point_offsets,
                                                                   uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    __shared__ uint2 sw[8];
    const int base = blockIdx.x * CS_TILE + threadIdx.x * CS_ITEMS;
    uint32_t t[CS_ITEMS];
    uint2 s = make_uint2(0, 0);
#pragma unroll
    for (int i = 0; i < CS_ITEMS; ++i) {
        t[i] = (base + i < n) ? tiles_touched[base + i] : 0u;
        s.x += t[i]; s.y += (t[i] != 0u);
    }
    uint2 total;
    uint2 run = block_exclusive_scan2(s, sw, total);
    const uint2 off = tile_offsets[blockIdx.x];
    run.x += off.x; run.y += off.y;
#pragma unroll
    for (int i = 0; i < CS_ITEMS; ++i) {
        run.x += t[i];
        if (base + i < n) point_offsets[base + i] = run.x;     // inclusive, like cub::DeviceScan::InclusiveSum
        if (t[i] != 0u) {
            keys[run.y] = __float_as_uint(depths[base + i]);   // depth > 0.2: the bit pattern orders like the float
            vals[run.y] = (uint32_t)(base + i);
            ++run.y;
        }
    }
}

// inclusive scan of tiles_touched[order[i]]: same three-kernel scheme on one counter, one Gaussian per thread
constexpr int EMIT_ITEMS = 1;
constexpr int EMIT_TILE = CS_THREADS * EMIT_ITEMS;

__global__ void __launch_bounds__(CS_THREADS) ordered_tile_sums_kernel(const uint32_t* __restrict__ tiles_touched,
                                                                       const uint32_t* __restrict__ order, int n_cap,
                                                                       const uint32_t* __restrict__ n_ptr,
                                                                       uint2* __restrict__ tile_sums) {
    __shared__ uint2 sw[8];
    const int n = n_ptr ? (int)min(*n_ptr, (uint32_t)n_cap) : n_cap;
    if ((int)(blockIdx.x * EMIT_TILE) >= n) return;
    const int base = blockIdx.x * EMIT_TILE + threadIdx.x * EMIT_ITEMS;
    uint2 s = make_uint2(0, 0);
#pragma unroll
    for (int i = 0; i < EMIT_ITEMS; ++i)
        if (base + i < n) s.x += tiles_touched[order[base + i]];
    uint2 total;
    block_exclusive_scan2(s, sw, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// scan apply fused with the emission.  A thread owns one depth rank (its Gaussian's tile rectangle and instance count); the
// instances themselves are written WARP-COOPERATIVELY: the warp's 32 Gaussians occupy one contiguous slice of the output, lane l
// writes slots l, l+32, ... of that slice and finds the Gaussian a slot belongs to by a 5-step search over the lanes' exclusive
// offsets.  Every store instruction therefore covers 32 consecutive (tile id, index) pairs, and a screen-filling splat is spread
// over the whole warp instead of serialising one thread (the per-thread loops of duplicateWithKeys, rasterizer_impl.cu:63-103).
// The order inside the slice is unchanged: by Gaussian (depth rank), then row-major over its rectangle.
__global__ void __launch_bounds__(CS_THREADS) emit_in_depth_order_kernel(const uint32_t* __restrict__ tiles_touched,
                                                                         const uint32_t* __restrict__ order, int n_cap,
                                                                         const uint32_t* __restrict__ n_ptr,
                                                                         const uint2* __restrict__ tile_offsets,
                                                                         const float4* __restrict__ xy_conic_ab,
                                                                         const int* __restrict__ radii, int tiles_x, int tiles_y,
                                                                         uint32_t* __restrict__ tile_keys,
                                                                         uint32_t* __restrict__ vals) {
    __shared__ uint2 sw[8];
    const int n = n_ptr ? (int)min(*n_ptr, (uint32_t)n_cap) : n_cap;
    if ((int)(blockIdx.x * EMIT_TILE) >= n) return;
    const int rank = blockIdx.x * EMIT_TILE + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t id = (rank < n) ? order[rank] : 0u;
    const uint32_t count = (rank < n) ? tiles_touched[id] : 0u;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (count != 0u) {
        const float4 rec = xy_conic_ab[id];
        tile_rect(rec.x, rec.y, radii[id], tiles_x, tiles_y, x0, y0, x1, y1);
    }
    uint2 total;
    const uint32_t off = block_exclusive_scan2(make_uint2(count, 0u), sw, total).x + tile_offsets[blockIdx.x].x;
    // the warp's slice: [base, base + warp_total); `rel` = this lane's exclusive offset inside it
    const uint32_t base = __shfl_sync(0xffffffffu, off, 0);
    const uint32_t rel = off - base;
    const uint32_t warp_total = __shfl_sync(0xffffffffu, rel + count, 31);
    const int rect_w = x1 - x0;
    for (uint32_t j0 = 0; j0 < warp_total; j0 += 32) {       // warp-uniform trip count: every lane takes part in the shuffles
        const uint32_t j = j0 + (uint32_t)lane;
        const bool active = j < warp_total;
        // owner = the largest lane whose exclusive offset is <= j (lanes with no instances share their successor's offset and
        // can only be picked when they are followed by no instance at all, which j < warp_total excludes)
        int owner = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = owner + step;
            const uint32_t r = __shfl_sync(0xffffffffu, rel, cand & 31);
            if (cand < 32 && r <= j) owner = cand;
        }
        const uint32_t o_rel = __shfl_sync(0xffffffffu, rel, owner);
        const int o_x0 = __shfl_sync(0xffffffffu, x0, owner), o_y0 = __shfl_sync(0xffffffffu, y0, owner);
        const int o_w = __shfl_sync(0xffffffffu, rect_w, owner);
        const uint32_t o_id = __shfl_sync(0xffffffffu, id, owner);
        if (active) {
            const uint32_t m = j - o_rel;
            const uint32_t row = m / (uint32_t)o_w, col = m - row * (uint32_t)o_w;
            tile_keys[base + j] = (uint32_t)((o_y0 + (int)row) * tiles_x + o_x0 + (int)col);
            vals[base + j] = o_id;
        }
    }
}

}  // namespace

size_t compact_temp_bytes(int n) {
    const size_t tiles = (size_t)((n > 0 ? n : 1) + EMIT_TILE - 1) / EMIT_TILE;   // the finer of the two tilings
    return (tiles + 1) * sizeof(uint2) + 128;
}

// stage 1: point_offsets (inclusive scan of tiles_touched) + stable compaction of the visible Gaussians to (depth bits, index)
// pairs + the control block bin_info (R, V, discard flags; the caller has zeroed it before the preprocess kernel)
int binning_df_compact(int P, const GeomState& g, uint32_t* keys, uint32_t* vals, uint32_t* bin_info, uint32_t R_capacity,
                       cudaStream_t s) {
    if (P <= 0) return GS2M_OK;
    const int tiles = (P + CS_TILE - 1) / CS_TILE;
    uint2* tile_sums = reinterpret_cast<uint2*>(g.scan_temp);
    count_launches(3);
    compact_tile_sums_kernel<<<tiles, CS_THREADS, 0, s>>>(g.tiles_touched, P, tile_sums);
    compact_spine_kernel<true><<<1, CS_THREADS, 0, s>>>(tile_sums, P, nullptr, CS_TILE, bin_info, R_capacity);
    compact_apply_kernel<<<tiles, CS_THREADS, 0, s>>>(g.tiles_touched, g.depths, P, tile_sums, g.point_offsets, keys, vals);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

// stage 3: instances of the depth-ordered visible Gaussians -> (tile id, index) pairs.  The grids cover V_cap Gaussians; the
// kernels read the real count from n_ptr (device) when given.
int binning_df_emit(int V_cap, const uint32_t* n_ptr, const GeomState& g, const uint32_t* order, const int* radii, int tiles_x,
                    int tiles_y, uint32_t* tile_keys, uint32_t* vals, cudaStream_t s) {
    if (V_cap <= 0) return GS2M_OK;
    const int tiles = (V_cap + EMIT_TILE - 1) / EMIT_TILE;
    uint2* tile_sums = reinterpret_cast<uint2*>(g.scan_temp);
    count_launches(3);
    ordered_tile_sums_kernel<<<tiles, CS_THREADS, 0, s>>>(g.tiles_touched, order, V_cap, n_ptr, tile_sums);
    compact_spine_kernel<false><<<1, CS_THREADS, 0, s>>>(tile_sums, V_cap, n_ptr, EMIT_TILE, nullptr, 0u);
    emit_in_depth_order_kernel<<<tiles, CS_THREADS, 0, s>>>(g.tiles_touched, order, V_cap, n_ptr, tile_sums, g.xy_conic_ab, radii,
                                                            tiles_x, tiles_y, tile_keys, vals);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace gs2m
