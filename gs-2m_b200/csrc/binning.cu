// Binning stage: inclusive scan of tiles_touched, per-tile key duplication, stable LSD radix sort of
// (tile|depth) keys, and tile-range identification.
//
// Behavioural reference: cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:265), duplicateWithKeys (:63-103),
// cub::DeviceRadixSort::SortPairs on bits [0, 32+bit_length(n_tiles)) (:288-296), identifyTileRanges (:108-129).
// All integer work: the outputs are bit-identical to the reference's (stable sort, ties resolved by the emission
// order = ascending Gaussian index, then row-major tile order).
//
// The radix sort is a hand-written single-pass-per-digit ("onesweep") sort, templated on 32- and 64-bit keys: one
// up-front histogram kernel over all digits, then per 8-bit digit one kernel that ranks a tile of keys with warp
// match-any, resolves its global base by decoupled look-back over a per-tile status array, stages the tile in sorted
// order in shared memory and copies it out so that every digit leaves as a coalesced run.  The work is pure HBM/L2
// streaming, the roofline DESIGN.md charges it to.  The default binning path (binning_depthfirst.cu) uses the 32-bit
// instance twice (depth bits of the visible Gaussians, then tile ids of the instances); the scan, duplication and
// 64-bit sort of this file are the reference-shaped "sort64" path and the building blocks the C-ABI exports.
#include "common.cuh"

namespace gs2m {
namespace {

// ------------------------------------------------------------------ inclusive scan (u32) -----------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, writes total to `total`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem_warp /*[>=8]*/, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_inclusive_scan(v, lane);
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) {
        const uint32_t s = smem_warp[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint32_t* __restrict__ in, int n,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t sw[8];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += in[base + i];
    uint32_t total;
    block_exclusive_scan(s, sw, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums in place
__global__ void __launch_bounds__(1024) scan_spine_kernel(uint32_t* __restrict__ tile_sums, int n_tiles) {
    __shared__ uint32_t sw[32];
    uint32_t carry = 0;
    for (int start = 0; start < n_tiles; start += 1024) {
        const int i = start + threadIdx.x;
        const uint32_t v = (i < n_tiles) ? tile_sums[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, sw, total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                  int n, const uint32_t* __restrict__ tile_offsets) {
    __shared__ uint32_t sw[8];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        s += v[i];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, sw, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        run += v[i];
        if (base + i < n) out[base + i] = run;
    }
}

// ------------------------------------------------------------------ key duplication ---------------------------
__global__ void __launch_bounds__(256) duplicate_with_keys_kernel(int P, const float4* __restrict__ xy_conic_ab,
                                                                  const float* __restrict__ depths,
                                                                  const uint32_t* __restrict__ offsets,
                                                                  const int* __restrict__ radii, int tiles_x, int tiles_y,
                                                                  uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const int radius = radii[idx];
    if (radius <= 0) return;
    uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
    const float4 rec = xy_conic_ab[idx];
    int x0, y0, x1, y1;
    tile_rect(rec.x, rec.y, radius, tiles_x, tiles_y, x0, y0, x1, y1);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(depths[idx]);
    for (int y = y0; y < y1; ++y) {
        for (int x = x0; x < x1; ++x) {
            const uint64_t tile = (uint64_t)(uint32_t)(y * tiles_x + x);
            keys[off] = (tile << 32) | depth_bits;
            vals[off] = (uint32_t)idx;
            ++off;
        }
    }
}

// ------------------------------------------------------------------ tile ranges --------------------------------
__global__ void __launch_bounds__(256) identify_tile_ranges_kernel(int R, const uint64_t* __restrict__ keys,
                                                                   uint2* __restrict__ ranges) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t tile = (uint32_t)(keys[i] >> 32);
    if (i == 0) {
        ranges[tile].x = 0;
    } else {
        const uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (prev != tile) {
            ranges[prev].y = (uint32_t)i;
            ranges[tile].x = (uint32_t)i;
        }
    }
    if (i == R - 1) ranges[tile].y = (uint32_t)R;
}

// ------------------------------------------------------------------ onesweep radix sort ------------------------
// number of items a count-dependent kernel processes: the host-known capacity, clamped by the device-side count when given
__device__ __forceinline__ int device_count(int n_cap, const uint32_t* __restrict__ n_ptr) {
    return n_ptr ? (int)min(*n_ptr, (uint32_t)n_cap) : n_cap;
}

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 12;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // 3072 keys per CTA
constexpr int RS_RADIX = 256;
constexpr int RS_MAX_PASSES = 8;
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_VALUE_MASK = (1u << 30) - 1;

// all digit histograms in one pass over the keys
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_histogram_kernel(const KeyT* __restrict__ keys, int n_cap,
                                                                  const uint32_t* __restrict__ n_ptr, int n_passes,
                                                                  int end_bit, uint32_t* __restrict__ hist /*[passes][256]*/) {
    __shared__ uint32_t sh[RS_MAX_PASSES * RS_RADIX];
    const int n = device_count(n_cap, n_ptr);
    if ((int)(blockIdx.x * RS_THREADS) >= n) return;
    for (int i = threadIdx.x; i < n_passes * RS_RADIX; i += RS_THREADS) sh[i] = 0;
    __syncthreads();
    const int stride = gridDim.x * RS_THREADS;
    for (int i = blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        const KeyT k = keys[i];
        for (int p = 0; p < n_passes; ++p) {
            const int bits = min(8, end_bit - 8 * p);
            const uint32_t d = (uint32_t)(k >> (8 * p)) & ((1u << bits) - 1u);
            atomicAdd(&sh[p * RS_RADIX + d], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_passes * RS_RADIX; i += RS_THREADS)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// One digit pass. Keys are held warp-striped: warp w owns keys [w*32*ITEMS, (w+1)*32*ITEMS) of the tile and lane l
// holds items l, l+32, ...; stable order inside the tile is therefore (warp, item, lane).
template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS) rs_onesweep_kernel(const KeyT* __restrict__ keys_in,
                                                                 KeyT* __restrict__ keys_out,
                                                                 const uint32_t* __restrict__ vals_in,
                                                                 uint32_t* __restrict__ vals_out, int n_cap,
                                                                 const uint32_t* __restrict__ n_ptr, int shift, int bits,
                                                                 const uint32_t* __restrict__ digit_hist /*[256] counts of this digit*/,
                                                                 volatile uint32_t* __restrict__ status /*[tiles][256]*/,
                                                                 uint32_t* __restrict__ ticket) {
    // The grid covers the capacity; CTAs beyond the real count leave before taking a ticket, so tickets 0..ceil(n/TILE)-1 are
    // taken by CTAs that do the work (in arrival order) and the look-back never waits for a CTA that exits.
    const int n = device_count(n_cap, n_ptr);
    if ((long long)blockIdx.x * RS_TILE >= (long long)n) return;
    __shared__ uint32_t s_warp_hist[RS_WARPS][RS_RADIX];
    __shared__ uint32_t s_base[RS_RADIX];       // global index of the tile's first key of a digit, minus its tile-local offset
    __shared__ uint32_t s_tile_off[RS_RADIX];   // tile-local offset of a digit's keys in the staged order
    __shared__ uint32_t s_scan[8];
    __shared__ KeyT s_keys[RS_TILE];            // the tile in sorted order: scattered here first, then copied out so that
    __shared__ uint32_t s_vals[RS_TILE];        // keys of one digit leave as contiguous runs (coalesced sectors)
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t mask = (1u << bits) - 1u;
    const int tile_start = (int)tile * RS_TILE + warp * 32 * RS_ITEMS;

    KeyT key[RS_ITEMS];
    uint32_t val[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int g = tile_start + i * 32 + lane;
        key[i] = (g < n) ? keys_in[g] : (KeyT)~(KeyT)0;
        val[i] = (g < n) ? vals_in[g] : 0u;
    }
    // rank inside the warp, item by item (stable)
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int g = tile_start + i * 32 + lane;
        const bool valid = g < n;
        const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        // lanes with the same digit (invalid lanes get a private pseudo-digit so they never match a valid one)
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : (0x10000u + lane));
        const int leader = __ffs(peers) - 1;
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = s_warp_hist[warp][d];
            s_warp_hist[warp][d] = prev + __popc(peers);
        }
        prev = __shfl_sync(0xffffffffu, prev, leader);
        rank[i] = prev + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();

    // per-digit: exclusive prefix over warps, tile total, decoupled look-back for the global base
    {
        const int d = threadIdx.x;  // RS_THREADS == RS_RADIX
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = s_warp_hist[w][d];
            s_warp_hist[w][d] = run;
            run += c;
        }
        const uint32_t count = run;
        volatile uint32_t* st = status + (size_t)tile * RS_RADIX + d;
        uint32_t exclusive = 0;
        if (tile == 0) {
            *st = RS_FLAG_PREFIX | count;
        } else {
            *st = RS_FLAG_AGG | count;
            int j = (int)tile - 1;
            while (true) {
                const uint32_t v = status[(size_t)j * RS_RADIX + d];
                const uint32_t f = v & ~RS_VALUE_MASK;
                if (f == 0) continue;  // predecessor not published yet
                exclusive += v & RS_VALUE_MASK;
                if (f == RS_FLAG_PREFIX) break;
                --j;                   // aggregate only: keep walking (tile 0 always publishes PREFIX)
            }
            *st = RS_FLAG_PREFIX | (exclusive + count);
        }
        uint32_t tile_total;
        const uint32_t tile_off = block_exclusive_scan(count, s_scan, tile_total);
        // first output slot of digit d = exclusive scan of the digit's global histogram (every CTA redoes this 256-element scan:
        // cheaper than a separate one-block kernel between the histogram and the first pass)
        const uint32_t digit_base = block_exclusive_scan(digit_hist[d], s_scan, tile_total);
        s_tile_off[d] = tile_off;
        s_base[d] = digit_base + exclusive - tile_off;
    }
    __syncthreads();

    // stage the tile in sorted order in shared memory ...
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int g = tile_start + i * 32 + lane;
        if (g < n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const uint32_t pos = s_tile_off[d] + s_warp_hist[warp][d] + rank[i];
            s_keys[pos] = key[i];
            s_vals[pos] = val[i];
        }
    }
    __syncthreads();
    // ... and copy it out: consecutive threads hold consecutive keys of a digit, which are consecutive in the output
    const int n_tile = min(RS_TILE, n - (int)tile * RS_TILE);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int pos = i * RS_THREADS + threadIdx.x;
        if (pos < n_tile) {
            const KeyT k = s_keys[pos];
            const uint32_t dst = s_base[(uint32_t)(k >> shift) & mask] + (uint32_t)pos;
            keys_out[dst] = k;
            vals_out[dst] = s_vals[pos];
        }
    }
}

inline int rs_num_tiles(int n) { return (n + RS_TILE - 1) / RS_TILE; }
inline int rs_num_passes(int end_bit) { return (end_bit + 7) / 8; }

}  // namespace

// ------------------------------------------------------------------ host launchers -----------------------------
size_t scan_temp_bytes(int n) {
    const size_t tiles = (size_t)(n + SCAN_TILE - 1) / SCAN_TILE;
    return (tiles + 1) * sizeof(uint32_t) + 128;
}

int inclusive_sum_u32(const uint32_t* in, uint32_t* out, int n, char* temp, cudaStream_t s) {
    if (n <= 0) return GS2M_OK;
    const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* tile_sums = reinterpret_cast<uint32_t*>(temp);
    count_launches(3);
    scan_tile_sums_kernel<<<tiles, SCAN_THREADS, 0, s>>>(in, n, tile_sums);
    scan_spine_kernel<<<1, 1024, 0, s>>>(tile_sums, tiles);
    scan_apply_kernel<<<tiles, SCAN_THREADS, 0, s>>>(in, out, n, tile_sums);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

size_t sort_temp_bytes(int n) {
    const size_t tiles = (size_t)rs_num_tiles(n > 0 ? n : 1);
    // histograms [8][256] + tickets [8] + status [8 passes][tiles][256]
    return (size_t)RS_MAX_PASSES * RS_RADIX * 4 + 128 + RS_MAX_PASSES * 4 + 128 +
           (size_t)RS_MAX_PASSES * tiles * RS_RADIX * 4 + 256;
}

// Sorts on key bits [0,end_bit). If `result_in_input` is non-null no final copy is made and it reports whether the
// sorted data ended in the *_in buffers (even number of digit passes) — the forward uses this to avoid a copy.
template <typename KeyT>
static int sort_pairs_pingpong_t(KeyT* keys_in, KeyT* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n,
                                 const uint32_t* n_ptr, int end_bit, char* temp, cudaStream_t s, int* result_in_input) {
    if (result_in_input) *result_in_input = 0;
    if (n <= 0) return GS2M_OK;
    if (end_bit <= 0 || end_bit > (int)(8 * sizeof(KeyT))) { set_error("sort_pairs_u64: end_bit %d outside 1..64", end_bit); return GS2M_ERR_INVALID_ARGUMENT; }
    if ((unsigned)n >= RS_VALUE_MASK) { set_error("sort_pairs_u64: %d pairs exceed the 30-bit look-back counters", n); return GS2M_ERR_TOO_LARGE; }
    const int passes = rs_num_passes(end_bit);
    const int tiles = rs_num_tiles(n);
    char* p = temp;
    uint32_t *hist, *tickets, *status;
    carve_array(p, hist, (size_t)RS_MAX_PASSES * RS_RADIX);
    carve_array(p, tickets, (size_t)RS_MAX_PASSES);
    carve_array(p, status, (size_t)passes * tiles * RS_RADIX);
    // one memset covers histograms, tickets and all status words (they are carved contiguously)
    GS2M_CUDA(cudaMemsetAsync(hist, 0, (size_t)((char*)(status + (size_t)passes * tiles * RS_RADIX) - (char*)hist), s));
    int hist_blocks = (n + RS_THREADS * 16 - 1) / (RS_THREADS * 16);
    if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
    count_launches(1 + passes);
    rs_histogram_kernel<KeyT><<<hist_blocks, RS_THREADS, 0, s>>>(keys_in, n, n_ptr, passes, end_bit, hist);
    KeyT* kbuf[2] = {keys_in, keys_out};
    uint32_t* vbuf[2] = {vals_in, vals_out};
    int src = 0;
    if (!(passes & 1) && result_in_input == nullptr) {
        // even number of passes and the caller insists on the *_out buffers: start from a copy in *_out
        GS2M_CUDA(cudaMemcpyAsync(keys_out, keys_in, (size_t)n * sizeof(KeyT), cudaMemcpyDeviceToDevice, s));
        GS2M_CUDA(cudaMemcpyAsync(vals_out, vals_in, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        src = 1;
    }
    for (int pass = 0; pass < passes; ++pass) {
        const int bits = (end_bit - 8 * pass) < 8 ? (end_bit - 8 * pass) : 8;
        rs_onesweep_kernel<KeyT><<<tiles, RS_THREADS, 0, s>>>(kbuf[src], kbuf[src ^ 1], vbuf[src], vbuf[src ^ 1], n, n_ptr, 8 * pass, bits,
                                                        hist + pass * RS_RADIX, status + (size_t)pass * tiles * RS_RADIX,
                                                        tickets + pass);
        src ^= 1;
    }
    GS2M_CUDA(cudaGetLastError());
    if (result_in_input) *result_in_input = (src == 0);
    return GS2M_OK;
}

int sort_pairs_u64_pingpong(uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n,
                            int end_bit, char* temp, cudaStream_t s, int* result_in_input) {
    return sort_pairs_pingpong_t<uint64_t>(keys_in, keys_out, vals_in, vals_out, n, nullptr, end_bit, temp, s, result_in_input);
}

// 32-bit keys (depth ranking of the Gaussians): same kernels, 4 digit passes.
int sort_pairs_u32_pingpong(uint32_t* keys_in, uint32_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n_cap,
                            const uint32_t* n_ptr, int end_bit, char* temp, cudaStream_t s, int* result_in_input) {
    return sort_pairs_pingpong_t<uint32_t>(keys_in, keys_out, vals_in, vals_out, n_cap, n_ptr, end_bit, temp, s, result_in_input);
}

int sort_pairs_u64(uint64_t* keys_in, uint64_t* keys_out, uint32_t* vals_in, uint32_t* vals_out, int n, int end_bit,
                   char* temp, cudaStream_t s) {
    return sort_pairs_u64_pingpong(keys_in, keys_out, vals_in, vals_out, n, end_bit, temp, s, nullptr);
}

int launch_duplicate_with_keys(int P, const GeomState& g, const int* radii, int tiles_x, int tiles_y, uint64_t* keys,
                               uint32_t* vals, cudaStream_t s) {
    if (P == 0) return GS2M_OK;
    count_launches(1);
    duplicate_with_keys_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.xy_conic_ab, g.depths, g.point_offsets, radii, tiles_x,
                                                               tiles_y, keys, vals);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int launch_identify_tile_ranges(int R, const uint64_t* keys_sorted, uint2* ranges, int n_tiles, cudaStream_t s) {
    GS2M_CUDA(cudaMemsetAsync(ranges, 0, (size_t)n_tiles * sizeof(uint2), s));
    if (R > 0) {
        count_launches(1);
        identify_tile_ranges_kernel<<<(R + 255) / 256, 256, 0, s>>>(R, keys_sorted, ranges);
        GS2M_CUDA(cudaGetLastError());
    }
    return GS2M_OK;
}

}  // namespace gs2m
