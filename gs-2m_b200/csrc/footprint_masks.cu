// Footprint masks and per-warp-block lists: for every entry of every tile list, which of the tile's eight 8x4-pixel warp
// blocks can the Gaussian reach with alpha >= 1/255 (conservative test of blend_common.cuh)?  Computed ONCE per
// (Gaussian, tile) instance after binning, stored as one byte per instance, and then used to compact the instance list into
// one dense list per warp block, which is what the forward and the backward blend kernels walk (see dense_fill_kernel).
// (Round 1 had each of the 8 warps gather every record and redo the test: 18 % of the forward's instructions and a third of
// its stall samples; round 2a filtered the tile list by the mask byte inside the blend kernels; now the filter runs once.)
#include "blend_common.cuh"

namespace gs2m {
namespace {

// One thread per instance, fused with the tile-range identification
// (identifyTileRanges, rasterizer_impl.cu:108-129): the sorted key gives the tile, the sorted value the Gaussian.
// KeyT = uint64_t: the reference's (tile << 32 | depth) keys.  KeyT = uint32_t: bare tile ids of the depth-first binning
// path; the 64-bit key of every instance is then re-materialised into keys64_out (the parity surface of the sort).
constexpr int RM_ITEMS = 2;      // instances per thread: the record gathers of both are in flight together (the kernel is
                                 // bound by the latency of those gathers, ncu: long_scoreboard 54 % with one instance per thread)
template <typename KeyT>
__global__ void __launch_bounds__(256) ranges_and_masks_kernel(int R_cap, const uint32_t* __restrict__ n_ptr, int tiles_x,
                                                               const KeyT* __restrict__ keys,
                                                               const uint32_t* __restrict__ point_list,
                                                               const float4* __restrict__ rec_a, const float4* __restrict__ rec_b,
                                                               const float* __restrict__ depths, uint64_t* __restrict__ keys64_out,
                                                               uint2* __restrict__ ranges, uint8_t* __restrict__ masks,
                                                               uint32_t* __restrict__ block_totals, int n_blocks_cap) {
    const int R = n_ptr ? (int)min(*n_ptr, (uint32_t)R_cap) : R_cap;   // grid covers the capacity, the count lives on the device
    constexpr int SHIFT = sizeof(KeyT) == 8 ? 32 : 0;
    const int i0 = blockIdx.x * (256 * RM_ITEMS) + threadIdx.x;
    if (blockIdx.x * (256 * RM_ITEMS) >= R) return;
    __shared__ uint32_t s_tot[BLEND_WARPS];
    if (threadIdx.x < BLEND_WARPS) s_tot[threadIdx.x] = 0u;
    __syncthreads();
    uint32_t my_count = 0;           // lane w of a warp counts the instances of its chunks whose mask has bit w
    uint32_t tile[RM_ITEMS], prev[RM_ITEMS], gid[RM_ITEMS];
    float4 ra[RM_ITEMS], rb[RM_ITEMS];
    float dep[RM_ITEMS];
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) {
        const int i = i0 + k * 256;
        tile[k] = prev[k] = gid[k] = 0u;
        if (i < R) {
            tile[k] = (uint32_t)(keys[i] >> SHIFT);
            prev[k] = (i > 0) ? (uint32_t)(keys[i - 1] >> SHIFT) : 0u;
            gid[k] = point_list[i];
        }
    }
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) {
        const int i = i0 + k * 256;
        ra[k] = rb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        dep[k] = 0.f;
        if (i < R) {
            ra[k] = __ldg(rec_a + gid[k]);
            rb[k] = __ldg(rec_b + gid[k]);
            if (sizeof(KeyT) == 4) dep[k] = __ldg(depths + gid[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) {
        const int i = i0 + k * 256;
        uint32_t m = 0u;
        if (i < R) {
            if (i == 0) {
                ranges[tile[k]].x = 0;
            } else if (prev[k] != tile[k]) {
                ranges[prev[k]].y = (uint32_t)i;
                ranges[tile[k]].x = (uint32_t)i;
            }
            if (i == R - 1) ranges[tile[k]].y = (uint32_t)R;
            if (sizeof(KeyT) == 4) keys64_out[i] = ((uint64_t)tile[k] << 32) | (uint64_t)__float_as_uint(dep[k]);
            const CullRecord cr = make_cull_record(ra[k], rb[k]);
            const int ty = (int)(tile[k] / (uint32_t)tiles_x), tx = (int)(tile[k] - (uint32_t)ty * (uint32_t)tiles_x);
            m = warp_block_mask(cr, tx * GS2M_TILE_X, ty * GS2M_TILE_Y);
            masks[i] = (uint8_t)m;
        }
        // per-block totals of every mask bit: the first level of the scan behind the per-warp-block lists (dense_fill_kernel)
#pragma unroll
        for (int w = 0; w < BLEND_WARPS; ++w) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (m >> w) & 1u);
            if ((threadIdx.x & 31) == w) my_count += __popc(bal);
        }
    }
    if ((threadIdx.x & 31) < BLEND_WARPS) atomicAdd(&s_tot[threadIdx.x & 31], my_count);
    __syncthreads();
    if (threadIdx.x < BLEND_WARPS) block_totals[(size_t)threadIdx.x * n_blocks_cap + blockIdx.x] = s_tot[threadIdx.x];
}

// ---- per-warp-block lists ----
// A blend warp owns one 8x4 pixel block of a tile and only ever blends the list entries whose footprint mask has its bit:
// 19 % of the entries on the benchmark scene.  Walking the tile's list with the mask as a filter cost the blend kernels 15 %
// (forward) / 10 % (backward) of their instructions and a fifth of their stall samples in per-32-entries bookkeeping that found
// 7 hits on average.  So the filter is applied once, here: for each of the eight block positions w the whole instance list is
// compacted (stable, so every per-pixel blending order is untouched) by mask bit w into dense_gid[w][.]; the list of
// (tile t, block w) is the slice [block_ranges[t][w].x, .y) of it.  dense_pos carries each entry's position in its tile list,
// which the forward needs once per pixel to report n_contrib in the reference's coordinates.
// Three small kernels: per-block bit totals (above, inside the mask kernel), their exclusive scan (8 sequences of R/512
// values), and the fill, which redoes the ballots of its 512 instances and scatters.
constexpr int DL_BLOCK = 256 * RM_ITEMS;      // instances per block of the mask kernel and of the fill kernel

__global__ void __launch_bounds__(1024) dense_scan_kernel(int R_cap, const uint32_t* __restrict__ n_ptr, uint32_t* __restrict__ block_totals,
                                                          int n_blocks_cap) {
    // block w: exclusive scan (in place) of block_totals[w][0 .. n_blocks)
    const int R = n_ptr ? (int)min(*n_ptr, (uint32_t)R_cap) : R_cap;
    const int n_blocks = (R + DL_BLOCK - 1) / DL_BLOCK;
    uint32_t* __restrict__ v = block_totals + (size_t)blockIdx.x * n_blocks_cap;
    __shared__ uint32_t s_part[1024];
    const int per = (n_blocks + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(n_blocks, b0 + per);
    uint32_t sum = 0;
    for (int b = b0; b < b1; ++b) sum += v[b];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {         // Hillis-Steele over the 1024 partial sums
        const uint32_t t = (threadIdx.x >= o) ? s_part[threadIdx.x - o] : 0u;
        __syncthreads();
        s_part[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = s_part[threadIdx.x] - sum;     // exclusive prefix of this thread's run
    for (int b = b0; b < b1; ++b) {
        const uint32_t t = v[b];
        v[b] = run;
        run += t;
    }
}

__global__ void __launch_bounds__(256) dense_fill_kernel(int R_cap, const uint32_t* __restrict__ n_ptr, const uint64_t* __restrict__ keys64,
                                                         const uint32_t* __restrict__ point_list, const uint8_t* __restrict__ masks,
                                                         const uint2* __restrict__ ranges, const uint32_t* __restrict__ block_offsets,
                                                         int n_blocks_cap, uint32_t* __restrict__ dense_gid,
                                                         uint32_t* __restrict__ dense_pos, uint2* __restrict__ block_ranges) {
    const int R = n_ptr ? (int)min(*n_ptr, (uint32_t)R_cap) : R_cap;
    if (blockIdx.x * DL_BLOCK >= R) return;
    constexpr int CHUNKS = DL_BLOCK / 32;
    __shared__ uint32_t s_base[CHUNKS][BLEND_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t bal[RM_ITEMS][BLEND_WARPS], m[RM_ITEMS], tile[RM_ITEMS], prev[RM_ITEMS], gid[RM_ITEMS], start[RM_ITEMS];
    // every global load of the thread is issued before anything waits on one (the kernel is latency-, not bandwidth-sized)
    // (the high words of the 64-bit keys are the tile ids; only they are read)
    const uint32_t* __restrict__ key_hi = reinterpret_cast<const uint32_t*>(keys64) + 1;
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) {
        const int i = blockIdx.x * DL_BLOCK + k * 256 + threadIdx.x;
        m[k] = tile[k] = prev[k] = gid[k] = 0u;
        if (i < R) {
            m[k] = masks[i];
            tile[k] = key_hi[2 * (size_t)i];
            prev[k] = (i > 0) ? key_hi[2 * (size_t)i - 2] : 0u;
            gid[k] = point_list[i];
        }
    }
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) start[k] = ranges[tile[k]].x;
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) {
#pragma unroll
        for (int w = 0; w < BLEND_WARPS; ++w) {
            bal[k][w] = __ballot_sync(0xffffffffu, (m[k] >> w) & 1u);
            if (lane == w) s_base[k * 8 + warp][w] = __popc(bal[k][w]);
        }
    }
    __syncthreads();
    if (threadIdx.x < BLEND_WARPS) {        // exclusive scan over the block's chunks (instance order: k major, warp minor)
        uint32_t run = block_offsets[(size_t)threadIdx.x * n_blocks_cap + blockIdx.x];
        for (int c = 0; c < CHUNKS; ++c) {
            const uint32_t t = s_base[c][threadIdx.x];
            s_base[c][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RM_ITEMS; ++k) {
        const int i = blockIdx.x * DL_BLOCK + k * 256 + threadIdx.x;
        if (i >= R) continue;
        const bool first = (i == 0) || (prev[k] != tile[k]), last = (i == R - 1);
        if (m[k]) {
            const uint32_t pos = (uint32_t)i - start[k];
#pragma unroll
            for (int w = 0; w < BLEND_WARPS; ++w) {
                if ((m[k] >> w) & 1u) {
                    const size_t dst = (size_t)w * R_cap + s_base[k * 8 + warp][w] + __popc(bal[k][w] & lt);   // rank among bit w
                    dense_gid[dst] = gid[k];
                    dense_pos[dst] = pos;
                }
            }
        }
        if (first || last) {      // a tile's slices start / end where the running counts stand at its first / behind its last entry
#pragma unroll
            for (int w = 0; w < BLEND_WARPS; ++w) {
                const uint32_t before = s_base[k * 8 + warp][w] + __popc(bal[k][w] & lt);
                if (first) {
                    block_ranges[(size_t)tile[k] * BLEND_WARPS + w].x = before;
                    if (i > 0) block_ranges[(size_t)prev[k] * BLEND_WARPS + w].y = before;
                }
                if (last) block_ranges[(size_t)tile[k] * BLEND_WARPS + w].y = before + ((m[k] >> w) & 1u);
            }
        }
    }
}

// Longest lists first.  The blend kernels' CTAs take the tiles in this order (the hardware starts CTAs in index order), so the
// tiles that run longest start first and the short ones fill in behind them.  It matters when list lengths are very unequal AND
// pixels do not saturate early (e.g. after GS-2M's periodic opacity reset to 0.01, when every pixel walks its whole list):
// a 292x-the-mean list that starts in the middle of the grid would otherwise finish long after everything else.  Exact order
// is not needed: tiles are bucketed by the bit length of their list length (one CTA, a 33-bin counting sort, 13 us).
// Measured against row-major order (B200, blend forward / backward): DTU-shaped 300 k scene (lists up to 5.5x the mean)
// 0.172 / 0.253 -> 0.138 / 0.214 ms; clustered scene after an opacity reset 2.74 / 2.78 -> 2.66 / 2.57 ms; uniform 3 M scene
// 1.116 / 1.893 -> 1.099 / 1.883 ms.
__global__ void __launch_bounds__(1024) tile_order_kernel(int n_tiles, const uint2* __restrict__ ranges,
                                                          uint32_t* __restrict__ order) {
    __shared__ uint32_t s_count[33], s_start[33];
    if (threadIdx.x < 33) s_count[threadIdx.x] = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += 1024) {
        const uint2 r = ranges[t];
        atomicAdd(&s_count[32 - __clz(r.y - r.x)], 1u);       // bucket = bit length of the list length (0 for an empty tile)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = 32; b >= 0; --b) { s_start[b] = run; run += s_count[b]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_tiles; t += 1024) {
        const uint2 r = ranges[t];
        order[atomicAdd(&s_start[32 - __clz(r.y - r.x)], 1u)] = (uint32_t)t;
    }
}

static int build_dense_lists(int R_cap, const uint32_t* n_ptr, int tiles_x, int tiles_y, const BinState& b, const ImageState& im,
                            cudaStream_t s) {
    const int n_blocks_cap = (R_cap + DL_BLOCK - 1) / DL_BLOCK;
    GS2M_CUDA(cudaMemsetAsync(im.block_ranges, 0, (size_t)tiles_x * tiles_y * BLEND_WARPS * sizeof(uint2), s));
    count_launches(2);
    dense_scan_kernel<<<BLEND_WARPS, 1024, 0, s>>>(R_cap, n_ptr, b.dense_block_totals, n_blocks_cap);
    dense_fill_kernel<<<n_blocks_cap, 256, 0, s>>>(R_cap, n_ptr, b.keys_sorted, b.point_list, b.masks, im.ranges, b.dense_block_totals,
                                                  n_blocks_cap, b.dense_gid, b.dense_pos, im.block_ranges);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace

int launch_tile_order(int n_tiles, const uint2* ranges, uint32_t* tile_order, cudaStream_t s) {
    count_launches(1);
    tile_order_kernel<<<1, 1024, 0, s>>>(n_tiles, ranges, tile_order);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int launch_ranges_and_masks(int R, int tiles_x, int tiles_y, const BinState& b, const GeomState& g, const ImageState& im, cudaStream_t s) {
    GS2M_CUDA(cudaMemsetAsync(im.ranges, 0, (size_t)tiles_x * tiles_y * sizeof(uint2), s));
    if (R > 0) {
        const int n_blocks_cap = (R + DL_BLOCK - 1) / DL_BLOCK;
        count_launches(1);
        ranges_and_masks_kernel<uint64_t><<<n_blocks_cap, 256, 0, s>>>(R, nullptr, tiles_x, b.keys_sorted, b.point_list, g.xy_conic_ab,
                                                                      g.conic_c_opac, nullptr, nullptr, im.ranges, b.masks,
                                                                      b.dense_block_totals, n_blocks_cap);
        GS2M_CUDA(cudaGetLastError());
        return build_dense_lists(R, nullptr, tiles_x, tiles_y, b, im, s);
    }
    GS2M_CUDA(cudaMemsetAsync(im.block_ranges, 0, (size_t)tiles_x * tiles_y * BLEND_WARPS * sizeof(uint2), s));
    return GS2M_OK;
}

// depth-first binning: sorted 32-bit tile ids in, ranges + masks + the 64-bit (tile | depth) keys + the per-warp-block lists out
int launch_ranges_masks_keys(int R_cap, const uint32_t* n_ptr, int tiles_x, int tiles_y, const uint32_t* tile_keys_sorted,
                             const BinState& b, const GeomState& g, const ImageState& im, cudaStream_t s) {
    GS2M_CUDA(cudaMemsetAsync(im.ranges, 0, (size_t)tiles_x * tiles_y * sizeof(uint2), s));
    if (R_cap > 0) {
        const int n_blocks_cap = (R_cap + DL_BLOCK - 1) / DL_BLOCK;
        count_launches(1);
        ranges_and_masks_kernel<uint32_t><<<n_blocks_cap, 256, 0, s>>>(R_cap, n_ptr, tiles_x, tile_keys_sorted, b.point_list,
                                                                      g.xy_conic_ab, g.conic_c_opac, g.depths, b.keys_sorted,
                                                                      im.ranges, b.masks, b.dense_block_totals, n_blocks_cap);
        GS2M_CUDA(cudaGetLastError());
        return build_dense_lists(R_cap, n_ptr, tiles_x, tiles_y, b, im, s);
    }
    GS2M_CUDA(cudaMemsetAsync(im.block_ranges, 0, (size_t)tiles_x * tiles_y * BLEND_WARPS * sizeof(uint2), s));
    return GS2M_OK;
}

}  // namespace gs2m
