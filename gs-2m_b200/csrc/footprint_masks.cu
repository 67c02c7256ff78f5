// Footprint masks: for every entry of every tile list, which of the tile's eight 8x4-pixel warp blocks can the
// Gaussian reach with alpha >= 1/255 (conservative test of blend_common.cuh)?  Computed ONCE per (Gaussian, tile)
// instance after binning and stored as one byte per instance; the forward and the backward blend kernels then read
// 32 bytes per step instead of each of the 8 warps gathering every record and redoing the test (an ncu source-level
// profile showed 18 % of the forward's instructions and a third of its stall samples in that per-warp gather + test).
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
                                                               uint2* __restrict__ ranges, uint8_t* __restrict__ masks) {
    const int R = n_ptr ? (int)min(*n_ptr, (uint32_t)R_cap) : R_cap;   // grid covers the capacity, the count lives on the device
    constexpr int SHIFT = sizeof(KeyT) == 8 ? 32 : 0;
    const int i0 = blockIdx.x * (256 * RM_ITEMS) + threadIdx.x;
    if (blockIdx.x * (256 * RM_ITEMS) >= R) return;
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
        if (i >= R) continue;
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
        masks[i] = (uint8_t)warp_block_mask(cr, tx * GS2M_TILE_X, ty * GS2M_TILE_Y);
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

}  // namespace

int launch_tile_order(int n_tiles, const uint2* ranges, uint32_t* tile_order, cudaStream_t s) {
    count_launches(1);
    tile_order_kernel<<<1, 1024, 0, s>>>(n_tiles, ranges, tile_order);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

int launch_ranges_and_masks(int R, int tiles_x, int tiles_y, const uint64_t* keys_sorted, const uint32_t* point_list,
                            const GeomState& g, uint2* ranges, uint8_t* masks, cudaStream_t s) {
    GS2M_CUDA(cudaMemsetAsync(ranges, 0, (size_t)tiles_x * tiles_y * sizeof(uint2), s));
    if (R > 0) {
        count_launches(1);
        ranges_and_masks_kernel<uint64_t><<<(R + 256 * RM_ITEMS - 1) / (256 * RM_ITEMS), 256, 0, s>>>(R, nullptr, tiles_x, keys_sorted, point_list, g.xy_conic_ab,
                                                                          g.conic_c_opac, nullptr, nullptr, ranges, masks);
        GS2M_CUDA(cudaGetLastError());
    }
    return GS2M_OK;
}

// depth-first binning: sorted 32-bit tile ids in, ranges + masks + the 64-bit (tile | depth) keys out
int launch_ranges_masks_keys(int R_cap, const uint32_t* n_ptr, int tiles_x, int tiles_y, const uint32_t* tile_keys_sorted,
                             const uint32_t* point_list, const GeomState& g, uint64_t* keys64_out, uint2* ranges, uint8_t* masks,
                             cudaStream_t s) {
    GS2M_CUDA(cudaMemsetAsync(ranges, 0, (size_t)tiles_x * tiles_y * sizeof(uint2), s));
    if (R_cap > 0) {
        count_launches(1);
        ranges_and_masks_kernel<uint32_t><<<(R_cap + 256 * RM_ITEMS - 1) / (256 * RM_ITEMS), 256, 0, s>>>(R_cap, n_ptr, tiles_x, tile_keys_sorted, point_list,
                                                                              g.xy_conic_ab, g.conic_c_opac, g.depths, keys64_out,
                                                                              ranges, masks);
        GS2M_CUDA(cudaGetLastError());
    }
    return GS2M_OK;
}

}  // namespace gs2m
