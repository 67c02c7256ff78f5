// Binning, B200 version: depth-rank the Gaussians once, then bin per tile.
//
// The reference sorts R (Gaussian, tile) instances on 64-bit (tile | depth) keys (rasterizer_impl.cu:63-103,288-296):
// with CUB that is 6 digit passes over 12-byte pairs, ~170 B of traffic per instance.  The same final order — by tile,
// then depth bits, then Gaussian index (stable sort of an index-ordered emission) — is produced here from far less
// data movement:
//   1. rank      stable radix sort of the P Gaussians on their 32-bit depth bits (culled ones carry 0xFFFFFFFF and sort to
//                the end): rank r orders by (depth, index).  4 passes over P pairs instead of 6 over R.
//   2. count     one atomic per instance into a per-tile counter; an exclusive scan of the T counters IS the tile-range
//                table (empty tiles keep (0,0) exactly like the reference's memset + identifyTileRanges, :298-305).
//   3. emit      each Gaussian claims a slot in every tile it touches (atomic cursor) and writes (rank << 32 | index)
//                there: unordered inside a tile, but ranks are unique ...
//   4. tile sort ... so a per-tile shared-memory radix sort on the rank makes the order deterministic and identical to
//                the reference's.  The sorted index list and the 64-bit keys ((tile << 32) | depth bits) are written once.
// ~32 B of traffic per instance.  Rectangles with more than 32 tiles are spread over the warp (the reference walks
// them with one thread).  Tiles whose list exceeds the shared-memory capacity make the caller fall back to the global
// 64-bit sort of binning.cu (same results).
#include <atomic>
#include "common.cuh"

namespace gs2m {
namespace {

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += n;
    }
    return v;
}

// visits every tile of the rectangles owned by the 32 lanes of a warp; small rectangles are walked by their own lane,
// large ones by the whole warp
template <typename Visit>
__device__ __forceinline__ void for_each_tile(bool visible, int x0, int y0, int x1, int y1, int tiles_x, uint32_t tag,
                                              Visit visit) {
    const int lane = threadIdx.x & 31;
    const int w = x1 - x0;
    const int cnt = visible ? w * (y1 - y0) : 0;
    if (cnt > 0 && cnt <= 32) {
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) visit((uint32_t)(y * tiles_x + x), tag);
    }
    uint32_t big = __ballot_sync(0xffffffffu, cnt > 32);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bw = __shfl_sync(0xffffffffu, w, src), bc = __shfl_sync(0xffffffffu, cnt, src);
        const uint32_t btag = __shfl_sync(0xffffffffu, tag, src);
        for (int t = lane; t < bc; t += 32) visit((uint32_t)((by0 + t / bw) * tiles_x + bx0 + t % bw), btag);
    }
}

__global__ void __launch_bounds__(256) tile_count_kernel(int P, const float4* __restrict__ rec_a, const int* __restrict__ radii,
                                                         int tiles_x, int tiles_y, uint32_t* __restrict__ counts) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int radius = (idx < P) ? radii[idx] : 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (radius > 0) {
        const float4 r = rec_a[idx];
        tile_rect(r.x, r.y, radius, tiles_x, tiles_y, x0, y0, x1, y1);
    }
    for_each_tile(radius > 0, x0, y0, x1, y1, tiles_x, 0u, [&](uint32_t tile, uint32_t) { atomicAdd(counts + tile, 1u); });
}

// single CTA: exclusive scan of the per-tile counts -> starts + ranges; info = {total, max count}
__global__ void __launch_bounds__(1024) tile_scan_kernel(int n_tiles, const uint32_t* __restrict__ counts,
                                                         uint32_t* __restrict__ starts, uint2* __restrict__ ranges,
                                                         uint32_t* __restrict__ info) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_max[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = 0, vmax = 0;
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t c = (t < n_tiles) ? counts[t] : 0u;
        vmax = max(vmax, c);
        const uint32_t inc = warp_incl_scan(c, lane);
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const uint32_t v = s_warp[w];
            if (w < warp) wbase += v;
            total += v;
        }
        __syncthreads();
        if (t < n_tiles) {
            const uint32_t start = carry + wbase + inc - c;
            starts[t] = start;
            ranges[t] = c ? make_uint2(start, start + c) : make_uint2(0u, 0u);
        }
        carry += total;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) s_max[warp] = vmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        for (int w = 0; w < 32; ++w) m = max(m, s_max[w]);
        info[0] = carry;
        info[1] = m;
    }
}

__global__ void __launch_bounds__(256) emit_ranked_kernel(int P, const uint32_t* __restrict__ order,
                                                          const float4* __restrict__ rec_a, const int* __restrict__ radii,
                                                          int tiles_x, int tiles_y, const uint32_t* __restrict__ starts,
                                                          uint32_t* __restrict__ cursor, uint64_t* __restrict__ tmp) {
    const int r = blockIdx.x * 256 + threadIdx.x;
    const uint32_t gid = (r < P) ? order[r] : 0u;
    const int radius = (r < P) ? radii[gid] : 0;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (radius > 0) {
        const float4 rec = rec_a[gid];
        tile_rect(rec.x, rec.y, radius, tiles_x, tiles_y, x0, y0, x1, y1);
    }
    // the 64-bit payload (rank << 32 | index) is rebuilt from the two broadcast words for warp-shared rectangles
    const uint32_t rank = (uint32_t)r;
    const int lane = threadIdx.x & 31;
    const int w = x1 - x0;
    const int cnt = (radius > 0) ? w * (y1 - y0) : 0;
    if (cnt > 0 && cnt <= 32) {
        const uint64_t payload = ((uint64_t)rank << 32) | gid;
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                const uint32_t tile = (uint32_t)(y * tiles_x + x);
                tmp[starts[tile] + atomicAdd(cursor + tile, 1u)] = payload;
            }
    }
    uint32_t big = __ballot_sync(0xffffffffu, cnt > 32);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int bx0 = __shfl_sync(0xffffffffu, x0, src), by0 = __shfl_sync(0xffffffffu, y0, src);
        const int bw = __shfl_sync(0xffffffffu, w, src), bc = __shfl_sync(0xffffffffu, cnt, src);
        const uint64_t payload = ((uint64_t)__shfl_sync(0xffffffffu, rank, src) << 32) | __shfl_sync(0xffffffffu, gid, src);
        for (int t = lane; t < bc; t += 32) {
            const uint32_t tile = (uint32_t)((by0 + t / bw) * tiles_x + bx0 + t % bw);
            tmp[starts[tile] + atomicAdd(cursor + tile, 1u)] = payload;
        }
    }
}

// One CTA per tile: stable LSD radix sort (8-bit digits) of the tile's (rank, index) pairs in shared memory.
// dynamic smem: u32 k[2][cap], v[2][cap]; u16 lrank[cap]; u32 whist[8][256]; u32 dbase[256]
__global__ void __launch_bounds__(256) tile_sort_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ tmp,
                                                        const float* __restrict__ depths, uint64_t* __restrict__ keys_out,
                                                        uint32_t* __restrict__ vals_out, int key_bits, int cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* kbuf = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* vbuf = kbuf + 2 * (size_t)cap;
    uint32_t* whist = vbuf + 2 * (size_t)cap;               // [8][256]
    uint32_t* dbase = whist + 8 * 256;                       // [256]
    uint16_t* lrank = reinterpret_cast<uint16_t*>(dbase + 256);
    __shared__ uint32_t s_scan[8];

    const uint32_t tile = blockIdx.x;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    if (n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += 256) {
        const uint64_t e = tmp[range.x + i];
        kbuf[i] = (uint32_t)(e >> 32);
        vbuf[i] = (uint32_t)e;
    }
    const int chunk = (((n + 7) >> 3) + 31) & ~31;           // contiguous, warp-ordered, multiple of 32
    const int wbeg = warp * chunk;
    int cur = 0;
    for (int shift = 0; shift < key_bits; shift += 8) {
        uint32_t* kin = kbuf + cur * cap;
        uint32_t* vin = vbuf + cur * cap;
        uint32_t* kout = kbuf + (cur ^ 1) * cap;
        uint32_t* vout = vbuf + (cur ^ 1) * cap;
        for (int i = tid; i < 8 * 256; i += 256) whist[i] = 0;
        __syncthreads();
        for (int it = 0; it < chunk; it += 32) {
            const int i = wbeg + it + lane;
            const bool valid = i < n;
            const uint32_t d = valid ? ((kin[i] >> shift) & 255u) : (0x1000u + lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            uint32_t prev = 0;
            if (valid && lane == leader) {
                prev = whist[warp * 256 + d];
                whist[warp * 256 + d] = prev + __popc(peers);
            }
            prev = __shfl_sync(0xffffffffu, prev, leader);
            if (valid) lrank[i] = (uint16_t)(prev + __popc(peers & ((1u << lane) - 1u)));
            __syncwarp();
        }
        __syncthreads();
        {   // per digit: exclusive prefix over warps, then exclusive scan over the 256 digit totals
            const int d = tid;
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = whist[w * 256 + d];
                whist[w * 256 + d] = run;
                run += c;
            }
            const uint32_t inc = warp_incl_scan(run, lane);
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            uint32_t wb = 0;
            for (int w = 0; w < warp; ++w) wb += s_scan[w];
            dbase[d] = wb + inc - run;
        }
        __syncthreads();
        for (int it = 0; it < chunk; it += 32) {
            const int i = wbeg + it + lane;
            if (i < n) {
                const uint32_t k = kin[i];
                const uint32_t d = (k >> shift) & 255u;
                const uint32_t dst = dbase[d] + whist[warp * 256 + d] + lrank[i];
                kout[dst] = k;
                vout[dst] = vin[i];
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    const uint32_t* vfin = vbuf + cur * cap;
    for (int i = tid; i < n; i += 256) {
        const uint32_t gid = vfin[i];
        vals_out[range.x + i] = gid;
        keys_out[range.x + i] = ((uint64_t)tile << 32) | (uint64_t)__float_as_uint(depths[gid]);
    }
}

}  // namespace

size_t tile_sort_smem_bytes(int cap) { return (size_t)cap * (16 + 2) + (8 * 256 + 256) * 4 + 16; }

// stage A (needs no instance buffer, runs before the host learns R): per-tile counts -> ranges, and the depth ranking
int binning2_rank_and_count(int P, const GeomState& g, const int* radii, int tiles_x, int tiles_y, const ImageState& im,
                            cudaStream_t s, const uint32_t** order_out) {
    const int n_tiles = tiles_x * tiles_y;
    GS2M_CUDA(cudaMemsetAsync(im.tile_counts, 0, (size_t)n_tiles * 2 * sizeof(uint32_t), s));   // counts + cursor
    count_launches(2);
    tile_count_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, g.xy_conic_ab, radii, tiles_x, tiles_y, im.tile_counts);
    tile_scan_kernel<<<1, 1024, 0, s>>>(n_tiles, im.tile_counts, im.tile_starts, im.ranges, im.bin_info);
    GS2M_CUDA(cudaGetLastError());
    int in_input = 0;
    const int rc = sort_pairs_u32_pingpong(g.depth_keys, g.depth_keys_alt, g.order_a, g.order_b, P, 32, g.rank_temp, s, &in_input);
    if (rc != GS2M_OK) return rc;
    *order_out = in_input ? g.order_a : g.order_b;
    return GS2M_OK;
}

// stage B: emit (rank, index) pairs into their tiles and sort every tile in shared memory
int binning2_emit_and_sort(int P, const GeomState& g, const int* radii, int tiles_x, int tiles_y, const ImageState& im,
                           const uint32_t* order, uint64_t* tmp, uint64_t* keys_out, uint32_t* vals_out, int max_count,
                           cudaStream_t s) {
    const int n_tiles = tiles_x * tiles_y;
    int key_bits = 1;
    while ((1ll << key_bits) < (long long)P) ++key_bits;   // ranks are < P
    const int cap = ((max_count + 63) / 64) * 64;
    const size_t smem = tile_sort_smem_bytes(cap);
    static std::atomic<size_t> configured{0};
    if (smem > configured.load()) {
        GS2M_CUDA(cudaFuncSetAttribute(tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    count_launches(2);
    emit_ranked_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, order, g.xy_conic_ab, radii, tiles_x, tiles_y, im.tile_starts,
                                                       im.tile_cursor, tmp);
    tile_sort_kernel<<<n_tiles, 256, smem, s>>>(im.ranges, tmp, g.depths, keys_out, vals_out, key_bits, cap);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace gs2m
