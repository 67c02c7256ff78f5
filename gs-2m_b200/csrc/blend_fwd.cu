// Forward blend: every 16x16 tile's depth-sorted Gaussian list is walked front to back and RGB plus the first F feature
// columns (alpha, distance, normal, albedo, roughness, metallic) are alpha-blended into every pixel.
//
// Behavioural reference: renderCUDA (cuda_rasterizer/forward.cu:246-372).  Per pixel the sequence of
// (power, alpha, test_T, T) values, the termination point, `n_contrib` and the `observe` counts are bit-identical to
// the reference; what differs is how the work is organised:
//   * warp-autonomous walk: a tile is 8 warps, each owning an 8x4 pixel block, with no block-wide barrier anywhere (the
//     CTA is only a scheduling unit: one warp per CTA).  A warp walks ITS OWN list: footprint_masks.cu decides once per
//     (Gaussian, tile) instance which of the tile's eight warp blocks the Gaussian can reach with alpha >= 1/255 and compacts
//     the tile lists by that bit (dense_gid / block_ranges; shared with the backward), so every entry a warp loads is one it
//     evaluates.  The 32-byte blend record, colour and feature vector of an entry are copied global -> shared with cp.async
//     (LDGSTS) sixteen entries at a time, one half of a 32-slot ring being filled while the other is blended; the Gaussian
//     indices are read two 32-entry steps ahead.  No per-pair global re-fetches as in the reference;
//   * two entries are evaluated per iteration with branch-free alpha code; the blend itself stays in list order;
//   * a warp stops as soon as all of its 32 pixels have terminated (the reference only leaves when all 256 have);
//   * `observe` increments are aggregated per warp (ballot + popc): one integer reduction per (entry, warp) — sums of
//     integers, so the totals are exact;
//   * F is a template parameter: accumulators stay in registers and the loops unroll.
#include "blend_common.cuh"

namespace gs2m {
namespace {

// Warps are autonomous, so the CTA is only a scheduling unit: a tile's 8 warp blocks are spread over 8 / GS2M_FWD_WARPS CTAs.
#ifndef GS2M_FWD_WARPS
#define GS2M_FWD_WARPS 1
#endif
constexpr int FWD_CTA_WARPS = GS2M_FWD_WARPS;
constexpr int FWD_CTAS_PER_TILE = BLEND_WARPS / FWD_CTA_WARPS;
static_assert(BLEND_WARPS % FWD_CTA_WARPS == 0, "CTA must hold a divisor of the tile's 8 warp blocks");

template <int F>
__global__ void __launch_bounds__(FWD_CTA_WARPS * 32, 32 / FWD_CTA_WARPS) blend_forward_kernel(
    const uint2* __restrict__ block_ranges, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ dense_gid,
    const uint32_t* __restrict__ dense_pos, int R_cap, int W, int H, int tiles_x, const float4* __restrict__ rec_a,
    const float4* __restrict__ rec_b, const float4* __restrict__ rgb,
    const float* __restrict__ features, const float* __restrict__ bg, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, uint32_t* __restrict__ n_contrib_dense, float* __restrict__ out_color,
    int* __restrict__ out_observe, float* __restrict__ out_buffer) {
    __shared__ StagedRing<F> sm_all[FWD_CTA_WARPS];
    constexpr int NV = StagedRing<F>::NV;
    constexpr int NP = StagedRing<F>::NPAIR;

    const int lane = threadIdx.x & 31;
    const int warp = (int)(blockIdx.x % FWD_CTAS_PER_TILE) * FWD_CTA_WARPS + (int)(threadIdx.x >> 5);   // warp block in the tile
    StagedRing<F>& sm = sm_all[threadIdx.x >> 5];
    const int tile = (int)tile_order[blockIdx.x / FWD_CTAS_PER_TILE];      // CTAs take the tiles longest list first
    const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
    int px, py;
    pixel_of_thread(tile_x, tile_y, warp * 32 + lane, px, py);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;

    // this warp block's own list
    const uint2 br = block_ranges[(size_t)tile * BLEND_WARPS + warp];
    const int n_list = (int)(br.y - br.x);
    const uint32_t* __restrict__ list = dense_gid + (size_t)warp * R_cap + br.x;

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;    // in the coordinates of this list
    // colour + feature accumulators as aligned pairs in the staged order (r,g) (b,-) (f0,f1) ...: one FMUL2 + one FFMA2 per
    // pair of channels, each half rounding exactly like the reference's scalar fma(T, alpha * c, C)
    float2 A[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) A[i] = make_float2(0.f, 0.f);

    bool warp_done = __all_sync(0xffffffffu, done);
    // Software pipeline: lane l holds the Gaussian index of entry 32 j + l for the steps j + 1 and j + 2 (registers, coalesced
    // loads two steps ahead); the records of a half-step are copied while the previous half-step is being blended.
    const int half = lane >> 4;
    int g1 = (32 + lane < n_list) ? (int)list[32 + lane] : 0;
    int g2 = (64 + lane < n_list) ? (int)list[64 + lane] : 0;
    {
        const int g0 = (lane < n_list) ? (int)list[lane] : 0;
        stage_half<F>(sm, lane, half == 0 && lane < n_list, g0, rec_a, rec_b, rgb, features);
        stage_half<F>(sm, lane, half == 1 && lane < n_list, g0, rec_a, rec_b, rgb, features);
    }
    const int n_half = (n_list + 15) >> 4;
    for (int h = 0; h < n_half && !warp_done; ++h) {
        cp_async_wait<1>();       // half-step h has landed (h + 1 may still be in flight)
        __syncwarp();

        // ---- blend the half-step's entries in list order ----
        // two entries per iteration: both alpha evaluations are independent and branch-free (interleaved by the
        // scheduler); the blend itself stays strictly in list order
        const int e0 = h << 4, cnt = min(16, n_list - e0), s0 = (h & 1) << 4;
        for (int k = 0; k < cnt && !warp_done; k += 2) {
            const int slot0 = s0 + k;
            const bool two = k + 1 < cnt;
            const int slot1 = two ? slot0 + 1 : slot0;
            const float4 ra0 = sm.a[slot0], rb0 = sm.b[slot0];
            const float4 ra1 = sm.a[slot1], rb1 = sm.b[slot1];
            float G0, alpha0, G1, alpha1;
            bool v0, v1;
            {
                float dx, dy;
                v0 = pair_alpha_nb(ra0.x, ra0.y, ra0.z, ra0.w, rb0.x, rb0.y, pxf, pyf, dx, dy, G0, alpha0);
                v1 = pair_alpha_nb(ra1.x, ra1.y, ra1.z, ra1.w, rb1.x, rb1.y, pxf, pyf, dx, dy, G1, alpha1) && two;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                const bool v = u ? v1 : v0;
                const int slot = u ? slot1 : slot0;
                const float alpha = u ? alpha1 : alpha0;
                bool obs = false;
                if (!done && v) {
                    const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
#pragma unroll
                        for (int kk = 0; kk < NV; ++kk) {
                            const float4 t = sm.col[kk][slot];
                            if (2 * kk < NP) A[2 * kk] = fma2_rn(make_float2(T, T), mul2_rn(make_float2(alpha, alpha), make_float2(t.x, t.y)), A[2 * kk]);
                            if (2 * kk + 1 < NP) A[2 * kk + 1] = fma2_rn(make_float2(T, T), mul2_rn(make_float2(alpha, alpha), make_float2(t.z, t.w)), A[2 * kk + 1]);
                        }
                        obs = T > 0.5f;
                        T = test_T;
                        last_contributor = (uint32_t)(e0 + k + u) + 1u;
                    }
                }
                const uint32_t ob = __ballot_sync(0xffffffffu, obs);
                if (ob != 0 && lane == 0) atomicAdd(out_observe + __float_as_int(u ? rb1.z : rb0.z), __popc(ob));
            }
            warp_done = __all_sync(0xffffffffu, done);
        }
        __syncwarp();   // all lanes are done with this half of the ring

        // ---- refill it: half-step h + 2 = the entries of step (h >> 1) + 1 held by the lanes of half (h & 1) ----
        stage_half<F>(sm, lane, half == (h & 1) && ((h + 2) << 4) + (lane & 15) < n_list, g1, rec_a, rec_b, rgb, features);
        if (h & 1) {
            g1 = g2;
            const int e = (((h >> 1) + 3) << 5) + lane;
            g2 = (e < n_list) ? (int)list[e] : 0;
        }
    }
    cp_async_wait<0>();   // no copy may still be in flight when the warp's shared memory is released

    if (inside) {
        const size_t N = (size_t)W * H;
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        // the reference's n_contrib counts positions in the TILE's list: one gather per pixel from the list's position column
        n_contrib[pix] = last_contributor ? dense_pos[(size_t)warp * R_cap + br.x + last_contributor - 1u] + 1u : 0u;
        n_contrib_dense[pix] = last_contributor;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) out_color[ch * N + pix] = __fmaf_rn(T, bg[ch], (ch & 1) ? A[ch >> 1].y : A[ch >> 1].x);
#pragma unroll
        for (int ch = 0; ch < GS2M_NUM_FEATURES; ++ch) {
            const int i = ch < F ? staged_pos(3 + ch) : 0;
            out_buffer[ch * N + pix] = (ch < F) ? ((i & 1) ? A[i >> 1].y : A[i >> 1].x) : 0.f;
        }
    }
}

template <int F>
int launch_f(const FwdParams& p, const GeomState& g, const BinState& b, int R_cap, const ImageState& im,
             float* out_color, int* out_observe, float* out_buffer, cudaStream_t s) {
    const unsigned grid = (unsigned)(p.tiles_x * p.tiles_y) * FWD_CTAS_PER_TILE;
    count_launches(1);
    blend_forward_kernel<F><<<grid, FWD_CTA_WARPS * 32, 0, s>>>(im.block_ranges, im.tile_order, b.dense_gid, b.dense_pos, R_cap, p.W, p.H,
                                                           p.tiles_x, g.xy_conic_ab, g.conic_c_opac, g.rgb, p.features, p.background,
                                                           im.final_T, im.n_contrib, im.n_contrib_dense, out_color, out_observe,
                                                           out_buffer);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace

int launch_blend_forward(const FwdParams& p, const GeomState& g, const BinState& b, int R_cap,
                         const ImageState& im, float* out_color, int* out_observe, float* out_buffer, cudaStream_t s) {
    switch (p.F) {
#define GS2M_CASE(N) case N: return launch_f<N>(p, g, b, R_cap, im, out_color, out_observe, out_buffer, s);
        GS2M_CASE(0) GS2M_CASE(1) GS2M_CASE(2) GS2M_CASE(3) GS2M_CASE(4) GS2M_CASE(5)
        GS2M_CASE(6) GS2M_CASE(7) GS2M_CASE(8) GS2M_CASE(9) GS2M_CASE(10)
#undef GS2M_CASE
    }
    set_error("feature_count %d outside 0..10", p.F);
    return GS2M_ERR_INVALID_ARGUMENT;
}

}  // namespace gs2m
