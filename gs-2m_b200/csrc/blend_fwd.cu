// Forward blend: every 16x16 tile's depth-sorted Gaussian list is walked front to back and RGB plus the first F feature
// columns (alpha, distance, normal, albedo, roughness, metallic) are alpha-blended into every pixel.
//
// Behavioural reference: renderCUDA (cuda_rasterizer/forward.cu:246-372).  Per pixel the sequence of
// (power, alpha, test_T, T) values, the termination point, `n_contrib` and the `observe` counts are bit-identical to
// the reference; what differs is how the work is organised:
//   * warp-autonomous walk: a tile is 8 warps, each owning an 8x4 pixel block and walking the tile list on its own,
//     32 entries per step, with no block-wide barrier anywhere (the CTA is only a scheduling unit: one warp per CTA).
//     Lane l reads the footprint-mask byte of entry l (footprint_masks.cu: which of the tile's eight warp blocks the
//     Gaussian can reach with alpha >= 1/255, computed once per instance and shared with the backward); the ballot of
//     the warp's bit is its work list, and only the hit lanes gather the 32-byte blend record (two 128-bit loads, an
//     L1/L2 hit for the other warps of the tile) and stage colour + feature vector as 16-byte shared records, instead
//     of the reference's per-pair global re-fetches.  Indices and masks are fetched two steps ahead, records one;
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
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ point_list,
    const uint8_t* __restrict__ masks, int W, int H, int tiles_x, const float4* __restrict__ rec_a, const float4* __restrict__ rec_b, const float4* __restrict__ rgb,
    const float* __restrict__ features, const float* __restrict__ bg, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, int* __restrict__ out_observe,
    float* __restrict__ out_buffer) {
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

    const uint2 range = ranges[tile];
    const int n_list = (int)(range.y - range.x);

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    // colour + feature accumulators as aligned pairs in the staged order (r,g) (b,-) (f0,f1) ...: one FMUL2 + one FFMA2 per
    // pair of channels, each half rounding exactly like the reference's scalar fma(T, alpha * c, C)
    float2 A[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) A[i] = make_float2(0.f, 0.f);

    bool warp_done = __all_sync(0xffffffffu, done);
    // Software pipeline: list indices + footprint-mask bytes are fetched two steps ahead (registers), the records of the
    // hits one step ahead (cp.async into the ring), so both latencies are covered by the blending of the current step.
    const uint32_t* __restrict__ list = point_list + range.x;
    const uint8_t* __restrict__ mlist = masks + range.x;
    auto fetch = [&](int step, int& g, uint32_t& m) {
        const int li = 32 * step + lane;
        g = (li < n_list) ? (int)list[li] : 0;
        m = (li < n_list) ? mlist[li] : 0u;
    };
    int gq[LIST_AHEAD];            // gq[i], mq[i]: index and mask byte of this lane's entry in step (current + 1 + i)
    uint32_t mq[LIST_AHEAD];
    int tail = 0, h_cur;
    {
        int g0;
        uint32_t m0;
        fetch(0, g0, m0);
#pragma unroll
        for (int i = 0; i < LIST_AHEAD; ++i) fetch(1 + i, gq[i], mq[i]);
        const bool hit = (m0 >> warp) & 1u;
        const uint32_t word = __ballot_sync(0xffffffffu, hit);
        h_cur = __popc(word);
        stage_step<F>(sm, lane, hit, word, g0, lane, 0, rec_a, rec_b, rgb, features);
    }
    for (int base = 0; base < n_list && !warp_done; base += 32) {
        // ---- issue the next step's records behind the current step's if the ring has room for both ----
        const int g1 = gq[0];
        const bool hit1 = (mq[0] >> warp) & 1u;
        const uint32_t word1 = __ballot_sync(0xffffffffu, hit1);
        const int h1 = __popc(word1);
        const bool fits = h_cur + h1 <= 32;
        if (fits) {
            stage_step<F>(sm, lane, hit1, word1, g1, base + 32 + lane, tail + h_cur, rec_a, rec_b, rgb, features);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();

        // ---- blend the current step's hits in list order ----
        // two entries per iteration: both alpha evaluations are independent and branch-free (interleaved by the
        // scheduler); the blend itself stays strictly in list order
        for (int k = 0; k < h_cur && !warp_done; k += 2) {
            const int slot0 = (tail + k) & 31;
            const bool two = k + 1 < h_cur;
            const int slot1 = two ? ((tail + k + 1) & 31) : slot0;
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
                        last_contributor = (uint32_t)__float_as_int(u ? rb1.w : rb0.w) + 1u;
                    }
                }
                const uint32_t ob = __ballot_sync(0xffffffffu, obs);
                if (ob != 0 && lane == 0) atomicAdd(out_observe + __float_as_int(u ? rb1.z : rb0.z), __popc(ob));
            }
            warp_done = __all_sync(0xffffffffu, done);
        }
        __syncwarp();   // all lanes are done with this step's slots
        tail = (tail + h_cur) & 31;
        if (!fits) stage_step<F>(sm, lane, hit1, word1, g1, base + 32 + lane, tail, rec_a, rec_b, rgb, features);
        h_cur = h1;
#pragma unroll
        for (int i = 0; i + 1 < LIST_AHEAD; ++i) { gq[i] = gq[i + 1]; mq[i] = mq[i + 1]; }
        fetch(base / 32 + 1 + LIST_AHEAD, gq[LIST_AHEAD - 1], mq[LIST_AHEAD - 1]);
    }
    cp_async_wait<0>();   // no copy may still be in flight when the warp's shared memory is released

    if (inside) {
        const size_t N = (size_t)W * H;
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last_contributor;
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
int launch_f(const FwdParams& p, const GeomState& g, const uint32_t* point_list, const uint8_t* masks, const ImageState& im,
             float* out_color, int* out_observe, float* out_buffer, cudaStream_t s) {
    const unsigned grid = (unsigned)(p.tiles_x * p.tiles_y) * FWD_CTAS_PER_TILE;
    count_launches(1);
    blend_forward_kernel<F><<<grid, FWD_CTA_WARPS * 32, 0, s>>>(im.ranges, im.tile_order, point_list, masks, p.W, p.H, p.tiles_x, g.xy_conic_ab,
                                                           g.conic_c_opac, g.rgb, p.features, p.background, im.final_T,
                                                           im.n_contrib, out_color, out_observe, out_buffer);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace

int launch_blend_forward(const FwdParams& p, const GeomState& g, const uint32_t* point_list, const uint8_t* masks,
                         const ImageState& im, float* out_color, int* out_observe, float* out_buffer, cudaStream_t s) {
    switch (p.F) {
#define GS2M_CASE(N) case N: return launch_f<N>(p, g, point_list, masks, im, out_color, out_observe, out_buffer, s);
        GS2M_CASE(0) GS2M_CASE(1) GS2M_CASE(2) GS2M_CASE(3) GS2M_CASE(4) GS2M_CASE(5)
        GS2M_CASE(6) GS2M_CASE(7) GS2M_CASE(8) GS2M_CASE(9) GS2M_CASE(10)
#undef GS2M_CASE
    }
    set_error("feature_count %d outside 0..10", p.F);
    return GS2M_ERR_INVALID_ARGUMENT;
}

}  // namespace gs2m
