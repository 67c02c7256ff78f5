// Backward blend: per tile, walk the depth-sorted list back to front, recover the per-pair blending weights and
// produce per-Gaussian gradients of the 2-D mean (signed + absolute), conic, opacity, colour and feature vector.
//
// Behavioural reference: renderCUDA backward (cuda_rasterizer/backward.cu:413-598), which issues 11+F same-address
// float atomics per (pixel, Gaussian) pair.  This kernel produces the same sums (to fp32 re-association) with
// *no per-pixel atomics* and *no block-wide barriers*:
//
//   warp-autonomous walk      A tile is 8 warps, each owning an 8x4 pixel block and walking its own list back to front
//                             (the CTA is only a scheduling unit: two warps per CTA, 96 registers, 20 warps per SM).  The
//                             list is the one the forward walked: the tile's entries whose footprint mask has the block's
//                             bit, compacted by footprint_masks.cu (dense_gid / block_ranges), so every entry loaded is
//                             evaluated.  Records, colour and features of sixteen entries at a time are copied global ->
//                             shared with cp.async into one half of a 32-slot ring while the other half is evaluated.
//                             Entries behind the warp's deepest last-contributor (n_contrib_dense, written by the forward
//                             in list coordinates) are never touched.  (The round-1 first cut staged batches per CTA
//                             behind __syncthreads; ncu showed 32% of the stall samples on those barriers because the 8
//                             warps of a tile have very unequal work.)
//   evaluate (lane = pixel)   For a surviving entry every lane recomputes alpha bit-exactly like the forward, steps
//                             T <- T/(1-alpha) (MUFU.RCP + one Newton step), and needs only two scalars per pair:
//                             w = alpha*T  (weight of the colour/feature gradients) and  Q = G * dL/dalpha  (weight of
//                             every geometric gradient).  dL/dalpha uses the scalar recurrence
//                             S <- a_prev*cd_prev + (1-a_prev)*S  with  cd = <colour+features, dL/dpixel>,
//                             algebraically identical to the reference's per-channel accum_rec recurrences
//                             (backward.cu:546-560).  Two entries are evaluated per iteration (independent alpha
//                             chains).  (w, Q) of 32 pixels x 16 entries are parked in a per-warp shared tile; of the
//                             entry itself only the Gaussian index is kept (lane e holds entry e's).
//   reduce (lane = entry x pixel-half)  When 16 entries are parked the warp transposes roles: lane (e, h) owns entry e
//                             and sums over the 16 pixels of half h — 3+F colour/feature sums, six geometric moments
//                             and the two AbsGS sums — privately in registers: a shared-memory transpose instead of
//                             21 shuffle butterflies per pair.  The record is re-gathered here (an L1/L2 hit).
//   accumulate                The halves are combined with one shuffle per value and lane e adds its 11+F sums to the
//                             Gaussian's packed 96-byte accumulator row with 128-bit vector reductions
//                             (red.global.add.v4.f32 -> REDG.E.ADD.F32x4): 6 per (Gaussian, warp block) instead of
//                             21 x 32 scalar atomics.
#include <atomic>
#include "blend_common.cuh"

namespace gs2m {
namespace {

constexpr int PARK = 16;         // entries parked per warp before a reduce (2 lanes per entry in the reduce)
constexpr int PARK_STRIDE = 33;  // padded row -> conflict-free for both write (lane = pixel) and read (lane = entry,half)

template <int F>
struct WarpSmemB {
    static constexpr int NV = StagedRing<F>::NV;
    StagedRing<F> ring;            // staged records of the hits (blend_common.cuh)
    float4 dpix[2][16 * NV + 1];   // dL/d(colour,features) of the warp's 32 pixels in the staged channel order, two 16-pixel
                                   // halves (+16 B skew so that the two halves never share a bank in the reduce)
    float park_w[PARK * PARK_STRIDE];
    float park_q[PARK * PARK_STRIDE];
};

__device__ __forceinline__ float __frcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Reduce phase: lane (e = lane & 15, h = lane >> 4) owns parked entry e and sums over the 16 pixels of half h;
// the two halves are combined with one shuffle per value and lanes 0..15 issue the vector reductions.
// Only the Gaussian index of a parked entry is carried in registers (lane e holds entry e's); the 32-byte record is
// re-gathered here (an L1/L2 hit) while the colour sums, which do not need it, are running.
// The geometric sums are linear in the six moments sum(q), sum(q dx), sum(q dy), sum(q dx dx), sum(q dx dy),
// sum(q dy dy); only the AbsGS sums need the per-pixel value.
template <int F>
__device__ __forceinline__ void reduce_parked(WarpSmemB<F>& sm, int lane, int n_parked, int my_gid,
                                              const float4* __restrict__ rec_a, const float4* __restrict__ rec_b,
                                              float wpx0, float wpy0, float half_w, float half_h,
                                              float* __restrict__ grad_acc) {
    constexpr int NV = WarpSmemB<F>::NV;
    constexpr int NC = 3 + F;
    constexpr int NG = 11 + F;
    const int e = lane & 15, h = lane >> 4;
    const int gid = __shfl_sync(0xffffffffu, my_gid, e);
    float out[GS2M_ACC_STRIDE];
#pragma unroll
    for (int i = 0; i < GS2M_ACC_STRIDE; ++i) out[i] = 0.f;
    if (e < n_parked) {
        const float4 ra = __ldg(rec_a + gid);
        const float4 rb = __ldg(rec_b + gid);
        // colour / feature sums as aligned pairs: one FFMA2 per two channels (the 128-bit shared loads deliver the pairs)
        constexpr int NP = StagedRing<F>::NPAIR;
        float2 gc2[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) gc2[i] = make_float2(0.f, 0.f);
        const float* pw = &sm.park_w[e * PARK_STRIDE + h * 16];
        const float* pq = &sm.park_q[e * PARK_STRIDE + h * 16];
        const float4* dp = sm.dpix[h];
#pragma unroll 8
        for (int j = 0; j < 16; ++j) {
            const float2 ww = make_float2(pw[j], pw[j]);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 t = dp[j * NV + k];
                if (2 * k < NP) gc2[2 * k] = fma2_rn(ww, make_float2(t.x, t.y), gc2[2 * k]);
                if (2 * k + 1 < NP) gc2[2 * k + 1] = fma2_rn(ww, make_float2(t.z, t.w), gc2[2 * k + 1]);
            }
        }
        const float ca = ra.z, cb = ra.w, cc = rb.x, op = rb.y;
        const float gxr = ra.x - wpx0;                         // exact: both are multiples of ulp(mean) and the result is smaller
        const float gyr = ra.y - (wpy0 + (float)(2 * h));
        float m0 = 0.f, mx = 0.f, my = 0.f, mxx = 0.f, mxy = 0.f, myy = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float q = pq[j];
            const float dx = gxr - (float)(j & 7);
            const float dy = gyr - (float)(j >> 3);
            const float qdx = q * dx, qdy = q * dy;
            m0 += q; mx += qdx; my += qdy;
            mxx = fmaf(qdx, dx, mxx);
            mxy = fmaf(qdx, dy, mxy);
            myy = fmaf(qdy, dy, myy);
            ax += fabsf(fmaf(ca, qdx, cb * qdy));
            ay += fabsf(fmaf(cc, qdy, cb * qdx));
        }
        const float kx = op * half_w, ky = op * half_h;
        out[0] = -kx * fmaf(ca, mx, cb * my);
        out[1] = -ky * fmaf(cc, my, cb * mx);
        out[2] = fabsf(kx) * ax;
        out[3] = fabsf(ky) * ay;
        out[4] = -0.5f * op * mxx;
        out[5] = -0.5f * op * mxy;
        out[6] = -0.5f * op * myy;
        out[7] = m0;
#pragma unroll
        for (int i = 0; i < NC; ++i) out[8 + i] = (staged_pos(i) & 1) ? gc2[staged_pos(i) >> 1].y : gc2[staged_pos(i) >> 1].x;
    }
#pragma unroll
    for (int i = 0; i < NG; ++i) out[i] += __shfl_xor_sync(0xffffffffu, out[i], 16);
    if (h == 0 && e < n_parked) {
        float* dst = grad_acc + (size_t)gid * GS2M_ACC_STRIDE;
#pragma unroll
        for (int v = 0; v < (NG + 3) / 4; ++v) red_add_v4(dst + 4 * v, out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
    }
    __syncwarp();
}

// Warps are autonomous, so the CTA is only a scheduling unit: a tile's 8 warp blocks may be spread over 8 / GS2M_BWD_WARPS
// CTAs (finer register-file granularity: occupancy is floor(64K / (regs * 32 * GS2M_BWD_WARPS)) CTAs).
#ifndef GS2M_BWD_WARPS
#define GS2M_BWD_WARPS 2
#endif
#ifndef GS2M_BWD_WARPS_PER_SM
#define GS2M_BWD_WARPS_PER_SM 20     // 96 registers per thread (4 bytes of spill); 16 (128 registers) is 6 % slower, 22 (88) 1.5 % slower
#endif
constexpr int BWD_CTA_WARPS = GS2M_BWD_WARPS;
constexpr int BWD_CTAS_PER_TILE = BLEND_WARPS / BWD_CTA_WARPS;
static_assert(BLEND_WARPS % BWD_CTA_WARPS == 0, "CTA must hold a divisor of the tile's 8 warp blocks");

#ifdef GS2M_BWD_MAXNREG
#define GS2M_BWD_BOUNDS __maxnreg__(GS2M_BWD_MAXNREG)
#else
#define GS2M_BWD_BOUNDS __launch_bounds__(BWD_CTA_WARPS * 32, GS2M_BWD_WARPS_PER_SM / BWD_CTA_WARPS)
#endif
template <int F>
__global__ void GS2M_BWD_BOUNDS blend_backward_kernel(
    const uint2* __restrict__ block_ranges, const uint32_t* __restrict__ tile_order, const uint32_t* __restrict__ dense_gid,
    int R_cap, int W, int H, int tiles_x, const float4* __restrict__ rec_a, const float4* __restrict__ rec_b, const float4* __restrict__ rgb,
    const float* __restrict__ features, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib_dense, const float* __restrict__ grad_color, const float* __restrict__ grad_buffer,
    float* __restrict__ grad_acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NV = WarpSmemB<F>::NV;
    constexpr int NC = 3 + F;

    const int lane = threadIdx.x & 31;
    const int warp = (int)(blockIdx.x % BWD_CTAS_PER_TILE) * BWD_CTA_WARPS + (int)(threadIdx.x >> 5);   // warp block in the tile
    WarpSmemB<F>& sm = reinterpret_cast<WarpSmemB<F>*>(smem_raw)[threadIdx.x >> 5];
    const int tile = (int)tile_order[blockIdx.x / BWD_CTAS_PER_TILE];      // CTAs take the tiles longest list first
    const int tile_y = tile / tiles_x, tile_x = tile - tile_y * tiles_x;
    int px, py;
    pixel_of_thread(tile_x, tile_y, warp * 32 + lane, px, py);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;
    const size_t N = (size_t)W * H;
    const size_t pix = (size_t)py * W + px;

    // this warp block's own list (footprint_masks.cu: the tile's entries whose mask has the block's bit, compacted)
    const uint2 br = block_ranges[(size_t)tile * BLEND_WARPS + warp];
    const int n_list = (int)(br.y - br.x);

    // per-pixel state
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t my_contrib = inside ? n_contrib_dense[pix] : 0u;      // last contributor + 1, in this list's coordinates
    float T = T_final;
    float dL[4 * NV];              // staged channel order: colour 0..2, 0, features 4..
#pragma unroll
    for (int i = 0; i < 4 * NV; ++i) dL[i] = 0.f;
    if (inside) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dL[ch] = grad_color[ch * N + pix];
#pragma unroll
        for (int ch = 0; ch < F; ++ch) dL[4 + ch] = grad_buffer[ch * N + pix];
    }
    const float bg_dot = bg[0] * dL[0] + bg[1] * dL[1] + bg[2] * dL[2];
#pragma unroll
    for (int k = 0; k < NV; ++k)
        sm.dpix[lane >> 4][(lane & 15) * NV + k] = make_float4(dL[4 * k], dL[4 * k + 1], dL[4 * k + 2], dL[4 * k + 3]);

    // deepest contributor of this warp: nothing behind it is ever touched
    uint32_t wmax = my_contrib;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    const int n_back = (int)min(wmax, (uint32_t)n_list);   // forward guarantees wmax <= n_list; defensive
    __syncwarp();

    const int wpx0i = tile_x * GS2M_TILE_X + (warp & 1) * WARP_PIX_X;
    const int wpy0i = tile_y * GS2M_TILE_Y + (warp >> 1) * WARP_PIX_Y;
    const float wpx0 = (float)wpx0i, wpy0 = (float)wpy0i;
    const float half_w = 0.5f * (float)W, half_h = 0.5f * (float)H;

    float S = 0.f, last_alpha = 0.f, last_cd = 0.f;
    int n_parked = 0, my_gid = 0;

    // Software pipeline: the list is walked back to front; reversed index r = 32 j + l (lane l, step j) is list entry
    // n_back - 1 - r.  Lane l holds the Gaussian indices of its entries of steps j + 1 and j + 2 (coalesced loads two steps
    // ahead); the records of a half-step (16 entries) are copied while the previous half-step is evaluated.
    const uint32_t* __restrict__ list = dense_gid + (size_t)warp * R_cap + br.x;
    StagedRing<F>& ring = sm.ring;
    const int half = lane >> 4;
    auto fetch = [&](int step) { const int f = n_back - 1 - (32 * step + lane); return (f >= 0) ? (int)list[f] : 0; };
    int g1 = fetch(1), g2 = fetch(2);
    {
        const int g0 = fetch(0);
        stage_half<F>(ring, lane, half == 0 && lane < n_back, g0, rec_a, rec_b, rgb, features);
        stage_half<F>(ring, lane, half == 1 && lane < n_back, g0, rec_a, rec_b, rgb, features);
    }
    const int n_half = (n_back + 15) >> 4;
    for (int h = 0; h < n_half; ++h) {
        cp_async_wait<1>();       // half-step h has landed (h + 1 may still be in flight)
        __syncwarp();

        // ---- evaluate (lane = pixel) / reduce (lane = parked entry) ----
        // Two list entries per iteration: their alpha evaluations (the long dependent chain with the expf) are
        // independent and written branch-free so the scheduler can interleave them; only the T / S recurrences and
        // the parking are sequential.
        const int r0 = h << 4, cnt = min(16, n_back - r0), s0 = (h & 1) << 4;
        for (int k = 0; k < cnt; k += 2) {
            const int slot0 = s0 + k;
            const bool two = k + 1 < cnt;
            const int slot1 = two ? slot0 + 1 : slot0;
            const float4 ra0 = ring.a[slot0], rb0 = ring.b[slot0];
            const float4 ra1 = ring.a[slot1], rb1 = ring.b[slot1];
            float G0, alpha0, G1, alpha1;
            bool v0, v1;
            {
                float dx, dy;
                v0 = pair_alpha_nb(ra0.x, ra0.y, ra0.z, ra0.w, rb0.x, rb0.y, pxf, pyf, dx, dy, G0, alpha0);
                v1 = pair_alpha_nb(ra1.x, ra1.y, ra1.z, ra1.w, rb1.x, rb1.y, pxf, pyf, dx, dy, G1, alpha1);
            }
            // a pixel blended exactly the entries in front of its last contributor
            v0 = v0 && ((uint32_t)(n_back - 1 - (r0 + k)) < my_contrib);
            v1 = v1 && two && ((uint32_t)(n_back - 2 - (r0 + k)) < my_contrib);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const bool v = u ? v1 : v0;
                if (u == 1 && !two) break;
                // an entry no pixel of the block actually blends contributes nothing: skip it entirely
                if (!__any_sync(0xffffffffu, v)) continue;
                const int slot = u ? slot1 : slot0;
                const float alpha = u ? alpha1 : alpha0;
                const float G = u ? G1 : G0;
                float pw = 0.f, pq = 0.f;
                if (v) {
                    // one reciprocal serves both T / (1 - alpha) and T_final / (1 - alpha) (backward.cu:532,572 divide
                    // twice).  1 - alpha is in [0.01, 1], so MUFU.RCP + one Newton step is within 1 ulp without the
                    // range checks of a correctly rounded reciprocal; the difference is far inside the 1e-4 gate.
                    const float one_m_alpha = 1.0f - alpha;
                    float inv_one_m_alpha = __frcp_approx(one_m_alpha);
                    inv_one_m_alpha = fmaf(inv_one_m_alpha, fmaf(-one_m_alpha, inv_one_m_alpha, 1.0f), inv_one_m_alpha);
                    T *= inv_one_m_alpha;
                    // cd = <colour+features, dL/dpixel>, one sequential chain in channel order.  (Two interleaved partial
                    // sums on FFMA2 pairs were measured: 0.7 % faster, but the different association moved the raw-parameter
                    // gradients from 2e-6 to 1.2e-5 of the reference's, which also accumulates channel by channel —
                    // backward.cu:546-560 — so the scalar chain stays.)  Staged words behind the last used channel are
                    // not ours and are never multiplied.
                    float cd = 0.f;
#pragma unroll
                    for (int kk = 0; kk < NV; ++kk) {
                        const float4 t = ring.col[kk][slot];
                        const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int ch = kk == 0 ? q : 4 * kk - 1 + q;       // blended channel of this staged word
                            if ((kk == 0 && q < 3) || (kk > 0 && ch < NC)) cd = fmaf(tv[q], dL[4 * kk + q], cd);
                        }
                    }
                    S = fmaf(last_alpha, last_cd - S, S);          // a_prev*cd_prev + (1-a_prev)*S
                    last_alpha = alpha;
                    last_cd = cd;
                    float dL_dalpha = (cd - S) * T;
                    dL_dalpha = fmaf(-T_final * inv_one_m_alpha, bg_dot, dL_dalpha);
                    pw = alpha * T;
                    pq = dL_dalpha * G;
                }
                if (lane == n_parked) my_gid = __float_as_int(u ? rb1.z : rb0.z);
                sm.park_w[n_parked * PARK_STRIDE + lane] = pw;
                sm.park_q[n_parked * PARK_STRIDE + lane] = pq;
                ++n_parked;
                if (n_parked == PARK) {
                    __syncwarp();
                    reduce_parked<F>(sm, lane, n_parked, my_gid, rec_a, rec_b, wpx0, wpy0, half_w, half_h, grad_acc);
                    n_parked = 0;
                }
            }
        }
        __syncwarp();   // every lane is done reading this half of the ring

        // ---- refill it: half-step h + 2 = the entries of step (h >> 1) + 1 held by the lanes of half (h & 1) ----
        stage_half<F>(ring, lane, half == (h & 1) && ((h + 2) << 4) + (lane & 15) < n_back, g1, rec_a, rec_b, rgb, features);
        if (h & 1) {
            g1 = g2;
            g2 = fetch((h >> 1) + 3);
        }
    }
    cp_async_wait<0>();   // no copy may still be in flight when the warp's shared memory is released
    if (n_parked > 0) {
        __syncwarp();
        reduce_parked<F>(sm, lane, n_parked, my_gid, rec_a, rec_b, wpx0, wpy0, half_w, half_h, grad_acc);
    }
}

template <int F>
int launch_b(const BwdParams& p, const GeomState& g, const BinState& b, int R_cap, const ImageState& im, cudaStream_t s) {
    const unsigned grid = (unsigned)(p.tiles_x * p.tiles_y) * BWD_CTAS_PER_TILE;
    const size_t smem = sizeof(WarpSmemB<F>) * BWD_CTA_WARPS;
    static PerDeviceOnce configured;   // per kernel instantiation and per device; callers may use several host threads
    int dev;
    if (configured.need(dev)) {
        GS2M_CUDA(cudaFuncSetAttribute(blend_backward_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef GS2M_BWD_CARVEOUT
        GS2M_CUDA(cudaFuncSetAttribute(blend_backward_kernel<F>, cudaFuncAttributePreferredSharedMemoryCarveout, GS2M_BWD_CARVEOUT));
#endif
        configured.done(dev);
    }
    count_launches(1);
    blend_backward_kernel<F><<<grid, BWD_CTA_WARPS * 32, smem, s>>>(im.block_ranges, im.tile_order, b.dense_gid, R_cap, p.W, p.H, p.tiles_x,
                                                               g.xy_conic_ab, g.conic_c_opac, g.rgb, p.features, p.background,
                                                               im.final_T, im.n_contrib_dense, p.grad_color, p.grad_buffer,
                                                               g.grad_acc);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace

int launch_blend_backward(const BwdParams& p, const GeomState& g, const BinState& b, int R_cap, const ImageState& im, cudaStream_t s) {
    switch (p.F) {
#define GS2M_CASE(N) case N: return launch_b<N>(p, g, b, R_cap, im, s);
        GS2M_CASE(0) GS2M_CASE(1) GS2M_CASE(2) GS2M_CASE(3) GS2M_CASE(4) GS2M_CASE(5)
        GS2M_CASE(6) GS2M_CASE(7) GS2M_CASE(8) GS2M_CASE(9) GS2M_CASE(10)
#undef GS2M_CASE
    }
    set_error("feature_count %d outside 0..10", p.F);
    return GS2M_ERR_INVALID_ARGUMENT;
}

}  // namespace gs2m
