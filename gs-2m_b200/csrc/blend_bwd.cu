// Backward blend: per tile, walk the depth-sorted list back to front, recover the per-pair blending weights and
// produce per-Gaussian gradients of the 2-D mean (signed + absolute), conic, opacity, colour and feature vector.
//
// Behavioural reference: renderCUDA backward (cuda_rasterizer/backward.cu:413-598), which issues 11+F same-address
// float atomics per (pixel, Gaussian) pair.  This kernel produces the same sums (to fp32 re-association) with
// *no per-pixel atomics*:
//
//   evaluate (lane = pixel)   For a list entry whose footprint reaches the warp's 8x4 block, every lane recomputes
//                             alpha bit-exactly like the forward, steps T <- T/(1-alpha), and needs only two scalars
//                             per pair:  w = alpha*T  (weight of the colour/feature gradients) and
//                             Q = G * dL/dalpha       (weight of every geometric gradient).
//                             dL/dalpha uses the scalar recurrence  S <- a_prev*cd_prev + (1-a_prev)*S  with
//                             cd = <colour+features, dL/dpixel>, algebraically identical to the reference's
//                             per-channel accum_rec/accum_buf recurrences (backward.cu:546-560).
//                             (w, Q) of 32 pixels x up to 32 entries are parked in a per-warp shared tile.
//   reduce (lane = Gaussian)  When 32 entries are parked the warp transposes roles: lane g owns entry g and sums
//                             over the 32 pixels — 3+F colour/feature sums and 8 geometric moments — privately in
//                             registers.  This is the "warp-aggregated accumulation": a shared-memory transpose
//                             instead of 21 shuffle butterflies per pair.
//   accumulate                Lane g adds its 11+F sums into a per-CTA shared accumulator row of its list entry
//                             (8 warps share a row), and after the batch one thread per entry issues the 11+F
//                             global atomics: one set per (Gaussian, tile) instead of per (Gaussian, pixel).
//
// Entries behind every pixel's last contributor are never staged (the reference stages and skips them), and the
// same conservative footprint masks as in the forward keep warps away from entries that cannot touch them.
#include "blend_common.cuh"

namespace gs2m {
namespace {

constexpr int BWD_BATCH = 128;   // list entries staged per round
constexpr int PARK = 32;         // entries parked per warp before a reduce
constexpr int PARK_STRIDE = 33;  // padded row -> conflict-free for both write (lane = pixel) and read (lane = entry)

template <int F>
struct BwdSmem {
    static constexpr int NV = (3 + F + 3) / 4;
    static constexpr int NG = 11 + F;                       // gradient sums per Gaussian
    float4 a[BWD_BATCH];
    float4 b[BWD_BATCH];
    float4 col[NV][BWD_BATCH];
    float4 dpix[BLEND_WARPS][32][NV];                       // dL/d(colour,features) of every pixel, per warp
    float park_w[BLEND_WARPS][PARK * PARK_STRIDE];
    float park_q[BLEND_WARPS][PARK * PARK_STRIDE];
    float acc[BWD_BATCH][NG];
    uint32_t words[BLEND_WARPS][BWD_BATCH / 32];
    uint32_t warp_max[BLEND_WARPS];
    uint32_t touched[BWD_BATCH / 32];
};

template <int F>
__device__ __forceinline__ void reduce_parked(BwdSmem<F>& sm, int warp, int lane, int n_parked, int my_slot, float wpx0,
                                              float wpy0, float half_w, float half_h) {
    constexpr int NV = BwdSmem<F>::NV;
    constexpr int NC = 3 + F;
    if (lane < n_parked) {
        const float4 ra = sm.a[my_slot];
        const float4 rb = sm.b[my_slot];
        const float gx = ra.x, gy = ra.y, ca = ra.z, cb = ra.w, cc = rb.x, op = rb.y;
        float gc[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) gc[i] = 0.f;
        float sx = 0.f, sy = 0.f, ax = 0.f, ay = 0.f, cxx = 0.f, cxy = 0.f, cyy = 0.f, so = 0.f;
        const float* pw = &sm.park_w[warp][lane * PARK_STRIDE];
        const float* pq = &sm.park_q[warp][lane * PARK_STRIDE];
#pragma unroll 8
        for (int p = 0; p < 32; ++p) {
            const float w = pw[p];
            const float q = pq[p];
            float d[4 * NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const float4 t = sm.dpix[warp][p][k];
                d[4 * k] = t.x; d[4 * k + 1] = t.y; d[4 * k + 2] = t.z; d[4 * k + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) gc[i] = fmaf(w, d[i], gc[i]);
            const float dx = gx - (wpx0 + (float)(p & 7));
            const float dy = gy - (wpy0 + (float)(p >> 3));
            const float qx = q * fmaf(ca, dx, cb * dy);
            const float qy = q * fmaf(cc, dy, cb * dx);
            sx += qx; sy += qy; ax += fabsf(qx); ay += fabsf(qy);
            const float qdx = q * dx, qdy = q * dy;
            cxx = fmaf(qdx, dx, cxx);
            cxy = fmaf(qdx, dy, cxy);
            cyy = fmaf(qdy, dy, cyy);
            so += q;
        }
        float* acc = sm.acc[my_slot];
        const float kx = op * half_w, ky = op * half_h;
        atomicAdd(acc + 0, -kx * sx);
        atomicAdd(acc + 1, -ky * sy);
        atomicAdd(acc + 2, fabsf(kx) * ax);
        atomicAdd(acc + 3, fabsf(ky) * ay);
        atomicAdd(acc + 4, -0.5f * op * cxx);
        atomicAdd(acc + 5, -0.5f * op * cxy);
        atomicAdd(acc + 6, -0.5f * op * cyy);
        atomicAdd(acc + 7, so);
#pragma unroll
        for (int i = 0; i < NC; ++i) atomicAdd(acc + 8 + i, gc[i]);
    }
    __syncwarp();
}

template <int F>
__global__ void __launch_bounds__(BLEND_THREADS) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int tiles_x,
    const float4* __restrict__ rec_a, const float4* __restrict__ rec_b, const float4* __restrict__ rgb,
    const float* __restrict__ features, const float* __restrict__ bg, const float* __restrict__ final_T,
    const uint32_t* __restrict__ n_contrib, const float* __restrict__ grad_color, const float* __restrict__ grad_buffer,
    float* __restrict__ grad_acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem<F>& sm = *reinterpret_cast<BwdSmem<F>*>(smem_raw);
    constexpr int NV = BwdSmem<F>::NV;
    constexpr int NC = 3 + F;
    constexpr int NG = BwdSmem<F>::NG;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    int px, py;
    pixel_of_thread(tile_x, tile_y, tid, px, py);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;
    const size_t N = (size_t)W * H;
    const size_t pix = (size_t)py * W + px;

    const uint2 range = ranges[tile_y * tiles_x + tile_x];
    const int n_list = (int)(range.y - range.x);

    // per-pixel state
    const float T_final = inside ? final_T[pix] : 0.f;
    const uint32_t my_contrib = inside ? n_contrib[pix] : 0u;
    float T = T_final;
    float dL[4 * NV];
#pragma unroll
    for (int i = 0; i < 4 * NV; ++i) dL[i] = 0.f;
    if (inside) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dL[ch] = grad_color[ch * N + pix];
#pragma unroll
        for (int ch = 0; ch < F; ++ch) dL[3 + ch] = grad_buffer[ch * N + pix];
    }
    const float bg_dot = bg[0] * dL[0] + bg[1] * dL[1] + bg[2] * dL[2];
#pragma unroll
    for (int k = 0; k < NV; ++k) sm.dpix[warp][lane][k] = make_float4(dL[4 * k], dL[4 * k + 1], dL[4 * k + 2], dL[4 * k + 3]);

    // deepest contributor of this warp / of the tile: nothing behind it is ever staged
    uint32_t wmax = my_contrib;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) sm.warp_max[warp] = wmax;
    __syncthreads();
    uint32_t tile_max = 0;
#pragma unroll
    for (int w = 0; w < BLEND_WARPS; ++w) tile_max = max(tile_max, sm.warp_max[w]);
    if (tile_max > (uint32_t)n_list) tile_max = (uint32_t)n_list;  // defensive; forward guarantees <=
    const int n_back = (int)tile_max;                                // entries [0, n_back) in forward order matter
    const int rounds = (n_back + BWD_BATCH - 1) / BWD_BATCH;

    const float wpx0 = (float)(tile_x * GS2M_TILE_X + (warp & 1) * WARP_PIX_X);
    const float wpy0 = (float)(tile_y * GS2M_TILE_Y + (warp >> 1) * WARP_PIX_Y);
    const float half_w = 0.5f * (float)W, half_h = 0.5f * (float)H;

    float S = 0.f, last_alpha = 0.f, last_cd = 0.f;
    int n_parked = 0, my_slot = 0;

    for (int batch = 0; batch < rounds; ++batch) {
        // ---- stage entries f = n_back-1-(batch*BATCH+t), t < BATCH, in back-to-front order ----
        uint32_t mask = 0;
        int staged_id = -1;
        if (tid < BWD_BATCH) {
            const int f = n_back - 1 - (batch * BWD_BATCH + tid);
            if (f >= 0) {
                staged_id = (int)point_list[range.x + f];
                const float4 ra = __ldg(rec_a + staged_id);
                const float4 rb = __ldg(rec_b + staged_id);
                const CullRecord cr = make_cull_record(ra, rb);
                mask = warp_block_mask(cr, tile_x * GS2M_TILE_X, tile_y * GS2M_TILE_Y);
#pragma unroll
                for (int w = 0; w < BLEND_WARPS; ++w)
                    if ((uint32_t)f >= sm.warp_max[w]) mask &= ~(1u << w);
                if (mask) {
                    sm.a[tid] = ra;
                    sm.b[tid] = rb;
                    const float4 c = __ldg(rgb + staged_id);
                    float v[4 * NV];
                    v[0] = c.x; v[1] = c.y; v[2] = c.z;
#pragma unroll
                    for (int i = 3; i < 4 * NV; ++i) v[i] = 0.f;
                    if (F > 0) {
                        const float2* f2 = reinterpret_cast<const float2*>(features + (size_t)staged_id * GS2M_NUM_FEATURES);
#pragma unroll
                        for (int i = 0; i < (F + 1) / 2; ++i) {
                            const float2 t = __ldg(f2 + i);
                            v[3 + 2 * i] = t.x;
                            if (2 * i + 1 < F) v[3 + 2 * i + 1] = t.y;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NV; ++k)
                        sm.col[k][tid] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
#pragma unroll
                    for (int i = 0; i < NG; ++i) sm.acc[tid][i] = 0.f;
                }
            }
#pragma unroll
            for (int w = 0; w < BLEND_WARPS; ++w) {
                const uint32_t word = __ballot_sync(0xffffffffu, (mask >> w) & 1u);
                if (lane == 0) sm.words[w][warp] = word;
            }
            const uint32_t any = __ballot_sync(0xffffffffu, mask != 0);
            if (lane == 0) sm.touched[warp] = any;
        }
        __syncthreads();

        // ---- evaluate (lane = pixel) / reduce (lane = parked entry) ----
        for (int sw = 0; sw < BWD_BATCH / 32; ++sw) {
            uint32_t word = sm.words[warp][sw];
            while (word != 0) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                const int slot = sw * 32 + bit;
                const uint32_t f = (uint32_t)(n_back - 1 - (batch * BWD_BATCH + slot));
                float pw = 0.f, pq = 0.f;
                if (f < my_contrib) {
                    const float4 ra = sm.a[slot];
                    const float4 rb = sm.b[slot];
                    float dx, dy, G, alpha;
                    if (pair_alpha(ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, pxf, pyf, dx, dy, G, alpha)) {
                        const float one_m_alpha = 1.0f - alpha;
                        T = __fdiv_rn(T, one_m_alpha);
                        float c[4 * NV];
#pragma unroll
                        for (int k = 0; k < NV; ++k) {
                            const float4 t = sm.col[k][slot];
                            c[4 * k] = t.x; c[4 * k + 1] = t.y; c[4 * k + 2] = t.z; c[4 * k + 3] = t.w;
                        }
                        float cd = 0.f;
#pragma unroll
                        for (int i = 0; i < NC; ++i) cd = fmaf(c[i], dL[i], cd);
                        S = fmaf(last_alpha, last_cd - S, S);          // a_prev*cd_prev + (1-a_prev)*S
                        last_alpha = alpha;
                        last_cd = cd;
                        float dL_dalpha = (cd - S) * T;
                        dL_dalpha += (-T_final / one_m_alpha) * bg_dot;
                        pw = alpha * T;
                        pq = dL_dalpha * G;
                    }
                }
                if (lane == n_parked) my_slot = slot;
                sm.park_w[warp][n_parked * PARK_STRIDE + lane] = pw;
                sm.park_q[warp][n_parked * PARK_STRIDE + lane] = pq;
                ++n_parked;
                if (n_parked == PARK) {
                    __syncwarp();
                    reduce_parked<F>(sm, warp, lane, n_parked, my_slot, wpx0, wpy0, half_w, half_h);
                    n_parked = 0;
                }
            }
        }
        if (n_parked > 0) {
            __syncwarp();
            reduce_parked<F>(sm, warp, lane, n_parked, my_slot, wpx0, wpy0, half_w, half_h);
            n_parked = 0;
        }
        __syncthreads();

        // ---- one set of global atomics per (Gaussian, tile) ----
        if (tid < BWD_BATCH && mask != 0) {
            float* dst = grad_acc + (size_t)staged_id * GS2M_ACC_STRIDE;
#pragma unroll
            for (int i = 0; i < NG; ++i) atomicAdd(dst + i, sm.acc[tid][i]);
        }
        // the next round's staging writes a/b/col/acc of slots whose flush (same thread) is complete; the words are
        // rewritten by the staging warps only after every consumer passed the barrier above.
    }
}

template <int F>
int launch_b(const BwdParams& p, const GeomState& g, const uint32_t* point_list, const ImageState& im, cudaStream_t s) {
    dim3 grid(p.tiles_x, p.tiles_y);
    const size_t smem = sizeof(BwdSmem<F>);
    static bool configured = false;
    if (!configured) {
        GS2M_CUDA(cudaFuncSetAttribute(blend_backward_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    count_launches(1);
    blend_backward_kernel<F><<<grid, BLEND_THREADS, smem, s>>>(im.ranges, point_list, p.W, p.H, p.tiles_x, g.xy_conic_ab,
                                                               g.conic_c_opac, g.rgb, p.features, p.background,
                                                               im.final_T, im.n_contrib, p.grad_color, p.grad_buffer,
                                                               g.grad_acc);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace

int launch_blend_backward(const BwdParams& p, const GeomState& g, const uint32_t* point_list, const ImageState& im,
                          cudaStream_t s) {
    GS2M_CUDA(cudaMemsetAsync(g.grad_acc, 0, (size_t)p.P * GS2M_ACC_STRIDE * sizeof(float), s));
    switch (p.F) {
#define GS2M_CASE(N) case N: return launch_b<N>(p, g, point_list, im, s);
        GS2M_CASE(0) GS2M_CASE(1) GS2M_CASE(2) GS2M_CASE(3) GS2M_CASE(4) GS2M_CASE(5)
        GS2M_CASE(6) GS2M_CASE(7) GS2M_CASE(8) GS2M_CASE(9) GS2M_CASE(10)
#undef GS2M_CASE
    }
    set_error("feature_count %d outside 0..10", p.F);
    return GS2M_ERR_INVALID_ARGUMENT;
}

}  // namespace gs2m
