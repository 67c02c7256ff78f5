// Forward blend: one CTA per 16x16 tile walks the tile's depth-sorted Gaussian list front to back and alpha-blends
// RGB plus the first F feature columns (alpha, distance, normal, albedo, roughness, metallic) into every pixel.
//
// Behavioural reference: renderCUDA (cuda_rasterizer/forward.cu:246-372).  Per pixel the sequence of
// (power, alpha, test_T, T) values, the termination point, `n_contrib` and the `observe` counts are bit-identical to
// the reference; what differs is how the work is organised:
//   * each batch of 256 list entries is staged once into shared memory as 16-byte records (position/conic,
//     opacity, colour + feature vector) with 128-bit loads, instead of being re-fetched from global per pair;
//   * while staging, every thread proves for its Gaussian which of the tile's eight 8x4-pixel warp blocks can
//     possibly receive alpha >= 1/255 (rect_may_contribute); the eight ballots become per-warp bit masks and a
//     warp only ever evaluates the entries whose bit is set;
//   * a warp leaves the batch as soon as all of its 32 pixels have terminated; the CTA stops staging when all 256
//     have (the reference's only exit);
//   * `observe` increments are aggregated per warp (ballot+popc) into shared counters and flushed once per staged
//     entry — integer sums, so the totals are exact;
//   * F is a template parameter: accumulators stay in registers and the loops unroll.
#include "blend_common.cuh"

namespace gs2m {
namespace {

template <int F>
struct FwdSmem {
    static constexpr int NV = (3 + F + 3) / 4;  // float4s per staged colour+feature vector
    float4 a[BLEND_THREADS];                    // mean.x, mean.y, conic.a, conic.b
    float4 b[BLEND_THREADS];                    // conic.c, opacity, -, -
    float4 col[NV][BLEND_THREADS];              // r,g,b,f0 | f1..f4 | f5..f8 | f9,0,0,0
    uint32_t words[BLEND_WARPS][BLEND_WARPS];   // [consumer warp][staging warp] -> 32 entry bits
    int obs[BLEND_THREADS];
};

template <int F>
__global__ void __launch_bounds__(BLEND_THREADS) blend_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int tiles_x,
    const float4* __restrict__ rec_a, const float4* __restrict__ rec_b, const float4* __restrict__ rgb,
    const float* __restrict__ features, const float* __restrict__ bg, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, int* __restrict__ out_observe,
    float* __restrict__ out_buffer) {
    __shared__ FwdSmem<F> sm;
    constexpr int NV = FwdSmem<F>::NV;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    int px, py;
    pixel_of_thread(tile_x, tile_y, tid, px, py);
    const bool inside = (px < W) && (py < H);
    const float pxf = (float)px, pyf = (float)py;

    const uint2 range = ranges[tile_y * tiles_x + tile_x];
    const int n_list = (int)(range.y - range.x);
    const int rounds = (n_list + BLEND_THREADS - 1) / BLEND_THREADS;

    bool done = !inside;
    float T = 1.0f;
    uint32_t last_contributor = 0;
    float C[3] = {0.f, 0.f, 0.f};
    float Fv[F > 0 ? F : 1];
#pragma unroll
    for (int i = 0; i < (F > 0 ? F : 1); ++i) Fv[i] = 0.f;

    int staged_id = 0;
    for (int batch = 0;; ++batch) {
        const int num_done = __syncthreads_count(done);  // also: every warp has left the previous batch
        if (batch > 0) {
            const int c = sm.obs[tid];
            if (c != 0) atomicAdd(out_observe + staged_id, c);
        }
        if (batch == rounds || num_done == BLEND_THREADS) break;

        // ---- stage one batch: 128-bit gathers of the blend records, footprint masks ----
        const int li = batch * BLEND_THREADS + tid;
        uint32_t mask = 0;
        if (li < n_list) {
            staged_id = (int)point_list[range.x + li];
            const float4 ra = __ldg(rec_a + staged_id);
            const float4 rb = __ldg(rec_b + staged_id);
            const CullRecord cr = make_cull_record(ra, rb);
            mask = warp_block_mask(cr, tile_x * GS2M_TILE_X, tile_y * GS2M_TILE_Y);
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
                for (int k = 0; k < NV; ++k) sm.col[k][tid] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
        }
        sm.obs[tid] = 0;
#pragma unroll
        for (int w = 0; w < BLEND_WARPS; ++w) {
            const uint32_t word = __ballot_sync(0xffffffffu, (mask >> w) & 1u);
            if (lane == 0) sm.words[w][warp] = word;
        }
        __syncthreads();

        // ---- blend: each warp walks only the entries whose footprint reaches its 8x4 block ----
        bool warp_done = __all_sync(0xffffffffu, done);
        const uint32_t base_contrib = (uint32_t)(batch * BLEND_THREADS) + 1u;
        for (int sw = 0; sw < BLEND_WARPS && !warp_done; ++sw) {
            uint32_t word = sm.words[warp][sw];
            while (word != 0 && !warp_done) {
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                const int slot = sw * 32 + bit;
                const float4 ra = sm.a[slot];
                const float4 rb = sm.b[slot];
                bool obs = false;
                if (!done) {
                    float dx, dy, G, alpha;
                    if (pair_alpha(ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, pxf, pyf, dx, dy, G, alpha)) {
                        const float test_T = __fmul_rn(T, __fadd_rn(1.0f, -alpha));
                        if (test_T < 0.0001f) {
                            done = true;
                        } else {
                            float v[4 * NV];
#pragma unroll
                            for (int k = 0; k < NV; ++k) {
                                const float4 t = sm.col[k][slot];
                                v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
                            }
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) C[ch] = __fmaf_rn(T, __fmul_rn(alpha, v[ch]), C[ch]);
#pragma unroll
                            for (int ch = 0; ch < F; ++ch) Fv[ch] = __fmaf_rn(T, __fmul_rn(alpha, v[3 + ch]), Fv[ch]);
                            obs = T > 0.5f;
                            T = test_T;
                            last_contributor = base_contrib + (uint32_t)slot;
                        }
                    }
                }
                const uint32_t ob = __ballot_sync(0xffffffffu, obs);
                if (ob != 0 && lane == 0) atomicAdd(&sm.obs[slot], __popc(ob));
                warp_done = __all_sync(0xffffffffu, done);
            }
        }
    }

    if (inside) {
        const size_t N = (size_t)W * H;
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last_contributor;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) out_color[ch * N + pix] = __fmaf_rn(T, bg[ch], C[ch]);
#pragma unroll
        for (int ch = 0; ch < GS2M_NUM_FEATURES; ++ch) out_buffer[ch * N + pix] = (ch < F) ? Fv[ch < F ? ch : 0] : 0.f;
    }
}

template <int F>
int launch_f(const FwdParams& p, const GeomState& g, const uint32_t* point_list, const ImageState& im, float* out_color,
             int* out_observe, float* out_buffer, cudaStream_t s) {
    dim3 grid(p.tiles_x, p.tiles_y);
    count_launches(1);
    blend_forward_kernel<F><<<grid, BLEND_THREADS, 0, s>>>(im.ranges, point_list, p.W, p.H, p.tiles_x, g.xy_conic_ab,
                                                           g.conic_c_opac, g.rgb, p.features, p.background, im.final_T,
                                                           im.n_contrib, out_color, out_observe, out_buffer);
    GS2M_CUDA(cudaGetLastError());
    return GS2M_OK;
}

}  // namespace

int launch_blend_forward(const FwdParams& p, const GeomState& g, const uint32_t* point_list, const ImageState& im,
                         float* out_color, int* out_observe, float* out_buffer, cudaStream_t s) {
    switch (p.F) {
#define GS2M_CASE(N) case N: return launch_f<N>(p, g, point_list, im, out_color, out_observe, out_buffer, s);
        GS2M_CASE(0) GS2M_CASE(1) GS2M_CASE(2) GS2M_CASE(3) GS2M_CASE(4) GS2M_CASE(5)
        GS2M_CASE(6) GS2M_CASE(7) GS2M_CASE(8) GS2M_CASE(9) GS2M_CASE(10)
#undef GS2M_CASE
    }
    set_error("feature_count %d outside 0..10", p.F);
    return GS2M_ERR_INVALID_ARGUMENT;
}

}  // namespace gs2m
