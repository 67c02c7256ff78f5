// Pieces shared by the forward and backward blend kernels: pixel <-> thread mapping, the bit-exact per-pair alpha
// evaluation, and the conservative footprint test that lets whole warps skip Gaussians.
#pragma once
#include "common.cuh"

namespace gs2m {

// One CTA = one 16x16 tile = 8 warps; a warp covers an 8x4 pixel block (more square than the reference's 16x2 rows,
// so a splat's footprint intersects fewer warps).  Warp w sits at block column (w & 1), block row (w >> 1).
constexpr int WARP_PIX_X = 8, WARP_PIX_Y = 4;
constexpr int BLEND_THREADS = 256;
constexpr int BLEND_WARPS = 8;

__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int tid, int& px, int& py) {
    const int warp = tid >> 5, lane = tid & 31;
    px = tile_x * GS2M_TILE_X + (warp & 1) * WARP_PIX_X + (lane & 7);
    py = tile_y * GS2M_TILE_Y + (warp >> 1) * WARP_PIX_Y + (lane >> 3);
}

// Packed fp32 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2 work on an aligned register pair and take ONE issue slot for two
// IEEE round-to-nearest results, so each half is bit-identical to the scalar __fmaf_rn / __fmul_rn / __fadd_rn).  The blend
// kernels are issue-bound, and their channel loops (colour + features: 13 values per pair) are where the pairs come for free:
// the staged colour vectors arrive as 128-bit shared loads, i.e. already in aligned pairs.
#ifndef GS2M_F32X2
#define GS2M_F32X2 1
#endif
__device__ __forceinline__ float2 fma2_rn(float2 a, float2 b, float2 c) {
#if GS2M_F32X2
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
#else
    return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
#if GS2M_F32X2
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
#endif
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
#if GS2M_F32X2
    float2 d;
    asm("{ .reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
#endif
}

// Asynchronous global -> shared copies (LDGSTS): the staging of a list step's records never passes through registers,
// so it can be issued a whole step ahead of its use without raising the register budget.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Staged records of the blend kernels.  A warp walks ITS list — the tile's entries whose footprint mask has the warp block's bit,
// compacted once per view by footprint_masks.cu (dense_gid / block_ranges) — so every entry it loads is one it evaluates.
// Entry e of the list is staged by lane e % 32 into slot e % 32 of a 32-slot ring, sixteen entries ("half-step") at a time:
// while the warp evaluates one half of the ring the copies of the other half are in flight.  Slot layout:
//   a   = (mean.x, mean.y, conic.a, conic.b)                    cp.async 16 B from the blend record
//   b   = (conic.c, opacity | Gaussian index, -)                 cp.async 8 B + one store by the owning lane
//   col = (r, g, b, 0) | f0..f3 | f4..f7 | f8, f9, -, -          cp.async 16 B (rgb) + 8 B per feature pair
// Channel i of the blended vector therefore sits at position i for the colour and 4 + i for feature i; the '-' words are
// never read as a used channel.
template <int F>
struct StagedRing {
    static constexpr int NV = 1 + (F + 3) / 4;   // float4s per staged colour + feature vector
    static constexpr int NPAIR = 2 + (F + 1) / 2; // aligned channel pairs: (r,g) (b,0) (f0,f1) ...
    float4 a[32];
    float4 b[32];
    float4 col[NV][32];
};
// position of blended channel ch (0..2 colour, 3.. features) inside the staged vector / the pair accumulators
__host__ __device__ constexpr int staged_pos(int ch) { return ch < 3 ? ch : ch + 1; }

// Issues the copies of one half-step: the lanes with `active` stage their entry (Gaussian `gid`) into slot `lane`.  Always
// commits one group (possibly empty) so that every lane's group count stays in step.
template <int F>
__device__ __forceinline__ void stage_half(StagedRing<F>& sm, int lane, bool active, int gid,
                                           const float4* __restrict__ rec_a, const float4* __restrict__ rec_b,
                                           const float4* __restrict__ rgb, const float* __restrict__ features) {
    if (active) {
        cp_async16(&sm.a[lane], rec_a + gid);
        cp_async8(&sm.b[lane], rec_b + gid);
        reinterpret_cast<int*>(&sm.b[lane])[2] = gid;
        cp_async16(&sm.col[0][lane], rgb + gid);
        if (F > 0) {
            const float* frow = features + (size_t)gid * GS2M_NUM_FEATURES;
#pragma unroll
            for (int i = 0; i < (F + 1) / 2; ++i)
                cp_async8(reinterpret_cast<float2*>(&sm.col[1 + i / 2][lane]) + (i & 1), frow + 2 * i);
        }
    }
    cp_async_commit();
}

// Per-pair evaluation, bit-identical to the reference's renderCUDA as compiled for sm_100
// (cuda_rasterizer/forward.cu:325-339; SASS: dx*a, dy*c, dy*(dy*c), fma(dx, dx*a, .), dy*(dx*b), fma(., -0.5, -.),
// accurate expf, min(0.99, o*E)).  Returns false when the pair is skipped (power > 0 or alpha < 1/255).
// Branch-free (same arithmetic, same decisions as the reference's early-outs): lets two independent evaluations be interleaved.
__device__ __forceinline__ bool pair_alpha_nb(float gx, float gy, float ca, float cb, float cc, float opacity,
                                              float pxf, float pyf, float& dx, float& dy, float& G, float& alpha) {
    dx = __fadd_rn(gx, -pxf);
    dy = __fadd_rn(gy, -pyf);
    const float s = __fmaf_rn(dx, __fmul_rn(dx, ca), __fmul_rn(dy, __fmul_rn(dy, cc)));
    const float u = __fmul_rn(dy, __fmul_rn(dx, cb));
    const float power = __fmaf_rn(s, -0.5f, -u);
    G = expf(power);
    alpha = fminf(__fmul_rn(opacity, G), 0.99f);
    return !(power > 0.0f) & !(alpha < 0.00392156885936856f);
}

// Lower bound (over an axis-aligned pixel rectangle [x0,x1]x[y0,y1]) of the conic quadratic
//   q(d) = a dx^2 + 2 b dx dy + c dy^2,   d = pixel - mean,
// valid when the conic is positive definite.  A pair can only pass the alpha >= 1/255 test where q <= thr
// (thr = 2 ln(255 opacity), stored by preprocess), so a lower bound above thr proves that no pixel of the
// rectangle is blended and the reference's per-pixel tests would all have skipped it.  The margin absorbs the
// rounding of both this bound and the reference's fp32 power/exp evaluation (scaled by the largest term magnitude
// that can occur inside the rectangle), so the test only ever errs toward "may contribute".
struct CullRecord {
    float mx, my, a, b, c, thr, inv_a, inv_c;
    bool cullable;  // false: degenerate / indefinite conic -> never cull
};

__device__ __forceinline__ CullRecord make_cull_record(float4 ra, float4 rb) {
    CullRecord r;
    r.mx = ra.x; r.my = ra.y; r.a = ra.z; r.b = ra.w; r.c = rb.x; r.thr = rb.z;
    const float det = r.a * r.c - r.b * r.b;
    r.cullable = (r.a > 0.f) && (r.c > 0.f) && (det > 1e-5f * r.a * r.c);
    r.inv_a = __fdividef(1.0f, r.a);   // approximate reciprocals are fine: only the conservative bound uses them and its
    r.inv_c = __fdividef(1.0f, r.c);   // margin is orders of magnitude above their 2-ulp error
    return r;
}

// 8-bit mask: bit w set when warp w's 8x4 pixel block may receive this Gaussian.  Per block: thr < 0 (opacity < 1/255) can
// never contribute (exact); a degenerate conic is never culled; the mean inside the block gives q = 0; otherwise the minimum
// of q over the rectangle lies on the edge(s) nearest to the mean, where q is a 1-D quadratic minimised in closed form with
// the minimiser clamped to the edge.  Everything that only depends on the block's column (2 of them) or row (4) is hoisted.
__device__ __forceinline__ uint32_t warp_block_mask(const CullRecord& r, int tile_px0, int tile_py0) {
    if (r.thr < 0.f) return 0u;
    if (!r.cullable) return 0xFFu;
    const float b2 = 2.f * r.b, ab = fabsf(b2);
    // per column c (x range [lx,hx]) / per row w (y range [ly,hy])
    float lx[2], hx[2], ex[2], qx0[2], qx1[2], tx[2], mx2[2], mxb[2];
    bool inx[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        lx[c] = (float)(tile_px0 + c * WARP_PIX_X) - r.mx;
        hx[c] = lx[c] + (float)(WARP_PIX_X - 1);
        inx[c] = (lx[c] <= 0.f) && (hx[c] >= 0.f);
        ex[c] = (lx[c] > 0.f) ? lx[c] : hx[c];          // nearest vertical edge (used when !inx)
        qx0[c] = r.a * ex[c] * ex[c];                   // q on that edge = qx0 + (qx1 + c*dy)*dy
        qx1[c] = b2 * ex[c];
        tx[c] = -r.b * ex[c] * r.inv_c;                 // unconstrained minimiser in dy
        const float ax = fmaxf(fabsf(lx[c]), fabsf(hx[c]));
        mx2[c] = r.a * ax * ax;
        mxb[c] = ab * ax;
    }
    uint32_t m = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const float ly = (float)(tile_py0 + w * WARP_PIX_Y) - r.my, hy = ly + (float)(WARP_PIX_Y - 1);
        const bool iny = (ly <= 0.f) && (hy >= 0.f);
        const float ey = (ly > 0.f) ? ly : hy;
        const float qy0 = r.c * ey * ey, qy1 = b2 * ey, ty = -r.b * ey * r.inv_a;
        const float ay = fmaxf(fabsf(ly), fabsf(hy));
        const float my2 = r.c * ay * ay;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            bool may = inx[c] && iny;
            if (!may) {
                float qmin = 3.0e38f;
                if (!inx[c]) {
                    const float dy = fminf(fmaxf(tx[c], ly), hy);
                    qmin = qx0[c] + (qx1[c] + r.c * dy) * dy;
                }
                if (!iny) {
                    const float dx = fminf(fmaxf(ty, lx[c]), hx[c]);
                    qmin = fminf(qmin, qy0 + (qy1 + r.a * dx) * dx);
                }
                const float mag = mx2[c] + my2 + mxb[c] * ay;
                may = !(qmin > r.thr + 1e-5f * mag + 1e-3f);
            }
            if (may) m |= 1u << (2 * w + c);
        }
    }
    return m;
}

}  // namespace gs2m
