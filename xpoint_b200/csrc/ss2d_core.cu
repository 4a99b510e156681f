// ss2d_core.cu -- SS2D core with CrossMerge fused into the selective scan (xp_ss2d_core), d_state 1 | 2.
//
// Reference path replaced: VMamba.py:603-632 (cross_scan_fn -> x_proj / dt_proj -> selective_scan_fn -> cross_merge_fn)
// and csm_triton.py:56-85 (CrossMerge).  The op-level scan (selective_scan.cu) writes one fp32 y plane per direction
// (4 x 4 B per element) that a second kernel reads back to sum them.  Here the four directions of a channel meet in
// SHARED memory and the merged y is written once (4 B per element):
//
//   * a CTA owns the token planes of CH channels of one image: accumulator Y0 (row-major token order, directions
//     "row forward" + "row backward") and Y1 (column-major order, "column forward" + "column backward"), fp32;
//   * four scan chains (one per direction) run concurrently, each on WPC consumer warps.  A chain's (step, channel)
//     items are dealt round-robin to its warps; a step is 32*C consecutive tokens (lanes scan C tokens in registers, the
//     32 lane chunks are combined by a warp-shuffle scan of the affine maps, like scan_lanes_kernel), and the running
//     state is handed from step to step through a 16-byte shared-memory slot {h[0], h[1], sequence} -- so ONE long row
//     (20 480 tokens at 512x640 stage 0) is scanned by several warps in a software pipeline;
//   * every warp requests its own items (u / delta / B / C chunks, 1-D bulk TMA copies completing on an mbarrier) into a
//     private ring, one item ahead of the one it scans, also across plane boundaries;
//   * forward and backward chains of a layout write the same accumulator: the first visitor of a 32*C-token block
//     stores, the second adds (a 3-state flag per block; the sum of two terms is order-independent, so results are
//     bit-reproducible);
//   * epilogue: out[h][w] = Y0[h*W + w] + Y1[w*H + h], coalesced fp32 stores of the (B, D, H, W) merged plane.  Y0 / Y1
//     use a 16-byte-vector XOR swizzle (and Y1 a padded column pitch) so that the accumulate stores, the second-visitor
//     loads and the transposing epilogue reads are free of bank conflicts.
//
// Inputs are what the copy-free SS2D path already produces: xx = [x ; x^T] (xp_ss2d_dwconv_pack), delta (xp_ss2d_dt_proj)
// and B / C as strided views of the x_proj output, all in the fused direction order [row fwd, row bwd, col fwd, col bwd].
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace xp {

struct CoreParams {
    const void* xx; const void* delta; const void* Bm; const void* Cm;
    const float* A; const float* Ds; const float* bias;
    float* out;
    int64_t batch, D, L;
    int H, W;
    int64_t B_bs, B_gs, B_ss, C_bs, C_gs, C_ss;
    int softplus;
    int CH;          // channels per CTA (D % CH == 0)
    int nblk;        // token blocks (steps) per row
    int y1_pitch;    // words between columns of the Y1 accumulator (H + 4 | H + 8: (pitch / 4) odd)
    int spw;         // ring slots per consumer warp
    int64_t nplanes; // batch * D / CH
};

constexpr int SC_MAX_SPW = 3;

// Every warp owns a PRIVATE ring of SPW item slots and issues its own bulk-TMA loads: right after it has copied item j into
// registers, its lane 0 requests item j + SPW into the slot just drained (also across plane boundaries), so the next item
// is in flight while this one is scanned.  No producer warp (one warp feeding four chains was bound by its own instruction
// stream: ncu showed the consumers waiting on the "full" barriers 45 % of the time) and no "empty" barriers.  A warp
// consumes its items strictly in order, so it never waits on an mbarrier phase it is more than one fill ahead of.
// An item carries everything one (step, channel) needs: u, delta and the step's B / C rows (rows of one step re-read B / C
// through L2 when a CTA holds several channels).
// WPC = warps per chain (direction); the CTA has 4 * WPC warps
template <int NST, typename IN_T, int C, int WPC> struct CoreCfg {
    static constexpr int NCW = 4 * WPC;
    static constexpr int TOK = 32 * C;
    static constexpr int CHUNK = TOK * (int)sizeof(IN_T);
    static constexpr int ITEM = (2 + 2 * NST) * CHUNK;        // u, delta, B rows, C rows
    static constexpr int RCF = 8;                             // floats of per-(chain, channel) constants
    __host__ __device__ static int y0_bytes(int L) { return L * 4; }
    __host__ __device__ static int y1_bytes(int W, int pitch) { return W * pitch * 4; }
    __host__ __device__ static int ring_bytes(int spw) { return NCW * spw * ITEM; }
    __host__ __device__ static int nbars(int spw) { return NCW * spw; }
    // layout: [Y0 x CH][Y1 x CH][rings: warp-major][rc 4 x CH x RCF][carry 4 x CH x 16 B][vis 2 x CH x nblk][bars]
    __host__ __device__ static int smem(int CH, int L, int W, int pitch, int nblk, int spw) {
        return CH * (y0_bytes(L) + y1_bytes(W, pitch)) + ring_bytes(spw) + 4 * CH * RCF * 4 + 4 * CH * 16
               + ((2 * CH * nblk * 4 + 15) & ~15) + nbars(spw) * 8 + 16;
    }
};

__device__ __forceinline__ uint4 lds128_volatile(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32u_volatile(uint32_t saddr, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ float lds32f(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
// accumulator swizzle on a word (token) index: XOR the 16-byte-vector index with bits of the 128-byte line index
__device__ __forceinline__ uint32_t acc_swz(uint32_t t) { return t ^ (((t >> 5) & 3u) << 2); }

template <typename IN_T> __device__ __forceinline__ void widen16c(const uint4& v, float (&f)[16 / sizeof(IN_T)]);
template <> __device__ __forceinline__ void widen16c<float>(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void widen16c<__half>(const uint4& v, float (&f)[8]) { VecIO<__half, 8>::widen(v, f); }
template <> __device__ __forceinline__ void widen16c<__nv_bfloat16>(const uint4& v, float (&f)[8]) { VecIO<__nv_bfloat16, 8>::widen(v, f); }

// C consecutive elements from shared memory, widened to fp32; REV hands them back in reversed order
template <typename IN_T, int C, bool REV> __device__ __forceinline__ void lds_chunk(uint32_t saddr, float (&f)[C]) {
    constexpr int PER = 16 / (int)sizeof(IN_T);
#pragma unroll
    for (int v = 0; v < C / PER; ++v) {
        float w[PER];
        widen16c<IN_T>(lds128(saddr + 16 * v), w);
#pragma unroll
        for (int i = 0; i < PER; ++i) f[REV ? C - 1 - (v * PER + i) : v * PER + i] = w[i];
    }
}

// ------------------------------------------------------------------------------------------------ loads
// position of a warp's load stream: its items (wi, wi + WPC, ...) of plane after plane, with the derived indices kept
// incrementally (no integer divisions per item)
struct CoreCursor {
    int64_t plane;
    int step, r;                   // item = step * CH + r
    int64_t u0, dl0, b0, c0;       // element offsets of the plane's first channel row in xx / delta, of its image in B / C
};
__device__ __forceinline__ void cursor_set_plane(CoreCursor& c, const CoreParams& p, int64_t plane, int wi, int q) {
    const int cpb = (int)(p.D / p.CH);
    c.plane = plane;
    const int b = (int)(plane / cpb);
    const int d0 = ((int)plane - b * cpb) * p.CH;
    c.u0 = (((int64_t)b * 2 + (q >> 1)) * p.D + d0) * p.L;
    c.dl0 = (((int64_t)b * 4 + q) * p.D + d0) * p.L;
    c.b0 = b * p.B_bs + q * p.B_gs;
    c.c0 = b * p.C_bs + q * p.C_gs;
    if (p.CH == 1) { c.step = wi; c.r = 0; }
    else { c.step = wi / p.CH; c.r = wi - c.step * p.CH; }
}
__device__ __forceinline__ void cursor_advance(CoreCursor& c, const CoreParams& p, int wi, int q, int grid, int wpc) {
    if (p.CH == 1) {
        c.step += wpc;
    } else {
        c.r += wpc;
        while (c.r >= p.CH) { c.r -= p.CH; ++c.step; }
    }
    if (c.step >= p.nblk) cursor_set_plane(c, p, c.plane + grid, wi, q);
}

// (called by lane 0 of the owning warp) request the cursor's item for chain q into the ring slot at `dst`
template <int NST, typename IN_T, int C, int WPC>
__device__ __forceinline__ void core_issue(const CoreParams& p, int q, const CoreCursor& c, uint8_t* dst, uint64_t* full) {
    using Cfg = CoreCfg<NST, IN_T, C, WPC>;
    constexpr int TOK = Cfg::TOK, ES = (int)sizeof(IN_T);
    const int blk = (q & 1) ? p.nblk - 1 - c.step : c.step;
    const int64_t m0 = (int64_t)blk * TOK;
    const uint32_t bytes = (uint32_t)(min((int64_t)TOK, p.L - m0) * ES);
    const int64_t row = (int64_t)c.r * p.L + m0;
    const IN_T* u = (const IN_T*)p.xx + c.u0 + row;
    const IN_T* dl = (const IN_T*)p.delta + c.dl0 + row;
    const IN_T* Bb = (const IN_T*)p.Bm + c.b0 + m0;
    const IN_T* Cb = (const IN_T*)p.Cm + c.c0 + m0;
    mbar_arrive_expect_tx(full, (2 + 2 * NST) * bytes);
    bulk_load(dst, u, bytes, full);
    bulk_load(dst + Cfg::CHUNK, dl, bytes, full);
#pragma unroll
    for (int n = 0; n < NST; ++n) {
        bulk_load(dst + (2 + n) * Cfg::CHUNK, Bb + n * p.B_ss, bytes, full);
        bulk_load(dst + (2 + NST + n) * Cfg::CHUNK, Cb + n * p.C_ss, bytes, full);
    }
}

// ------------------------------------------------------------------------------------------------ consumer
// All (step, channel) items of chain q that belong to warp wi, for the current plane.
template <int NST, typename IN_T, int C, int WPC, bool SOFTPLUS, bool REV>
__device__ __forceinline__ void core_chain(const CoreParams& p, uint8_t* ring_g, uint64_t* bars_g, uint32_t rcs, uint32_t carry,
                                           uint32_t vis, uint32_t yacc, int yacc_stride, int q, int wi, int lane, uint32_t& mycnt,
                                           CoreCursor& next) {
    using Cfg = CoreCfg<NST, IN_T, C, WPC>;
    constexpr int TOK = Cfg::TOK, RCF = Cfg::RCF;
    constexpr bool FAST = SOFTPLUS && sizeof(IN_T) == 2;
    const int CH = p.CH, nblk = p.nblk, spw = p.spw;
    const int layout = q >> 1;
    const int ci = REV ? 31 - lane : lane;                      // memory chunk of this lane inside a block
    const uint32_t lane_off = (uint32_t)(ci * C * (int)sizeof(IN_T));
    const int H = p.H, pitch = p.y1_pitch;
    const float invH = 1.0f / (float)H;
    const uint32_t ring = smem_u32(ring_g), bars = smem_u32(bars_g);
    const int grid = (int)gridDim.x;

    int step = CH == 1 ? wi : wi / CH, r = CH == 1 ? 0 : wi - step * CH;     // this warp's items: wi, wi + WPC, ...
    for (; step < nblk; ++mycnt) {
        const int blk = REV ? nblk - 1 - step : step;
        const int valid = (int)min((int64_t)TOK, p.L - (int64_t)blk * TOK);      // tokens of the block (multiple of 4)
        const uint32_t slot = mycnt % (uint32_t)spw;
        const uint32_t full = bars + slot * 8;
        mbar_wait_s(full, (mycnt / (uint32_t)spw) & 1u);
        const uint32_t item = ring + slot * Cfg::ITEM + lane_off;
        float uv[C], dv[C], Bv[NST][C], Cv[NST][C];
        lds_chunk<IN_T, C, REV>(item, uv);
        lds_chunk<IN_T, C, REV>(item + Cfg::CHUNK, dv);
#pragma unroll
        for (int n = 0; n < NST; ++n) {
            lds_chunk<IN_T, C, REV>(item + (2 + n) * Cfg::CHUNK, Bv[n]);
            lds_chunk<IN_T, C, REV>(item + (2 + NST + n) * Cfg::CHUNK, Cv[n]);
        }
        float kc[RCF];
#pragma unroll
        for (int k4 = 0; k4 < RCF / 4; ++k4) {
            const uint4 t = lds128(rcs + ((q * CH + r) * RCF + 4 * k4) * 4);
            kc[4 * k4] = __uint_as_float(t.x); kc[4 * k4 + 1] = __uint_as_float(t.y);
            kc[4 * k4 + 2] = __uint_as_float(t.z); kc[4 * k4 + 3] = __uint_as_float(t.w);
        }
        __syncwarp();                                    // every lane has drained the slot
        if (lane == 0 && next.plane < p.nplanes)         // refill it: item j + SPW of this warp loads while this one is scanned
            core_issue<NST, IN_T, C, WPC>(p, q, next, ring_g + slot * Cfg::ITEM, bars_g + slot);
        cursor_advance(next, p, wi, q, grid, WPC);
        const float kb = kc[NST], kD = kc[NST + 1];
        const uint32_t cslot = carry + (uint32_t)(q * CH + r) * 16;
        const uint32_t ybase = yacc + (uint32_t)r * (uint32_t)yacc_stride;
        const uint32_t vflag = vis + (uint32_t)(((layout * CH + r) * nblk + blk) * 4);

        // Partial block (the last block of a row): tokens past the end become identity steps by zeroing their operands --
        // u = B = C = 0 and delta' forced to 0 below -- so ONE body serves full and partial blocks (a separate predicate-free
        // copy of the body doubled the code and cost more in instruction-cache misses than the predicates it saved).
        const bool partial = valid != TOK;
        const int nvm = partial ? max(0, min(C, valid - ci * C)) : C;            // valid memory elements of this lane
        const int jlo = REV ? C - nvm : 0, jhi = REV ? C : nvm;                   // valid scan indices [jlo, jhi)
        if (partial) {
#pragma unroll
            for (int j = 0; j < C; ++j)
                if (j < jlo || j >= jhi) {                                        // stale smem may hold NaN / Inf
                    uv[j] = 0.0f; dv[j] = 0.0f;
#pragma unroll
                    for (int n = 0; n < NST; ++n) { Bv[n][j] = 0.0f; Cv[n][j] = 0.0f; }
                }
        }
        {
            float y[C], lw[C];
            if constexpr (FAST) {
                // 16-bit inputs: softplus(x) = ln2 * lg2(1 + 2^(x*log2e)); ln2 / log2e are folded into A (unscaled), bias
                // (x log2e) and the B*u product (x ln2); token-independent arithmetic on packed fp32 pairs
                const float2 kl2 = make_float2(kLog2e, kLog2e), kb2 = make_float2(kb, kb), one2 = make_float2(1.0f, 1.0f);
                const float2 kD2 = make_float2(kD, kD), ln2 = make_float2(kLn2, kLn2);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    const float2 t2 = fma2(make_float2(dv[j], dv[j + 1]), kl2, kb2);
                    const float2 s2 = add2(make_float2(ex2_approx(fminf(t2.x, 100.0f)), ex2_approx(fminf(t2.y, 100.0f))), one2);
                    float2 l2 = make_float2(fmaxf(lg2_approx(s2.x), t2.x), fmaxf(lg2_approx(s2.y), t2.y));
                    float2 u2 = make_float2(uv[j], uv[j + 1]);
                    if (partial) {                                               // identity steps: a = 2^0 = 1, b = 0
                        if (j < jlo || j >= jhi) l2.x = 0.0f;
                        if (j + 1 < jlo || j + 1 >= jhi) l2.y = 0.0f;
                    }
                    const float2 y2 = mul2(kD2, u2);
                    u2 = mul2(mul2(u2, l2), ln2);
                    lw[j] = l2.x; lw[j + 1] = l2.y;
                    y[j] = y2.x; y[j + 1] = y2.y;
                    uv[j] = u2.x; uv[j + 1] = u2.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const float d = dv[j] + kb;
                    float l = SOFTPLUS ? softplus_f(d) : d;
                    if (partial && (j < jlo || j >= jhi)) l = 0.0f;                 // identity step
                    lw[j] = l;
                    y[j] = kD * uv[j];
                    uv[j] *= l;
                }
            }
            // local scans of all states.  y[j] = sum_n (pl[n][j] * hin[n] + hl[n][j]) * C[n][j] + D u[j] is split so that only
            // one FMA per (token, state) is left behind the carry hand-off:  pc = pl * C (kept), y += hl * C (folded now)
            float P[NST], S_[NST], pc[NST][C];
#pragma unroll
            for (int n = 0; n < NST; ++n) {
                const float kA = kc[n];
                float Pn = 1.0f, Sn = 0.0f;
                // token pairs: the products that do not depend on the running state issue as packed FMUL2 / FFMA2
                const float2 kA2 = make_float2(kA, kA);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    const float2 x2 = mul2(make_float2(lw[j], lw[j + 1]), kA2);
                    const float2 b2 = mul2(make_float2(uv[j], uv[j + 1]), make_float2(Bv[n][j], Bv[n][j + 1]));
                    const float2 c2 = make_float2(Cv[n][j], Cv[n][j + 1]);
                    const float a0 = ex2_approx(x2.x), a1 = ex2_approx(x2.y);
                    float2 s2, p2;
                    Sn = fmaf(a0, Sn, b2.x); Pn *= a0; s2.x = Sn; p2.x = Pn;
                    Sn = fmaf(a1, Sn, b2.y); Pn *= a1; s2.y = Sn; p2.y = Pn;
                    const float2 pc2 = mul2(p2, c2);
                    const float2 y2 = fma2(s2, c2, make_float2(y[j], y[j + 1]));
                    pc[n][j] = pc2.x; pc[n][j + 1] = pc2.y;
                    y[j] = y2.x; y[j + 1] = y2.y;
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float Pp = __shfl_up_sync(0xffffffffu, Pn, o), Sp = __shfl_up_sync(0xffffffffu, Sn, o);
                    if (lane >= o) { Sn = fmaf(Pn, Sp, Sn); Pn *= Pp; }
                }
                P[n] = Pn; S_[n] = Sn;
            }
            float Pe[NST], Se[NST];
#pragma unroll
            for (int n = 0; n < NST; ++n) {
                Pe[n] = __shfl_up_sync(0xffffffffu, P[n], 1); Se[n] = __shfl_up_sync(0xffffffffu, S_[n], 1);
                if (lane == 0) { Pe[n] = 1.0f; Se[n] = 0.0f; }
            }
            // plane addresses of this lane's C / 4 vectors (before the hand-off: off the chain's critical path)
            uint32_t addr[C / 4];
            {
                uint32_t t = (uint32_t)(blk * TOK + ci * C);                        // first plane token of this lane
                if (layout == 0) {
#pragma unroll
                    for (int v = 0; v < C / 4; ++v) addr[v] = ybase + acc_swz(t + 4 * v) * 4u;
                } else {                                                            // column-major accumulator, padded pitch
                    int col = __float2int_rz((float)t * invH);
                    int row = (int)t - col * H;
                    if (row >= H) { row -= H; ++col; }
                    if (row < 0) { row += H; --col; }
#pragma unroll
                    for (int v = 0; v < C / 4; ++v) {
                        addr[v] = ybase + acc_swz((uint32_t)(col * pitch + row)) * 4u;
                        row += 4;
                        if (row >= H) { row = 0; ++col; }
                    }
                }
            }
            // carry slot of (chain, channel): {h[0], h[1], sequence = number of steps folded in, -}
            uint4 cv;
            {
                uint32_t spins = 0;
                do {
                    cv = lds128_volatile(cslot);
                    if (cv.z != (uint32_t)step && ++spins > (1u << 24)) __trap();
                } while (cv.z != (uint32_t)step);
            }
            const float hc[2] = {__uint_as_float(cv.x), __uint_as_float(cv.y)};
            __syncwarp();                                   // every lane has read the slot before lane 31 overwrites it
            if (lane == 31) {
                const float h0 = fmaf(P[0], hc[0], S_[0]);
                const float h1 = NST > 1 ? fmaf(P[NST - 1], hc[NST - 1], S_[NST - 1]) : 0.0f;
                sts128u(cslot, __float_as_uint(h0), __float_as_uint(h1), (uint32_t)step + 1u, 0u);
            }
#pragma unroll
            for (int n = 0; n < NST; ++n) {
                const float hin = fmaf(Pe[n], hc[n], Se[n]);
                const float2 hin2 = make_float2(hin, hin);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    const float2 y2 = fma2(make_float2(pc[n][j], pc[n][j + 1]), hin2, make_float2(y[j], y[j + 1]));
                    y[j] = y2.x; y[j + 1] = y2.y;
                }
            }
            // ---- accumulate into the plane: memory order m[i] = y[REV ? C-1-i : i]
            float m[C];
#pragma unroll
            for (int j = 0; j < C; ++j) m[REV ? C - 1 - j : j] = y[j];
            // visitor protocol per (layout, channel, block): 0 untouched, 1 first visitor storing, 2 stored
            uint32_t old = 0;
            if (lane == 0) asm volatile("atom.shared.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "r"(vflag) : "memory");
            old = __shfl_sync(0xffffffffu, old, 0);
            if (old == 0) {
#pragma unroll
                for (int v = 0; v < C / 4; ++v)
                    if (4 * v < nvm) sts128(addr[v], m[4 * v], m[4 * v + 1], m[4 * v + 2], m[4 * v + 3]);
                __syncwarp();
                if (lane == 0) { __threadfence_block(); sts32u_volatile(vflag, 2u); }
            } else {
                if (lane == 0) {
                    uint32_t spins = 0;
                    while (lds32_volatile(vflag) != 2u)
                        if (++spins > (1u << 24)) __trap();
                    __threadfence_block();
                }
                __syncwarp();
#pragma unroll
                for (int v = 0; v < C / 4; ++v)
                    if (4 * v < nvm) {
                        const uint4 o = lds128(addr[v]);
                        sts128(addr[v], m[4 * v] + __uint_as_float(o.x), m[4 * v + 1] + __uint_as_float(o.y),
                               m[4 * v + 2] + __uint_as_float(o.z), m[4 * v + 3] + __uint_as_float(o.w));
                    }
            }
        }
        if (CH == 1) {
            step += WPC;
        } else {
            r += WPC;
            while (r >= CH) { r -= CH; ++step; }
        }
    }
}

template <int NST, typename IN_T, int C, int WPC, bool SOFTPLUS>
__global__ void __launch_bounds__(4 * WPC * 32, 1) ss2d_core_kernel(const CoreParams p) {
    using Cfg = CoreCfg<NST, IN_T, C, WPC>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    const int CH = p.CH, nblk = p.nblk;
    const int y0b = Cfg::y0_bytes((int)p.L), y1b = Cfg::y1_bytes(p.W, p.y1_pitch);
    uint8_t* y0 = smem_raw;
    uint8_t* y1 = y0 + CH * y0b;
    uint8_t* rings = y1 + CH * y1b;
    uint8_t* rc = rings + Cfg::ring_bytes(p.spw);
    uint8_t* carry = rc + 4 * CH * Cfg::RCF * 4;
    uint8_t* vis = carry + 4 * CH * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(vis + ((2 * CH * nblk * 4 + 15) & ~15));
    if (tid == 0) {
        for (int i = 0; i < Cfg::nbars(p.spw); ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    const int q = warp / WPC, wi = warp % WPC;
    const uint32_t s_rc = smem_u32(rc), s_carry = smem_u32(carry), s_vis = smem_u32(vis);
    uint8_t* my_ring = rings + (size_t)warp * p.spw * Cfg::ITEM;
    uint64_t* my_bars = bars + warp * p.spw;
    const uint32_t s_yacc = (q >> 1) ? smem_u32(y1) : smem_u32(y0);
    const int yacc_stride = (q >> 1) ? y1b : y0b;
    const int64_t cpb = p.D / CH;
    constexpr bool FAST = SOFTPLUS && sizeof(IN_T) == 2;
    const int H = p.H, W = p.W, L = (int)p.L, pitch = p.y1_pitch;
    const int nitems = nblk * CH;
    // prime the ring: the first SPW items of this warp
    CoreCursor next;
    cursor_set_plane(next, p, wi < nitems ? (int64_t)blockIdx.x : p.nplanes, wi, q);   // (a warp without items never loads)
    for (int s0 = 0; s0 < p.spw; ++s0) {
        if (lane == 0 && next.plane < p.nplanes) core_issue<NST, IN_T, C, WPC>(p, q, next, my_ring + s0 * Cfg::ITEM, my_bars + s0);
        if (next.plane < p.nplanes) cursor_advance(next, p, wi, q, (int)gridDim.x, WPC);
    }
    uint32_t mycnt = 0;                                 // items this warp has consumed (ring slot / mbarrier phase)
    for (int64_t plane = blockIdx.x; plane < p.nplanes; plane += gridDim.x) {
        const int64_t b = plane / cpb, d0 = (plane % cpb) * CH;
        // ---- per-plane tables: constants {A[n] * s, bias * s', D}, carry slots, visitor flags
        for (int i = tid; i < 4 * CH; i += Cfg::NCW * 32) {
            const int qq = i / CH, r = i % CH;
            const int64_t row = (int64_t)qq * p.D + d0 + r;
            float* k = reinterpret_cast<float*>(rc) + i * Cfg::RCF;
#pragma unroll
            for (int n = 0; n < NST; ++n) k[n] = p.A[row * NST + n] * (FAST ? 1.0f : kLog2e);
            k[NST] = (p.bias ? p.bias[row] : 0.0f) * (FAST ? kLog2e : 1.0f);
            k[NST + 1] = p.Ds ? p.Ds[row] : 0.0f;
            reinterpret_cast<uint4*>(carry)[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int i = tid; i < 2 * CH * nblk; i += Cfg::NCW * 32) reinterpret_cast<uint32_t*>(vis)[i] = 0u;
        named_barrier(1, Cfg::NCW * 32);
        // (two instantiations only: the kernel's working set of instructions must stay small -- with one instantiation per
        //  direction the four chains of an SM thrashed the instruction cache and the kernel ran 1.5x slower)
        if (q & 1) core_chain<NST, IN_T, C, WPC, SOFTPLUS, true>(p, my_ring, my_bars, s_rc, s_carry, s_vis, s_yacc, yacc_stride, q, wi, lane, mycnt, next);
        else core_chain<NST, IN_T, C, WPC, SOFTPLUS, false>(p, my_ring, my_bars, s_rc, s_carry, s_vis, s_yacc, yacc_stride, q, wi, lane, mycnt, next);
        named_barrier(1, Cfg::NCW * 32);
        // ---- epilogue: out[h][w] = Y0[h*W + w] + Y1[w*pitch + h]; a warp takes (channel, 4 rows) strips, lanes along w
        const int hq = H >> 2;
        const uint32_t s_y0 = smem_u32(y0), s_y1 = smem_u32(y1);
        for (int strip = warp; strip < CH * hq; strip += Cfg::NCW) {
            const int r = strip / hq, rv = strip - r * hq;
            const uint32_t y0r = s_y0 + (uint32_t)r * y0b, y1r = s_y1 + (uint32_t)r * y1b;
            float* orow = p.out + ((b * p.D + d0 + r) * (int64_t)L) + (int64_t)(4 * rv) * W;
            const uint32_t t0 = (uint32_t)(4 * rv * W);
            for (int w = lane; w < W; w += 32) {
                const uint4 c = lds128(y1r + acc_swz((uint32_t)(w * pitch + 4 * rv)) * 4u);
                const float a0 = lds32f(y0r + acc_swz(t0 + w) * 4u);
                const float a1 = lds32f(y0r + acc_swz(t0 + W + w) * 4u);
                const float a2 = lds32f(y0r + acc_swz(t0 + 2 * W + w) * 4u);
                const float a3 = lds32f(y0r + acc_swz(t0 + 3 * W + w) * 4u);
                orow[w] = a0 + __uint_as_float(c.x);
                orow[W + w] = a1 + __uint_as_float(c.y);
                orow[2 * W + w] = a2 + __uint_as_float(c.z);
                orow[3 * W + w] = a3 + __uint_as_float(c.w);
            }
        }
        // the next plane's tables and accumulators are written after every warp has left the epilogue
        named_barrier(1, Cfg::NCW * 32);
    }
}

// ------------------------------------------------------------------------------------------------ host
static int core_y1_pitch(int H) { return H + ((H % 8 == 0) ? 4 : 8); }

// channels per CTA (largest divisor of D up to 12 that fits with one ring slot per warp) and ring depth (as deep as fits)
template <int NST, typename IN_T, int C, int WPC>
static int core_pick_ch(int64_t D, int64_t L, int H, int W, int smem_limit, int* spw_out = nullptr) {
    using Cfg = CoreCfg<NST, IN_T, C, WPC>;
    const int nblk = (int)ceil_div(L, Cfg::TOK), pitch = core_y1_pitch(H);
    int best = 0;
    for (int ch = 1; ch <= 12 && ch <= D; ++ch)
        if (D % ch == 0 && Cfg::smem(ch, (int)L, W, pitch, nblk, 1) <= smem_limit) best = ch;
    if (best > 0 && spw_out) {
        int spw = 1;
        while (spw < SC_MAX_SPW && Cfg::smem(best, (int)L, W, pitch, nblk, spw + 1) <= smem_limit) ++spw;
        *spw_out = spw;
    }
    return best;
}

template <int NST, typename IN_T, int C, int WPC>
static int core_launch(CoreParams p, cudaStream_t st) {
    using Cfg = CoreCfg<NST, IN_T, C, WPC>;
    int dev = 0, smem_limit = 0;
    XP_CUDA_OK(cudaGetDevice(&dev));
    XP_CUDA_OK(cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    p.nblk = (int)ceil_div(p.L, Cfg::TOK);
    p.y1_pitch = core_y1_pitch(p.H);
    p.CH = core_pick_ch<NST, IN_T, C, WPC>(p.D, p.L, p.H, p.W, smem_limit, &p.spw);
    XP_REQUIRE(p.CH > 0, "xp_ss2d_core: a %d x %d token plane does not fit in shared memory", p.H, p.W);
    p.nplanes = p.batch * (p.D / p.CH);
    const int smem = Cfg::smem(p.CH, (int)p.L, p.W, p.y1_pitch, p.nblk, p.spw);
    const unsigned grid = (unsigned)min((int64_t)num_sms(), p.nplanes);
    if (p.softplus) {
        auto kern = ss2d_core_kernel<NST, IN_T, C, WPC, true>;
        XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, Cfg::NCW * 32, smem, st>>>(p);
    } else {
        auto kern = ss2d_core_kernel<NST, IN_T, C, WPC, false>;
        XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, Cfg::NCW * 32, smem, st>>>(p);
    }
    XP_LAUNCH_CHECK("ss2d_core_kernel");
    return XP_OK;
}

// (8 tokens per lane and 6 warps per chain -- 24 warps at 80 registers -- was measured and is slower: 3.15 vs 2.50 ms at stage 0)
template <typename IN_T> static int core_dispatch(const CoreParams& p, int64_t N, cudaStream_t st) {
    if (N == 1) {
        if constexpr (sizeof(IN_T) == 2) return core_launch<1, IN_T, 16, 3>(p, st);
        else return core_launch<1, IN_T, 8, 3>(p, st);
    }
    return core_launch<2, IN_T, 8, 3>(p, st);
}

template <typename IN_T> static int core_channels(int64_t N, int64_t D, int64_t L, int H, int W, int smem_limit) {
    if (N == 1) {
        if constexpr (sizeof(IN_T) == 2) return core_pick_ch<1, IN_T, 16, 3>(D, L, H, W, smem_limit);
        else return core_pick_ch<1, IN_T, 8, 3>(D, L, H, W, smem_limit);
    }
    return core_pick_ch<2, IN_T, 8, 3>(D, L, H, W, smem_limit);
}

}  // namespace xp

using namespace xp;

// Channels per CTA the fused core would use for this shape (0 = the planes do not fit: use xp_selective_scan_fwd +
// xp_ss2d_merge_norm instead).  No launch.
extern "C" int32_t xp_ss2d_core_channels(int64_t d_inner, int64_t dstate, int64_t H, int64_t W, int32_t in_dtype) {
    if (d_inner <= 0 || H <= 0 || W <= 0 || H % 4 || W % 4 || dstate < 1 || dstate > 2 || in_dtype < XP_F32 || in_dtype > XP_BF16) return 0;
    if (H * W >= (1 << 22)) return 0;
    int dev = 0, smem_limit = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    if (in_dtype == XP_F32) return core_channels<float>(dstate, d_inner, H * W, (int)H, (int)W, smem_limit);
    if (in_dtype == XP_F16) return core_channels<__half>(dstate, d_inner, H * W, (int)H, (int)W, smem_limit);
    return core_channels<__nv_bfloat16>(dstate, d_inner, H * W, (int)H, (int)W, smem_limit);
}

extern "C" int xp_ss2d_core(const xp_ss2d_core_args* a, xp_stream_t stream) {
    XP_REQUIRE(a != nullptr, "xp_ss2d_core: args is NULL");
    XP_REQUIRE(a->xx && a->delta && a->B && a->C && a->A && a->out, "xp_ss2d_core: xx/delta/B/C/A/out must be non-NULL");
    XP_REQUIRE(a->batch >= 0 && a->d_inner > 0 && a->H > 0 && a->W > 0, "xp_ss2d_core: bad shape");
    XP_REQUIRE(a->dstate == 1 || a->dstate == 2, "xp_ss2d_core: d_state must be 1 or 2 (got %lld)", (long long)a->dstate);
    XP_REQUIRE(a->H % 4 == 0 && a->W % 4 == 0, "xp_ss2d_core: H and W must be multiples of 4 (got %lld x %lld)", (long long)a->H,
               (long long)a->W);
    XP_REQUIRE(a->H * a->W < (1 << 22), "xp_ss2d_core: at most 2^22 tokens per plane");
    XP_REQUIRE(a->in_dtype >= XP_F32 && a->in_dtype <= XP_BF16, "xp_ss2d_core: unsupported dtype %d", a->in_dtype);
    const int64_t L = a->H * a->W, va = a->in_dtype == XP_F32 ? 4 : 8;
    XP_REQUIRE(L % va == 0, "xp_ss2d_core: H*W must be a multiple of %lld for this dtype", (long long)va);
    const int64_t strides[] = {a->B_batch_stride, a->B_group_stride, a->B_state_stride, a->C_batch_stride, a->C_group_stride,
                               a->C_state_stride};
    for (int64_t s : strides) XP_REQUIRE(s % va == 0, "xp_ss2d_core: B/C strides must keep rows 16-byte aligned");
    const void* ptrs[] = {a->xx, a->delta, a->B, a->C, a->out};
    for (const void* q : ptrs) XP_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "xp_ss2d_core: tensors must be 16-byte aligned");
    if (a->batch == 0) return XP_OK;
    CoreParams p;
    p.xx = a->xx; p.delta = a->delta; p.Bm = a->B; p.Cm = a->C; p.A = a->A; p.Ds = a->D; p.bias = a->delta_bias; p.out = (float*)a->out;
    p.batch = a->batch; p.D = a->d_inner; p.L = L; p.H = (int)a->H; p.W = (int)a->W;
    p.B_bs = a->B_batch_stride; p.B_gs = a->B_group_stride; p.B_ss = a->B_state_stride;
    p.C_bs = a->C_batch_stride; p.C_gs = a->C_group_stride; p.C_ss = a->C_state_stride;
    p.softplus = a->delta_softplus; p.CH = 0; p.nblk = 0; p.y1_pitch = 0; p.nplanes = 0; p.spw = 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->in_dtype) {
        case XP_F32: return core_dispatch<float>(p, a->dstate, st);
        case XP_F16: return core_dispatch<__half>(p, a->dstate, st);
        default: return core_dispatch<__nv_bfloat16>(p, a->dstate, st);
    }
}
