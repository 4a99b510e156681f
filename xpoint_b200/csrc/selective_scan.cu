// selective_scan.cu -- selective-scan forward for sm_100a (xp_selective_scan_fwd).
//
// Replaces selective_scan_cuda_oflex.fwd of the reference
// (kernels/selective_scan/csrc/selective_scan/cusoflex/selective_scan_fwd_kernel_oflex.cuh:67-212), which
// runs one CTA per (batch, channel) row and, per 2048-token chunk, one CUB block-scan per state.
// This file is a different decomposition with three kernels, chosen by shape on the host:
//
//  * scan_lanes_kernel   (dstate 1|2, the shipped XPoint config; HBM-bound)
//      a consumer warp owns up to 32 channel rows of one (batch, group) and walks the sequence in steps of 256 or 512
//      tokens; a producer warp streams the rows' u / delta chunks (1-D bulk TMA copies) into per-warp shared-memory
//      rings, B/C once per step.  Every lane scans 8 or 16 consecutive tokens sequentially in registers and the 32
//      lane chunks are combined with a warp-shuffle prefix scan over the affine maps (a, b) -> a*h + b.
//      With xp_scan_args.dt_weight (fused dt_proj) delta is not loaded but formed per step from the low-rank dts_r rows
//      by a tensor-core micro-GEMM (mma.sync) inside the consumer warp.
//  * scan_rows_kernel    (dstate 4|8|16; MUFU/FP32-issue-bound)
//      a channel row is owned by two adjacent lanes (states n % 2) that run the recurrences sequentially (one ex2 and
//      four FP32 ops per (token, state) -- no redundant scan work).  Each warp runs a private TMA pipeline:
//      u/delta (and z) tiles [16 rows x 128 B] land in 128B-swizzled shared memory, the CTA shares one ring of B/C tiles
//      [dstate x 128 B] that are read as broadcasts, y leaves through 32-byte stores per row and step.
//  * scan_generic_kernel  (any dstate <= 256, any alignment / strides, delta groups, fused dt_proj of any rank)
//      warp per row, lanes along the sequence, one state at a time, scalar loads.
//
// All three keep fp32 state and accumulation (reference: csms6s.py:52-68).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace xp {

struct ScanParams {
    const void* u; const void* delta; const float* A; const void* Bm; const void* Cm;
    const float* D; const void* z; const float* bias; void* out; float* last;
    int64_t batch, dim, delta_dim, groups, dstate, L;
    int64_t u_bs, u_ds, dl_bs, dl_ds, B_bs, B_gs, B_ss, C_bs, C_gs, C_ss, z_bs, z_ds, o_bs, o_ds;
    int64_t u_gs, u_gdiv;   // u row of (b, g, dg) = u + b*u_bs + (g / u_gdiv)*u_gs + dg*u_ds
    uint64_t rev_mask;      // bit g: group g walks memory backwards (all tensors, including out)
    const void* wdt;        // fused dt_proj: (dim, R) weights in IN_T; delta then points at dts_r (batch, groups, R, L)
    int64_t R, dl_gs;       //   rank (0 = delta is materialised), group stride of dts_r (dl_ds = its rank-row stride)
    int softplus;
    int rows_per_warp;   // scan_lanes only
};

__device__ __forceinline__ float delta_act(float d, float bias, int softplus) {
    d += bias;
    return softplus ? softplus_f(d) : d;
}

// delta of one token from the low-rank factors (fused dt_proj): sum_r w[r] * x[r*xs + l], fp32 accumulation, rounded to
// the input dtype like the output of the reference's dt_proj GEMM (VMamba.py:607-608 under autocast)
template <typename IN_T>
__device__ __forceinline__ float delta_lowrank(const IN_T* x, int64_t xs, const IN_T* w, int R, int64_t l) {
    float acc = 0.0f;
    for (int r = 0; r < R; ++r) acc = fmaf(to_f32(w[r]), to_f32(x[r * xs + l]), acc);
    return to_f32(from_f32<IN_T>(acc));
}

// ============================================================================================
// generic kernel
// ============================================================================================
constexpr int GEN_WARPS = 4;
constexpr int GEN_C = 4;  // tokens per lane per step

template <typename IN_T, typename OUT_T>
__global__ void __launch_bounds__(GEN_WARPS * 32) scan_generic_kernel(const ScanParams p) {
    __shared__ float carry_s[GEN_WARPS][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * GEN_WARPS + warp;
    if (row >= p.batch * p.dim) return;
    const int64_t b = row / p.dim, d = row % p.dim;
    const int64_t g = d / (p.dim / p.groups);
    const int64_t dd = d / (p.dim / p.delta_dim);
    const int N = (int)p.dstate;
    const int64_t Dg = p.dim / p.groups;
    const bool rev = g < 64 && ((p.rev_mask >> g) & 1);
    const IN_T* u = (const IN_T*)p.u + b * p.u_bs + (g / p.u_gdiv) * p.u_gs + (d - g * Dg) * p.u_ds;
    const int R = (int)p.R;
    const IN_T* dl = R ? (const IN_T*)p.delta + b * p.dl_bs + g * p.dl_gs : (const IN_T*)p.delta + b * p.dl_bs + dd * p.dl_ds;
    const IN_T* wdt = R ? (const IN_T*)p.wdt + d * R : nullptr;
    const IN_T* Bm = (const IN_T*)p.Bm + b * p.B_bs + g * p.B_gs;
    const IN_T* Cm = (const IN_T*)p.Cm + b * p.C_bs + g * p.C_gs;
    const IN_T* z = p.z ? (const IN_T*)p.z + b * p.z_bs + d * p.z_ds : nullptr;
    OUT_T* out = (OUT_T*)p.out + b * p.o_bs + d * p.o_ds;
    const float* A = p.A + d * N;
    const float bias = p.bias ? p.bias[dd] : 0.0f;
    const float Dv = p.D ? p.D[d] : 0.0f;
    float* carry = carry_s[warp];
    for (int n = lane; n < N; n += 32) carry[n] = 0.0f;
    __syncwarp();

    for (int64_t t0 = 0; t0 < p.L; t0 += 32 * GEN_C) {
        float uv[GEN_C], dv[GEN_C], y[GEN_C];
        const int64_t l0 = t0 + lane * GEN_C;
        auto mem = [&](int64_t l) { return rev ? p.L - 1 - l : l; };   // scan position -> memory token
#pragma unroll
        for (int j = 0; j < GEN_C; ++j) {
            const bool ok = l0 + j < p.L;
            uv[j] = ok ? to_f32(u[mem(l0 + j)]) : 0.0f;
            const float draw = !ok ? 0.0f : R ? delta_lowrank(dl, p.dl_ds, wdt, R, mem(l0 + j)) : to_f32(dl[mem(l0 + j)]);
            dv[j] = ok ? delta_act(draw, bias, p.softplus) : 0.0f;                       // 0 -> identity step
            y[j] = Dv * uv[j];
        }
        for (int n = 0; n < N; ++n) {
            const float An = A[n];
            float hl[GEN_C], pl[GEN_C];
            float P = 1.0f, S = 0.0f;
#pragma unroll
            for (int j = 0; j < GEN_C; ++j) {
                const bool ok = l0 + j < p.L;
                const float a = __expf(dv[j] * An);
                const float bu = ok ? (dv[j] * to_f32(Bm[n * p.B_ss + mem(l0 + j)])) * uv[j] : 0.0f;
                S = fmaf(a, S, bu);
                P *= a;
                hl[j] = S; pl[j] = P;
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float Pp = __shfl_up_sync(0xffffffffu, P, o), Sp = __shfl_up_sync(0xffffffffu, S, o);
                if (lane >= o) { S = fmaf(P, Sp, S); P *= Pp; }
            }
            float Pe = __shfl_up_sync(0xffffffffu, P, 1), Se = __shfl_up_sync(0xffffffffu, S, 1);
            if (lane == 0) { Pe = 1.0f; Se = 0.0f; }
            const float hc = carry[n];
            const float hin = fmaf(Pe, hc, Se);
#pragma unroll
            for (int j = 0; j < GEN_C; ++j) {
                const bool ok = l0 + j < p.L;
                const float h = fmaf(pl[j], hin, hl[j]);
                if (ok) y[j] = fmaf(h, to_f32(Cm[n * p.C_ss + mem(l0 + j)]), y[j]);
            }
            __syncwarp();
            if (lane == 31) carry[n] = fmaf(P, hc, S);
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < GEN_C; ++j)
            if (l0 + j < p.L) {
                float v = y[j];
                if (z) v *= silu_f(to_f32(z[mem(l0 + j)]));
                out[mem(l0 + j)] = from_f32<OUT_T>(v);
            }
    }
    if (p.last)
        for (int n = lane; n < N; n += 32) p.last[row * N + n] = carry[n];
}

// ============================================================================================
// lanes kernel (dstate 1 | 2)
// ============================================================================================
// CTA = LN_CONSUMERS consumer warps + 1 producer warp; every consumer warp runs a private pipeline over its own
// (row, step) items, there is no CTA-wide synchronisation after start-up.
//   * producer warp: lane w feeds consumer warp w.  Per item it waits (non-blocking poll) for the ring slot's
//     "empty" mbarrier and streams u/delta[/z] of one channel row and TOK = 32*C consecutive tokens with 1-D bulk
//     async copies (cp.async.bulk -> UBLKCP, the TMA engine) into the consumer's shared-memory ring; B/C chunks go
//     into a separate double buffer once per step and are shared by the RW rows of the warp.  The consumers spend
//     no issue slots on address arithmetic and no registers / scoreboard slots on loads in flight.
//   * consumer warp: lanes read C tokens each (LDS.128), scan them sequentially in registers, and the 32 lane
//     chunks are combined with a warp-shuffle prefix scan of the affine maps h -> P*h + S; y leaves through
//     128-bit stores.  Per-row constants and the running state live in a small per-warp shared-memory table.
//   * full steps (all 32*C tokens valid) run a predicate-free body; only the last step of a row is masked.
//   * reversed groups (xp_scan_args.reverse_group_mask) walk memory backwards: the producer fetches the mirrored
//     window, lanes take mirrored chunks and reverse the element order in registers -- no extra instructions.
//   * 16-bit inputs use softplus(x) = ln2 * lg2(1 + 2^(x*log2e)) with ln2 / log2e folded into A, bias and B
//     (4 FP32 + 2 MUFU per token instead of ~14 instructions); its error (~2e-7 absolute on delta') is far
//     below the 16-bit input quantisation.  fp32 inputs keep the series-corrected softplus_f.
constexpr int LN_STAGES = 4;     // u/delta ring depth per consumer warp
constexpr int LN_BCS = 2;        // B/C buffer depth per consumer warp
constexpr int LN_MAXROWS = 32;   // max channel rows per warp (row table size)
constexpr int LN_DT_ROWS = 8;    // rows per warp with the fused dt_proj (height of its delta tile)

template <typename IN_T> __device__ __forceinline__ void widen16(const uint4& v, float (&f)[16 / sizeof(IN_T)]);
template <> __device__ __forceinline__ void widen16<float>(const uint4& v, float (&f)[4]) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void widen16<__half>(const uint4& v, float (&f)[8]) { VecIO<__half, 8>::widen(v, f); }
template <> __device__ __forceinline__ void widen16<__nv_bfloat16>(const uint4& v, float (&f)[8]) {
    VecIO<__nv_bfloat16, 8>::widen(v, f);
}

// C consecutive elements from shared memory, widened to fp32; REV hands them back in reversed order
template <typename IN_T, int C, bool REV> __device__ __forceinline__ void lds_tokens(uint32_t saddr, float (&f)[C]) {
    constexpr int PER = 16 / (int)sizeof(IN_T);
    static_assert(C % PER == 0, "lane chunk must be whole 16-byte vectors");
#pragma unroll
    for (int v = 0; v < C / PER; ++v) {
        float w[PER];
        widen16<IN_T>(lds128(saddr + 16 * v), w);
#pragma unroll
        for (int i = 0; i < PER; ++i) f[REV ? C - 1 - (v * PER + i) : v * PER + i] = w[i];
    }
}

template <typename OUT_T> __device__ __forceinline__ void stg16(OUT_T* p, const float* v);   // one 16-byte store
template <> __device__ __forceinline__ void stg16<float>(float* p, const float* v) {
    const float a[4] = {v[0], v[1], v[2], v[3]};
    VecIO<float, 4>::store(p, a);
}
template <> __device__ __forceinline__ void stg16<__half>(__half* p, const float* v) {
    const float a[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
    VecIO<__half, 8>::store(p, a);
}
template <> __device__ __forceinline__ void stg16<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
    const float a[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
    VecIO<__nv_bfloat16, 8>::store(p, a);
}

// y[0..C) in scan order -> C consecutive memory elements starting at gp (reversed when REV).  nv = number of valid
// scan-order tokens of this lane (C in full steps); vectors are all-or-nothing (host guarantees L % vector == 0).
template <typename OUT_T, int C, bool REV, bool FULL>
__device__ __forceinline__ void store_tokens(OUT_T* gp, const float (&y)[C], int nv) {
    constexpr int PER = 16 / (int)sizeof(OUT_T);
    float m[C];
#pragma unroll
    for (int j = 0; j < C; ++j) m[REV ? C - 1 - j : j] = y[j];
#pragma unroll
    for (int v = 0; v < C / PER; ++v) {
        const int jmin = REV ? C - (v + 1) * PER : v * PER;   // lowest scan index inside memory vector v
        if (FULL || jmin < nv) stg16<OUT_T>(gp + v * PER, &m[v * PER]);
    }
}

// Fused dt_proj (DTF, 16-bit inputs only): delta = W_dt x dts_r is a rank-R micro-GEMM.  Per step a consumer warp
// multiplies its 8 rows' weights (A operand, constant registers) with the step's dts_r rows (B operand, ldmatrix.trans
// straight from the TMA-filled tile) on the tensor cores (mma.sync m16n8k8 / m16n8k16, fp32 accumulate -- rows 8..15 of
// the M=16 tile are zero padding), rounds the result to the input dtype exactly where the reference's dt_proj GEMM does
// (F.conv1d / einsum under autocast, VMamba.py:607-608) and parks it in a per-warp [8 rows][TOK] tile that the row loop
// reads like a materialised delta chunk.  Row stride = CHUNK + 16 bytes: conflict-free for ldmatrix, the packed
// stores and the row loop's LDS.128.
template <int NST, typename IN_T, int C, bool HAS_Z, int NW, bool DTF> struct LanesCfg {
    static constexpr int TOK = 32 * C;                                   // tokens per step
    static constexpr int CHUNK = TOK * (int)sizeof(IN_T);                // bytes of one row chunk
    static constexpr int ITEM = CHUNK * ((HAS_Z ? 2 : 1) + (DTF ? 0 : 1));   // u [, delta] [, z]
    static constexpr int ZOFF = CHUNK * (DTF ? 1 : 2);
    static constexpr int BC = 2 * NST * CHUNK;                           // B rows then C rows
    static constexpr int RCF = NST == 1 ? 4 : 8;                         // floats of per-row constants
    static constexpr int RC_BYTES = LN_MAXROWS * RCF * 4;
    static constexpr int DROW = CHUNK + 16;                              // padded row stride of the dts_r / delta tiles
    static constexpr int DTILE = LN_DT_ROWS * DROW;                      // delta tile: one row per channel row of the warp
    static constexpr int NBARS = 2 * LN_STAGES + 2 * LN_BCS + (DTF ? 2 : 0);   // full/empty, bcfull/bcempty [, dfull/dempty]
    // dts_r tile: the rank padded to the MMA's K (8 or 16); R = 0 unless DTF, everything folds to constants then
    __host__ __device__ static constexpr int rpad(int R) { return R <= 8 ? 8 : 16; }
    __host__ __device__ static constexpr int dts_bytes(int R) { return DTF ? rpad(R) * DROW : 0; }
    __host__ __device__ static constexpr int warp_bytes(int R) {
        return LN_STAGES * ITEM + LN_BCS * BC + dts_bytes(R) + (DTF ? DTILE : 0) + RC_BYTES;
    }
    __host__ __device__ static constexpr int smem(int R) { return NW * (warp_bytes(R) + NBARS * 8) + 16; }
};

struct LanesWarp {   // decoded work assignment of one consumer warp
    int64_t b, g, d0;
    int nrows;
    bool valid;
};
__device__ __forceinline__ LanesWarp lanes_decode(const ScanParams& p, int64_t wid) {
    LanesWarp w;
    const int RW = p.rows_per_warp;
    const int64_t Dg = p.dim / p.groups;
    const int64_t rbpg = (Dg + RW - 1) / RW;
    w.valid = wid < p.batch * p.groups * rbpg;
    const int64_t rb = wid % rbpg;
    w.g = (wid / rbpg) % p.groups;
    w.b = wid / (rbpg * p.groups);
    w.d0 = w.g * Dg + rb * RW;
    w.nrows = w.valid ? (int)min((int64_t)RW, Dg - rb * RW) : 0;
    return w;
}

template <int NST, typename IN_T, int C, bool HAS_Z, int NW, bool DTF>
__device__ __forceinline__ void lanes_producer(const ScanParams& p, uint8_t* base, int lane) {
    using Cfg = LanesCfg<NST, IN_T, C, HAS_Z, NW, DTF>;
    constexpr int TOK = Cfg::TOK;
    constexpr int ES = (int)sizeof(IN_T);
    const int R = DTF ? (int)p.R : 0;
    const int WB = Cfg::warp_bytes(R);
    const bool mine = lane < NW;
    const LanesWarp w = lanes_decode(p, (int64_t)blockIdx.x * NW + (mine ? lane : 0));
    uint8_t* ring = base + (mine ? lane : 0) * WB;
    uint8_t* bcbuf = ring + LN_STAGES * Cfg::ITEM;
    uint8_t* dts = bcbuf + LN_BCS * Cfg::BC;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + NW * WB) + (mine ? lane : 0) * Cfg::NBARS;
    uint64_t* empty = full + LN_STAGES;
    uint64_t* bcfull = empty + LN_STAGES;
    uint64_t* bcempty = bcfull + LN_BCS;
    uint64_t* dfull = bcempty + LN_BCS;      // DTF only
    uint64_t* dempty = dfull + 1;
    const int64_t Dg = p.dim / p.groups;
    const int64_t dg0 = w.d0 - w.g * Dg;
    const bool rev = (p.rev_mask >> (w.g & 63)) & 1 && w.g < 64;
    const IN_T* ub = (const IN_T*)p.u + w.b * p.u_bs + (w.g / p.u_gdiv) * p.u_gs + dg0 * p.u_ds;
    // materialised delta: the warp's first row; fused dt_proj: the R rank rows of the warp's (batch, group)
    const IN_T* db = DTF ? (const IN_T*)p.delta + w.b * p.dl_bs + w.g * p.dl_gs : (const IN_T*)p.delta + w.b * p.dl_bs + w.d0 * p.dl_ds;
    const IN_T* zb = HAS_Z ? (const IN_T*)p.z + w.b * p.z_bs + w.d0 * p.z_ds : nullptr;
    const IN_T* Bb = (const IN_T*)p.Bm + w.b * p.B_bs + w.g * p.B_gs;
    const IN_T* Cb = (const IN_T*)p.Cm + w.b * p.C_bs + w.g * p.C_gs;
    const int nsteps = (int)((p.L + TOK - 1) / TOK);
    const int total = (mine && w.valid) ? nsteps * w.nrows : 0;
    int it = 0, r = 0, step = 0, slot = 0;
    uint32_t fill = 0;          // how many times the ring has wrapped
    while (true) {
        const bool active = it < total;
        if (!__any_sync(0xffffffffu, active)) break;
        bool issued = false;
        if (active) {
            bool ok = mbar_test_wait(&empty[slot], (fill - 1u) & 1u);          // fill 0: parity 1 passes on a fresh barrier
            const int bs = step % LN_BCS;
            if (ok && r == 0) ok = mbar_test_wait(&bcempty[bs], ((uint32_t)(step / LN_BCS) - 1u) & 1u);
            if (DTF && ok && r == 0) ok = mbar_test_wait(dempty, ((uint32_t)step - 1u) & 1u);
            if (ok) {
                const int64_t t0 = (int64_t)step * TOK;
                const int64_t valid = min((int64_t)TOK, p.L - t0);
                const int64_t m0 = rev ? p.L - t0 - valid : t0;                 // first memory token of the window
                const uint32_t off = rev ? (uint32_t)((TOK - valid) * ES) : 0u; // mirrored windows are right-aligned
                const uint32_t bytes = (uint32_t)(valid * ES);
                uint8_t* dst = ring + slot * Cfg::ITEM + off;
                mbar_arrive_expect_tx(&full[slot], bytes * (Cfg::ITEM / Cfg::CHUNK));
                bulk_load(dst, ub + r * p.u_ds + m0, bytes, &full[slot]);
                if (!DTF) bulk_load(dst + Cfg::CHUNK, db + r * p.dl_ds + m0, bytes, &full[slot]);
                if (HAS_Z) bulk_load(dst + Cfg::ZOFF, zb + r * p.z_ds + m0, bytes, &full[slot]);
                if (r == 0) {
                    uint8_t* bc = bcbuf + bs * Cfg::BC + off;
                    mbar_arrive_expect_tx(&bcfull[bs], bytes * 2 * NST);
#pragma unroll
                    for (int n = 0; n < NST; ++n) {
                        bulk_load(bc + n * Cfg::CHUNK, Bb + n * p.B_ss + m0, bytes, &bcfull[bs]);
                        bulk_load(bc + (NST + n) * Cfg::CHUNK, Cb + n * p.C_ss + m0, bytes, &bcfull[bs]);
                    }
                    if (DTF) {                                                  // the step's dts_r rows (single buffer)
                        mbar_arrive_expect_tx(dfull, bytes * R);
#pragma unroll 1
                        for (int q = 0; q < R; ++q) bulk_load(dts + q * Cfg::DROW + off, db + q * p.dl_ds + m0, bytes, dfull);
                    }
                }
                if (++r == w.nrows) { r = 0; ++step; }
                if (++slot == LN_STAGES) { slot = 0; ++fill; }
                ++it;
                issued = true;
            }
        }
        if (!__any_sync(0xffffffffu, issued)) __nanosleep(40);
    }
}

// ---- tensor-core pieces of the fused dt_proj ----
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// D(16x8, fp32) = A(16xK) * B(Kx8); only rows 0..7 of A are populated (a1 / a3 = rows 8..15 = 0), so d[2], d[3] stay 0
template <typename IN_T> __device__ __forceinline__ void mma_k8(float (&d)[4], uint32_t a0, uint32_t b0);
template <> __device__ __forceinline__ void mma_k8<__half>(float (&d)[4], uint32_t a0, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a0), "r"(0u), "r"(b0), "f"(0.0f));
}
template <> __device__ __forceinline__ void mma_k8<__nv_bfloat16>(float (&d)[4], uint32_t a0, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a0), "r"(0u), "r"(b0), "f"(0.0f));
}
template <typename IN_T> __device__ __forceinline__ void mma_k16(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma_k16<__half>(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1), "f"(0.0f));
}
template <> __device__ __forceinline__ void mma_k16<__nv_bfloat16>(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1), "f"(0.0f));
}
template <typename IN_T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<float>(float, float) { return 0u; }   // never used (DTF is 16-bit only)
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}

// delta tile [8 rows][TOK] (row stride DROW) = W (A fragments a_lo, a_hi: this lane's row lane/4, ranks 2*(lane%4)+{0,1}
// and +8) x dts_r tile [rpad rows][TOK]
template <typename IN_T, int TOK, int DROW>
__device__ __forceinline__ void dt_tile_mma(uint32_t dts, uint32_t dtile, uint32_t a_lo, uint32_t a_hi, bool k16, int lane) {
    const int g = lane >> 2, t = lane & 3;
    const uint32_t out = dtile + g * DROW + t * 4;           // tokens 2t, 2t+1 of row g
    const int mi = lane >> 3, kr = lane & 7;                 // ldmatrix: lane supplies row kr of matrix mi
    if (!k16) {
        // four 8x8 matrices = ranks 0..7 x tokens tb + 8*mi
        const uint32_t src = dts + kr * DROW + mi * 16;
#pragma unroll 4
        for (int tb = 0; tb < TOK; tb += 32) {
            uint32_t b[4];
            ldmatrix_x4_trans(src + tb * 2, b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float d[4];
                mma_k8<IN_T>(d, a_lo, b[i]);
                sts_u32(out + (tb + 8 * i) * 2, pack2<IN_T>(d[0], d[1]));
            }
        }
    } else {
        // matrices: (ranks 0..7, tokens tb..), (ranks 8..15, tokens tb..), (ranks 0..7, tokens tb+8..), (ranks 8..15, tokens tb+8..)
        const uint32_t src = dts + (kr + 8 * (mi & 1)) * DROW + (mi >> 1) * 16;
#pragma unroll 4
        for (int tb = 0; tb < TOK; tb += 16) {
            uint32_t b[4];
            ldmatrix_x4_trans(src + tb * 2, b);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float d[4];
                mma_k16<IN_T>(d, a_lo, a_hi, b[2 * i], b[2 * i + 1]);
                sts_u32(out + (tb + 8 * i) * 2, pack2<IN_T>(d[0], d[1]));
            }
        }
    }
}

template <int NST, typename IN_T, typename OUT_T, int C, bool HAS_Z, bool SOFTPLUS, bool REV, int NW, bool DTF>
__device__ __forceinline__ void lanes_consumer(const ScanParams& p, uint8_t* base, const LanesWarp& w, int warp, int lane) {
    using Cfg = LanesCfg<NST, IN_T, C, HAS_Z, NW, DTF>;
    constexpr int TOK = Cfg::TOK, RCF = Cfg::RCF;
    constexpr bool FAST = SOFTPLUS && sizeof(IN_T) == 2;
    const int R = DTF ? (int)p.R : 0;
    const int WB = Cfg::warp_bytes(R);
    // everything below addresses shared memory through 32-bit shared-window addresses (LDS/STS, not generic LD/ST)
    const uint32_t ring = smem_u32(base) + warp * WB;
    const uint32_t bcbuf = ring + LN_STAGES * Cfg::ITEM;
    const uint32_t dts = bcbuf + LN_BCS * Cfg::BC;
    const uint32_t dtile = dts + Cfg::dts_bytes(R);
    const uint32_t rcs = dtile + (DTF ? Cfg::DTILE : 0);
    float* rc = reinterpret_cast<float*>(base + warp * WB + (rcs - ring));
    const uint32_t full = smem_u32(base) + NW * WB + warp * Cfg::NBARS * 8;
    const uint32_t empty = full + LN_STAGES * 8;
    const uint32_t bcfull = empty + LN_STAGES * 8;
    const uint32_t bcempty = bcfull + LN_BCS * 8;
    const uint32_t dfull = bcempty + LN_BCS * 8;      // DTF only
    const uint32_t dempty = dfull + 8;

    // per-row constants {A[n]*s, bias*s', D, carry[n]}  (s = log2e, s' = 1 unless FAST: s = 1, s' = log2e)
    if (lane < w.nrows) {
        float* k = rc + lane * RCF;
#pragma unroll
        for (int n = 0; n < NST; ++n) {
            k[n] = p.A[(w.d0 + lane) * NST + n] * (FAST ? 1.0f : kLog2e);
            k[NST + 2 + n] = 0.0f;
        }
        k[NST] = (p.bias ? p.bias[w.d0 + lane] : 0.0f) * (FAST ? kLog2e : 1.0f);
        k[NST + 1] = p.D ? p.D[w.d0 + lane] : 0.0f;
    }
    uint32_t a_lo = 0, a_hi = 0;      // fused dt_proj: this lane's A fragments (row lane/4, ranks 2*(lane%4)+{0,1} [+8])
    if constexpr (DTF) {
        const int g = lane >> 2, k0 = 2 * (lane & 3);
        const unsigned short* ws = reinterpret_cast<const unsigned short*>(p.wdt) + (w.d0 + g) * R;
        auto wv = [&](int k) -> uint32_t { return (g < w.nrows && k < R) ? (uint32_t)ws[k] : 0u; };
        a_lo = wv(k0) | (wv(k0 + 1) << 16);
        a_hi = wv(k0 + 8) | (wv(k0 + 9) << 16);
        // rank rows R .. rpad-1 of the dts_r tile are never written by the producer: zero them once (0 x stale NaN = NaN)
        for (int q = R * Cfg::DROW + lane * 4; q < Cfg::dts_bytes(R); q += 128) sts_u32(dts + q, 0u);
    }
    __syncwarp();

    const int ci = REV ? 31 - lane : lane;
    const uint32_t lane_off = (uint32_t)(ci * C * (int)sizeof(IN_T));
    OUT_T* orow0 = (OUT_T*)p.out + w.b * p.o_bs + w.d0 * p.o_ds;
    const int nsteps = (int)((p.L + TOK - 1) / TOK);
    const int total = nsteps * w.nrows;

    float Bv[NST][C], Cv[NST][C];
    int r = 0, step = 0, slot = 0;
    uint32_t phase = 0;
    for (int it = 0; it < total; ++it) {
        const int64_t tok0 = (int64_t)step * TOK + lane * C;       // first scan-order token of this lane
        const bool full_step = (int64_t)(step + 1) * TOK <= p.L;
        const int nv = full_step ? C : (int)max((int64_t)0, min((int64_t)C, p.L - tok0));
        if (r == 0) {
            const int bs = step % LN_BCS;
            mbar_wait_s(bcfull + bs * 8, (uint32_t)(step / LN_BCS) & 1u);
            const uint32_t bc = bcbuf + bs * Cfg::BC + lane_off;
#pragma unroll
            for (int n = 0; n < NST; ++n) {
                lds_tokens<IN_T, C, REV>(bc + n * Cfg::CHUNK, Bv[n]);
                lds_tokens<IN_T, C, REV>(bc + (NST + n) * Cfg::CHUNK, Cv[n]);
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    if (FAST) Bv[n][j] *= kLn2;
                    if (!full_step && j >= nv) { Bv[n][j] = 0.0f; Cv[n][j] = 0.0f; }   // stale smem may hold NaN/Inf
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_s(bcempty + bs * 8);
            if constexpr (DTF) {
                // the step's delta tile for all rows of the warp (the previous step's tile was drained before the
                // __syncwarp of its last item)
                mbar_wait_s(dfull, (uint32_t)step & 1u);
                dt_tile_mma<IN_T, TOK, Cfg::DROW>(dts, dtile, a_lo, a_hi, R > 8, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive_s(dempty);
            }
        }
        mbar_wait_s(full + slot * 8, phase);
        const uint32_t item = ring + slot * Cfg::ITEM + lane_off;
        float uv[C], dv[C], zv[HAS_Z ? C : 1];
        lds_tokens<IN_T, C, REV>(item, uv);
        if constexpr (DTF) lds_tokens<IN_T, C, REV>(dtile + r * Cfg::DROW + lane_off, dv);
        else lds_tokens<IN_T, C, REV>(item + Cfg::CHUNK, dv);
        if constexpr (HAS_Z) lds_tokens<IN_T, C, REV>(item + Cfg::ZOFF, zv);
        float kc[RCF];
#pragma unroll
        for (int q = 0; q < RCF / 4; ++q) {
            const uint4 t = lds128(rcs + (r * RCF + 4 * q) * 4);
            kc[4 * q] = __uint_as_float(t.x); kc[4 * q + 1] = __uint_as_float(t.y);
            kc[4 * q + 2] = __uint_as_float(t.z); kc[4 * q + 3] = __uint_as_float(t.w);
        }
        __syncwarp();                                    // every lane has drained the slot (and read the row table)
        if (lane == 0) mbar_arrive_s(empty + slot * 8);  // hand it back to the producer
        if (++slot == LN_STAGES) { slot = 0; phase ^= 1u; }

        const float kb = kc[NST], kD = kc[NST + 1];
        auto body = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            float y[C], lw[C];
            if constexpr (FAST) {
                // 16-bit inputs: the per-token arithmetic that is independent between neighbouring tokens runs on packed
                // fp32 pairs (FFMA2 / FMUL2 / FADD2: two IEEE operations per issue slot, bit-identical to the scalar forms)
                const float2 kl2 = make_float2(kLog2e, kLog2e), kb2 = make_float2(kb, kb), one2 = make_float2(1.0f, 1.0f);
                const float2 kD2 = make_float2(kD, kD);
#pragma unroll
                for (int j = 0; j < C; j += 2) {
                    const float2 t2 = fma2(make_float2(dv[j], dv[j + 1]), kl2, kb2);      // log2e * (delta + bias)
                    const float2 s2 = add2(make_float2(ex2_approx(fminf(t2.x, 100.0f)), ex2_approx(fminf(t2.y, 100.0f))), one2);
                    float2 l2 = make_float2(fmaxf(lg2_approx(s2.x), t2.x), fmaxf(lg2_approx(s2.y), t2.y));   // log2e * softplus
                    float2 u2 = make_float2(uv[j], uv[j + 1]);
                    if (!FULL) {                                                          // identity steps past the end
                        if (j >= nv) { l2.x = 0.0f; u2.x = 0.0f; }
                        if (j + 1 >= nv) { l2.y = 0.0f; u2.y = 0.0f; }
                    }
                    const float2 y2 = mul2(kD2, u2);
                    u2 = mul2(u2, l2);
                    lw[j] = l2.x; lw[j + 1] = l2.y;
                    y[j] = y2.x; y[j + 1] = y2.y;
                    uv[j] = u2.x; uv[j + 1] = u2.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const float d = dv[j] + kb;
                    float l = SOFTPLUS ? softplus_f(d) : d;
                    if (!FULL && j >= nv) { l = 0.0f; uv[j] = 0.0f; }   // identity step past the end
                    lw[j] = l;
                    y[j] = kD * uv[j];
                    uv[j] *= l;
                }
            }
#pragma unroll
            for (int n = 0; n < NST; ++n) {
                const float kA = kc[n], hc = kc[NST + 2 + n];
                float hl[C], pl[C];
                float P = 1.0f, S_ = 0.0f;
                if constexpr (FAST) {
                    const float2 kA2 = make_float2(kA, kA);
#pragma unroll
                    for (int j = 0; j < C; j += 2) {
                        const float2 x2 = mul2(make_float2(lw[j], lw[j + 1]), kA2);
                        const float2 b2 = mul2(make_float2(uv[j], uv[j + 1]), make_float2(Bv[n][j], Bv[n][j + 1]));
                        const float a0 = ex2_approx(x2.x), a1 = ex2_approx(x2.y);
                        S_ = fmaf(a0, S_, b2.x); P *= a0; hl[j] = S_; pl[j] = P;
                        S_ = fmaf(a1, S_, b2.y); P *= a1; hl[j + 1] = S_; pl[j + 1] = P;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        const float a = ex2_approx(lw[j] * kA);
                        S_ = fmaf(a, S_, uv[j] * Bv[n][j]);
                        P *= a;
                        hl[j] = S_; pl[j] = P;
                    }
                }
                // warp-level inclusive scan of the affine maps h -> P*h + S across the 32 lane chunks
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float Pp = __shfl_up_sync(0xffffffffu, P, o), Sp = __shfl_up_sync(0xffffffffu, S_, o);
                    if (lane >= o) { S_ = fmaf(P, Sp, S_); P *= Pp; }
                }
                float Pe = __shfl_up_sync(0xffffffffu, P, 1), Se = __shfl_up_sync(0xffffffffu, S_, 1);
                if (lane == 0) { Pe = 1.0f; Se = 0.0f; }
                const float hin = fmaf(Pe, hc, Se);
                if constexpr (FAST) {
                    const float2 hin2 = make_float2(hin, hin);
#pragma unroll
                    for (int j = 0; j < C; j += 2) {
                        const float2 h2 = fma2(make_float2(pl[j], pl[j + 1]), hin2, make_float2(hl[j], hl[j + 1]));
                        const float2 y2 = fma2(h2, make_float2(Cv[n][j], Cv[n][j + 1]), make_float2(y[j], y[j + 1]));
                        y[j] = y2.x; y[j + 1] = y2.y;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < C; ++j) y[j] = fmaf(fmaf(pl[j], hin, hl[j]), Cv[n][j], y[j]);
                }
                if (lane == 31) sts32(rcs + (r * RCF + NST + 2 + n) * 4, fmaf(P, hc, S_));
            }
            if constexpr (HAS_Z) {
#pragma unroll
                for (int j = 0; j < C; ++j) y[j] *= silu_f(zv[j]);
            }
            const int64_t m0 = REV ? p.L - tok0 - C : tok0;      // lowest memory token of this lane's chunk
            store_tokens<OUT_T, C, REV, FULL>(orow0 + r * p.o_ds + m0, y, nv);
        };
        if (full_step) body(std::true_type{}); else body(std::false_type{});
        if (++r == w.nrows) { r = 0; ++step; }
    }
    __syncwarp();
    if (p.last && lane < w.nrows) {
#pragma unroll
        for (int n = 0; n < NST; ++n) p.last[(w.b * p.dim + w.d0 + lane) * NST + n] = rc[lane * RCF + NST + 2 + n];
    }
}

template <int NST, typename IN_T, typename OUT_T, int C, bool HAS_Z, bool SOFTPLUS, int NW, bool DTF>
__global__ void __launch_bounds__((NW + 1) * 32) scan_lanes_kernel(const ScanParams p) {
    using Cfg = LanesCfg<NST, IN_T, C, HAS_Z, NW, DTF>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* base = smem_raw;
    const int WB = Cfg::warp_bytes(DTF ? (int)p.R : 0);
    if (threadIdx.x == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(base + NW * WB);
        for (int i = 0; i < NW * Cfg::NBARS; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (warp == NW) {
        lanes_producer<NST, IN_T, C, HAS_Z, NW, DTF>(p, base, lane);
        return;
    }
    const LanesWarp w = lanes_decode(p, (int64_t)blockIdx.x * NW + warp);
    if (!w.valid) return;
    const bool rev = w.g < 64 && ((p.rev_mask >> w.g) & 1);
    if (rev) lanes_consumer<NST, IN_T, OUT_T, C, HAS_Z, SOFTPLUS, true, NW, DTF>(p, base, w, warp, lane);
    else lanes_consumer<NST, IN_T, OUT_T, C, HAS_Z, SOFTPLUS, false, NW, DTF>(p, base, w, warp, lane);
}

// ============================================================================================
// rows kernel (dstate 4 | 8 | 16)
// ============================================================================================
// At dstate 16 the scan is bound by the MUFU pipe and FP32 issue, not HBM: one ex2 and four FP32
// ops per (token, state) is the minimum, so the design spends nothing else per state and makes the per-token work
// (softplus, delta*u, D*u, stores) cooperative.
//   * a channel row is owned by TWO adjacent lanes; lane half hf keeps the states n with n % 2 == hf in registers and
//     runs their recurrences sequentially (no redundant scan arithmetic).  A warp covers 16 rows, a CTA 4 warps = 64
//     rows of one (batch, group) plus one producer warp.
//   * per 8-token step the lanes of a pair each evaluate softplus / delta*u for 4 of the 8 tokens and swap them with
//     SHFL, so the two MUFU ops of softplus are paid once per token, not once per lane.
//   * operands arrive by TMA: per-warp rings of u / delta [/ z] tiles [16 rows x 128 B] and ONE CTA-wide ring of
//     B / C tiles [dstate x 128 B] (shared by the 64 rows), all in 128B-swizzled shared memory so that the 16-byte
//     reads of a pair-interleaved warp are conflict-free; completion through mbarriers, refill by the producer warp.
//   * y: the pair's partial sums are exchanged with SHFL and each lane stores its 4 tokens (32 contiguous bytes per
//     row per step); no staging tile.
//   * reversed groups / shared-u addressing as in the lanes kernel (tile walk mirrored, elements reversed in registers).
constexpr int RT_ROWS = 16;      // channel rows per consumer warp (two lanes each)
constexpr int RT_NW = 4;         // consumer warps per CTA
constexpr int RT_STAGES = 3;     // u/delta ring depth per warp
constexpr int RT_BCS = 3;        // B/C ring depth per CTA

template <int NST, typename IN_T, bool HAS_Z> struct RowsCfg {
    static constexpr int VEC = 16 / (int)sizeof(IN_T);        // tokens per 16-byte chunk
    static constexpr int T = 8 * VEC;                          // tokens per tile (128 B per row)
    static constexpr int TILE = RT_ROWS * 128;                 // u / delta / z tile bytes (2 KiB)
    static constexpr int ITEM = TILE * (HAS_Z ? 3 : 2);
    static constexpr int TB = (NST * 128 + 1023) / 1024 * 1024;   // B (or C) tile bytes, 1 KiB granules keep the swizzle phase
    static constexpr int BC = 2 * TB;
    static constexpr int RING = RT_STAGES * ITEM;
    static constexpr int NBARS = RT_NW * 2 * RT_STAGES + 2 * RT_BCS;
    static constexpr int SMEM = RT_NW * RING + RT_BCS * BC + NBARS * 8 + 1024 /*align slack*/;
};

// 8 consecutive scan-order tokens (group c8 of the tile) of one tile row, fp32, from 128B-swizzled shared memory
template <typename IN_T, bool REV> __device__ __forceinline__ void lds_group8(uint32_t tile_saddr, int row, int c8, float (&f)[8]) {
    constexpr int VEC = 16 / (int)sizeof(IN_T), NCH = 8 / VEC;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int ch = REV ? 7 - (c8 * NCH + k) : c8 * NCH + k;
        float w[VEC];
        widen16<IN_T>(lds128(tile_saddr + row * 128 + ((ch ^ (row & 7)) << 4)), w);
#pragma unroll
        for (int i = 0; i < VEC; ++i) f[k * VEC + i] = w[REV ? VEC - 1 - i : i];
    }
}

template <int NST, typename IN_T, typename OUT_T, bool HAS_Z, bool SOFTPLUS, bool REV>
__device__ __forceinline__ void rows_consumer(const ScanParams& p, uint32_t sbase, int warp, int lane, int64_t b, int64_t g,
                                              int64_t dg0) {
    using Cfg = RowsCfg<NST, IN_T, HAS_Z>;
    constexpr int T = Cfg::T, SPL = NST / 2;
    const uint32_t ring = sbase + warp * Cfg::RING;
    const uint32_t bcring = sbase + RT_NW * Cfg::RING;
    const uint32_t bars = bcring + RT_BCS * Cfg::BC;
    const uint32_t full = bars + warp * 2 * RT_STAGES * 8, empty = full + RT_STAGES * 8;
    const uint32_t bcfull = bars + RT_NW * 2 * RT_STAGES * 8, bcempty = bcfull + RT_BCS * 8;

    const int r = lane >> 1, hf = lane & 1;
    const int64_t Dg = p.dim / p.groups;
    const bool row_ok = dg0 + r < Dg;
    const int64_t d = g * Dg + dg0 + r;
    float A2[SPL], h[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) { A2[i] = row_ok ? p.A[d * NST + 2 * i + hf] * kLog2e : 0.0f; h[i] = 0.0f; }
    const float bias = (row_ok && p.bias) ? p.bias[d] : 0.0f;
    const float Dv = (row_ok && p.D) ? p.D[d] : 0.0f;
    const float Dm0 = hf == 0 ? Dv : 0.0f, Dm1 = hf == 1 ? Dv : 0.0f;   // D*u enters the pair sum exactly once
    const int src0 = lane & ~1, src1 = lane | 1;
    OUT_T* orow = (OUT_T*)p.out + b * p.o_bs + d * p.o_ds;

    const int ntiles = (int)((p.L + T - 1) / T);
    for (int tile = 0; tile < ntiles; ++tile) {
        const int slot = tile % RT_STAGES, bslot = tile % RT_BCS;
        mbar_wait_s(full + slot * 8, (uint32_t)(tile / RT_STAGES) & 1u);
        mbar_wait_s(bcfull + bslot * 8, (uint32_t)(tile / RT_BCS) & 1u);
        const uint32_t su = ring + slot * Cfg::ITEM, sd = su + Cfg::TILE, sz = sd + Cfg::TILE;
        const uint32_t sB = bcring + bslot * Cfg::BC, sC = sB + Cfg::TB;
        const int valid = (int)min((int64_t)T, p.L - (int64_t)tile * T);   // multiple of 8 (host-checked)
#pragma unroll 1
        for (int c8 = 0; c8 * 8 < valid; ++c8) {
            float u8[8], d8[8];
            lds_group8<IN_T, REV>(su, r, c8, u8);
            lds_group8<IN_T, REV>(sd, r, c8, d8);
            // this lane evaluates tokens 2j + hf, then the pair swaps: 4 softplus per lane instead of 8.
            // Token pairs (2j, 2j+1) stay packed as float2 so the per-state multiplies and the y update issue as
            // FMUL2 / FFMA2 (two fp32 ops per issue slot); only the recurrence itself is scalar.
            float2 dl2[4], du2[4], yp2[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float uo = hf ? u8[2 * j + 1] : u8[2 * j];
                const float x = (hf ? d8[2 * j + 1] : d8[2 * j]) + bias;
                const float dlo = SOFTPLUS ? softplus1_f(x) : x;      // one MUFU: this kernel is bound by the XU pipe
                const float duo = dlo * uo;
                dl2[j] = make_float2(__shfl_sync(0xffffffffu, dlo, src0), __shfl_sync(0xffffffffu, dlo, src1));
                du2[j] = make_float2(__shfl_sync(0xffffffffu, duo, src0), __shfl_sync(0xffffffffu, duo, src1));
                yp2[j] = make_float2(Dm0 * uo, Dm1 * uo);
            }
            // two states at a time: their recurrences are independent, so the dependent FFMA pairs of one state fill the
            // latency slots of the other (ncu: 1.2 "wait" stall cycles per issued instruction with one state at a time)
#pragma unroll
            for (int i = 0; i < SPL; i += 2) {
                float bv0[8], cv0[8], bv1[8], cv1[8];
                lds_group8<IN_T, REV>(sB, 2 * i + hf, c8, bv0);
                lds_group8<IN_T, REV>(sB, 2 * (i + 1) + hf, c8, bv1);
                lds_group8<IN_T, REV>(sC, 2 * i + hf, c8, cv0);
                lds_group8<IN_T, REV>(sC, 2 * (i + 1) + hf, c8, cv1);
                const float2 A0 = make_float2(A2[i], A2[i]), A1 = make_float2(A2[i + 1], A2[i + 1]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 arg0 = mul2(dl2[q], A0), arg1 = mul2(dl2[q], A1);
                    const float2 xb0 = mul2(du2[q], make_float2(bv0[2 * q], bv0[2 * q + 1]));
                    const float2 xb1 = mul2(du2[q], make_float2(bv1[2 * q], bv1[2 * q + 1]));
                    const float e0 = ex2_approx(arg0.x), e1 = ex2_approx(arg1.x), e2 = ex2_approx(arg0.y), e3 = ex2_approx(arg1.y);
                    float2 h0, h1;
                    h0.x = fmaf(e0, h[i], xb0.x);
                    h1.x = fmaf(e1, h[i + 1], xb1.x);
                    h0.y = fmaf(e2, h0.x, xb0.y);
                    h1.y = fmaf(e3, h1.x, xb1.y);
                    h[i] = h0.y; h[i + 1] = h1.y;
                    yp2[q] = fma2(h0, make_float2(cv0[2 * q], cv0[2 * q + 1]), yp2[q]);
                    yp2[q] = fma2(h1, make_float2(cv1[2 * q], cv1[2 * q + 1]), yp2[q]);
                }
            }
            float yp[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { yp[2 * q] = yp2[q].x; yp[2 * q + 1] = yp2[q].y; }
            const int64_t tok = (int64_t)tile * T + c8 * 8;          // first scan-order token of the group
            if constexpr (sizeof(OUT_T) == 4) {
                // lane hf finishes tokens 4*hf .. 4*hf+3: send the partner's half, keep mine
                float tot[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float send = hf ? yp[j] : yp[4 + j], keep = hf ? yp[4 + j] : yp[j];
                    tot[j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                }
                if constexpr (HAS_Z) {
                    float z8[8];
                    lds_group8<IN_T, REV>(sz, r, c8, z8);
#pragma unroll
                    for (int j = 0; j < 4; ++j) tot[j] *= silu_f(hf ? z8[4 + j] : z8[j]);
                }
                if (row_ok) {
                    const int64_t t4 = tok + 4 * hf;
                    float m[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) m[REV ? 3 - j : j] = tot[j];
                    VecIO<float, 4>::store(reinterpret_cast<float*>(orow) + (REV ? p.L - t4 - 4 : t4), m);
                }
            } else {
                float tot[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) tot[t] = yp[t] + __shfl_xor_sync(0xffffffffu, yp[t], 1);
                if constexpr (HAS_Z) {
                    float z8[8];
                    lds_group8<IN_T, REV>(sz, r, c8, z8);
#pragma unroll
                    for (int t = 0; t < 8; ++t) tot[t] *= silu_f(z8[t]);
                }
                if (row_ok && hf == 0) {
                    float m[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) m[REV ? 7 - t : t] = tot[t];
                    VecIO<OUT_T, 8>::store(orow + (REV ? p.L - tok - 8 : tok), m);
                }
            }
        }
        __syncwarp();                                   // every lane is done with both stages
        if (lane == 0) { mbar_arrive_s(empty + slot * 8); mbar_arrive_s(bcempty + bslot * 8); }
    }
    if (p.last && row_ok) {
#pragma unroll
        for (int i = 0; i < SPL; ++i) p.last[(b * p.dim + d) * NST + 2 * i + hf] = h[i];
    }
}

template <int NST, typename IN_T, bool HAS_Z>
__device__ __forceinline__ void rows_producer(const ScanParams& p, uint8_t* base, int lane, int nvalid, int64_t b, int64_t g,
                                              int64_t dg_cta, bool rev, const CUtensorMap* map_u, const CUtensorMap* map_d,
                                              const CUtensorMap* map_z, const CUtensorMap* map_B, const CUtensorMap* map_C) {
    using Cfg = RowsCfg<NST, IN_T, HAS_Z>;
    constexpr int T = Cfg::T;
    uint8_t* bcring = base + RT_NW * Cfg::RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bcring + RT_BCS * Cfg::BC);
    const bool is_bc = lane == RT_NW;                    // lane RT_NW feeds the shared B/C ring, lanes < nvalid the warps
    const bool mine = lane < nvalid || is_bc;
    const int depth = is_bc ? RT_BCS : RT_STAGES;
    uint64_t* full = is_bc ? bars + RT_NW * 2 * RT_STAGES : bars + (lane < RT_NW ? lane : 0) * 2 * RT_STAGES;
    uint64_t* empty = full + depth;
    uint8_t* ring = is_bc ? bcring : base + (lane < RT_NW ? lane : 0) * Cfg::RING;
    const int stage_bytes = is_bc ? Cfg::BC : Cfg::ITEM;
    const int ntiles = (int)((p.L + T - 1) / T);
    const int total = mine ? ntiles : 0;
    const int gu = (int)(g / p.u_gdiv), gi = (int)g, bi = (int)b;
    const int dg0 = (int)dg_cta + (lane < RT_NW ? lane : 0) * RT_ROWS;
    int tile = 0, slot = 0;
    uint32_t fill = 0;
    while (true) {
        const bool active = tile < total;
        if (!__any_sync(0xffffffffu, active)) break;
        bool issued = false;
        if (active && mbar_test_wait(&empty[slot], (fill - 1u) & 1u)) {
            // mirrored walk for reversed groups: scan tile i covers memory tokens [L - (i+1)T, L - iT); a negative start is
            // out of bounds for TMA and zero-filled, the consumer only touches the valid (high) end
            const int t0 = rev ? (int)(p.L - (int64_t)(tile + 1) * T) : tile * T;
            uint8_t* st = ring + slot * stage_bytes;
            // TMA always delivers the whole box (out-of-bounds rows / tokens are zero-filled), so the byte count is fixed
            mbar_arrive_expect_tx(&full[slot], (uint32_t)(is_bc ? 2 * NST * 128 : Cfg::ITEM));
            if (is_bc) {
                tma_load_4d(st, map_B, &full[slot], t0, 0, gi, bi);
                tma_load_4d(st + Cfg::TB, map_C, &full[slot], t0, 0, gi, bi);
            } else {
                tma_load_4d(st, map_u, &full[slot], t0, dg0, gu, bi);
                tma_load_4d(st + Cfg::TILE, map_d, &full[slot], t0, dg0, gi, bi);
                if (HAS_Z) tma_load_4d(st + 2 * Cfg::TILE, map_z, &full[slot], t0, dg0, gi, bi);
            }
            if (++slot == depth) { slot = 0; ++fill; }
            ++tile;
            issued = true;
        }
        if (!__any_sync(0xffffffffu, issued)) __nanosleep(40);
    }
}

template <int NST, typename IN_T, typename OUT_T, bool HAS_Z, bool SOFTPLUS>
__global__ void __launch_bounds__((RT_NW + 1) * 32, 3)
scan_rows_kernel(const ScanParams p, const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_d,
                 const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_B,
                 const __grid_constant__ CUtensorMap map_C) {
    using Cfg = RowsCfg<NST, IN_T, HAS_Z>;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s0 = smem_u32(smem_raw);
    const uint32_t pad = ((s0 + 1023u) & ~1023u) - s0;   // 1 KiB alignment: the 128B swizzle is a function of address bits 4..9
    uint8_t* base = smem_raw + pad;

    const int64_t Dg = p.dim / p.groups;
    const int64_t rbpg = (Dg + RT_NW * RT_ROWS - 1) / (RT_NW * RT_ROWS);
    const int64_t cta = blockIdx.x;
    const int64_t rb = cta % rbpg, g = (cta / rbpg) % p.groups, b = cta / (rbpg * p.groups);
    const int64_t dg_cta = rb * RT_NW * RT_ROWS;
    const int nvalid = (int)min((int64_t)RT_NW, (Dg - dg_cta + RT_ROWS - 1) / RT_ROWS);
    const bool rev = g < 64 && ((p.rev_mask >> g) & 1);
    if (threadIdx.x == 0) {
        uint64_t* bars = reinterpret_cast<uint64_t*>(base + RT_NW * Cfg::RING + RT_BCS * Cfg::BC);
        for (int i = 0; i < RT_NW * 2 * RT_STAGES; ++i) mbar_init(&bars[i], 1);
        for (int i = 0; i < RT_BCS; ++i) {
            mbar_init(&bars[RT_NW * 2 * RT_STAGES + i], 1);                       // bcfull: the producer's expect_tx
            mbar_init(&bars[RT_NW * 2 * RT_STAGES + RT_BCS + i], nvalid);          // bcempty: one arrival per live warp
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    __syncthreads();
    if (warp == RT_NW) {
        rows_producer<NST, IN_T, HAS_Z>(p, base, lane, nvalid, b, g, dg_cta, rev, &map_u, &map_d, &map_z, &map_B, &map_C);
        return;
    }
    if (warp >= nvalid) return;
    const uint32_t sbase = s0 + pad;
    if (rev) rows_consumer<NST, IN_T, OUT_T, HAS_Z, SOFTPLUS, true>(p, sbase, warp, lane, b, g, dg_cta + warp * RT_ROWS);
    else rows_consumer<NST, IN_T, OUT_T, HAS_Z, SOFTPLUS, false>(p, sbase, warp, lane, b, g, dg_cta + warp * RT_ROWS);
}

// ============================================================================================
// host dispatch
// ============================================================================================
template <typename IN_T, typename OUT_T> static int launch_generic(const ScanParams& p, cudaStream_t st) {
    const int64_t rows = p.batch * p.dim;
    scan_generic_kernel<IN_T, OUT_T><<<(unsigned)ceil_div(rows, GEN_WARPS), GEN_WARPS * 32, 0, st>>>(p);
    XP_LAUNCH_CHECK("scan_generic_kernel");
    return XP_OK;
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <int NST, typename IN_T, typename OUT_T, int C, bool HAS_Z, bool SOFTPLUS, int NW, bool DTF>
static int launch_lanes_cfg(const ScanParams& p, int64_t warps, cudaStream_t st) {
    const int smem = LanesCfg<NST, IN_T, C, HAS_Z, NW, DTF>::smem(DTF ? (int)p.R : 0);
    auto kern = scan_lanes_kernel<NST, IN_T, OUT_T, C, HAS_Z, SOFTPLUS, NW, DTF>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)ceil_div(warps, NW), (NW + 1) * 32, smem, st>>>(p);
    XP_LAUNCH_CHECK("scan_lanes_kernel");
    return XP_OK;
}

template <int NST, typename IN_T, typename OUT_T, int C> static int launch_lanes_c(const ScanParams& p, int64_t warps, cudaStream_t st) {
    constexpr int NW = 4;
    if constexpr (sizeof(IN_T) == 2) {
        if (p.R > 0) {      // fused dt_proj (16-bit inputs, no z gate: lanes_supports_dt)
            return p.softplus ? launch_lanes_cfg<NST, IN_T, OUT_T, C, false, true, NW, true>(p, warps, st)
                              : launch_lanes_cfg<NST, IN_T, OUT_T, C, false, false, NW, true>(p, warps, st);
        }
    }
    if (p.z) {
        return p.softplus ? launch_lanes_cfg<NST, IN_T, OUT_T, C, true, true, NW, false>(p, warps, st)
                          : launch_lanes_cfg<NST, IN_T, OUT_T, C, true, false, NW, false>(p, warps, st);
    }
    return p.softplus ? launch_lanes_cfg<NST, IN_T, OUT_T, C, false, true, NW, false>(p, warps, st)
                      : launch_lanes_cfg<NST, IN_T, OUT_T, C, false, false, NW, false>(p, warps, st);
}

constexpr int LN_MAX_DT_RANK = 16;   // fused dt_proj on the lanes kernel: one m16n8k8 / m16n8k16 MMA along the rank
template <typename IN_T> static bool lanes_supports_dt(const ScanParams& p) {
    return sizeof(IN_T) == 2 && p.R <= LN_MAX_DT_RANK && !p.z;
}

template <int NST, typename IN_T, typename OUT_T> static int launch_lanes(ScanParams p, cudaStream_t st) {
    const int64_t Dg = p.dim / p.groups;
    // rows per warp: share B/C across as many rows as possible while keeping >= ~2 waves of warps
    const int64_t target_warps = (int64_t)num_sms() * 16 * 2;
    int rw = p.R > 0 ? LN_DT_ROWS : LN_MAXROWS;
    while (rw > 1 && p.batch * p.groups * ceil_div(Dg, rw) < target_warps) rw >>= 1;
    static const int rw_override = env_int("XP_LANES_RW", 0);   // tuning knob
    if (rw_override > 0 && rw_override <= (p.R > 0 ? LN_DT_ROWS : LN_MAXROWS)) rw = rw_override;
    p.rows_per_warp = rw;
    const int64_t warps = p.batch * p.groups * ceil_div(Dg, rw);
    // tokens per lane per step: 16 for single-state 16-bit inputs (halves the per-step overhead) unless the
    // 512-token steps would pad the sequence more than 256-token steps do (L = 1280: 3 x 512 vs 5 x 256)
    static const int c_override = env_int("XP_LANES_C", 0);   // tuning knob: 8 | 16
    if constexpr (NST == 1 && sizeof(IN_T) == 2) {
        const bool pads_more = ceil_div(p.L, 512) * 512 > ceil_div(p.L, 256) * 256;
        // fused dt_proj: the dts_r / delta tiles grow with the step; keep at least two CTAs per SM
        static const int dt_c16 = env_int("XP_LANES_DT_C16", 0);   // measured: 256-token steps (4 CTAs/SM) beat 512-token steps (2 CTAs/SM)
        const bool too_big = p.R > 0 && (LanesCfg<NST, IN_T, 16, false, 4, true>::smem((int)p.R) > 110 * 1024 || !dt_c16);
        if (c_override == 16 || (c_override != 8 && !pads_more && !too_big)) return launch_lanes_c<NST, IN_T, OUT_T, 16>(p, warps, st);
    }
    return launch_lanes_c<NST, IN_T, OUT_T, 8>(p, warps, st);
}

template <int NST, typename IN_T, typename OUT_T, bool HAS_Z, bool SOFTPLUS>
static int launch_rows_cfg(const ScanParams& p, int in_dt, cudaStream_t st) {
    using Cfg = RowsCfg<NST, IN_T, HAS_Z>;
    const uint64_t Dg = (uint64_t)(p.dim / p.groups);
    const uint64_t es = sizeof(IN_T);
    CUtensorMap mu, md, mz, mB, mC;
    {
        const uint32_t box[4] = {(uint32_t)Cfg::T, RT_ROWS, 1, 1};
        const uint64_t udims[4] = {(uint64_t)p.L, Dg, (uint64_t)ceil_div(p.groups, p.u_gdiv), (uint64_t)p.batch};
        const uint64_t su[3] = {(uint64_t)p.u_ds * es, (uint64_t)p.u_gs * es, (uint64_t)p.u_bs * es};
        int rc = make_tensor_map(&mu, in_dt, 4, p.u, udims, su, box, 1);
        if (rc) return rc;
        const uint64_t dims[4] = {(uint64_t)p.L, Dg, (uint64_t)p.groups, (uint64_t)p.batch};
        const uint64_t sd[3] = {(uint64_t)p.dl_ds * es, Dg * p.dl_ds * es, (uint64_t)p.dl_bs * es};
        if ((rc = make_tensor_map(&md, in_dt, 4, p.delta, dims, sd, box, 1))) return rc;
        if (HAS_Z) {
            const uint64_t sz[3] = {(uint64_t)p.z_ds * es, Dg * p.z_ds * es, (uint64_t)p.z_bs * es};
            if ((rc = make_tensor_map(&mz, in_dt, 4, p.z, dims, sz, box, 1))) return rc;
        } else {
            mz = md;
        }
        const uint64_t bdims[4] = {(uint64_t)p.L, (uint64_t)NST, (uint64_t)p.groups, (uint64_t)p.batch};
        const uint32_t bbox[4] = {(uint32_t)Cfg::T, NST, 1, 1};
        const uint64_t sB[3] = {(uint64_t)p.B_ss * es, (uint64_t)p.B_gs * es, (uint64_t)p.B_bs * es};
        if ((rc = make_tensor_map(&mB, in_dt, 4, p.Bm, bdims, sB, bbox, 1))) return rc;
        const uint64_t sC[3] = {(uint64_t)p.C_ss * es, (uint64_t)p.C_gs * es, (uint64_t)p.C_bs * es};
        if ((rc = make_tensor_map(&mC, in_dt, 4, p.Cm, bdims, sC, bbox, 1))) return rc;
    }
    auto kern = scan_rows_kernel<NST, IN_T, OUT_T, HAS_Z, SOFTPLUS>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int64_t ctas = p.batch * p.groups * ceil_div((int64_t)Dg, RT_NW * RT_ROWS);
    XP_REQUIRE(ctas < ((int64_t)1 << 31), "selective scan: too many row blocks for one launch");
    kern<<<(unsigned)ctas, (RT_NW + 1) * 32, Cfg::SMEM, st>>>(p, mu, md, mz, mB, mC);
    XP_LAUNCH_CHECK("scan_rows_kernel");
    return XP_OK;
}

template <int NST, typename IN_T, typename OUT_T>
static int launch_rows(const ScanParams& p, int in_dt, cudaStream_t st) {
    if (p.z) {
        return p.softplus ? launch_rows_cfg<NST, IN_T, OUT_T, true, true>(p, in_dt, st)
                          : launch_rows_cfg<NST, IN_T, OUT_T, true, false>(p, in_dt, st);
    }
    return p.softplus ? launch_rows_cfg<NST, IN_T, OUT_T, false, true>(p, in_dt, st)
                      : launch_rows_cfg<NST, IN_T, OUT_T, false, false>(p, in_dt, st);
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

template <typename IN_T, typename OUT_T> static int dispatch(const ScanParams& p, const xp_scan_args* a, cudaStream_t st) {
    const int64_t va = 16 / (int64_t)sizeof(IN_T);    // elements per 16 bytes (input side)
    const int64_t vo = 16 / (int64_t)sizeof(OUT_T);
    bool vec_ok = !a->force_generic && p.delta_dim == p.dim && p.L % va == 0 && p.L % vo == 0;
    vec_ok = vec_ok && aligned16(p.u) && aligned16(p.delta) && aligned16(p.Bm) && aligned16(p.Cm) && aligned16(p.out) &&
             (!p.z || aligned16(p.z));
    const int64_t in_strides[] = {p.u_bs, p.u_ds, p.dl_bs, p.dl_ds, p.B_bs, p.B_gs, p.B_ss, p.C_bs, p.C_gs, p.C_ss,
                                  p.z ? p.z_bs : 0, p.z ? p.z_ds : 0};
    for (int64_t s : in_strides) vec_ok = vec_ok && (s % va == 0);
    vec_ok = vec_ok && (p.o_bs % vo == 0) && (p.o_ds % vo == 0);
    vec_ok = vec_ok && (p.u_gs % va == 0);
    if (p.R > 0) {            // fused dt_proj: lanes kernel or the generic kernel
        vec_ok = vec_ok && p.dl_gs % va == 0 && lanes_supports_dt<IN_T>(p);
        if (vec_ok && p.dstate == 1) return launch_lanes<1, IN_T, OUT_T>(p, st);
        if (vec_ok && p.dstate == 2) return launch_lanes<2, IN_T, OUT_T>(p, st);
        return launch_generic<IN_T, OUT_T>(p, st);
    }
    if (vec_ok && p.dstate == 1) return launch_lanes<1, IN_T, OUT_T>(p, st);
    if (vec_ok && p.dstate == 2) return launch_lanes<2, IN_T, OUT_T>(p, st);
    if (vec_ok && p.L % 8 == 0 && p.batch < 32768 * 65536LL) {   // whole 8-token steps; TMA coordinates are int32
        switch (p.dstate) {
            case 4: return launch_rows<4, IN_T, OUT_T>(p, a->in_dtype, st);
            case 8: return launch_rows<8, IN_T, OUT_T>(p, a->in_dtype, st);
            case 16: return launch_rows<16, IN_T, OUT_T>(p, a->in_dtype, st);
            default: break;
        }
    }
    return launch_generic<IN_T, OUT_T>(p, st);
}

}  // namespace xp

using namespace xp;

extern "C" int xp_selective_scan_fwd(const xp_scan_args* a, xp_stream_t stream) {
    XP_REQUIRE(a != nullptr, "xp_selective_scan_fwd: args is NULL");
    XP_REQUIRE(a->u && a->delta && a->A && a->B && a->C && a->out, "xp_selective_scan_fwd: u/delta/A/B/C/out must be non-NULL");
    XP_REQUIRE(a->batch >= 0 && a->dim > 0 && a->seqlen >= 0 && a->dstate > 0 && a->groups > 0 && a->delta_dim > 0,
               "xp_selective_scan_fwd: sizes must be positive");
    XP_REQUIRE(a->dstate <= 256, "selective_scan only supports state dimension <= 256 (got %lld)", (long long)a->dstate);
    XP_REQUIRE(a->dim % a->groups == 0, "dim (%lld) must be divisible by the number of B/C groups (%lld)",
               (long long)a->dim, (long long)a->groups);
    XP_REQUIRE(a->dim % a->delta_dim == 0, "dim (%lld) must be divisible by delta_dim (%lld)", (long long)a->dim,
               (long long)a->delta_dim);
    XP_REQUIRE(a->in_dtype >= XP_F32 && a->in_dtype <= XP_BF16, "unsupported input dtype %d", a->in_dtype);
    XP_REQUIRE(a->out_dtype == XP_F32 || a->out_dtype == a->in_dtype, "out dtype must be fp32 or the input dtype");
    if (a->batch == 0 || a->seqlen == 0) return XP_OK;
    ScanParams p;
    p.u = a->u; p.delta = a->delta; p.A = a->A; p.Bm = a->B; p.Cm = a->C; p.D = a->D; p.z = a->z; p.bias = a->delta_bias;
    p.out = a->out; p.last = a->last_state;
    p.batch = a->batch; p.dim = a->dim; p.delta_dim = a->delta_dim; p.groups = a->groups; p.dstate = a->dstate; p.L = a->seqlen;
    p.u_bs = a->u_batch_stride; p.u_ds = a->u_dim_stride; p.dl_bs = a->delta_batch_stride; p.dl_ds = a->delta_dim_stride;
    p.B_bs = a->B_batch_stride; p.B_gs = a->B_group_stride; p.B_ss = a->B_state_stride;
    p.C_bs = a->C_batch_stride; p.C_gs = a->C_group_stride; p.C_ss = a->C_state_stride;
    p.z_bs = a->z_batch_stride; p.z_ds = a->z_dim_stride; p.o_bs = a->out_batch_stride; p.o_ds = a->out_dim_stride;
    p.softplus = a->delta_softplus; p.rows_per_warp = 1;
    if (a->u_group_div > 0) { p.u_gdiv = a->u_group_div; p.u_gs = a->u_group_stride; }
    else { p.u_gdiv = 1; p.u_gs = (a->dim / a->groups) * a->u_dim_stride; }
    p.rev_mask = a->reverse_group_mask;
    p.wdt = nullptr; p.R = 0; p.dl_gs = 0;
    if (a->dt_rank != 0) {
        XP_REQUIRE(a->dt_rank > 0 && a->dt_rank <= 64, "fused dt_proj: dt_rank must be in 1..64 (got %lld)", (long long)a->dt_rank);
        XP_REQUIRE(a->dt_weight != nullptr, "fused dt_proj: dt_weight is NULL");
        XP_REQUIRE(a->delta_dim == a->dim, "fused dt_proj needs delta_dim == dim");
        p.wdt = a->dt_weight; p.R = a->dt_rank; p.dl_gs = a->dt_group_stride;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int key = a->in_dtype * 4 + a->out_dtype;
    switch (key) {
        case XP_F32 * 4 + XP_F32: return dispatch<float, float>(p, a, st);
        case XP_F16 * 4 + XP_F32: return dispatch<__half, float>(p, a, st);
        case XP_F16 * 4 + XP_F16: return dispatch<__half, __half>(p, a, st);
        case XP_BF16 * 4 + XP_F32: return dispatch<__nv_bfloat16, float>(p, a, st);
        case XP_BF16 * 4 + XP_BF16: return dispatch<__nv_bfloat16, __nv_bfloat16>(p, a, st);
        default: break;
    }
    set_error("xp_selective_scan_fwd: unsupported dtype combination in=%d out=%d", a->in_dtype, a->out_dtype);
    return XP_ERR_INVALID_ARG;
}

extern "C" int xp_selective_scan_bwd(void) {
    set_error("xp_selective_scan_bwd: this library is inference-only (forward kernels only)");
    return XP_ERR_UNSUPPORTED;
}
