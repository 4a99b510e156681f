// mlp_tc.cu -- the whole MLP branch of a VSSBlock in ONE tcgen05 kernel (xp_mlp_res_ln):
//
//     x_new = x + fc2( GELU( fc1(n) + b1 ) ) + b2 ;   y = LayerNorm_next(x_new)            (VMamba.py:110-128, 1229-1233)
//
// The 4C-wide hidden tensor (2 + 2 GB of HBM traffic per stage-0 block when fc1 and fc2 are separate kernels) never leaves
// the SM: per 128-row tile the hidden activation is produced 64 columns at a time, GELU'd on its way out of TMEM and handed to the
// second GEMM as its A operand (double-buffered) -- in tensor memory at C = 96 (packed 16-bit pairs, tcgen05.mma [d], [a_tmem], b),
// as a K-major 128B-swizzled shared-memory k-block at C = 192, where the 512 TMEM columns are taken by the accumulators.
//
// Persistent CTAs over 128-row tiles, C = 96 or 192; 28 warps, each role with its own instruction stream so that the
// FMA-pipe-bound GELU and the latency-bound residual / LayerNorm epilogue overlap.  g = flat index of a 64-column hidden chunk:
//   warp 0      TMA producer (one lane): the tile's A operand n [128 x C], then per chunk W1(g) and W2(g - 2) k-block tiles
//   warp 1      GEMM1 issuer (one lane): acc1[g % 2] = n W1[g]^T (K = C), as soon as its epilogue-1 group has read chunk g - 2
//   warp 3      GEMM2 issuer (one lane): acc2[tile % 2] += H(g) W2[:, g]^T (K = 64) when H(g) is written.  Two issuers, because
//               one thread issuing both GEMMs (with ring arithmetic) was the critical path of the whole CTA
//   warp 2      TMEM allocation: acc1 2 x 64 columns + acc2 2 x C columns (512 at C = 192) [+ H 2 x 32 columns at C = 96]
//   warps 4-19  epilogue 1, thread = row: two groups of 8 warps (even / odd chunks, so one group's barrier waits hide behind the
//               other's arithmetic), two warps (32-column halves) per TMEM lane quarter: tcgen05.ld -> + b1 -> exact GELU ->
//               16-bit -> H[g % 2]: tcgen05.st into tensor memory, or the shared-memory k-block (+ fence.proxy.async) at C = 192
//   warps 20-27 epilogue 2, thread = row, two warps (C/2-column halves) per lane quarter: acc2 + b2 + residual -> one-pass
//               shifted sums -> x_new (fp32) and LayerNorm (16-bit) out.  The residual arrives by cp.async two 16-column units
//               ahead (across tiles) in a per-warp ring of TMA-swizzled slots; the new residual rows are written back in place
//               and leave as one TMA store per unit (cp.async.bulk.tensor), the 16-bit rows likewise from two small slots.
#include "common.cuh"
#include "tcgen05.cuh"

namespace xp {

constexpr int MP_BM = 128, MP_BK = 64, MP_HC = 64;
constexpr int MP_KB_TILE = MP_BM * 128;                // one 64-wide k-block of a 128-row operand: 16 KiB
constexpr int MP_E1_WARPS = 16, MP_E2_WARPS = 8;     // epilogue 1: two groups of 8 warps (even / odd chunks)
constexpr int MP_E1G = MP_E1_WARPS / 2;                // warps per epilogue-1 group
constexpr int MP_E1C = MP_HC / (MP_E1G / 4);           // hidden columns per epilogue-1 warp and chunk (32)
constexpr int MP_THREADS = (4 + MP_E1_WARPS + MP_E2_WARPS) * 32;
constexpr int MP_PF = 2;                               // residual units in flight per epilogue-2 warp
// epilogue-2 staging, all of it in the layouts TMA stores read: a ring of three fp32 units [32 rows x 64 B] (64B swizzle: 16-byte
// piece ^= (row >> 1) & 3 -- conflict-free for thread = row) and two 16-bit units [32 rows x 32 B] (32B swizzle: piece ^= (row >> 2) & 1)
constexpr int MP_XSLOT = 32 * 64, MP_XSLOTS = MP_PF + 1, MP_YSLOT = 32 * 32;
constexpr int MP_E2_SMEM = MP_XSLOTS * MP_XSLOT + 2 * MP_YSLOT;
static_assert(MP_E2_SMEM % 1024 == 0, "the swizzle patterns are functions of the absolute shared-memory address");
constexpr int MP_BAR_BYTES = 256;

template <int C> struct MpCfg {
    static constexpr int HD = 4 * C;
    static constexpr int NCHUNK = HD / MP_HC;
    static constexpr int KB1 = (C + MP_BK - 1) / MP_BK;            // k-blocks of GEMM1 (K = C; the tail is TMA zero fill)
    static constexpr int A1_BYTES = KB1 * MP_KB_TILE;
    static constexpr int NA1 = C == 96 ? 2 : 1;                    // A-operand buffers (what shared memory allows)
    static constexpr int W1_STAGE = MP_HC * 128, S1 = C == 96 ? 4 : 3;
    static constexpr int W2_STAGE = C * 128, S2 = C == 96 ? 3 : 2;
    static constexpr int ACC2_COL = 2 * MP_HC;
    // H (the GELU output = A operand of GEMM2) in tensor memory when the 512 columns allow it: no shared-memory round trip and
    // no proxy fence; 2 x 32 columns of packed 16-bit pairs behind the accumulators
    static constexpr bool H_TMEM = 2 * MP_HC + 2 * C + 2 * (MP_HC / 2) <= 512;
    static constexpr int H_COL = ACC2_COL + 2 * C;
    static constexpr int H_SMEM = H_TMEM ? 0 : 2 * MP_KB_TILE;
    static constexpr int UPW = C / 32;                             // 16-column units per epilogue-2 warp and tile
    static constexpr int VEC_FLOATS = HD + 3 * C;                  // b1 | b2 | gamma | beta
    static constexpr int SMEM = NA1 * A1_BYTES + S1 * W1_STAGE + H_SMEM + S2 * W2_STAGE + 1024 /*align*/ + MP_BAR_BYTES
                                + MP_E2_WARPS * MP_E2_SMEM + VEC_FLOATS * 4 + 2 * MP_E2_WARPS * 32 * 8;
    static_assert(2 * MP_HC + 2 * C <= 512, "TMEM budget");
    static_assert((2 * NA1 + 2 * S1 + 2 * S2 + 12) * 8 + 4 <= MP_BAR_BYTES, "barrier block");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {     // src_bytes 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {       // explicit shared-space load (a generic LD costs a long scoreboard)
    const uint4 r = lds128(saddr);
    return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ float4 lds_const_f4(uint32_t saddr) {   // data that never changes after set-up: the compiler may hoist it
    float4 r;
    asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
    return r;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int C, bool BF16>
__global__ void __launch_bounds__(MP_THREADS, 1)
mlp_res_ln_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w1,
                     const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_x,
                     const __grid_constant__ CUtensorMap map_y, const float* __restrict__ b1, const float* __restrict__ b2,
                     const float* __restrict__ res, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ xnew, void* __restrict__ yout, int M, float eps) {
    using Cfg = MpCfg<C>;
    constexpr int NCHUNK = Cfg::NCHUNK, KB1 = Cfg::KB1, S1 = Cfg::S1, S2 = Cfg::S2, NA1 = Cfg::NA1, UPW = Cfg::UPW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a1 = base;
    uint8_t* w1 = a1 + NA1 * Cfg::A1_BYTES;
    uint8_t* hbuf = w1 + S1 * Cfg::W1_STAGE;
    uint8_t* w2 = hbuf + Cfg::H_SMEM;
    uint8_t* e2_smem = w2 + S2 * Cfg::W2_STAGE;                                     // 1024-aligned: every tile above is n x 1 KiB
    uint64_t* bars = reinterpret_cast<uint64_t*>(e2_smem + MP_E2_WARPS * MP_E2_SMEM);
    uint64_t* a1_full = bars;                   // [NA1]
    uint64_t* a1_empty = a1_full + NA1;         // [NA1]
    uint64_t* w1_full = a1_empty + NA1;         // [S1]
    uint64_t* w1_empty = w1_full + S1;
    uint64_t* w2_full = w1_empty + S1;          // [S2]
    uint64_t* w2_empty = w2_full + S2;
    uint64_t* acc1_full = w2_empty + S2;        // [2] each from here on
    uint64_t* acc1_empty = acc1_full + 2;
    uint64_t* h_full = acc1_empty + 2;
    uint64_t* h_empty = h_full + 2;
    uint64_t* acc2_full = h_empty + 2;
    uint64_t* acc2_empty = acc2_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc2_empty + 2);
    float* vec_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + MP_BAR_BYTES);   // b1 [HD] | b2 [C] | gamma [C] | beta [C]
    float2* sums_s = reinterpret_cast<float2*>(vec_s + Cfg::VEC_FLOATS);           // [2][MP_E2_WARPS][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_mb = (M + MP_BM - 1) / MP_BM;
    const int ntl = ((int)blockIdx.x < n_mb) ? (n_mb - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nchunks = ntl * NCHUNK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w1); tma_prefetch_desc(&map_w2);
        for (int s = 0; s < NA1; ++s) { mbar_init(&a1_full[s], 1); mbar_init(&a1_empty[s], 1); }
        for (int s = 0; s < S1; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
        for (int s = 0; s < S2; ++s) { mbar_init(&w2_full[s], 1); mbar_init(&w2_empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc1_full[b], 1); mbar_init(&acc1_empty[b], MP_E1G);
            mbar_init(&h_full[b], MP_E1G); mbar_init(&h_empty[b], 1);
            mbar_init(&acc2_full[b], 1); mbar_init(&acc2_empty[b], MP_E2_WARPS);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < Cfg::HD; i += MP_THREADS) vec_s[i] = b1 ? __ldg(b1 + i) : 0.0f;
    for (int i = threadIdx.x; i < C; i += MP_THREADS) {
        vec_s[Cfg::HD + i] = b2 ? __ldg(b2 + i) : 0.0f;
        vec_s[Cfg::HD + C + i] = gamma ? __ldg(gamma + i) : 1.0f;                     // gamma == NULL: y = x_new, no LayerNorm
        vec_s[Cfg::HD + 2 * C + i] = (gamma && beta) ? __ldg(beta + i) : 0.0f;
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t vec_ss = smem_u32(vec_s);

    // The three single-lane roles below keep running ring indices / parities instead of deriving them from the chunk number:
    // an issuer thread that spends ~350 dependent instructions per chunk on divisions and descriptor arithmetic is what the
    // whole CTA waits for (measured: the MMA warp never waited, and every epilogue warp waited for it).
    if (warp == 0 && lane == 0) {
        // ===================== TMA producer: A per tile, W1(g) and W2(g - 2) per chunk =====================
        uint32_t s1 = 0, p1 = 1, s2 = 0, p2 = 1, as = 0, ap = 1;                       // ring slot + parity of its "empty" barrier
        int j2 = -2;                                                                    // W2 runs two chunks behind W1
        auto load_w2 = [&]() {
            if (j2 >= 0) {
                mbar_wait_parked(&w2_empty[s2], p2);
                mbar_arrive_expect_tx(&w2_full[s2], Cfg::W2_STAGE);
                tma_load_2d(w2 + s2 * Cfg::W2_STAGE, &map_w2, &w2_full[s2], j2 * MP_HC, 0);
                if (++s2 == S2) { s2 = 0; p2 ^= 1u; }
            }
            if (++j2 == NCHUNK) j2 = 0;
        };
        int i0 = (int)blockIdx.x * MP_BM;
        for (int tl = 0; tl < ntl; ++tl, i0 += (int)gridDim.x * MP_BM) {
            mbar_wait_parked(&a1_empty[as], ap);
            mbar_arrive_expect_tx(&a1_full[as], Cfg::A1_BYTES);
#pragma unroll
            for (int kb = 0; kb < KB1; ++kb)
                tma_load_2d(a1 + as * Cfg::A1_BYTES + kb * MP_KB_TILE, &map_a, &a1_full[as], kb * MP_BK, i0);
            if (++as == NA1) { as = 0; ap ^= 1u; }
#pragma unroll 1
            for (int j = 0; j < NCHUNK; ++j) {
#pragma unroll
                for (int kb = 0; kb < KB1; ++kb) {
                    mbar_wait_parked(&w1_empty[s1], p1);
                    mbar_arrive_expect_tx(&w1_full[s1], Cfg::W1_STAGE);
                    tma_load_2d(w1 + s1 * Cfg::W1_STAGE, &map_w1, &w1_full[s1], kb * MP_BK, j * MP_HC);
                    if (++s1 == S1) { s1 = 0; p1 ^= 1u; }
                }
                load_w2();
            }
        }
        if (ntl) { load_w2(); load_w2(); }
    } else if (warp == 1 && lane == 0) {
        // ===================== GEMM1 issuer: acc1[g % 2] = n W1[g]^T =====================
        constexpr uint32_t idesc1 = make_idesc_f16(MP_BM, MP_HC, BF16);
        const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(a1)), w_desc0 = make_smem_desc_sw128(smem_u32(w1));
        uint32_t s1 = 0, f1 = 0, as = 0, af = 0, epar = 1;
        for (int tl = 0; tl < ntl; ++tl) {
            mbar_wait_parked(&a1_full[as], af);
            const uint64_t a_desc = a_desc0 + (uint64_t)((as * Cfg::A1_BYTES) >> 4);
#pragma unroll 1
            for (int j = 0; j < NCHUNK; j += 2) {
#pragma unroll
                for (int buf = 0; buf < 2; ++buf) {                                    // NCHUNK is even: chunk j + buf -> acc1[buf]
                    mbar_wait_parked(&acc1_empty[buf], epar);                                 // its epilogue-1 group has read chunk g - 2
                    tc_fence_after();
#pragma unroll
                    for (int kb = 0; kb < KB1; ++kb) {
                        mbar_wait_parked(&w1_full[s1], f1);
                        tc_fence_after();
                        const uint64_t ad = a_desc + (uint64_t)((kb * MP_KB_TILE) >> 4);
                        const uint64_t wd = w_desc0 + (uint64_t)((s1 * Cfg::W1_STAGE) >> 4);
#pragma unroll
                        for (int k = 0; k < MP_BK / 16; ++k)
                            if (kb * MP_BK + k * 16 < C)                               // K steps past C are TMA zero fill on both sides
                                umma_f16(tmem_base + buf * MP_HC, ad + (uint64_t)(k * 2), wd + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
                        umma_commit(&w1_empty[s1]);
                        if (++s1 == S1) { s1 = 0; f1 ^= 1u; }
                    }
                    umma_commit(&acc1_full[buf]);
                }
                epar ^= 1u;
            }
            umma_commit(&a1_empty[as]);                                                // the tile's A operand is free again
            if (++as == NA1) { as = 0; af ^= 1u; }
        }
    } else if (warp == 3 && lane == 0) {
        // ===================== GEMM2 issuer: acc2[tile % 2] += H(g) W2[:, g]^T =====================
        constexpr uint32_t idesc2 = make_idesc_f16(MP_BM, C, BF16);
        const uint64_t h_desc0 = make_smem_desc_sw128(smem_u32(hbuf)), w_desc0 = make_smem_desc_sw128(smem_u32(w2));
        uint32_t s2 = 0, f2 = 0, hpar = 0, cpar = 1;
        for (int tl = 0; tl < ntl; ++tl) {
            const uint32_t ab = (uint32_t)tl & 1u;
            const uint32_t d_tmem = tmem_base + Cfg::ACC2_COL + ab * C;
#pragma unroll 1
            for (int j = 0; j < NCHUNK; j += 2) {
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    mbar_wait_parked(&h_full[hb], hpar);                                      // epilogue 1 has written H of this chunk
                    if (j == 0 && hb == 0) mbar_wait_parked(&acc2_empty[ab], cpar);           // epilogue 2 has drained this accumulator
                    mbar_wait_parked(&w2_full[s2], f2);
                    tc_fence_after();
                    const uint64_t ad = h_desc0 + (uint64_t)((hb * MP_KB_TILE) >> 4);
                    const uint64_t wd = w_desc0 + (uint64_t)((s2 * Cfg::W2_STAGE) >> 4);
#pragma unroll
                    for (int k = 0; k < MP_BK / 16; ++k) {
                        if constexpr (Cfg::H_TMEM)
                            umma_f16_ts(d_tmem, tmem_base + Cfg::H_COL + hb * (MP_HC / 2) + k * 8, wd + (uint64_t)(k * 2), idesc2, (j | hb | k) != 0);
                        else
                            umma_f16(d_tmem, ad + (uint64_t)(k * 2), wd + (uint64_t)(k * 2), idesc2, (j | hb | k) != 0);
                    }
                    umma_commit(&w2_empty[s2]);
                    umma_commit(&h_empty[hb]);                                         // H[hb] may be overwritten once these retire
                    if (++s2 == S2) { s2 = 0; f2 ^= 1u; }
                }
                hpar ^= 1u;
            }
            umma_commit(&acc2_full[ab]);
            if (ab) cpar ^= 1u;
        }
    } else if (warp >= 4 && warp < 4 + MP_E1_WARPS) {
        // ===================== epilogue 1: hidden chunk -> + b1 -> GELU -> 16-bit -> H k-block =====================
        const int e = warp - 4, q = e & 3, grp = (e >> 2) & 1, part = e >> 3;          // (warp % 4) == q: TMEM lane quarter
        const uint32_t lane_t = (uint32_t)(q * 32) << 16;
        const int row = q * 32 + lane;
        const uint32_t h_row = smem_u32(hbuf) + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t buf = (uint32_t)grp;
        uint32_t ph = 0;
        int j = grp;                                                                   // chunk index inside the tile
#pragma unroll 1
        for (int g = grp; g < nchunks; g += 2, ph ^= 1u, j = (j + 2 == NCHUNK + grp) ? grp : j + 2) {   // acc1[grp] -> H[grp]
            mbar_wait_parked(&acc1_full[buf], ph);
            tc_fence_after();
            float v[MP_E1C];
            tmem_ldn(tmem_base + lane_t + buf * MP_HC + (uint32_t)(part * MP_E1C), v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc1_empty[buf]);                              // the accumulator is in registers
            const uint32_t bb_s = vec_ss + (uint32_t)((j * MP_HC + part * MP_E1C) * 4);
            uint32_t pk[MP_E1C / 2];
#pragma unroll
            for (int p4 = 0; p4 < MP_E1C / 8; ++p4) {                                  // 16-byte pieces of this row
                const float4 b_lo = lds_const_f4(bb_s + p4 * 32), b_hi = lds_const_f4(bb_s + p4 * 32 + 16);
                const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int c = p4 * 8 + 2 * t;
                    const float2 g2 = gelu_erf2(make_float2(v[c] + bb[2 * t], v[c + 1] + bb[2 * t + 1]));
                    if (BF16) { const __nv_bfloat162 hh = __floats2bfloat162_rn(g2.x, g2.y); pk[p4 * 4 + t] = *reinterpret_cast<const uint32_t*>(&hh); }
                    else { const __half2 hh = __floats2half2_rn(g2.x, g2.y); pk[p4 * 4 + t] = *reinterpret_cast<const uint32_t*>(&hh); }
                }
            }
            mbar_wait_parked(&h_empty[buf], ph ^ 1u);            // GEMM2 of this group's previous chunk has read H[buf] (hidden by the math)
            if constexpr (Cfg::H_TMEM) {
                static_assert(MP_E1C == 32, "one 16-column packed store per warp and chunk");
                tc_fence_after();
                tmem_st16u(tmem_base + lane_t + Cfg::H_COL + buf * (MP_HC / 2) + (uint32_t)(part * (MP_E1C / 2)), pk);
                tc_fence_before();
            } else {
#pragma unroll
                for (int p4 = 0; p4 < MP_E1C / 8; ++p4) {
                    const uint32_t piece = (uint32_t)(part * (MP_E1C / 8) + p4);
                    sts128(h_row + buf * MP_KB_TILE + ((piece ^ sw) << 4), pk[p4 * 4], pk[p4 * 4 + 1], pk[p4 * 4 + 2], pk[p4 * 4 + 3]);
                }
                fence_proxy_async();                                                   // H is read by the tensor core (async proxy)
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_full[buf]);
        }
    } else if (warp >= 4 + MP_E1_WARPS) {
        // ===================== epilogue 2: + b2 + residual -> x_new, LayerNorm =====================
        const int e = warp - 4 - MP_E1_WARPS, q = e & 3, half = e >> 2;                // (warp % 4) == q
        const uint32_t ring_s = smem_u32(e2_smem + e * MP_E2_SMEM), ys_s = ring_s + MP_XSLOTS * MP_XSLOT;
        const uint32_t sums_ss = smem_u32(sums_s);
        const uint32_t lane_t = (uint32_t)(q * 32) << 16;
        const int lr = lane >> 2, lp = lane & 3;
        const uint32_t own64 = (uint32_t)(lane * 64), sw64 = (uint32_t)((lane >> 1) & 3);   // this thread's row of an fp32 unit
        const uint32_t own32 = (uint32_t)(lane * 32), sw32 = (uint32_t)((lane >> 2) & 1);   //                  of a 16-bit unit
        const float invC = 1.0f / (float)C;
        const int nunits = ntl * UPW;
        uint32_t ycount = 0;
        auto row0_of = [&](int tl) { return ((int)blockIdx.x + tl * (int)gridDim.x) * MP_BM + q * 32; };
        auto prefetch = [&](int n) {                                                   // residual unit n -> ring slot n % 3
            if (n < nunits) {
                const int r0 = row0_of(n / UPW), col = half * (C / 2) + (n % UPW) * 16 + lp * 4;
                const uint32_t dst = ring_s + (uint32_t)((n % MP_XSLOTS) * MP_XSLOT);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = lr + 8 * i;
                    const bool ok = r0 + r < M;
                    cp_async16(dst + (uint32_t)(r * 64) + (((uint32_t)lp ^ (uint32_t)((r >> 1) & 3)) << 4),
                               res + (int64_t)(ok ? r0 + r : 0) * C + col, ok ? 16u : 0u);
                }
            }
            cp_async_commit();
        };
        if (lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_y); }
#pragma unroll
        for (int n = 0; n < MP_PF; ++n) prefetch(n);
#pragma unroll 1
        for (int tl = 0; tl < ntl; ++tl) {
            const int row0 = row0_of(tl);
            const uint32_t ab = (uint32_t)tl & 1u;
            const float shift = (row0 + lane < M) ? __ldg(res + (int64_t)(row0 + lane) * C) : 0.0f;
            mbar_wait_parked(&acc2_full[ab], ((uint32_t)tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_t + Cfg::ACC2_COL + ab * C + (uint32_t)(half * (C / 2));
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 1
            for (int k = 0; k < UPW; ++k) {
                const int n = tl * UPW + k;
                const uint32_t slot_s = ring_s + (uint32_t)((n % MP_XSLOTS) * MP_XSLOT);
                const int col0 = half * (C / 2) + k * 16;
                float v[16];
                tmem_ld16(taddr + (uint32_t)(k * 16), v);
                cp_async_wait<MP_PF - 1>();
                __syncwarp();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const uint4 rr = lds128(slot_s + own64 + (((uint32_t)jj ^ sw64) << 4));
                    const float4 bb = lds_f4(vec_ss + (uint32_t)((Cfg::HD + col0 + 4 * jj) * 4));
                    v[4 * jj] += __uint_as_float(rr.x) + bb.x; v[4 * jj + 1] += __uint_as_float(rr.y) + bb.y;
                    v[4 * jj + 2] += __uint_as_float(rr.z) + bb.z; v[4 * jj + 3] += __uint_as_float(rr.w) + bb.w;
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) { const float d = v[jj] - shift; s1 += d; s2 = fmaf(d, d, s2); }
                tmem_st16(taddr + (uint32_t)(k * 16), v);
                if (xnew) {                                // the new residual rows leave from the slot they arrived in: one TMA store
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        sts128(slot_s + own64 + (((uint32_t)jj ^ sw64) << 4), __float_as_uint(v[4 * jj]), __float_as_uint(v[4 * jj + 1]),
                               __float_as_uint(v[4 * jj + 2]), __float_as_uint(v[4 * jj + 3]));
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d(&map_x, slot_s, col0, row0); tma_store_commit(); }
                }
                // the slot the next prefetch lands in held unit n - 1: its store is older than the one just committed
                if (lane == 0) tma_store_wait_read<1>();
                __syncwarp();
                prefetch(n + MP_PF);
            }
            // the two halves of a row meet through shared memory (parity buffer, one 64-thread named barrier per quarter)
            const uint32_t sums_t = sums_ss + (uint32_t)((tl & 1) * MP_E2_WARPS * 32 * 8);
            asm volatile("st.shared.v2.f32 [%0], {%1,%2};" :: "r"(sums_t + (e * 32 + lane) * 8), "f"(s1), "f"(s2) : "memory");
            asm volatile("bar.sync %0, 64;" :: "r"(1 + q) : "memory");
            float2 o;
            asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(o.x), "=f"(o.y) : "r"(sums_t + ((e ^ 4) * 32 + lane) * 8) : "memory");
            const float m1 = (s1 + o.x) * invC;
            const float mean = gamma ? shift + m1 : 0.0f;
            const float rstd = gamma ? rsqrtf(fmaxf(fmaf(s2 + o.y, invC, -m1 * m1), 0.0f) + eps) : 1.0f;
#pragma unroll 1
            for (int k = 0; k < UPW; ++k, ++ycount) {
                const int col0 = half * (C / 2) + k * 16;
                const uint32_t yslot = ys_s + (ycount & 1u) * MP_YSLOT;
                float v[16];
                tmem_ld16(taddr + (uint32_t)(k * 16), v);
                uint32_t pk[8];
#pragma unroll
                for (int t = 0; t < 16; t += 4) {
                    const float4 gg = lds_f4(vec_ss + (uint32_t)((Cfg::HD + C + col0 + t) * 4));
                    const float4 be = lds_f4(vec_ss + (uint32_t)((Cfg::HD + 2 * C + col0 + t) * 4));
                    const float a0 = fmaf((v[t] - mean) * rstd, gg.x, be.x), a1v = fmaf((v[t + 1] - mean) * rstd, gg.y, be.y);
                    const float a2 = fmaf((v[t + 2] - mean) * rstd, gg.z, be.z), a3 = fmaf((v[t + 3] - mean) * rstd, gg.w, be.w);
                    if (BF16) {
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a0, a1v), h1 = __floats2bfloat162_rn(a2, a3);
                        pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h0); pk[t / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    } else {
                        const __half2 h0 = __floats2half2_rn(a0, a1v), h1 = __floats2half2_rn(a2, a3);
                        pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h0); pk[t / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    }
                }
                if (lane == 0) tma_store_wait_read<1>();   // the store that last read this slot is two groups back
                __syncwarp();
                sts128(yslot + own32 + ((0u ^ sw32) << 4), pk[0], pk[1], pk[2], pk[3]);
                sts128(yslot + own32 + ((1u ^ sw32) << 4), pk[4], pk[5], pk[6], pk[7]);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { tma_store_2d(&map_y, yslot, col0, row0); tma_store_commit(); }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc2_empty[ab]);
        }
        cp_async_wait<0>();
        if (lane == 0) tma_store_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int C, bool BF16>
static int mlp_launch(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* res,
                      const float* gamma, const float* beta, float* xnew, void* y, int64_t M, float eps, cudaStream_t st) {
    using Cfg = MpCfg<C>;
    static_assert(Cfg::SMEM <= 227 * 1024, "operands + rings + epilogue buffers must fit one CTA");
    CUtensorMap ma, mw1, mw2;
    const int dt = BF16 ? XP_BF16 : XP_F16;
    int rc;
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, strides[1] = {(uint64_t)C * 2};
        const uint32_t box[2] = {MP_BK, MP_BM};
        if ((rc = make_tensor_map(&ma, dt, 2, A, dims, strides, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)Cfg::HD}, strides[1] = {(uint64_t)C * 2};         // W1 (4C, C)
        const uint32_t box[2] = {MP_BK, MP_HC};
        if ((rc = make_tensor_map(&mw1, dt, 2, W1, dims, strides, box, 1))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)Cfg::HD, (uint64_t)C}, strides[1] = {(uint64_t)Cfg::HD * 2};   // W2 (C, 4C)
        const uint32_t box[2] = {MP_BK, (uint32_t)C};
        if ((rc = make_tensor_map(&mw2, dt, 2, W2, dims, strides, box, 1))) return rc;
    }
    CUtensorMap mx, my;
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, strides[1] = {(uint64_t)C * 2};             // y (M, C) 16-bit
        const uint32_t box[2] = {16, 32};
        if ((rc = make_tensor_map(&my, dt, 2, y, dims, strides, box, 3))) return rc;
        mx = my;
        if (xnew) {
            const uint64_t xstrides[1] = {(uint64_t)C * 4};                                               // x_new (M, C) fp32
            if ((rc = make_tensor_map(&mx, XP_F32, 2, xnew, dims, xstrides, box, 2))) return rc;
        }
    }
    auto kern = mlp_res_ln_tc_kernel<C, BF16>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int64_t n_mb = ceil_div(M, MP_BM);
    kern<<<(unsigned)(n_mb < num_sms() ? n_mb : num_sms()), MP_THREADS, Cfg::SMEM, st>>>(ma, mw1, mw2, mx, my, b1, b2, res, gamma, beta,
                                                                                        xnew, y, (int)M, eps);
    XP_LAUNCH_CHECK("mlp_res_ln_tc_kernel");
    return XP_OK;
}

}  // namespace xp

using namespace xp;

extern "C" int xp_mlp_res_ln(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* residual,
                             const float* gamma, const float* beta, float* x_new, void* y, int64_t M, int64_t C, int32_t dtype,
                             float eps, xp_stream_t stream) {
    XP_REQUIRE(A && W1 && W2 && residual && y, "xp_mlp_res_ln: NULL tensor pointer");
    XP_REQUIRE(dtype == XP_F16 || dtype == XP_BF16, "xp_mlp_res_ln: 16-bit inputs only (got dtype %d)", dtype);
    XP_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), "xp_mlp_res_ln: bad shape");
    XP_REQUIRE(C == 96 || C == 192, "xp_mlp_res_ln: C must be 96 or 192 (hidden 4C; got %lld)", (long long)C);
    for (const void* p : {A, W1, W2, (const void*)residual, (const void*)y, (const void*)x_new})
        XP_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "xp_mlp_res_ln: tensors must be 16-byte aligned");
    if (M == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool bf = dtype == XP_BF16;
    if (C == 96) return bf ? mlp_launch<96, true>(A, W1, b1, W2, b2, residual, gamma, beta, x_new, y, M, eps, st)
                           : mlp_launch<96, false>(A, W1, b1, W2, b2, residual, gamma, beta, x_new, y, M, eps, st);
    return bf ? mlp_launch<192, true>(A, W1, b1, W2, b2, residual, gamma, beta, x_new, y, M, eps, st)
              : mlp_launch<192, false>(A, W1, b1, W2, b2, residual, gamma, beta, x_new, y, M, eps, st);
}
