// mlp_tc.cu -- the whole MLP branch of a VSSBlock in ONE tcgen05 kernel (xp_mlp_res_ln):
//
//     x_new = x + fc2( GELU( fc1(n) + b1 ) ) + b2 ;   y = LayerNorm_next(x_new)            (VMamba.py:110-128, 1229-1233)
//
// The 4C-wide hidden tensor (2 + 2 GB of HBM traffic per stage-0 block when fc1 and fc2 are separate kernels) never leaves
// the SM: per 128-row tile the hidden activation is produced 64 columns at a time, GELU'd on its way out of TMEM, written to
// shared memory as one K-major 128B-swizzled k-block of the second GEMM's A operand (double-buffered), and consumed there.
//
// Persistent CTAs over 128-row tiles, C = 96 or 192; 20 warps, each role with its own instruction stream so that the
// FMA-pipe-bound GELU (16 fp32 slots per value) and the latency-bound residual / LayerNorm epilogue overlap:
//   warp 0      TMA producer (one lane): the tile's A operand n [128 x C], then W1 / W2 k-block tiles in exactly the order the
//               MMA warp consumes them
//   warp 1      MMA issuer (one lane), one flat stream of 64-column hidden chunks g across tiles:
//               GEMM1(g): acc1[g % 2] = n W1[g]^T (K = C), then GEMM2(g - 1): acc2[tile % 2] += H(g - 1) W2[:, g - 1]^T (K = 64)
//               one chunk late, so GELU(g - 1) overlaps GEMM1(g) -- also across the tile boundary
//   warp 2      TMEM allocation: acc1 2 x 64 columns + acc2 2 x C columns (512 at C = 192)
//   warps 4-11  epilogue 1, thread = row, two warps (32-column halves) per TMEM lane quarter: tcgen05.ld -> + b1 -> exact GELU ->
//               16-bit -> H k-block in shared memory (fence.proxy.async before the MMA warp is told)
//   warps 12-19 epilogue 2, thread = row, two warps (C/2-column halves) per lane quarter: acc2 + b2 + residual -> one-pass
//               shifted sums -> x_new (fp32) and LayerNorm (16-bit) out.  The residual arrives by cp.async two 16-column units
//               ahead (across tiles) in a per-warp ring whose slots double as the transpose buffers of the coalesced stores.
#include "common.cuh"
#include "tcgen05.cuh"

namespace xp {

constexpr int MP_BM = 128, MP_BK = 64, MP_HC = 64;
constexpr int MP_KB_TILE = MP_BM * 128;                // one 64-wide k-block of a 128-row operand: 16 KiB
constexpr int MP_E1_WARPS = 8, MP_E2_WARPS = 8;
constexpr int MP_THREADS = (4 + MP_E1_WARPS + MP_E2_WARPS) * 32;
constexpr int MP_PF = 2;                               // residual units in flight per epilogue-2 warp
constexpr int MP_STG_PITCH = 80, MP_STG = 32 * MP_STG_PITCH;      // 32 rows x (64 B + 16 B pad): one fp32 unit
constexpr int MP_YSTG_PITCH = 48, MP_YSTG = 32 * MP_YSTG_PITCH;   // 32 rows x (32 B + 16 B pad): one 16-bit unit
constexpr int MP_E2_SMEM = MP_PF * MP_STG + MP_YSTG;
constexpr int MP_BAR_BYTES = 512;

template <int C> struct MpCfg {
    static constexpr int HD = 4 * C;
    static constexpr int NCHUNK = HD / MP_HC;
    static constexpr int KB1 = (C + MP_BK - 1) / MP_BK;            // k-blocks of GEMM1 (K = C; the tail is TMA zero fill)
    static constexpr int A1_BYTES = KB1 * MP_KB_TILE;
    static constexpr int NA1 = C == 96 ? 2 : 1;                    // A-operand buffers (what shared memory allows)
    static constexpr int W1_STAGE = MP_HC * 128, S1 = C == 96 ? 4 : 3;
    static constexpr int W2_STAGE = C * 128, S2 = C == 96 ? 3 : 2;
    static constexpr int ACC2_COL = 2 * MP_HC;
    static constexpr int UPW = C / 32;                             // 16-column units per epilogue-2 warp and tile
    static constexpr int VEC_FLOATS = HD + 3 * C;                  // b1 | b2 | gamma | beta
    static constexpr int SMEM = NA1 * A1_BYTES + S1 * W1_STAGE + 2 * MP_KB_TILE + S2 * W2_STAGE + 1024 /*align*/ + MP_BAR_BYTES
                                + MP_E2_WARPS * MP_E2_SMEM + VEC_FLOATS * 4 + 2 * MP_E2_WARPS * 32 * 8;
    static_assert(2 * MP_HC + 2 * C <= 512, "TMEM budget");
    static_assert((2 * NA1 + 2 * S1 + 2 * S2 + 12) * 8 + 4 <= MP_BAR_BYTES, "barrier block");
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {     // src_bytes 0: zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int C, bool BF16>
__global__ void __launch_bounds__(MP_THREADS, 1)
mlp_res_ln_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w1,
                     const __grid_constant__ CUtensorMap map_w2, const float* __restrict__ b1, const float* __restrict__ b2,
                     const float* __restrict__ res, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ xnew, void* __restrict__ yout, int M, float eps) {
    using Cfg = MpCfg<C>;
    constexpr int NCHUNK = Cfg::NCHUNK, KB1 = Cfg::KB1, S1 = Cfg::S1, S2 = Cfg::S2, NA1 = Cfg::NA1, UPW = Cfg::UPW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a1 = base;
    uint8_t* w1 = a1 + NA1 * Cfg::A1_BYTES;
    uint8_t* hbuf = w1 + S1 * Cfg::W1_STAGE;
    uint8_t* w2 = hbuf + 2 * MP_KB_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(w2 + S2 * Cfg::W2_STAGE);
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
    uint8_t* e2_smem = reinterpret_cast<uint8_t*>(bars) + MP_BAR_BYTES;
    float* vec_s = reinterpret_cast<float*>(e2_smem + MP_E2_WARPS * MP_E2_SMEM);   // b1 [HD] | b2 [C] | gamma [C] | beta [C]
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
            mbar_init(&acc1_full[b], 1); mbar_init(&acc1_empty[b], MP_E1_WARPS);
            mbar_init(&h_full[b], MP_E1_WARPS); mbar_init(&h_empty[b], 1);
            mbar_init(&acc2_full[b], 1); mbar_init(&acc2_empty[b], MP_E2_WARPS);
        }
        fence_mbar_init();
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < Cfg::HD; i += MP_THREADS) vec_s[i] = b1 ? __ldg(b1 + i) : 0.0f;
    for (int i = threadIdx.x; i < C; i += MP_THREADS) {
        vec_s[Cfg::HD + i] = b2 ? __ldg(b2 + i) : 0.0f;
        vec_s[Cfg::HD + C + i] = __ldg(gamma + i);
        vec_s[Cfg::HD + 2 * C + i] = __ldg(beta + i);
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer: loads in the MMA warp's consumption order =====================
        uint32_t i1 = 0, i2 = 0;
        auto load_w2 = [&](int g) {
            const uint32_t s = i2 % S2;
            mbar_wait(&w2_empty[s], ((i2 / S2) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&w2_full[s], Cfg::W2_STAGE);
            tma_load_2d(w2 + s * Cfg::W2_STAGE, &map_w2, &w2_full[s], (g % NCHUNK) * MP_HC, 0);
            ++i2;
        };
        for (int g = 0; g < nchunks; ++g) {
            const int tl = g / NCHUNK, j = g % NCHUNK;
            if (j == 0) {
                const int i0 = ((int)blockIdx.x + tl * (int)gridDim.x) * MP_BM;
                const uint32_t ab = (uint32_t)tl % NA1;
                mbar_wait(&a1_empty[ab], (((uint32_t)tl / NA1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&a1_full[ab], Cfg::A1_BYTES);
                for (int kb = 0; kb < KB1; ++kb)
                    tma_load_2d(a1 + ab * Cfg::A1_BYTES + kb * MP_KB_TILE, &map_a, &a1_full[ab], kb * MP_BK, i0);
            }
            for (int kb = 0; kb < KB1; ++kb, ++i1) {
                const uint32_t s = i1 % S1;
                mbar_wait(&w1_empty[s], ((i1 / S1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&w1_full[s], Cfg::W1_STAGE);
                tma_load_2d(w1 + s * Cfg::W1_STAGE, &map_w1, &w1_full[s], kb * MP_BK, j * MP_HC);
            }
            if (g >= 1) load_w2(g - 1);
        }
        if (nchunks) load_w2(nchunks - 1);
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = make_idesc_f16(MP_BM, MP_HC, BF16);
        constexpr uint32_t idesc2 = make_idesc_f16(MP_BM, C, BF16);
        uint32_t i1 = 0, i2 = 0;
        const uint32_t a1_s = smem_u32(a1), h_s = smem_u32(hbuf);
        auto gemm2 = [&](int g) {
            const int tl = g / NCHUNK, j = g % NCHUNK;
            const uint32_t hb = (uint32_t)g & 1u, ab = (uint32_t)tl & 1u;
            mbar_wait(&h_full[hb], ((uint32_t)g >> 1) & 1u);                          // epilogue 1 has written H of this chunk
            if (j == 0) mbar_wait(&acc2_empty[ab], (((uint32_t)tl >> 1) & 1u) ^ 1u);   // epilogue 2 has drained this accumulator
            const uint32_t s = i2 % S2;
            mbar_wait(&w2_full[s], (i2 / S2) & 1u);
            tc_fence_after();
            const uint64_t ad = make_smem_desc_sw128(h_s + hb * MP_KB_TILE);
            const uint64_t wd = make_smem_desc_sw128(smem_u32(w2 + s * Cfg::W2_STAGE));
#pragma unroll
            for (int k = 0; k < MP_BK / 16; ++k) {
                const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                umma_f16(tmem_base + Cfg::ACC2_COL + ab * C, ad + adv, wd + adv, idesc2, (j | k) != 0);
            }
            umma_commit(&w2_empty[s]);
            umma_commit(&h_empty[hb]);                                                 // H[hb] may be overwritten once these retire
            if (j == NCHUNK - 1) umma_commit(&acc2_full[ab]);
            ++i2;
        };
        for (int g = 0; g < nchunks; ++g) {
            const int tl = g / NCHUNK, j = g % NCHUNK;
            const uint32_t buf = (uint32_t)g & 1u, ab = (uint32_t)tl % NA1;
            if (j == 0) mbar_wait(&a1_full[ab], ((uint32_t)tl / NA1) & 1u);
            mbar_wait(&acc1_empty[buf], (((uint32_t)g >> 1) & 1u) ^ 1u);               // epilogue 1 of two chunks ago has read it
            tc_fence_after();
            for (int kb = 0; kb < KB1; ++kb, ++i1) {
                const uint32_t s = i1 % S1;
                mbar_wait(&w1_full[s], (i1 / S1) & 1u);
                tc_fence_after();
                const uint64_t ad = make_smem_desc_sw128(a1_s + ab * Cfg::A1_BYTES + kb * MP_KB_TILE);
                const uint64_t wd = make_smem_desc_sw128(smem_u32(w1 + s * Cfg::W1_STAGE));
#pragma unroll
                for (int k = 0; k < MP_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);
                    umma_f16(tmem_base + buf * MP_HC, ad + adv, wd + adv, idesc1, (kb | k) != 0);
                }
                umma_commit(&w1_empty[s]);
            }
            umma_commit(&acc1_full[buf]);
            if (j == NCHUNK - 1) umma_commit(&a1_empty[ab]);                            // the tile's A operand is free again
            if (g >= 1) gemm2(g - 1);
        }
        if (nchunks) gemm2(nchunks - 1);
    } else if (warp >= 4 && warp < 4 + MP_E1_WARPS) {
        // ===================== epilogue 1: hidden chunk -> + b1 -> GELU -> 16-bit -> H k-block =====================
        const int e = warp - 4, q = e & 3, part = e >> 2;                              // (warp % 4) == q: TMEM lane quarter
        const uint32_t lane_t = (uint32_t)(q * 32) << 16;
        const int row = q * 32 + lane;
        const uint32_t h_row = smem_u32(hbuf) + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll 1
        for (int g = 0; g < nchunks; ++g) {
            const uint32_t buf = (uint32_t)g & 1u, ph = ((uint32_t)g >> 1) & 1u;
            mbar_wait(&acc1_full[buf], ph);
            tc_fence_after();
            float v[32];
            tmem_ld32(tmem_base + lane_t + buf * MP_HC + (uint32_t)(part * 32), v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc1_empty[buf]);                              // the accumulator is in registers
            const float* bb = vec_s + (g % NCHUNK) * MP_HC + part * 32;
            mbar_wait(&h_empty[buf], ph ^ 1u);                                         // GEMM2 of two chunks ago has read H[buf]
#pragma unroll
            for (int p4 = 0; p4 < 4; ++p4) {                                           // 16-byte pieces part * 4 + p4 of this row
                uint32_t pk[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int c = p4 * 8 + 2 * t;
                    const float2 g2 = gelu_erf2(make_float2(v[c] + bb[c], v[c + 1] + bb[c + 1]));
                    if (BF16) { const __nv_bfloat162 hh = __floats2bfloat162_rn(g2.x, g2.y); pk[t] = *reinterpret_cast<const uint32_t*>(&hh); }
                    else { const __half2 hh = __floats2half2_rn(g2.x, g2.y); pk[t] = *reinterpret_cast<const uint32_t*>(&hh); }
                }
                const uint32_t piece = (uint32_t)(part * 4 + p4);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};"
                             :: "r"(h_row + buf * MP_KB_TILE + ((piece ^ sw) << 4)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            }
            fence_proxy_async();                                                       // H is read by the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&h_full[buf]);
        }
    } else if (warp >= 4 + MP_E1_WARPS) {
        // ===================== epilogue 2: + b2 + residual -> x_new, LayerNorm =====================
        const int e = warp - 4 - MP_E1_WARPS, q = e & 3, half = e >> 2;                // (warp % 4) == q
        uint8_t* ring = e2_smem + e * MP_E2_SMEM;
        uint8_t* ystg = ring + MP_PF * MP_STG;
        const uint32_t ring_s = smem_u32(ring);
        const uint32_t lane_t = (uint32_t)(q * 32) << 16;
        const int lr = lane >> 2, lp = lane & 3;
        const float invC = 1.0f / (float)C;
        const int nunits = ntl * UPW;
        auto row0_of = [&](int tl) { return ((int)blockIdx.x + tl * (int)gridDim.x) * MP_BM + q * 32; };
        auto prefetch = [&](int n) {                                                   // residual unit n -> ring slot n % PF
            if (n < nunits) {
                const int r0 = row0_of(n / UPW), col = half * (C / 2) + (n % UPW) * 16 + lp * 4;
                const uint32_t dst = ring_s + (uint32_t)((n % MP_PF) * MP_STG + lp * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = lr + 8 * i;
                    const bool ok = r0 + r < M;
                    cp_async16(dst + r * MP_STG_PITCH, res + (int64_t)(ok ? r0 + r : 0) * C + col, ok ? 16u : 0u);
                }
            }
            cp_async_commit();
        };
#pragma unroll
        for (int n = 0; n < MP_PF; ++n) prefetch(n);
#pragma unroll 1
        for (int tl = 0; tl < ntl; ++tl) {
            const int row0 = row0_of(tl);
            const uint32_t ab = (uint32_t)tl & 1u;
            const float shift = (row0 + lane < M) ? __ldg(res + (int64_t)(row0 + lane) * C) : 0.0f;
            mbar_wait(&acc2_full[ab], ((uint32_t)tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + lane_t + Cfg::ACC2_COL + ab * C + (uint32_t)(half * (C / 2));
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll 1
            for (int k = 0; k < UPW; ++k) {
                const int n = tl * UPW + k;
                uint8_t* slot = ring + (n % MP_PF) * MP_STG;
                const uint32_t slot_s = ring_s + (uint32_t)((n % MP_PF) * MP_STG);
                const int col0 = half * (C / 2) + k * 16;
                float v[16];
                tmem_ld16(taddr + (uint32_t)(k * 16), v);
                cp_async_wait<MP_PF - 1>();
                __syncwarp();
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const uint4 rr = lds128(slot_s + lane * MP_STG_PITCH + jj * 16);
                    const float4 bb = *reinterpret_cast<const float4*>(vec_s + Cfg::HD + col0 + 4 * jj);
                    v[4 * jj] += __uint_as_float(rr.x) + bb.x; v[4 * jj + 1] += __uint_as_float(rr.y) + bb.y;
                    v[4 * jj + 2] += __uint_as_float(rr.z) + bb.z; v[4 * jj + 3] += __uint_as_float(rr.w) + bb.w;
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) { const float d = v[jj] - shift; s1 += d; s2 = fmaf(d, d, s2); }
                tmem_st16(taddr + (uint32_t)(k * 16), v);
                if (xnew) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        *reinterpret_cast<float4*>(slot + lane * MP_STG_PITCH + jj * 16) = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = lr + 8 * i;
                        const float4 val = *reinterpret_cast<const float4*>(slot + r * MP_STG_PITCH + lp * 16);
                        if (row0 + r < M) *reinterpret_cast<float4*>(xnew + (int64_t)(row0 + r) * C + col0 + lp * 4) = val;
                    }
                }
                __syncwarp();
                prefetch(n + MP_PF);
            }
            // the two halves of a row meet through shared memory (parity buffer, one 64-thread named barrier per quarter)
            float2* sums_t = sums_s + (tl & 1) * MP_E2_WARPS * 32;
            sums_t[e * 32 + lane] = make_float2(s1, s2);
            asm volatile("bar.sync %0, 64;" :: "r"(1 + q) : "memory");
            const float2 o = sums_t[(e ^ 4) * 32 + lane];
            const float m1 = (s1 + o.x) * invC;
            const float mean = shift + m1;
            const float rstd = rsqrtf(fmaxf(fmaf(s2 + o.y, invC, -m1 * m1), 0.0f) + eps);
#pragma unroll 1
            for (int k = 0; k < UPW; ++k) {
                const int col0 = half * (C / 2) + k * 16;
                float v[16];
                tmem_ld16(taddr + (uint32_t)(k * 16), v);
                uint32_t pk[8];
#pragma unroll
                for (int t = 0; t < 16; t += 4) {
                    const float4 gg = *reinterpret_cast<const float4*>(vec_s + Cfg::HD + C + col0 + t);
                    const float4 be = *reinterpret_cast<const float4*>(vec_s + Cfg::HD + 2 * C + col0 + t);
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
                *reinterpret_cast<uint4*>(ystg + lane * MP_YSTG_PITCH) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(ystg + lane * MP_YSTG_PITCH + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = (lane >> 1) + 16 * i, piece = lane & 1;
                    const uint4 val = *reinterpret_cast<const uint4*>(ystg + r * MP_YSTG_PITCH + piece * 16);
                    if (row0 + r < M)
                        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(yout) + (int64_t)(row0 + r) * C + col0 + piece * 8) = val;
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc2_empty[ab]);
        }
        cp_async_wait<0>();
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
    auto kern = mlp_res_ln_tc_kernel<C, BF16>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int64_t n_mb = ceil_div(M, MP_BM);
    kern<<<(unsigned)(n_mb < num_sms() ? n_mb : num_sms()), MP_THREADS, Cfg::SMEM, st>>>(ma, mw1, mw2, b1, b2, res, gamma, beta, xnew, y,
                                                                                        (int)M, eps);
    XP_LAUNCH_CHECK("mlp_res_ln_tc_kernel");
    return XP_OK;
}

}  // namespace xp

using namespace xp;

extern "C" int xp_mlp_res_ln(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, const float* residual,
                             const float* gamma, const float* beta, float* x_new, void* y, int64_t M, int64_t C, int32_t dtype,
                             float eps, xp_stream_t stream) {
    XP_REQUIRE(A && W1 && W2 && residual && gamma && beta && y, "xp_mlp_res_ln: NULL tensor pointer");
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
