// linear_ln_tc.cu -- Linear + bias + residual add + LayerNorm in ONE tcgen05 GEMM (xp_linear_res_ln).
//
// Replaces, for the two projections that close a VSSBlock branch (SS2D.out_proj, VMamba.py:664; Mlp.fc2, :110-128), the
// sequence  pend = A W^T + b  (cuBLAS, 16-bit out)  ->  x = x + pend ; n = LayerNorm(x)  (xp_add_layer_norm) of the block
// wrapper x + branch(norm(x)) (VMamba.py:1222-1234): the 16-bit `pend` tensor is never written, the fp32 residual stream is
// read and written once, and the normalised copy that feeds the next GEMM leaves in the same pass.
//
// The whole output row must sit in one accumulator tile, so N = C is 96, 192 or 384 (preset E stages 0-2; wider rows keep
// the two-kernel path).  Persistent CTAs walk 128-row blocks; 8 warps:
//   warp 0     TMA producer: A [128 x 64] and W [N x 64] k-blocks (128B swizzle) into a 3/4-stage ring
//   warp 1     MMA issuer: tcgen05.mma.kind::f16 M128 x N (two N=192 instructions for N = 384) into TMEM; the accumulator is
//              double-buffered for N <= 256 (columns 0 / 256), single for N = 384
//   warp 2     TMEM allocation (512 columns)
//   warps 4-11 epilogue, THREAD = ROW (TMEM lane); the two warps of a lane quarter split every 32-column chunk into 16-column
//              halves.  Pass 1 per chunk: tcgen05.ld, + bias + residual (fetched one chunk ahead with coalesced 16-byte loads,
//              transposed through a padded per-warp shared-memory tile), shifted one-pass sums for mean / variance, the new
//              residual row goes back to TMEM (tcgen05.st) and out to HBM in fp32 through the same tile (whole 64-byte runs
//              per row); the two halves' sums meet in shared memory (one 64-thread named barrier per quarter); pass 2:
//              tcgen05.ld, normalise, gamma / beta, pack to 16 bit, coalesced stores.  No shuffles.
#include "common.cuh"
#include "tcgen05.cuh"

namespace xp {

constexpr int LN_BM = 128, LN_BK = 64;
constexpr int LN_A_TILE = LN_BM * 128;
constexpr int LN_EPI_WARPS = 8;
constexpr int LN_THREADS = (4 + LN_EPI_WARPS) * 32;
constexpr int LN_STG_PITCH = 80;                  // 16 fp32 + 16 B pad: conflict-free 16-byte row reads and transposed writes
constexpr int LN_STG = 32 * LN_STG_PITCH;

template <int N, int STAGES> struct LnCfg {
    static constexpr int W_TILE = N * 128;
    static constexpr int STAGE = LN_A_TILE + W_TILE;
    static constexpr int NBUF = N <= 256 ? 2 : 1;
    static constexpr int SMEM = STAGES * STAGE + 1024 /*align*/ + 256 /*barriers + tmem ptr*/ + LN_EPI_WARPS * LN_STG + 3 * N * 4
                                + 2 * LN_EPI_WARPS * 32 * 8 /*row sums, double-buffered by tile parity*/;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
           "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
           "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
           "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
           "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
           "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
           "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
           "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int N, bool BF16, int STAGES>
__global__ void __launch_bounds__(LN_THREADS, 1)
linear_res_ln_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                        const float* __restrict__ bias, const float* __restrict__ res, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ xnew, void* __restrict__ yout, int M, int K, float eps) {
    using Cfg = LnCfg<N, STAGES>;
    constexpr int NBUF = Cfg::NBUF;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + STAGES * Cfg::STAGE);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + STAGES;             // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;         // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;    // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint8_t* stg_all = base + STAGES * Cfg::STAGE + 256;                    // [LN_EPI_WARPS][LN_STG]
    float* vec_s = reinterpret_cast<float*>(stg_all + LN_EPI_WARPS * LN_STG);   // bias | gamma | beta, N floats each
    float2* sums_s = reinterpret_cast<float2*>(vec_s + 3 * N);             // [LN_EPI_WARPS][32] (s1, s2) of a row half

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_kb = (K + LN_BK - 1) / LN_BK;
    const int n_mb = (M + LN_BM - 1) / LN_BM;
    const int ntl = ((int)blockIdx.x < n_mb) ? (n_mb - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // row blocks of this CTA

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], LN_EPI_WARPS); }
        fence_mbar_init();
        fence_proxy_async();
    }
    for (int i = threadIdx.x; i < N; i += LN_THREADS) {
        vec_s[i] = bias ? __ldg(bias + i) : 0.0f;
        vec_s[N + i] = __ldg(gamma + i);
        vec_s[2 * N + i] = __ldg(beta + i);
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int it = 0;
        for (int tl = 0; tl < ntl; ++tl) {
            const int i0 = ((int)blockIdx.x + tl * (int)gridDim.x) * LN_BM;
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&empty[s], (uint32_t)(((it / STAGES) & 1) ^ 1));
                uint8_t* st = base + s * Cfg::STAGE;
                mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
                tma_load_2d(st, &map_a, &full[s], kb * LN_BK, i0);
                constexpr int WB = N <= 256 ? N : 192;                         // rows per W box (TMA boxes hold at most 256 rows)
#pragma unroll
                for (int h = 0; h < N / WB; ++h)
                    tma_load_2d(st + LN_A_TILE + h * WB * 128, &map_w, &full[s], kb * LN_BK, h * WB);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        constexpr int NI = N <= 256 ? N : 192;                               // columns per MMA instruction
        constexpr uint32_t idesc = make_idesc_f16(LN_BM, NI, BF16);
        int it = 0;
        for (int tl = 0; tl < ntl; ++tl) {
            const int buf = tl % NBUF;
            mbar_wait(&tempty[buf], (uint32_t)((((tl / NBUF) & 1)) ^ 1));      // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (uint32_t)((it / STAGES) & 1));
                tc_fence_after();
                const uint32_t st = smem_u32(base + s * Cfg::STAGE);
                const uint64_t ad = make_smem_desc_sw128(st);
#pragma unroll
                for (int k = 0; k < LN_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);       // advance K inside the swizzle atom (bytes >> 4)
#pragma unroll
                    for (int h = 0; h < N / NI; ++h) {
                        const uint64_t wd = make_smem_desc_sw128(st + LN_A_TILE + h * NI * 128);
                        umma_f16(d_tmem + (uint32_t)(h * NI), ad + adv, wd + adv, idesc, (kb | k) != 0);
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(&tfull[buf]);
        }
    } else if (warp >= 4) {
        // ===================== epilogue: thread = row, two warps (column halves) per TMEM lane quarter =====================
        const int e = warp - 4, q = e & 3, half = e >> 2;                    // (warp % 4) == q: the TMEM lane quarter it may read
        uint8_t* stg = stg_all + e * LN_STG;
        const uint32_t stg_s = smem_u32(stg);
        const float invN = 1.0f / (float)N;
        constexpr int NCH = N / 32;
        const int lr = lane >> 2, lp = lane & 3;                             // coalesced fp32 I/O: lane -> (row lr + 8i, 16-byte piece lp)
        // Global traffic goes through a padded per-warp shared-memory tile so that every instruction moves whole 64-byte runs
        // per row (lane -> (row lr + 8i, 16-byte piece lp)).  Measured alternative: every thread streaming its own row
        // directly (no staging, no __syncwarp) is 30 % slower (0.95 vs 0.73 ms at stage 0) -- the half-filled sectors cost
        // more than the STS / LDS chain.  The residual is fetched PF chunks ahead into registers, across tile boundaries
        // (the kernel is bound by bytes in flight: 1 -> 3 chunks ahead took stage 0 from 0.73 to 0.66 ms).
        constexpr int PF = NCH < 3 ? NCH : 3;                                 // residual chunks in flight per warp (registers)
        static_assert(NCH % PF == 0, "the prefetch ring must line up across tiles");
        float4 pre[PF][4];
        auto tile_row0 = [&](int tl) { return ((int)blockIdx.x + tl * (int)gridDim.x) * LN_BM + q * 32; };
        auto fetch = [&](int r0, int c, int slot) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = lr + 8 * i;
                pre[slot][i] = (r0 + r < M) ? __ldg(reinterpret_cast<const float4*>(res + (int64_t)(r0 + r) * N + half * 16 + lp * 4 + c * 32))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        if (ntl > 0) {
#pragma unroll
            for (int c = 0; c < PF; ++c) fetch(tile_row0(0), c, c);
        }
        for (int tl = 0; tl < ntl; ++tl) {
            const int buf = tl % NBUF;
            const int row0 = tile_row0(tl);                                   // first row of this warp
            // shift of the one-pass variance: the row's first residual value (same for both halves of the row)
            const float shift = (row0 + lane < M) ? __ldg(res + (int64_t)(row0 + lane) * N) : 0.0f;
            mbar_wait(&tfull[buf], (uint32_t)((tl / NBUF) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256 + half * 16);
            float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
#pragma unroll
                for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(stg + (lr + 8 * i) * LN_STG_PITCH + lp * 16) = pre[c % PF][i];
                if (c + PF < NCH) fetch(row0, c + PF, c % PF);
                else if (tl + 1 < ntl) fetch(tile_row0(tl + 1), c + PF - NCH, c % PF);     // the next tile's first chunks
                __syncwarp();
                float v[16];
                tmem_ld16(taddr + (uint32_t)(c * 32), v);
                const int col = c * 32 + half * 16;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 rr = lds128(stg_s + lane * LN_STG_PITCH + j * 16);
                    const float4 bb = *reinterpret_cast<const float4*>(vec_s + col + 4 * j);
                    v[4 * j] += __uint_as_float(rr.x) + bb.x;
                    v[4 * j + 1] += __uint_as_float(rr.y) + bb.y;
                    v[4 * j + 2] += __uint_as_float(rr.z) + bb.z;
                    v[4 * j + 3] += __uint_as_float(rr.w) + bb.w;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) { const float d = v[j] - shift; s1 += d; s2 = fmaf(d, d, s2); }
                tmem_st16(taddr + (uint32_t)(c * 32), v);                    // keep the new residual row for pass 2
                __syncwarp();                                                 // everyone has read its residual row
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(stg + lane * LN_STG_PITCH + j * 16) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                __syncwarp();
                if (xnew) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int r = lr + 8 * i;
                        const float4 val = *reinterpret_cast<const float4*>(stg + r * LN_STG_PITCH + lp * 16);
                        if (row0 + r < M) *reinterpret_cast<float4*>(xnew + (int64_t)(row0 + r) * N + c * 32 + half * 16 + lp * 4) = val;
                    }
                }
                __syncwarp();
            }
            // the two halves of a row meet: (s1, s2) through shared memory, one named barrier per lane quarter
            float2* sums_t = sums_s + (tl & 1) * LN_EPI_WARPS * 32;           // parity buffer: rewritten two tiles later, i.e. after
            sums_t[e * 32 + lane] = make_float2(s1, s2);                      // the partner has passed the NEXT tile's barrier
            asm volatile("bar.sync %0, 64;" :: "r"(1 + q) : "memory");
            const float2 other = sums_t[(e ^ 4) * 32 + lane];
            const float m1 = (s1 + other.x) * invN;
            const float mean = shift + m1;
            const float rstd = rsqrtf(fmaxf(fmaf(s2 + other.y, invN, -m1 * m1), 0.0f) + eps);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                float v[16];
                tmem_ld16(taddr + (uint32_t)(c * 32), v);
                const int col = c * 32 + half * 16;
                uint32_t pk[8];
#pragma unroll
                for (int t = 0; t < 16; t += 4) {
                    const float4 gg = *reinterpret_cast<const float4*>(vec_s + N + col + t);
                    const float4 be = *reinterpret_cast<const float4*>(vec_s + 2 * N + col + t);
                    const float a0 = fmaf((v[t] - mean) * rstd, gg.x, be.x), a1 = fmaf((v[t + 1] - mean) * rstd, gg.y, be.y);
                    const float a2 = fmaf((v[t + 2] - mean) * rstd, gg.z, be.z), a3 = fmaf((v[t + 3] - mean) * rstd, gg.w, be.w);
                    if (BF16) {
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(a0, a1), h1 = __floats2bfloat162_rn(a2, a3);
                        pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h0); pk[t / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    } else {
                        const __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3);
                        pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h0); pk[t / 2 + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    }
                }
                *reinterpret_cast<uint4*>(stg + lane * LN_STG_PITCH) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4*>(stg + lane * LN_STG_PITCH + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) {                                  // lane -> (row lane/2 + 16i, 16-byte piece lane % 2)
                    const int r = (lane >> 1) + 16 * i, piece = lane & 1;
                    const uint4 val = *reinterpret_cast<const uint4*>(stg + r * LN_STG_PITCH + piece * 16);
                    if (row0 + r < M)
                        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(yout) + (int64_t)(row0 + r) * N + col + piece * 8) = val;
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <int N, bool BF16, int STAGES>
static int linear_ln_launch(const void* A, const void* W, const float* bias, const float* res, const float* gamma, const float* beta,
                            float* xnew, void* y, int64_t M, int64_t K, float eps, cudaStream_t st) {
    using Cfg = LnCfg<N, STAGES>;
    static_assert(Cfg::SMEM <= 227 * 1024, "ring + epilogue buffers must fit one CTA");
    CUtensorMap ma, mw;
    const int dt = BF16 ? XP_BF16 : XP_F16;
    const uint64_t adims[2] = {(uint64_t)K, (uint64_t)M}, wdims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t abox[2] = {LN_BK, LN_BM};
    int rc;
    if ((rc = make_tensor_map(&ma, dt, 2, A, adims, strides, abox, 1))) return rc;
    // the W tile is fetched with boxes of at most 256 rows (TMA limit): N = 384 uses two boxes of 192 rows = one [N x 64] tile
    const uint32_t wbox[2] = {LN_BK, (uint32_t)(N <= 256 ? N : 192)};
    if ((rc = make_tensor_map(&mw, dt, 2, W, wdims, strides, wbox, 1))) return rc;
    auto kern = linear_res_ln_tc_kernel<N, BF16, STAGES>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int64_t n_mb = ceil_div(M, LN_BM);
    kern<<<(unsigned)(n_mb < num_sms() ? n_mb : num_sms()), LN_THREADS, Cfg::SMEM, st>>>(ma, mw, bias, res, gamma, beta, xnew, y, (int)M,
                                                                                        (int)K, eps);
    XP_LAUNCH_CHECK("linear_res_ln_tc_kernel");
    return XP_OK;
}

}  // namespace xp

using namespace xp;

extern "C" int xp_linear_res_ln(const void* A, const void* W, const float* bias, const float* residual, const float* gamma,
                                const float* beta, float* x_new, void* y, int64_t M, int64_t N, int64_t K, int32_t dtype, float eps,
                                xp_stream_t stream) {
    XP_REQUIRE(A && W && residual && gamma && beta && y, "xp_linear_res_ln: NULL tensor pointer");
    XP_REQUIRE(dtype == XP_F16 || dtype == XP_BF16, "xp_linear_res_ln: 16-bit inputs only (got dtype %d)", dtype);
    XP_REQUIRE(M >= 0 && K > 0 && M < ((int64_t)1 << 31), "xp_linear_res_ln: bad shape");
    XP_REQUIRE(N == 96 || N == 192 || N == 384, "xp_linear_res_ln: the output row must fit one accumulator tile: N in {96, 192, 384} (got %lld)",
               (long long)N);
    XP_REQUIRE(K % 8 == 0, "xp_linear_res_ln: need K %% 8 == 0 (got %lld)", (long long)K);
    for (const void* q : {A, W, (const void*)residual, (const void*)y, (const void*)x_new})
        XP_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "xp_linear_res_ln: tensors must be 16-byte aligned");
    if (M == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool bf = dtype == XP_BF16;
#define XP_LL(N_, B_, S_) linear_ln_launch<N_, B_, S_>(A, W, bias, residual, gamma, beta, x_new, y, M, K, eps, st)
    if (N == 96) return bf ? XP_LL(96, true, 4) : XP_LL(96, false, 4);
    if (N == 192) return bf ? XP_LL(192, true, 4) : XP_LL(192, false, 4);
    return bf ? XP_LL(384, true, 3) : XP_LL(384, false, 3);
#undef XP_LL
}
