// linear_tc.cu -- Linear (+ bias) (+ exact-erf GELU) on tcgen05 tensor cores: out = act(A W^T + b), 16-bit in/out, fp32
// accumulation in TMEM.  (the MLP of VSSBlock, VMamba.py:110-128,1229-1233.)
//
// The MLP's first Linear is followed by an exact GELU over a (tokens x 4C) tensor; as separate kernels that is one write
// and one read + write of the largest activation of the block (4.2 ms of a 46 ms step).  Here the activation is applied
// to the fp32 accumulator while it leaves TMEM, so the hidden tensor is written exactly once, already activated.
//
// Persistent CTAs (one per SM) walk (128-row block, BN-column tile) pairs, BN = 192 or 128; 16 warps, warp-specialised:
//   warp 0     TMA producer: per 64-wide k-block (128 B = one swizzle atom) A [128 x 64] and W [BN x 64] tiles into a
//              4-stage ring; K is zero-filled by TMA past its end (K % 64 != 0 costs idle MMA columns, no branches)
//   warp 1     MMA issuer (one lane): 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) per k-block into one
//              of two TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator
//   warp 2     TMEM allocation (512 columns)
//   warps 4-15 epilogue: warp e owns TMEM lane quarter e % 4 and every third 64-column chunk (e / 4): tcgen05.ld,
//              + bias (shared memory), GELU, pack to 16 bit; a 32 x 64 chunk is transposed through a padded per-warp
//              shared-memory buffer so that every store instruction writes whole 128-byte lines (4 rows x 128 B).
// GELU uses erf(x) = 1 - (1 + a1 x + .. + a6 x^6)^-16 (Abramowitz-Stegun 7.1.28, |err| <= 3e-7, far below the 16-bit
// output quantum) evaluated on packed fp32 pairs instead of libdevice erff: the epilogue keeps up with the store stream.
#include "common.cuh"
#include "tcgen05.cuh"

namespace xp {

constexpr int LT_BM = 128, LT_BK = 64;
constexpr int LT_A_TILE = LT_BM * 128;
constexpr int LT_EPI_WARPS = 12, LT_EPI_PARTS = LT_EPI_WARPS / 4, LT_THREADS = (4 + LT_EPI_WARPS) * 32;
constexpr int LT_STG_PITCH = 144, LT_STG = 32 * LT_STG_PITCH;  // per-warp transpose buffer: 32 rows x (128 B + 16 B pad)

template <int BN, int LT_STAGES> struct LtCfg {
    static constexpr int W_TILE = BN * 128;
    static constexpr int STAGE = LT_A_TILE + W_TILE;
    static constexpr int SMEM = LT_STAGES * STAGE + 1024 /*align*/ + 256 /*barriers + tmem ptr*/ + LT_EPI_WARPS * (LT_STG + 256 /*bias of the chunk*/);
};

template <int BN, bool GELU, bool BF16, int LT_STAGES>
__global__ void __launch_bounds__(LT_THREADS, 1)
linear_act_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     const float* __restrict__ bias, void* __restrict__ out, int M, int N, int K, int nsplit) {
    using Cfg = LtCfg<BN, LT_STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + LT_STAGES * Cfg::STAGE);
    uint64_t* full = bars;                         // [LT_STAGES]
    uint64_t* empty = bars + LT_STAGES;            // [LT_STAGES]
    uint64_t* tfull = bars + 2 * LT_STAGES;        // [2]
    uint64_t* tempty = bars + 2 * LT_STAGES + 2;   // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * LT_STAGES + 4);
    uint8_t* stg_all = base + LT_STAGES * Cfg::STAGE + 256;                          // [LT_EPI_WARPS][LT_STG]
    float* bias_all = reinterpret_cast<float*>(stg_all + LT_EPI_WARPS * LT_STG);    // [LT_EPI_WARPS][64]: bias of the warp's current chunk

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (N + BN - 1) / BN;
    const int n_kb = (K + LT_BK - 1) / LT_BK;
    const int n_mb = (M + LT_BM - 1) / LT_BM;
    // nsplit CTAs share a row block and take every nsplit-th column tile (small M: fewer row blocks than SMs); nsplit = 1 otherwise
    const int vb = (int)blockIdx.x / nsplit, js = (int)blockIdx.x % nsplit, vgrid = (int)gridDim.x / nsplit;
    const int my_mb = (vb < n_mb) ? (n_mb - 1 - vb) / vgrid + 1 : 0;                 // row blocks of this CTA
    const int my_nt = (n_tiles - js + nsplit - 1) / nsplit;                           // column tiles of this CTA per row block
    const int ntl = my_mb * my_nt;                     // output tiles of this CTA: tl -> (row block, column tile)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w);
        for (int s = 0; s < LT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], LT_EPI_WARPS); }
        fence_mbar_init();
        fence_proxy_async();
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
            const int i0 = (vb + (tl / my_nt) * vgrid) * LT_BM, jt = js + (tl % my_nt) * nsplit;
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const int s = it % LT_STAGES;
                mbar_wait(&empty[s], (uint32_t)(((it / LT_STAGES) & 1) ^ 1));
                uint8_t* st = base + s * Cfg::STAGE;
                mbar_arrive_expect_tx(&full[s], Cfg::STAGE);
                tma_load_2d(st, &map_a, &full[s], kb * LT_BK, i0);
                tma_load_2d(st + LT_A_TILE, &map_w, &full[s], kb * LT_BK, jt * BN);
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_f16(LT_BM, BN, BF16);
        int it = 0;
        for (int tl = 0; tl < ntl; ++tl) {
            const int buf = tl & 1;
            mbar_wait(&tempty[buf], (uint32_t)(((tl >> 1) & 1) ^ 1));       // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const int s = it % LT_STAGES;
                mbar_wait(&full[s], (uint32_t)((it / LT_STAGES) & 1));
                tc_fence_after();
                const uint32_t st = smem_u32(base + s * Cfg::STAGE);
                const uint64_t ad = make_smem_desc_sw128(st), wd = make_smem_desc_sw128(st + LT_A_TILE);
#pragma unroll
                for (int k = 0; k < LT_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);       // advance K inside the swizzle atom (bytes >> 4)
                    umma_f16(d_tmem, ad + adv, wd + adv, idesc, (kb | k) != 0);
                }
                umma_commit(&empty[s]);                                       // frees the smem stage when the MMAs retire
            }
            umma_commit(&tfull[buf]);                                         // accumulator complete
        }
    } else if (warp >= 4) {
        // ===================== epilogue: bias + GELU + store =====================
        const int e = warp - 4, q = e & 3, part = e >> 2;                    // (warp % 4) == q: the TMEM lane quarter it may read
        uint8_t* stg = stg_all + e * LT_STG;
        float* bias_s = bias_all + e * 64;
        for (int tl = 0; tl < ntl; ++tl) {
            const int buf = tl & 1, jt = js + (tl % my_nt) * nsplit;
            const int row0 = (vb + (tl / my_nt) * vgrid) * LT_BM + q * 32;                             // first row of this warp
            mbar_wait(&tfull[buf], (uint32_t)((tl >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);
#pragma unroll 1
            for (int c = part; c < BN / 64; c += LT_EPI_PARTS) {             // 64-column chunks: a row leaves as one 128-byte line
                const int n0 = jt * BN + c * 64;
                if (n0 >= N) break;                                           // N % 32 == 0 (host-checked)
                bias_s[lane] = (bias && n0 + lane < N) ? __ldg(bias + n0 + lane) : 0.0f;
                bias_s[32 + lane] = (bias && n0 + 32 + lane < N) ? __ldg(bias + n0 + 32 + lane) : 0.0f;
                __syncwarp();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int nh = n0 + hh * 32;
                    if (nh >= N) break;
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(c * 64 + hh * 32), v);
                    uint32_t pk[16];
#pragma unroll
                    for (int t = 0; t < 32; t += 2) {
                        float2 ab = make_float2(v[t] + bias_s[hh * 32 + t], v[t + 1] + bias_s[hh * 32 + t + 1]);
                        if (GELU) ab = gelu_erf2(ab);
                        if (BF16) { const __nv_bfloat162 h = __floats2bfloat162_rn(ab.x, ab.y); pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h); }
                        else { const __half2 h = __floats2half2_rn(ab.x, ab.y); pk[t / 2] = *reinterpret_cast<const uint32_t*>(&h); }
                    }
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4)
                        *reinterpret_cast<uint4*>(stg + lane * LT_STG_PITCH + hh * 64 + s4 * 16) =
                            make_uint4(pk[4 * s4], pk[4 * s4 + 1], pk[4 * s4 + 2], pk[4 * s4 + 3]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {                                  // lane -> (row lane/8 + 4 i, 16-byte piece lane % 8)
                    const int r = (lane >> 3) + 4 * i, piece = lane & 7;
                    const uint4 val = *reinterpret_cast<const uint4*>(stg + r * LT_STG_PITCH + piece * 16);
                    if (row0 + r < M && n0 + piece * 8 < N)
                        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out) + (int64_t)(row0 + r) * N + n0 + piece * 8) = val;
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

template <int BN, bool GELU, bool BF16, int LT_STAGES>
static int linear_launch_s(const void* A, const void* W, const float* bias, void* out, int64_t M, int64_t N, int64_t K,
                           cudaStream_t st) {
    using Cfg = LtCfg<BN, LT_STAGES>;
    CUtensorMap ma, mw;
    const int dt = BF16 ? XP_BF16 : XP_F16;
    const uint64_t adims[2] = {(uint64_t)K, (uint64_t)M}, wdims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t abox[2] = {LT_BK, LT_BM}, wbox[2] = {LT_BK, BN};
    int rc;
    if ((rc = make_tensor_map(&ma, dt, 2, A, adims, strides, abox, 1))) return rc;
    if ((rc = make_tensor_map(&mw, dt, 2, W, wdims, strides, wbox, 1))) return rc;
    const int smem = Cfg::SMEM;
    auto kern = linear_act_tc_kernel<BN, GELU, BF16, LT_STAGES>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t n_mb = ceil_div(M, LT_BM), n_tiles = ceil_div(N, BN);
    int64_t nsplit = 1, grid = num_sms();
    if (n_mb < num_sms()) {                        // few row blocks (single-pair latency): spread the column tiles over the idle SMs
        nsplit = std::max<int64_t>(1, std::min<int64_t>(n_tiles, num_sms() / n_mb));
        grid = n_mb * nsplit;
    }
    kern<<<(unsigned)grid, LT_THREADS, smem, st>>>(ma, mw, bias, out, (int)M, (int)N, (int)K, (int)nsplit);
    XP_LAUNCH_CHECK("linear_act_tc_kernel");
    return XP_OK;
}

template <int BN, bool GELU, bool BF16>
static int linear_launch(const void* A, const void* W, const float* bias, void* out, int64_t M, int64_t N, int64_t K,
                         cudaStream_t st) {
    static_assert(LtCfg<BN, 4>::SMEM <= 227 * 1024, "four stages + epilogue buffers must fit one CTA");
    return linear_launch_s<BN, GELU, BF16, 4>(A, W, bias, out, M, N, K, st);
}

}  // namespace xp

using namespace xp;

extern "C" int xp_linear_act(const void* A, const void* W, const float* bias, void* out, int64_t M, int64_t N, int64_t K,
                             int32_t dtype, int32_t gelu, xp_stream_t stream) {
    XP_REQUIRE(A && W && out, "xp_linear_act: NULL tensor pointer");
    XP_REQUIRE(dtype == XP_F16 || dtype == XP_BF16, "xp_linear_act: 16-bit inputs only (got dtype %d)", dtype);
    XP_REQUIRE(M >= 0 && N > 0 && K > 0 && M < ((int64_t)1 << 31), "xp_linear_act: bad shape");
    XP_REQUIRE(K % 8 == 0 && N % 32 == 0 && N <= 8192, "xp_linear_act: need K %% 8 == 0, N %% 32 == 0, N <= 8192 (got K=%lld N=%lld)",
               (long long)K, (long long)N);
    for (const void* q : {A, W, (const void*)out})
        XP_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "xp_linear_act: tensors must be 16-byte aligned");
    if (M == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool bf = dtype == XP_BF16;
    const bool wide = N % 192 == 0;            // 192-column tiles divide 4C for every XPoint stage (384 .. 3072)
#define XP_LT(BN_, G_, B_) linear_launch<BN_, G_, B_>(A, W, bias, out, M, N, K, st)
    if (wide) {
        if (gelu) return bf ? XP_LT(192, true, true) : XP_LT(192, true, false);
        return bf ? XP_LT(192, false, true) : XP_LT(192, false, false);
    }
    if (gelu) return bf ? XP_LT(128, true, true) : XP_LT(128, true, false);
    return bf ? XP_LT(128, false, true) : XP_LT(128, false, false);
#undef XP_LT
}
