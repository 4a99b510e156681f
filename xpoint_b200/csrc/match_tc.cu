// match_tc.cu -- descriptor-similarity GEMM on 5th-gen tensor cores (tcgen05) with the row arg-min fused into
// the epilogue.  This is the only tensor-core kernel of the hot path (BASELINE north star); everything else is
// HBM-bound byte/fp32 work.
//
//   nn[i] = argmin_j ( |y_j|^2 - 2 <x_i, y_j> )          x_i: row of X (n_x x C), y_j: row of Y (n_y x C)
//
// Exactness: tcgen05 has no fp32 MMA, and plain TF32 flips ~1e-3 of the arg-mins.  The
// operands are therefore split  v = hi + lo  (hi = tf32(v), lo = tf32(v - hi), both exact fp32 values) and the
// product is accumulated as  hi.hi + hi.lo + lo.hi  in the fp32 TMEM accumulator ("3xTF32"): the dropped lo.lo
// term and the rounding of lo are ~2^-22 relative, i.e. fp32 summation noise (measured 3.9e-7 abs on unit
// descriptors, zero flipped arg-mins).  The similarity matrix never leaves the SM.
// When C % 64 == 0 the same three-product scheme runs on kind::f16 at twice the MMA rate ("3xFP16"): the tensors are
// scaled by a power of two so that their largest row norm sits below 2^14 (exact), hi = fp16(v), lo = fp16(v - hi) carry
// 22 mantissa bits between them, the products accumulate in the same fp32 TMEM accumulator and the epilogue undoes the
// scale; element magnitudes far below the largest lose only what is far below the accumulator's own rounding.
//
// CTA = 128 X-rows x all Y-rows (column tiles of 256), 8 warps, warp-specialised:
//   warp 0   TMA producer: per k-block (32 fp32 = 128 B, one 128B-swizzle atom) loads Xhi, Xlo [128 x 32] and
//            Yhi, Ylo [256 x 32] into a 2-stage ring (96 KiB / stage), mbarrier full/empty.
//   warp 1   MMA issuer (one elected lane): 4 k-steps x 3 products of tcgen05.mma.cta_group::1.kind::tf32
//            (M=128, N=256, K=8) per k-block into one of two 256-column TMEM accumulators;
//            tcgen05.commit releases the smem stage / publishes the accumulator.
//   warp 2   TMEM allocation (512 columns) and release.
//   warps 4-11 epilogue: thread = accumulator row, four warps per TMEM lane quarter take every fourth 32-column chunk;
//            tcgen05.ld 32 columns at a time, key = |y_j|^2 - 2 acc, running (min, argmin) in registers (ascending j,
//            strict <: lowest index wins ties), the partial minima of a row are merged through shared memory at the end.
// The column arg-min  nn_y[j] = argmin_i ( |x_i|^2 - 2 <x_i, y_j> )  comes out of the SAME accumulator tile (the
// similarity GEMM runs once, not twice): per column the 32 rows of a warp are reduced with two REDUX.MIN
// (order-preserving uint key, then the lowest row holding it), lane t keeps column t, and one 64-bit
// red.global.min of (key << 32 | row) per lane and 32-column chunk merges warps and row blocks.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "tcgen05.cuh"

namespace xp {

constexpr int TC_BM = 128, TC_BN = 256;
constexpr int TC_BK = 32;                                    // tf32 operands: fp32 elements per 128-byte swizzle row
constexpr int TC_BK16 = 64;                                  // fp16 operands
constexpr int TC_STAGES = 2;
constexpr int TC_X_TILE = TC_BM * 128, TC_Y_TILE = TC_BN * 128;
constexpr int TC_STAGE_BYTES = 2 * TC_X_TILE + 2 * TC_Y_TILE;  // 96 KiB
constexpr int TC_EPI_WARPS = 16, TC_EPI_PARTS = TC_EPI_WARPS / 4, TC_THREADS = (4 + TC_EPI_WARPS) * 32;
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 2 * TC_BN * 4 + 1024 /*align*/ + 128 /*barriers + tmem ptr*/ + TC_EPI_PARTS * TC_BM * 8 /*row merge*/;

// ---------------------------------------------------------------------------------- hi / lo split
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;   // n % 4 == 0 (C % 32 == 0)
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    float h[4], l[4];
    const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(in[k]));
        h[k] = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(in[k] - h[k]));
        l[k] = __uint_as_float(lb);
    }
    *reinterpret_cast<float4*>(hi + i) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + i) = make_float4(l[0], l[1], l[2], l[3]);
}

// scale = 2^(14 - e) with sqrt(max row norm^2) < 2^e: every element of the scaled tensor is below 2^14 in magnitude
__device__ __forceinline__ float f16_scale(const unsigned* maxnorm2_bits) {
    const float m = sqrtf(__uint_as_float(*maxnorm2_bits));
    if (!(m > 0.0f) || !isfinite(m)) return 1.0f;
    return ldexpf(1.0f, 14 - (ilogbf(m) + 1));
}

__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                                        __half* __restrict__ lo, int64_t n, const unsigned* __restrict__ maxnorm2_bits) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;   // n % 4 == 0
    const float sc = f16_scale(maxnorm2_bits);
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    const float in[4] = {v.x * sc, v.y * sc, v.z * sc, v.w * sc};          // exact: power-of-two scale
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = __float2half_rn(in[k]);
        l[k] = __float2half_rn(in[k] - __half2float(h[k]));
    }
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l);
}

// ---------------------------------------------------------------------------------- fused GEMM + arg-min
template <bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
nn_argmin_tc_kernel(const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo,
                    const __grid_constant__ CUtensorMap map_yhi, const __grid_constant__ CUtensorMap map_ylo,
                    const float* __restrict__ ynorm, const int32_t* __restrict__ nx, const int32_t* __restrict__ ny,
                    int x_stride, int y_stride, int C, int32_t* __restrict__ nn, const float* __restrict__ xnorm,
                    unsigned long long* __restrict__ colkey, const unsigned* __restrict__ maxnorm2_bits) {
    constexpr int BK = F16 ? TC_BK16 : TC_BK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* yn_s = reinterpret_cast<float*>(base + TC_STAGES * TC_STAGE_BYTES);            // [2][TC_BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(yn_s + 2 * TC_BN);
    uint64_t* full = bars;                 // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;    // [TC_STAGES]
    uint64_t* tfull = bars + 2 * TC_STAGES;      // [2]
    uint64_t* tempty = bars + 2 * TC_STAGES + 2; // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
    float* mrg_key = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);     // [2][TC_BM] row arg-min of each column half
    int* mrg_idx = reinterpret_cast<int*>(mrg_key + TC_EPI_PARTS * TC_BM);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const int n_x = nx ? nx[pair] : x_stride, n_y = ny ? ny[pair] : y_stride;
    const int i0 = blockIdx.x * TC_BM;
    if (i0 >= n_x) return;                                   // uniform for the whole CTA
    const int n_tiles = (n_y + TC_BN - 1) / TC_BN;
    const int n_kb = C / BK;
    // 3xFP16: the accumulator holds sx * sy * <x, y>
    const float neg2 = F16 ? -2.0f / (f16_scale(maxnorm2_bits) * f16_scale(maxnorm2_bits + 1)) : -2.0f;
    const int xrow0 = pair * x_stride + i0, yrow0 = pair * y_stride;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_xhi); tma_prefetch_desc(&map_xlo); tma_prefetch_desc(&map_yhi); tma_prefetch_desc(&map_ylo);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], TC_EPI_WARPS); }
        fence_mbar_init();
        fence_proxy_async();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (n_tiles > 0) {
        if (warp == 0 && lane == 0) {
            // ===================== TMA producer =====================
            int it = 0;
            for (int jt = 0; jt < n_tiles; ++jt)
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(&empty[s], (uint32_t)(((it / TC_STAGES) & 1) ^ 1));
                    uint8_t* st = base + s * TC_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
                    tma_load_2d(st, &map_xhi, &full[s], kb * BK, xrow0);
                    tma_load_2d(st + TC_X_TILE, &map_xlo, &full[s], kb * BK, xrow0);
                    tma_load_2d(st + 2 * TC_X_TILE, &map_yhi, &full[s], kb * BK, yrow0 + jt * TC_BN);
                    tma_load_2d(st + 2 * TC_X_TILE + TC_Y_TILE, &map_ylo, &full[s], kb * BK, yrow0 + jt * TC_BN);
                }
        } else if (warp == 1 && lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = F16 ? make_idesc_f16(TC_BM, TC_BN, false) : make_idesc_tf32(TC_BM, TC_BN);
            int it = 0;
            for (int jt = 0; jt < n_tiles; ++jt) {
                const int buf = jt & 1;
                mbar_wait(&tempty[buf], (uint32_t)(((jt >> 1) & 1) ^ 1));   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * TC_BN);
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(&full[s], (uint32_t)((it / TC_STAGES) & 1));
                    tc_fence_after();
                    const uint32_t st = smem_u32(base + s * TC_STAGE_BYTES);
                    const uint64_t xhi = make_smem_desc_sw128(st), xlo = make_smem_desc_sw128(st + TC_X_TILE);
                    const uint64_t yhi = make_smem_desc_sw128(st + 2 * TC_X_TILE);
                    const uint64_t ylo = make_smem_desc_sw128(st + 2 * TC_X_TILE + TC_Y_TILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                            // 4 x 32 bytes of K per 128-byte swizzle row
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);      // advance K inside the swizzle atom (bytes >> 4)
                        if (F16) {                                           // K = 16 fp16 per MMA
                            umma_f16(d_tmem, xhi + adv, yhi + adv, idesc, (kb | k) != 0);
                            umma_f16(d_tmem, xhi + adv, ylo + adv, idesc, 1);
                            umma_f16(d_tmem, xlo + adv, yhi + adv, idesc, 1);
                        } else {                                             // K = 8 tf32 per MMA
                            umma_tf32(d_tmem, xhi + adv, yhi + adv, idesc, (kb | k) != 0);
                            umma_tf32(d_tmem, xhi + adv, ylo + adv, idesc, 1);
                            umma_tf32(d_tmem, xlo + adv, yhi + adv, idesc, 1);
                        }
                    }
                    umma_commit(&empty[s]);                                  // frees the smem stage when the MMAs retire
                }
                umma_commit(&tfull[buf]);                                    // accumulator complete
            }
        } else if (warp >= 4) {
            // ===================== epilogue: fused arg-min =====================
            const int q = warp & 3, part = (warp - 4) >> 2;                  // TMEM lane quarter of this warp; its half of the 32-column chunks
            const int row = q * 32 + lane;
            float best = INFINITY;
            int bestj = 0x7fffffff;
            const float* yn = ynorm + (int64_t)pair * y_stride;
            const bool row_valid = i0 + row < n_x;
            const float xn = (colkey && row_valid) ? xnorm[(int64_t)pair * x_stride + i0 + row] : INFINITY;   // +inf: never the min
            const unsigned my_row = row_valid ? (unsigned)(i0 + row) : 0x7fffffffu;
            unsigned long long* ck = colkey ? colkey + (int64_t)pair * y_stride : nullptr;
            for (int jt = 0; jt < n_tiles; ++jt) {
                const int buf = jt & 1;
                // stage |y_j|^2 of this column tile (the first 256 epilogue threads, one value each)
                const int et = threadIdx.x - 128;
                if (et < TC_BN) {
                    const int j = jt * TC_BN + et;
                    yn_s[buf * TC_BN + et] = j < n_y ? yn[j] : INFINITY;
                }
                asm volatile("bar.sync 1, %0;" :: "n"(TC_EPI_WARPS * 32) : "memory");
                mbar_wait(&tfull[buf], (uint32_t)((jt >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN);
#pragma unroll 1
                for (int c = part; c < TC_BN / 32; c += TC_EPI_PARTS) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const float key = fmaf(neg2, v[t], yn_s[buf * TC_BN + c * 32 + t]);    // +inf past n_y
                        if (key < best) { best = key; bestj = jt * TC_BN + c * 32 + t; }
                    }
                    if (ck) {
                        unsigned cm = 0xffffffffu, ci = 0x7fffffffu;
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const unsigned u = __float_as_uint(fmaf(neg2, v[t], xn));
                            const unsigned k = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);   // order-preserving float -> uint
                            const unsigned m = __reduce_min_sync(0xffffffffu, k);
                            const unsigned who = __reduce_min_sync(0xffffffffu, k == m ? my_row : 0x7fffffffu);
                            if (lane == t) { cm = m; ci = who; }
                        }
                        const int j = jt * TC_BN + c * 32 + lane;
                        if (j < n_y && ci != 0x7fffffffu) atomicMin(ck + j, ((unsigned long long)cm << 32) | ci);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[buf]);
            }
            // merge the column parts of every row: smaller key, then lower index (the rule of a single ascending scan)
            mrg_key[part * TC_BM + row] = best;
            mrg_idx[part * TC_BM + row] = bestj;
            asm volatile("bar.sync 1, %0;" :: "n"(TC_EPI_WARPS * 32) : "memory");
            if (part == 0 && i0 + row < n_x) {
#pragma unroll
                for (int pp = 1; pp < TC_EPI_PARTS; ++pp) {
                    const float k1 = mrg_key[pp * TC_BM + row];
                    const int j1 = mrg_idx[pp * TC_BM + row];
                    if (k1 < best || (k1 == best && j1 < bestj)) { best = k1; bestj = j1; }
                }
                nn[(int64_t)pair * x_stride + i0 + row] = bestj;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------- host
template <bool F16>
static int tc_direction(const void* Xhi, const void* Xlo, const void* Yhi, const void* Ylo, const float* ynorm,
                        const int32_t* nx, const int32_t* ny, int64_t P, int64_t x_stride, int64_t y_stride, int64_t C,
                        int32_t* nn, const float* xnorm, unsigned long long* colkey, const unsigned* maxnorm2_bits,
                        cudaStream_t st) {
    CUtensorMap mxh, mxl, myh, myl;
    const int es = F16 ? 2 : 4, dt = F16 ? XP_F16 : XP_F32;
    const uint64_t xdims[2] = {(uint64_t)C, (uint64_t)(P * x_stride)}, ydims[2] = {(uint64_t)C, (uint64_t)(P * y_stride)};
    const uint64_t strides[1] = {(uint64_t)C * es};
    const uint32_t bk = F16 ? TC_BK16 : TC_BK;
    const uint32_t xbox[2] = {bk, TC_BM}, ybox[2] = {bk, TC_BN};
    int rc;
    if ((rc = make_tensor_map(&mxh, dt, 2, Xhi, xdims, strides, xbox, 1))) return rc;
    if ((rc = make_tensor_map(&mxl, dt, 2, Xlo, xdims, strides, xbox, 1))) return rc;
    if ((rc = make_tensor_map(&myh, dt, 2, Yhi, ydims, strides, ybox, 1))) return rc;
    if ((rc = make_tensor_map(&myl, dt, 2, Ylo, ydims, strides, ybox, 1))) return rc;
    XP_CUDA_OK(cudaFuncSetAttribute(nn_argmin_tc_kernel<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    dim3 grid((unsigned)ceil_div(x_stride, TC_BM), (unsigned)P);
    nn_argmin_tc_kernel<F16><<<grid, TC_THREADS, TC_SMEM, st>>>(mxh, mxl, myh, myl, ynorm, nx, ny, (int)x_stride, (int)y_stride, (int)C, nn,
                                                         xnorm, colkey, maxnorm2_bits);
    XP_LAUNCH_CHECK("nn_argmin_tc_kernel");
    return XP_OK;
}

int64_t mnn_tc_workspace_bytes(int64_t P, int64_t x_stride, int64_t y_stride, int64_t C) {
    return 2 * (P * x_stride + P * y_stride) * C * 4 + P * y_stride * 8 + 1024 + 256;   // hi/lo copies + column keys + max norms
}

// column keys (key << 32 | row) -> nn_y; untouched columns (no valid row / column past n_y) stay -1
__global__ void __launch_bounds__(256) decode_colkey_kernel(const unsigned long long* __restrict__ colkey, int32_t* __restrict__ nn_y,
                                                            int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = colkey[i];
    nn_y[i] = k == ~0ull ? -1 : (int32_t)(k & 0xffffffffu);
}

// largest row norm^2 of a tensor (bits of a non-negative float order like unsigned integers)
__global__ void __launch_bounds__(256) max_norm2_kernel(const float* __restrict__ norm2, int64_t n, unsigned* __restrict__ out) {
    float m = 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = norm2[i];
        if (v > m && isfinite(v)) m = v;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

int mnn_argmin_tc(const float* X, const float* Y, const int32_t* nx, const int32_t* ny, int64_t P, int64_t x_stride,
                  int64_t y_stride, int64_t C, const float* xnorm, const float* ynorm, int32_t* nn_x, int32_t* nn_y,
                  void* split_ws, cudaStream_t st) {
    XP_REQUIRE(C % TC_BK == 0 && C >= TC_BK && C <= 1024, "xp_mnn_match (tensor cores): C must be a multiple of 32 in [32, 1024]");
    XP_REQUIRE(P * x_stride < (1LL << 31) && P * y_stride < (1LL << 31), "xp_mnn_match: too many descriptors");
    XP_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0,
               "xp_mnn_match: descriptor tensors must be 16-byte aligned");
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(split_ws) + 255) & ~uintptr_t(255));
    const int64_t nxe = P * x_stride * C, nye = P * y_stride * C;
    unsigned long long* colkey = reinterpret_cast<unsigned long long*>(ws + 2 * nxe + 2 * nye);
    unsigned* maxbits = reinterpret_cast<unsigned*>(colkey + P * y_stride);           // [2]: X, Y
    XP_CUDA_OK(cudaMemsetAsync(nn_x, 0xff, sizeof(int32_t) * P * x_stride, st));
    XP_CUDA_OK(cudaMemsetAsync(colkey, 0xff, sizeof(unsigned long long) * P * y_stride, st));
    static const bool force_tf32 = getenv("XP_MATCH_TF32") != nullptr;                 // testing / A-B knob
    int rc;
    if (C % TC_BK16 == 0 && !force_tf32) {
        // 3xFP16: per-tensor power-of-two scale from the largest row norm, hi / lo halves
        XP_CUDA_OK(cudaMemsetAsync(maxbits, 0, 2 * sizeof(unsigned), st));
        max_norm2_kernel<<<(unsigned)std::min<int64_t>(ceil_div(P * x_stride, 256), 1024), 256, 0, st>>>(xnorm, P * x_stride, maxbits);
        XP_LAUNCH_CHECK("max_norm2_kernel");
        max_norm2_kernel<<<(unsigned)std::min<int64_t>(ceil_div(P * y_stride, 256), 1024), 256, 0, st>>>(ynorm, P * y_stride, maxbits + 1);
        XP_LAUNCH_CHECK("max_norm2_kernel");
        __half* h = reinterpret_cast<__half*>(ws);
        __half *Xhi = h, *Xlo = h + nxe, *Yhi = h + 2 * nxe, *Ylo = h + 2 * nxe + nye;
        split_f16_kernel<<<(unsigned)ceil_div(nxe / 4, 256), 256, 0, st>>>(X, Xhi, Xlo, nxe, maxbits);
        XP_LAUNCH_CHECK("split_f16_kernel");
        split_f16_kernel<<<(unsigned)ceil_div(nye / 4, 256), 256, 0, st>>>(Y, Yhi, Ylo, nye, maxbits + 1);
        XP_LAUNCH_CHECK("split_f16_kernel");
        rc = tc_direction<true>(Xhi, Xlo, Yhi, Ylo, ynorm, nx, ny, P, x_stride, y_stride, C, nn_x, xnorm, colkey, maxbits, st);
    } else {
        float *Xhi = ws, *Xlo = ws + nxe, *Yhi = ws + 2 * nxe, *Ylo = ws + 2 * nxe + nye;
        split_tf32_kernel<<<(unsigned)ceil_div(nxe / 4, 256), 256, 0, st>>>(X, Xhi, Xlo, nxe);
        XP_LAUNCH_CHECK("split_tf32_kernel");
        split_tf32_kernel<<<(unsigned)ceil_div(nye / 4, 256), 256, 0, st>>>(Y, Yhi, Ylo, nye);
        XP_LAUNCH_CHECK("split_tf32_kernel");
        // one similarity GEMM: row arg-min in registers, column arg-min through REDUX + 64-bit atomic min
        rc = tc_direction<false>(Xhi, Xlo, Yhi, Ylo, ynorm, nx, ny, P, x_stride, y_stride, C, nn_x, xnorm, colkey, maxbits, st);
    }
    if (rc) return rc;
    const int64_t ncol = P * y_stride;
    if (ncol) {
        decode_colkey_kernel<<<(unsigned)ceil_div(ncol, 256), 256, 0, st>>>(colkey, nn_y, ncol);
        XP_LAUNCH_CHECK("decode_colkey_kernel");
    }
    return XP_OK;
}

}  // namespace xp
