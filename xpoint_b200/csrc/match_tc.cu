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
//
// CTA = 128 X-rows x all Y-rows (column tiles of 256), 8 warps, warp-specialised:
//   warp 0   TMA producer: per k-block (32 fp32 = 128 B, one 128B-swizzle atom) loads Xhi, Xlo [128 x 32] and
//            Yhi, Ylo [256 x 32] into a 2-stage ring (96 KiB / stage), mbarrier full/empty.
//   warp 1   MMA issuer (one elected lane): 4 k-steps x 3 products of tcgen05.mma.cta_group::1.kind::tf32
//            (M=128, N=256, K=8) per k-block into one of two 256-column TMEM accumulators;
//            tcgen05.commit releases the smem stage / publishes the accumulator.
//   warp 2   TMEM allocation (512 columns) and release.
//   warps 4-7 epilogue: thread = accumulator row; tcgen05.ld 32 columns at a time, key = |y_j|^2 - 2 acc,
//            running (min, argmin) in registers (ascending j, strict <: lowest index wins ties).
// The column arg-min  nn_y[j] = argmin_i ( |x_i|^2 - 2 <x_i, y_j> )  comes out of the SAME accumulator tile (the
// similarity GEMM runs once, not twice): per column the 32 rows of a warp are reduced with two REDUX.MIN
// (order-preserving uint key, then the lowest row holding it), lane t keeps column t, and one 64-bit
// red.global.min of (key << 32 | row) per lane and 32-column chunk merges warps and row blocks.
#include "common.cuh"
#include "tcgen05.cuh"

namespace xp {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 32;          // fp32 elements; BK * 4 B = one 128-byte swizzle row
constexpr int TC_STAGES = 2;
constexpr int TC_X_TILE = TC_BM * 128, TC_Y_TILE = TC_BN * 128;
constexpr int TC_STAGE_BYTES = 2 * TC_X_TILE + 2 * TC_Y_TILE;  // 96 KiB
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 2 * TC_BN * 4 + 1024 /*align*/ + 128 /*barriers + tmem ptr*/;

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

// ---------------------------------------------------------------------------------- fused GEMM + arg-min
__global__ void __launch_bounds__(256, 1)
nn_argmin_tc_kernel(const __grid_constant__ CUtensorMap map_xhi, const __grid_constant__ CUtensorMap map_xlo,
                    const __grid_constant__ CUtensorMap map_yhi, const __grid_constant__ CUtensorMap map_ylo,
                    const float* __restrict__ ynorm, const int32_t* __restrict__ nx, const int32_t* __restrict__ ny,
                    int x_stride, int y_stride, int C, int32_t* __restrict__ nn, const float* __restrict__ xnorm,
                    unsigned long long* __restrict__ colkey) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* yn_s = reinterpret_cast<float*>(base + TC_STAGES * TC_STAGE_BYTES);            // [2][TC_BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(yn_s + 2 * TC_BN);
    uint64_t* full = bars;                 // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;    // [TC_STAGES]
    uint64_t* tfull = bars + 2 * TC_STAGES;      // [2]
    uint64_t* tempty = bars + 2 * TC_STAGES + 2; // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = blockIdx.y;
    const int n_x = nx ? nx[pair] : x_stride, n_y = ny ? ny[pair] : y_stride;
    const int i0 = blockIdx.x * TC_BM;
    if (i0 >= n_x) return;                                   // uniform for the whole CTA
    const int n_tiles = (n_y + TC_BN - 1) / TC_BN;
    const int n_kb = C / TC_BK;
    const int xrow0 = pair * x_stride + i0, yrow0 = pair * y_stride;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_xhi); tma_prefetch_desc(&map_xlo); tma_prefetch_desc(&map_yhi); tma_prefetch_desc(&map_ylo);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
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
                    tma_load_2d(st, &map_xhi, &full[s], kb * TC_BK, xrow0);
                    tma_load_2d(st + TC_X_TILE, &map_xlo, &full[s], kb * TC_BK, xrow0);
                    tma_load_2d(st + 2 * TC_X_TILE, &map_yhi, &full[s], kb * TC_BK, yrow0 + jt * TC_BN);
                    tma_load_2d(st + 2 * TC_X_TILE + TC_Y_TILE, &map_ylo, &full[s], kb * TC_BK, yrow0 + jt * TC_BN);
                }
        } else if (warp == 1 && lane == 0) {
            // ===================== MMA issuer =====================
            constexpr uint32_t idesc = make_idesc_tf32(TC_BM, TC_BN);
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
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);   // advance K inside the swizzle atom (bytes >> 4)
                        umma_tf32(d_tmem, xhi + adv, yhi + adv, idesc, (kb | k) != 0);
                        umma_tf32(d_tmem, xhi + adv, ylo + adv, idesc, 1);
                        umma_tf32(d_tmem, xlo + adv, yhi + adv, idesc, 1);
                    }
                    umma_commit(&empty[s]);                                  // frees the smem stage when the MMAs retire
                }
                umma_commit(&tfull[buf]);                                    // accumulator complete
            }
        } else if (warp >= 4) {
            // ===================== epilogue: fused arg-min =====================
            const int q = warp & 3;                                          // TMEM lane quarter of this warp
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
                // stage |y_j|^2 of this column tile (128 epilogue threads x 2 values)
                const int et = threadIdx.x - 128;
#pragma unroll
                for (int t = 0; t < TC_BN / 128; ++t) {
                    const int j = jt * TC_BN + et + t * 128;
                    yn_s[buf * TC_BN + et + t * 128] = j < n_y ? yn[j] : INFINITY;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(&tfull[buf], (uint32_t)((jt >> 1) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN);
#pragma unroll 1
                for (int c = 0; c < TC_BN / 32; ++c) {
                    float v[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), v);
#pragma unroll
                    for (int t = 0; t < 32; ++t) {
                        const float key = fmaf(-2.0f, v[t], yn_s[buf * TC_BN + c * 32 + t]);   // +inf past n_y
                        if (key < best) { best = key; bestj = jt * TC_BN + c * 32 + t; }
                    }
                    if (ck) {
                        unsigned cm = 0xffffffffu, ci = 0x7fffffffu;
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const unsigned u = __float_as_uint(fmaf(-2.0f, v[t], xn));
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
            if (i0 + row < n_x) nn[(int64_t)pair * x_stride + i0 + row] = bestj;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------- host
static int tc_direction(const float* Xhi, const float* Xlo, const float* Yhi, const float* Ylo, const float* ynorm,
                        const int32_t* nx, const int32_t* ny, int64_t P, int64_t x_stride, int64_t y_stride, int64_t C,
                        int32_t* nn, const float* xnorm, unsigned long long* colkey, cudaStream_t st) {
    CUtensorMap mxh, mxl, myh, myl;
    const uint64_t xdims[2] = {(uint64_t)C, (uint64_t)(P * x_stride)}, ydims[2] = {(uint64_t)C, (uint64_t)(P * y_stride)};
    const uint64_t strides[1] = {(uint64_t)C * 4};
    const uint32_t xbox[2] = {TC_BK, TC_BM}, ybox[2] = {TC_BK, TC_BN};
    int rc;
    if ((rc = make_tensor_map(&mxh, XP_F32, 2, Xhi, xdims, strides, xbox, 1))) return rc;
    if ((rc = make_tensor_map(&mxl, XP_F32, 2, Xlo, xdims, strides, xbox, 1))) return rc;
    if ((rc = make_tensor_map(&myh, XP_F32, 2, Yhi, ydims, strides, ybox, 1))) return rc;
    if ((rc = make_tensor_map(&myl, XP_F32, 2, Ylo, ydims, strides, ybox, 1))) return rc;
    XP_CUDA_OK(cudaFuncSetAttribute(nn_argmin_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    dim3 grid((unsigned)ceil_div(x_stride, TC_BM), (unsigned)P);
    nn_argmin_tc_kernel<<<grid, 256, TC_SMEM, st>>>(mxh, mxl, myh, myl, ynorm, nx, ny, (int)x_stride, (int)y_stride, (int)C, nn,
                                                    xnorm, colkey);
    XP_LAUNCH_CHECK("nn_argmin_tc_kernel");
    return XP_OK;
}

int64_t mnn_tc_workspace_bytes(int64_t P, int64_t x_stride, int64_t y_stride, int64_t C) {
    return 2 * (P * x_stride + P * y_stride) * C * 4 + P * y_stride * 8 + 1024;   // hi/lo copies + column keys
}

// column keys (key << 32 | row) -> nn_y; untouched columns (no valid row / column past n_y) stay -1
__global__ void __launch_bounds__(256) decode_colkey_kernel(const unsigned long long* __restrict__ colkey, int32_t* __restrict__ nn_y,
                                                            int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = colkey[i];
    nn_y[i] = k == ~0ull ? -1 : (int32_t)(k & 0xffffffffu);
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
    float *Xhi = ws, *Xlo = ws + nxe, *Yhi = ws + 2 * nxe, *Ylo = ws + 2 * nxe + nye;
    split_tf32_kernel<<<(unsigned)ceil_div(nxe / 4, 256), 256, 0, st>>>(X, Xhi, Xlo, nxe);
    XP_LAUNCH_CHECK("split_tf32_kernel");
    split_tf32_kernel<<<(unsigned)ceil_div(nye / 4, 256), 256, 0, st>>>(Y, Yhi, Ylo, nye);
    XP_LAUNCH_CHECK("split_tf32_kernel");
    unsigned long long* colkey = reinterpret_cast<unsigned long long*>(ws + 2 * nxe + 2 * nye);
    XP_CUDA_OK(cudaMemsetAsync(nn_x, 0xff, sizeof(int32_t) * P * x_stride, st));
    XP_CUDA_OK(cudaMemsetAsync(colkey, 0xff, sizeof(unsigned long long) * P * y_stride, st));
    // one similarity GEMM: row arg-min in registers, column arg-min through REDUX + 64-bit atomic min
    int rc = tc_direction(Xhi, Xlo, Yhi, Ylo, ynorm, nx, ny, P, x_stride, y_stride, C, nn_x, xnorm, colkey, st);
    if (rc) return rc;
    const int64_t ncol = P * y_stride;
    if (ncol) {
        decode_colkey_kernel<<<(unsigned)ceil_div(ncol, 256), 256, 0, st>>>(colkey, nn_y, ncol);
        XP_LAUNCH_CHECK("decode_colkey_kernel");
    }
    return XP_OK;
}

}  // namespace xp
