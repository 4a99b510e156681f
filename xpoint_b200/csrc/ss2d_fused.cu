// ss2d_fused.cu -- the copy-free SS2D core around the selective scan (kernel K2/K3).
//
// The reference runs CrossScan (1 read + 4 writes of the activations, csm_triton.py:22-29), the scan, CrossMerge
// (4 reads + 1 write, csm_triton.py:56-62), a transpose and out_norm (VMamba.py:603-646).  Here the scan kernel
// itself routes the four directions (xp_scan_args.u_group_div / reverse_group_mask), so all that is left is
//   * xp_ss2d_pack:        x (B,D,H,W) -> xx (B,2,D,L) = [x ; x^T]: the two distinct token orders.  Directions
//                          l0 = h*W+w and L-1-l0 read xx[:,0] forwards / backwards, l1 = w*H+h and L-1-l1 read xx[:,1].
//   * xp_ss2d_dwconv_pack: the same with the depth-wise 3x3 convolution + SiLU in front (VMamba.py:651-655),
//                          reading the channel-last in_proj output directly.
//   * xp_ss2d_merge_norm:  y planes in NATURAL memory order [row-fwd, row-bwd, col-fwd, col-bwd] ->
//                          LayerNorm_D(y0 + y1 + (y2 + y3)^T) [* gate] in channel-last layout, ONE pass:
//                          16 B read + 2..4 B written per (token, channel), no intermediate buffer.
#include <stdlib.h>

#include "common.cuh"

namespace xp {

// ------------------------------------------------------------------------------------------ pack
// One CTA moves a 32x32 (h, w) tile of one (b, d) plane: straight copy + padded shared-memory transpose, every
// global access a full row of the tile.
template <typename T>
__global__ void __launch_bounds__(256) ss2d_pack_kernel(const T* __restrict__ x, T* __restrict__ xx, int64_t D, int H, int W,
                                                        int tiles_w, int tiles_h) {
    __shared__ T tile[32][33];
    const int64_t L = (int64_t)H * W;
    int64_t t = blockIdx.x;
    const int tw = (int)(t % tiles_w); t /= tiles_w;
    const int th = (int)(t % tiles_h); t /= tiles_h;
    const int64_t bd = t;                       // b * D + d
    const int64_t b = bd / D, d = bd % D;
    const int h0 = th * 32, w0 = tw * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* src = x + bd * L;
    T* d0 = xx + ((b * 2 + 0) * D + d) * L;
    T* d1 = xx + ((b * 2 + 1) * D + d) * L;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const T v = src[(int64_t)h * W + w];
            tile[r][tx] = v;
            d0[(int64_t)h * W + w] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;
        if (h < H && w < W) d1[(int64_t)w * H + h] = tile[tx][r];
    }
}

// ------------------------------------------------------------------------------------------ dwconv + SiLU + pack
// in  : (B, H, W, *) channel-last, channel c of token (h, w) at in[((b*H + h)*W + w)*in_stride + c], c < D
// out : xx (B, 2, D, L) = [act(conv(x)) ; its transpose]
// One CTA = 16 x 32 tokens (+1 halo) x 16 channels, 256 threads.
//   load   : 16-byte vectors along the channels (the in_proj GEMM output layout) -> shared memory as channel PAIRS
//   compute: thread = one channel pair x one row x 16 consecutive tokens; every input pair is read once per
//            kernel row (54 LDS for 32 outputs) and the 3x3 weights live in registers
//   store  : row-major plane straight from registers (32-byte runs); the column-major plane goes through a
//            swizzled shared-memory transpose (aliasing the dead input tile) and leaves as 32-byte runs along h
constexpr int DW_TH = 16, DW_TW = 32, DW_CB = 16, DW_NP = DW_CB / 2, DW_PITCH = DW_NP + 1;
constexpr int DW_HT = DW_TH + 2, DW_WT = DW_TW + 2;

template <typename T> struct DwPair;                      // two adjacent channels of one token in shared memory
template <> struct DwPair<float> {
    using type = float2;
    static __device__ __forceinline__ float2 unpack(float2 v) { return v; }
    static __device__ __forceinline__ float2 pack(float a, float b) { return make_float2(a, b); }
};
template <> struct DwPair<__half> {
    using type = __half2;
    static __device__ __forceinline__ float2 unpack(__half2 v) { return __half22float2(v); }
    static __device__ __forceinline__ __half2 pack(float a, float b) { return __floats2half2_rn(a, b); }
};
template <> struct DwPair<__nv_bfloat16> {
    using type = __nv_bfloat162;
    static __device__ __forceinline__ float2 unpack(__nv_bfloat162 v) { return __bfloat1622float2(v); }
    static __device__ __forceinline__ __nv_bfloat162 pack(float a, float b) { return __floats2bfloat162_rn(a, b); }
};

template <typename T, bool SILU>
__global__ void __launch_bounds__(256, 4) ss2d_dwconv_pack_kernel(const T* __restrict__ in, const float* __restrict__ wgt,
                                                               const float* __restrict__ bias, T* __restrict__ xx, int64_t D,
                                                               int H, int W, int64_t in_stride, int tiles_w, int tiles_h,
                                                               int chan_blocks, int vec_in, int vec_row, int vec_col) {
    using P = DwPair<T>;
    using PairT = typename P::type;
    constexpr int VEC = 16 / (int)sizeof(T);              // elements per 16-byte vector
    constexpr int TILE_PAIRS = DW_HT * DW_WT * DW_PITCH;
    constexpr int STAGE_ELEMS = DW_CB * DW_TW * DW_TH;
    static_assert(STAGE_ELEMS * sizeof(T) <= TILE_PAIRS * sizeof(PairT), "transpose staging must fit in the input tile");
    __shared__ __align__(16) PairT tile[TILE_PAIRS];
    const int64_t L = (int64_t)H * W;
    int64_t t = blockIdx.x;
    const int cb = (int)(t % chan_blocks); t /= chan_blocks;
    const int tw = (int)(t % tiles_w); t /= tiles_w;
    const int th = (int)(t % tiles_h); t /= tiles_h;
    const int64_t b = t;
    const int h0 = th * DW_TH, w0 = tw * DW_TW, c0 = cb * DW_CB;
    const int tid = threadIdx.x;

    // ---- load the halo tile
    const T* inb = in + b * H * (int64_t)W * in_stride;
    if (vec_in) {                                         // D % VEC == 0: vectors are all-or-nothing against D
        constexpr int VPT = DW_CB / VEC;                  // vectors per token
        constexpr int NV = DW_HT * DW_WT * VPT, NLD = (NV + 255) / 256;
        uint4 raw[NLD];                                   // every load of this thread in flight before the first store
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int i = tid + k * 256;
            const int v = i % VPT, tok = i / VPT;
            const int hh = tok / DW_WT, ww = tok % DW_WT;
            const int h = h0 + hh - 1, w = w0 + ww - 1;
            raw[k] = make_uint4(0u, 0u, 0u, 0u);          // zero padding (Conv2d padding=1)
            if (i < NV && h >= 0 && h < H && w >= 0 && w < W && c0 + v * VEC < D)
                raw[k] = __ldg(reinterpret_cast<const uint4*>(inb + ((int64_t)h * W + w) * in_stride + c0 + v * VEC));
        }
#pragma unroll
        for (int k = 0; k < NLD; ++k) {
            const int i = tid + k * 256;
            if (i < NV) {
                const int v = i % VPT, tok = i / VPT;
                const PairT* pr = reinterpret_cast<const PairT*>(&raw[k]);
#pragma unroll
                for (int j = 0; j < VEC / 2; ++j) tile[tok * DW_PITCH + v * (VEC / 2) + j] = pr[j];
            }
        }
    } else {
        for (int i = tid; i < DW_HT * DW_WT * DW_NP; i += 256) {
            const int pr = i % DW_NP, tok = i / DW_NP;
            const int hh = tok / DW_WT, ww = tok % DW_WT;
            const int h = h0 + hh - 1, w = w0 + ww - 1;
            float a = 0.0f, c = 0.0f;
            if (h >= 0 && h < H && w >= 0 && w < W) {
                const T* src = inb + ((int64_t)h * W + w) * in_stride + c0 + 2 * pr;
                if (c0 + 2 * pr < D) a = to_f32(src[0]);
                if (c0 + 2 * pr + 1 < D) c = to_f32(src[1]);
            }
            tile[tok * DW_PITCH + pr] = P::pack(a, c);
        }
    }
    // ---- this thread's pair / row / strip, weights in registers
    const int hh = tid % DW_TH, pr = (tid / DW_TH) % DW_NP, strip = tid / (DW_TH * DW_NP);   // strip: 16 tokens
    const int cA = c0 + 2 * pr, cB = cA + 1;
    float2 wab[9];                                        // (channel A, channel B) taps: one packed FFMA2 serves both
#pragma unroll
    for (int k = 0; k < 9; ++k)
        wab[k] = make_float2(cA < D ? __ldg(wgt + (int64_t)cA * 9 + k) : 0.0f, cB < D ? __ldg(wgt + (int64_t)cB * 9 + k) : 0.0f);
    const float ba = (bias && cA < D) ? __ldg(bias + cA) : 0.0f, bb = (bias && cB < D) ? __ldg(bias + cB) : 0.0f;
    __syncthreads();
    constexpr int NT = 16;
    float2 yab[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) yab[j] = make_float2(ba, bb);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const PairT* row = tile + ((hh + ky) * DW_WT + strip * NT) * DW_PITCH + pr;
#pragma unroll
        for (int q = 0; q < NT + 2; ++q) {
            const float2 v = P::unpack(row[q * DW_PITCH]);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int j = q - kx;
                if (j >= 0 && j < NT) yab[j] = fma2(v, wab[ky * 3 + kx], yab[j]);
            }
        }
    }
    if (SILU) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            // x * rcp(1 + 2^(-x log2 e)) on the channel pair: MUFU.EX2 + MUFU.RCP per value, the rest packed.  (An IEEE division
            // costs ~8 more instructions per element in this issue-bound kernel, __fdividef 3 more for its range check; the
            // approximate reciprocal is good to 1 ulp, far inside the fp32 / 16-bit tolerances of the SS2D block.  For x -> -inf
            // the denominator overflows to +inf and the product is -0, like the exact expression.)
            const float2 t = mul2(yab[j], make_float2(-1.4426950408889634f, -1.4426950408889634f));
            float2 e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
            const float2 den = fma2(e, make_float2(1.0f, 1.0f), make_float2(1.0f, 1.0f));
            float2 r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(den.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(den.y));
            yab[j] = mul2(yab[j], r);
        }
    }
    float ya[NT], yb[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) { ya[j] = yab[j].x; yb[j] = yab[j].y; }
    // ---- row-major plane: 16 consecutive tokens per channel straight from registers
    const int h = h0 + hh, wbeg = w0 + strip * NT;
    if (h < H) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int c = e ? cB : cA;
            if (c >= D) continue;
            const float* y = e ? yb : ya;
            T* dst = xx + ((b * 2 + 0) * D + c) * L + (int64_t)h * W + wbeg;
            if (vec_row) {                                 // W % VEC == 0: vectors are all-or-nothing against W
#pragma unroll
                for (int v = 0; v < NT / VEC; ++v) {
                    if (wbeg + v * VEC >= W) break;
                    uint4 raw;
                    T* o = reinterpret_cast<T*>(&raw);
#pragma unroll
                    for (int j = 0; j < VEC; ++j) o[j] = from_f32<T>(y[v * VEC + j]);
                    *reinterpret_cast<uint4*>(dst + v * VEC) = raw;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    if (wbeg + j < W) dst[j] = from_f32<T>(y[j]);
            }
        }
    }
    // ---- column-major plane: transpose through shared memory, layout [c][w'][h] with w' = (w + c/2) & 31 so that a
    //      warp (16 rows x 2 pairs) writes 16 different banks, two rows per 32-bit word
    __syncthreads();                                      // every thread is done reading the input tile
    T* stage = reinterpret_cast<T*>(tile);
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int wsw = (strip * NT + j + pr) & (DW_TW - 1);
        stage[((2 * pr) * DW_TW + wsw) * DW_TH + hh] = from_f32<T>(ya[j]);
        stage[((2 * pr + 1) * DW_TW + wsw) * DW_TH + hh] = from_f32<T>(yb[j]);
    }
    __syncthreads();
    constexpr int VPC = DW_TH / VEC;                       // vectors per (channel, column)
    for (int i = tid; i < DW_CB * DW_TW * VPC; i += 256) {
        const int v = i % VPC, ww = (i / VPC) % DW_TW, cl = i / (VPC * DW_TW);
        const int c = c0 + cl, w = w0 + ww, hb = h0 + v * VEC;
        if (c >= D || w >= W || hb >= H) continue;
        const T* src = stage + (cl * DW_TW + ((ww + cl / 2) & (DW_TW - 1))) * DW_TH + v * VEC;
        T* dst = xx + ((b * 2 + 1) * D + c) * L + (int64_t)w * H + hb;
        if (vec_col) {                                     // H % VEC == 0
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (hb + j < H) dst[j] = src[j];
        }
    }
}

// ------------------------------------------------------------------------------------------ merge + norm (+ gate)
// CTA = TH x TW token tile, all D channels, fp32 tile[token][D | 1] in shared memory.
//   phase 1: tile  = y0 + y1        row-major planes, float4 = 4 consecutive w
//   phase 2: tile += y2 + y3        column-major planes, float4 = 4 consecutive h
//   phase 3: warp per token: two-pass LayerNorm over the D channels (row held in registers when DPER > 0),
//            affine, optional gate, channel-last store
template <typename TO, int TH, int TW, int DPER, bool MERGED>
__global__ void __launch_bounds__(256) ss2d_merge_norm_kernel(const float* __restrict__ ys, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const TO* __restrict__ zact,
                                                              TO* __restrict__ out, int D, int H, int W, int tiles_w,
                                                              int tiles_h, float eps) {
    extern __shared__ __align__(16) float tile[];
    const int P = D | 1;                                   // odd pitch: conflict-free token-major writes
    const int L = H * W;                                   // D * L < 2^31 (host-checked): 32-bit offsets inside a plane set
    int64_t t = blockIdx.x;
    const int tw = (int)(t % tiles_w); t /= tiles_w;
    const int th = (int)(t % tiles_h); t /= tiles_h;
    const int64_t b = t;
    const int h0 = th * TH, w0 = tw * TW;
    const int tid = threadIdx.x;
    // MERGED: ys is the already merged (B, D, H, W) plane of xp_ss2d_core -- phase 1 copies it, phase 2 is skipped
    // (a compile-time switch: as a run-time flag inside the load loops it cost the four-plane kernel 0.8 ms per stage-0 launch)
    constexpr bool merged = MERGED;
    const float* y0 = ys + (b * (MERGED ? 1 : 4) + 0) * D * (int64_t)L;
    const float* y1 = y0 + (int64_t)D * L;
    const float* y2 = y1 + (int64_t)D * L;
    const float* y3 = y2 + (int64_t)D * L;
    constexpr int QW = TW / 4, QH = TH / 4;
    // phase 1
    {
        const int q = tid % QW, hh = (tid / QW) % TH, cstep = 256 / (QW * TH);
        const int h = h0 + hh, w = w0 + 4 * q;
        const bool ok = h < H && w < W;                    // W % 4 == 0: quads are all-or-nothing
        const int base = h * W + w;
        float* dst = tile + (hh * TW + 4 * q) * P;
#pragma unroll 4
        for (int c = tid / (QW * TH); c < D; c += cstep) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(y0 + c * L + base));
                if (merged) {
                    v = a;
                } else {
                    const float4 r = __ldg(reinterpret_cast<const float4*>(y1 + c * L + base));
                    v = make_float4(a.x + r.x, a.y + r.y, a.z + r.z, a.w + r.w);
                }
            }
            dst[c] = v.x; dst[P + c] = v.y; dst[2 * P + c] = v.z; dst[3 * P + c] = v.w;
        }
    }
    __syncthreads();
    // phase 2
    {
        const int q = tid % QH, ww = (tid / QH) % TW, cstep = 256 / (QH * TW);
        const int w = w0 + ww, h = h0 + 4 * q;
        const bool ok = h < H && w < W;
        const int base = w * H + h;
        float* dst = tile + ((4 * q) * TW + ww) * P;
        if (ok && !MERGED) {
#pragma unroll 4
            for (int c = tid / (QH * TW); c < D; c += cstep) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(y2 + c * L + base));
                const float4 r = __ldg(reinterpret_cast<const float4*>(y3 + c * L + base));
                dst[c] += a.x + r.x; dst[TW * P + c] += a.y + r.y; dst[2 * TW * P + c] += a.z + r.z;
                dst[3 * TW * P + c] += a.w + r.w;
            }
        }
    }
    __syncthreads();
    // phase 3
    const int warp = tid >> 5, lane = tid & 31;
    if constexpr (DPER > 0) {
        float gm[DPER], bt[DPER];
#pragma unroll
        for (int i = 0; i < DPER; ++i) {
            const int c = lane + 32 * i;
            gm[i] = c < D ? __ldg(gamma + c) : 0.0f;
            bt[i] = c < D ? __ldg(beta + c) : 0.0f;
        }
        const float invD = 1.0f / (float)D;
        for (int tok = warp; tok < TH * TW; tok += 8) {
            const int h = h0 + tok / TW, w = w0 + tok % TW;
            if (h >= H || w >= W) continue;                // warp-uniform
            const float* row = tile + tok * P;
            float v[DPER];
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < DPER; ++i) {
                const int c = lane + 32 * i;
                v[i] = c < D ? row[c] : 0.0f;
                s += v[i];
            }
            const float mean = warp_sum(s) * invD;
            float ss = 0.0f;
#pragma unroll
            for (int i = 0; i < DPER; ++i) {
                const float dlt = (lane + 32 * i < D) ? v[i] - mean : 0.0f;
                v[i] = dlt;
                ss = fmaf(dlt, dlt, ss);
            }
            const float rstd = rsqrtf(warp_sum(ss) * invD + eps);
            const int64_t o = (b * L + (int64_t)h * W + w) * D;
#pragma unroll
            for (int i = 0; i < DPER; ++i) {
                const int c = lane + 32 * i;
                if (c < D) {
                    float r = fmaf(v[i] * rstd, gm[i], bt[i]);
                    if (zact) r *= to_f32(zact[o + c]);
                    out[o + c] = from_f32<TO>(r);
                }
            }
        }
    } else {
        for (int tok = warp; tok < TH * TW; tok += 8) {
            const int h = h0 + tok / TW, w = w0 + tok % TW;
            if (h >= H || w >= W) continue;                // warp-uniform
            const float* row = tile + tok * P;
            float s = 0.0f;
            for (int c = lane; c < D; c += 32) s += row[c];
            const float mean = warp_sum(s) / (float)D;
            float ss = 0.0f;
            for (int c = lane; c < D; c += 32) { const float dlt = row[c] - mean; ss += dlt * dlt; }
            const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
            const int64_t o = (b * L + (int64_t)h * W + w) * D;
            for (int c = lane; c < D; c += 32) {
                float v = (row[c] - mean) * rstd * gamma[c] + beta[c];
                if (zact) v *= to_f32(zact[o + c]);
                out[o + c] = from_f32<TO>(v);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ dt projection (rank <= 8)
// delta[b, g, d, l] = sum_r W[g, d, r] * dts_r[b, g, r, l]   (VMamba.py:607-608 / :325: dt_proj as a grouped 1x1 conv).
// With dt_rank = ceil(d_model / 16) <= 8 this is an outer-product-sized contraction whose cost is writing delta: a
// library GEMM runs it at a quarter of the HBM rate, here every thread keeps its 8 tokens of the R rank rows in
// registers and streams 16-byte stores over the channel rows.
template <typename T>
__global__ void __launch_bounds__(256) ss2d_dt_proj_kernel(const T* __restrict__ xr, const float* __restrict__ Wt, T* __restrict__ out,
                                                           int G, int D, int R, int64_t L, int64_t x_bs, int64_t x_gs,
                                                           int64_t x_rs) {
    constexpr int VEC = 16 / (int)sizeof(T);          // tokens per thread (8 for 16-bit, 4 for fp32)
    // the 8 warps of a CTA cover ADJACENT token ranges and walk the channel rows together: at any moment the CTA writes
    // one 4 KiB (2 KiB for fp32 ... 16-byte x 256 threads) run of a delta row, which DRAM takes far better than 8 rows apart
    const int64_t bg = blockIdx.y;
    const int64_t b = bg / G, g = bg % G;
    const int64_t l0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * VEC;
    if (l0 >= L) return;                               // L % VEC == 0: token groups are all-or-nothing
    float x[8][VEC];
    const T* src = xr + b * x_bs + g * x_gs + l0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r < R) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + r * x_rs));
            const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
            for (int j = 0; j < VEC; ++j) x[r][j] = to_f32(e[j]);
        }
    }
    const float* Wg = Wt + g * (int64_t)D * R;
    T* dst = out + (bg * D) * L + l0;
    for (int d = 0; d < D; ++d) {
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r < R) {
                const float w = __ldg(Wg + (int64_t)d * R + r);     // the warp shares the row: one broadcast L1 hit
#pragma unroll
                for (int j = 0; j < VEC; ++j) acc[j] = fmaf(w, x[r][j], acc[j]);
            }
        }
        uint4 raw;
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int j = 0; j < VEC; ++j) e[j] = from_f32<T>(acc[j]);
        *reinterpret_cast<uint4*>(dst + (int64_t)d * L) = raw;
    }
}

// 16-bit inputs (all ranks up to 16; XPoint: dt_rank 6, 12): the 8 tokens x R rank rows of a thread stay PACKED (16-byte
// vectors, R * 4 registers) and feed the mixed-precision FMA of sm_100 (fma.rn.f32.f16 -> FHFMA: 16-bit operands taken
// from either half of a register, fp32 accumulator, full FFMA rate), so nothing is widened; the (D, R) weights of the
// group are rounded to the input dtype (what the reference's autocast GEMM multiplies with) and paired in shared memory.
template <typename T> __device__ __forceinline__ float fma_mixed16(uint32_t a16, uint32_t b16, float acc);
template <> __device__ __forceinline__ float fma_mixed16<__half>(uint32_t a16, uint32_t b16, float acc) {
    float d;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"((unsigned short)a16), "h"((unsigned short)b16), "f"(acc));
    return d;
}
template <> __device__ __forceinline__ float fma_mixed16<__nv_bfloat16>(uint32_t a16, uint32_t b16, float acc) {
    float d;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"((unsigned short)a16), "h"((unsigned short)b16), "f"(acc));
    return d;
}

template <typename T, int RP>     // RP: rank pairs held per thread (8 -> ranks up to 16)
__global__ void __launch_bounds__(256) ss2d_dt_proj16_kernel(const T* __restrict__ xr, const float* __restrict__ Wt,
                                                             T* __restrict__ out, int G, int D, int R, int64_t L, int64_t x_bs,
                                                             int64_t x_gs, int64_t x_rs) {
    extern __shared__ uint32_t wsm[];                  // [D][RP] weight pairs (W[d][2q], W[d][2q+1]) in T, zero-padded
    const int64_t bg = blockIdx.y;
    const int64_t b = bg / G, g = bg % G;
    const float* Wg = Wt + g * (int64_t)D * R;
    for (int i = threadIdx.x; i < D * RP; i += 256) {
        const int d = i / RP, q = i % RP;
        const T lo = from_f32<T>(2 * q < R ? Wg[(int64_t)d * R + 2 * q] : 0.0f);
        const T hi = from_f32<T>(2 * q + 1 < R ? Wg[(int64_t)d * R + 2 * q + 1] : 0.0f);
        wsm[i] = (uint32_t)*reinterpret_cast<const unsigned short*>(&lo) |
                 ((uint32_t)*reinterpret_cast<const unsigned short*>(&hi) << 16);
    }
    __syncthreads();
    const int64_t l0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;    // warps side by side along the tokens (see above)
    if (l0 >= L) return;                               // L % 8 == 0: token groups are all-or-nothing
    uint4 x[2 * RP];
    const T* src = xr + b * x_bs + g * x_gs + l0;
#pragma unroll
    for (int r = 0; r < 2 * RP; ++r) x[r] = r < R ? __ldg(reinterpret_cast<const uint4*>(src + r * x_rs)) : make_uint4(0u, 0u, 0u, 0u);
    T* dst = out + (bg * D) * L + l0;
    for (int d = 0; d < D; ++d) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const uint32_t w2 = wsm[d * RP + q];       // broadcast
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t w = h ? w2 >> 16 : w2 & 0xffffu;
                const uint4 v = x[2 * q + h];
                acc[0] = fma_mixed16<T>(v.x & 0xffffu, w, acc[0]); acc[1] = fma_mixed16<T>(v.x >> 16, w, acc[1]);
                acc[2] = fma_mixed16<T>(v.y & 0xffffu, w, acc[2]); acc[3] = fma_mixed16<T>(v.y >> 16, w, acc[3]);
                acc[4] = fma_mixed16<T>(v.z & 0xffffu, w, acc[4]); acc[5] = fma_mixed16<T>(v.z >> 16, w, acc[5]);
                acc[6] = fma_mixed16<T>(v.w & 0xffffu, w, acc[6]); acc[7] = fma_mixed16<T>(v.w >> 16, w, acc[7]);
            }
        }
        uint4 raw;
        T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = from_f32<T>(acc[j]);
        *reinterpret_cast<uint4*>(dst + (int64_t)d * L) = raw;
    }
}

template <typename TO, int TH, int TW, int DPER>
static int merge_norm_launch_d(const float* ys, const float* gamma, const float* beta, const void* zact, void* out, int64_t B,
                               int64_t D, int64_t H, int64_t W, float eps, cudaStream_t st, int planes) {
    const int tiles_w = (int)ceil_div(W, TW), tiles_h = (int)ceil_div(H, TH);
    const int smem = TH * TW * (int)(D | 1) * 4;
    if (planes == 1) {
        auto kern = ss2d_merge_norm_kernel<TO, TH, TW, DPER, true>;
        XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)(B * tiles_h * tiles_w), 256, smem, st>>>(ys, gamma, beta, (const TO*)zact, (TO*)out, (int)D, (int)H, (int)W,
                                                                     tiles_w, tiles_h, eps);
    } else {
        auto kern = ss2d_merge_norm_kernel<TO, TH, TW, DPER, false>;
        XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)(B * tiles_h * tiles_w), 256, smem, st>>>(ys, gamma, beta, (const TO*)zact, (TO*)out, (int)D, (int)H, (int)W,
                                                                     tiles_w, tiles_h, eps);
    }
    XP_LAUNCH_CHECK("ss2d_merge_norm_kernel");
    return XP_OK;
}

template <typename TO, int TH, int TW>
static int merge_norm_launch(const float* ys, const float* gamma, const float* beta, const void* zact, void* out, int64_t B,
                             int64_t D, int64_t H, int64_t W, float eps, cudaStream_t st, int planes) {
    const int per = (int)ceil_div(D, 32);                  // channels per lane in the LayerNorm phase
    if (per <= 3) return merge_norm_launch_d<TO, TH, TW, 3>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    if (per <= 6) return merge_norm_launch_d<TO, TH, TW, 6>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    if (per <= 12) return merge_norm_launch_d<TO, TH, TW, 12>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    if (per <= 24) return merge_norm_launch_d<TO, TH, TW, 24>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    return merge_norm_launch_d<TO, TH, TW, 0>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
}

template <typename TO>
static int merge_norm_dispatch(const float* ys, const float* gamma, const float* beta, const void* zact, void* out, int64_t B,
                               int64_t D, int64_t H, int64_t W, float eps, cudaStream_t st, int planes) {
    // 8x8 tokens for narrow blocks (25 KiB tile, 8 CTAs / SM at D = 96); from D = 192 on the 4x8 tile wins because it
    // doubles the resident CTAs (measured on B200: D = 192 1.01 -> 0.58 ms, D = 768 0.26 -> 0.18 ms); 4x4 for very wide blocks
    static const int tile_env = getenv("XP_MN_TILE") ? atoi(getenv("XP_MN_TILE")) : 0;     // tuning knob: 48 = 4x8, 44 = 4x4
    if (tile_env == 48 && D <= 1536) return merge_norm_launch<TO, 4, 8>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    if (tile_env == 44) return merge_norm_launch<TO, 4, 4>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    if (D <= 128) return merge_norm_launch<TO, 8, 8>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    if (D <= 1536) return merge_norm_launch<TO, 4, 8>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    return merge_norm_launch<TO, 4, 4>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
}

}  // namespace xp

using namespace xp;

extern "C" int xp_ss2d_pack(const void* x, void* xx, int64_t B, int64_t D, int64_t H, int64_t W, int32_t dtype,
                            xp_stream_t stream) {
    XP_REQUIRE(x && xx, "xp_ss2d_pack: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && D > 0 && H > 0 && W > 0, "xp_ss2d_pack: bad shape");
    XP_REQUIRE(dtype >= XP_F32 && dtype <= XP_BF16, "xp_ss2d_pack: unsupported dtype %d", dtype);
    if (B == 0) return XP_OK;
    const int tiles_w = (int)ceil_div(W, 32), tiles_h = (int)ceil_div(H, 32);
    const int64_t grid = B * D * tiles_w * tiles_h;
    XP_REQUIRE(grid < (int64_t)1 << 31, "xp_ss2d_pack: problem too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XP_F32) ss2d_pack_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)x, (float*)xx, D, (int)H, (int)W, tiles_w, tiles_h);
    else ss2d_pack_kernel<unsigned short><<<(unsigned)grid, 256, 0, st>>>((const unsigned short*)x, (unsigned short*)xx, D, (int)H, (int)W, tiles_w, tiles_h);
    XP_LAUNCH_CHECK("ss2d_pack_kernel");
    return XP_OK;
}

extern "C" int xp_ss2d_dwconv_pack(const void* in, const float* weight, const float* bias, void* xx, int64_t B, int64_t D,
                                   int64_t H, int64_t W, int64_t in_token_stride, int32_t dtype, int32_t silu,
                                   xp_stream_t stream) {
    XP_REQUIRE(in && weight && xx, "xp_ss2d_dwconv_pack: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && D > 0 && H > 0 && W > 0 && in_token_stride >= D, "xp_ss2d_dwconv_pack: bad shape");
    XP_REQUIRE(dtype >= XP_F32 && dtype <= XP_BF16, "xp_ss2d_dwconv_pack: unsupported dtype %d", dtype);
    if (B == 0) return XP_OK;
    const int tiles_w = (int)ceil_div(W, DW_TW), tiles_h = (int)ceil_div(H, DW_TH), cbs = (int)ceil_div(D, DW_CB);
    const int64_t grid = B * tiles_w * tiles_h * cbs;
    XP_REQUIRE(grid < (int64_t)1 << 31, "xp_ss2d_dwconv_pack: problem too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t vec = dtype == XP_F32 ? 4 : 8;           // elements per 16-byte vector
    const bool al_in = (reinterpret_cast<uintptr_t>(in) & 15) == 0, al_out = (reinterpret_cast<uintptr_t>(xx) & 15) == 0;
    const int vec_in = al_in && in_token_stride % vec == 0 && D % vec == 0;
    const int vec_row = al_out && W % vec == 0, vec_col = al_out && H % vec == 0 && (H * W) % vec == 0;
#define XP_DW(T, S) ss2d_dwconv_pack_kernel<T, S><<<(unsigned)grid, 256, 0, st>>>((const T*)in, weight, bias, (T*)xx, D, (int)H, \
                                                                                   (int)W, in_token_stride, tiles_w, tiles_h, cbs, \
                                                                                   vec_in, vec_row, vec_col)
    if (dtype == XP_F32) { if (silu) XP_DW(float, true); else XP_DW(float, false); }
    else if (dtype == XP_F16) { if (silu) XP_DW(__half, true); else XP_DW(__half, false); }
    else { if (silu) XP_DW(__nv_bfloat16, true); else XP_DW(__nv_bfloat16, false); }
#undef XP_DW
    XP_LAUNCH_CHECK("ss2d_dwconv_pack_kernel");
    return XP_OK;
}

static int merge_norm_entry(const char* who, const float* ys, const float* gamma, const float* beta, const void* zact, void* out,
                            int64_t B, int64_t D, int64_t H, int64_t W, int32_t out_dtype, float eps, xp_stream_t stream,
                            int planes) {
    XP_REQUIRE(ys && gamma && beta && out, "%s: NULL tensor pointer", who);
    XP_REQUIRE(B >= 0 && D > 0 && H > 0 && W > 0, "%s: bad shape", who);
    XP_REQUIRE(H % 4 == 0 && W % 4 == 0, "%s: H and W must be multiples of 4 (got %lld x %lld)", who, (long long)H, (long long)W);
    XP_REQUIRE(D <= 3072, "%s: D must be <= 3072 (got %lld)", who, (long long)D);
    XP_REQUIRE(D * H * W < ((int64_t)1 << 31), "%s: D*H*W must be < 2^31", who);
    XP_REQUIRE(out_dtype >= XP_F32 && out_dtype <= XP_BF16, "%s: unsupported dtype %d", who, out_dtype);
    XP_REQUIRE((reinterpret_cast<uintptr_t>(ys) & 15) == 0, "%s: the y planes must be 16-byte aligned", who);
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (out_dtype) {
        case XP_F32: return merge_norm_dispatch<float>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
        case XP_F16: return merge_norm_dispatch<__half>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
        default: return merge_norm_dispatch<__nv_bfloat16>(ys, gamma, beta, zact, out, B, D, H, W, eps, st, planes);
    }
}

extern "C" int xp_ss2d_merge_norm(const float* ys, const float* gamma, const float* beta, const void* zact, void* out,
                                  int64_t B, int64_t D, int64_t H, int64_t W, int32_t out_dtype, float eps,
                                  xp_stream_t stream) {
    return merge_norm_entry("xp_ss2d_merge_norm", ys, gamma, beta, zact, out, B, D, H, W, out_dtype, eps, stream, 4);
}

extern "C" int xp_ss2d_plane_norm(const float* y, const float* gamma, const float* beta, const void* zact, void* out,
                                  int64_t B, int64_t D, int64_t H, int64_t W, int32_t out_dtype, float eps,
                                  xp_stream_t stream) {
    return merge_norm_entry("xp_ss2d_plane_norm", y, gamma, beta, zact, out, B, D, H, W, out_dtype, eps, stream, 1);
}

extern "C" int xp_ss2d_dt_proj(const void* dts_r, const float* weight, void* delta, int64_t B, int64_t G, int64_t D, int64_t R,
                               int64_t L, int64_t x_batch_stride, int64_t x_group_stride, int64_t x_rank_stride, int32_t dtype,
                               xp_stream_t stream) {
    XP_REQUIRE(dts_r && weight && delta, "xp_ss2d_dt_proj: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && G > 0 && D > 0 && L > 0, "xp_ss2d_dt_proj: bad shape");
    XP_REQUIRE(R >= 1 && (R <= 8 || (R <= 16 && dtype != XP_F32 && D * 8 * 4 <= 48 * 1024)),
               "xp_ss2d_dt_proj: dt_rank must be in 1..8, or 9..16 with 16-bit inputs and D <= 1536 (got %lld)", (long long)R);
    XP_REQUIRE(R <= 8 || D * 8 * 4 <= 48 * 1024, "xp_ss2d_dt_proj: D too large for ranks above 8");
    XP_REQUIRE(dtype >= XP_F32 && dtype <= XP_BF16, "xp_ss2d_dt_proj: unsupported dtype %d", dtype);
    const int64_t vec = dtype == XP_F32 ? 4 : 8;
    XP_REQUIRE(L % vec == 0 && x_batch_stride % vec == 0 && x_group_stride % vec == 0 && x_rank_stride % vec == 0 &&
                   (reinterpret_cast<uintptr_t>(dts_r) & 15) == 0 && (reinterpret_cast<uintptr_t>(delta) & 15) == 0,
               "xp_ss2d_dt_proj: rows must be 16-byte aligned (L and strides multiples of %lld elements)", (long long)vec);
    XP_REQUIRE(B * G <= 65535, "xp_ss2d_dt_proj: batch * groups must be <= 65535");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)ceil_div(L, 256 * vec), (unsigned)(B * G));
    if (dtype != XP_F32 && D * 8 * 4 <= 48 * 1024) {        // 16-bit inputs: packed operands + mixed-precision FMA (FHFMA)
#define XP_DT16(T, RP) ss2d_dt_proj16_kernel<T, RP><<<grid, 256, (size_t)D * RP * 4, st>>>((const T*)dts_r, weight, (T*)delta, (int)G, \
                                                                                            (int)D, (int)R, L, x_batch_stride,     \
                                                                                            x_group_stride, x_rank_stride)
        if (dtype == XP_F16) { if (R <= 8) XP_DT16(__half, 4); else XP_DT16(__half, 8); }
        else { if (R <= 8) XP_DT16(__nv_bfloat16, 4); else XP_DT16(__nv_bfloat16, 8); }
#undef XP_DT16
        XP_LAUNCH_CHECK("ss2d_dt_proj16_kernel");
        return XP_OK;
    }
#define XP_DT(T) ss2d_dt_proj_kernel<T><<<grid, 256, 0, st>>>((const T*)dts_r, weight, (T*)delta, (int)G, (int)D, (int)R, L, \
                                                               x_batch_stride, x_group_stride, x_rank_stride)
    if (dtype == XP_F32) XP_DT(float);
    else if (dtype == XP_F16) XP_DT(__half);
    else XP_DT(__nv_bfloat16);
#undef XP_DT
    XP_LAUNCH_CHECK("ss2d_dt_proj_kernel");
    return XP_OK;
}
