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
// One CTA: a TH x TW token tile (+1 halo) of CB channels.  Loads are channel-contiguous (the in_proj GEMM output
// layout), stores are token-contiguous runs in both orders.
constexpr int DW_TH = 16, DW_TW = 32, DW_CB = 16;

template <typename T, bool SILU>
__global__ void __launch_bounds__(256) ss2d_dwconv_pack_kernel(const T* __restrict__ in, const float* __restrict__ wgt,
                                                               const float* __restrict__ bias, T* __restrict__ xx, int64_t D,
                                                               int H, int W, int64_t in_stride, int tiles_w, int tiles_h,
                                                               int chan_blocks) {
    constexpr int HT = DW_TH + 2, WT = DW_TW + 2;
    __shared__ float sin[HT][WT][DW_CB + 1];
    __shared__ float sw[DW_CB][9];
    __shared__ float sb[DW_CB];
    const int64_t L = (int64_t)H * W;
    int64_t t = blockIdx.x;
    const int cb = (int)(t % chan_blocks); t /= chan_blocks;
    const int tw = (int)(t % tiles_w); t /= tiles_w;
    const int th = (int)(t % tiles_h); t /= tiles_h;
    const int64_t b = t;
    const int h0 = th * DW_TH, w0 = tw * DW_TW, c0 = cb * DW_CB;
    const int tid = threadIdx.x;
    if (tid < DW_CB * 9) {
        const int c = tid / 9, k = tid % 9;
        sw[c][k] = c0 + c < D ? wgt[(int64_t)(c0 + c) * 9 + k] : 0.0f;
    }
    if (tid < DW_CB) sb[tid] = (bias && c0 + tid < D) ? bias[c0 + tid] : 0.0f;
    // load the halo tile: consecutive threads -> consecutive channels of one token
    for (int i = tid; i < HT * WT * DW_CB; i += 256) {
        const int c = i % DW_CB, tok = i / DW_CB;
        const int hh = tok / WT, ww = tok % WT;
        const int h = h0 + hh - 1, w = w0 + ww - 1;
        float v = 0.0f;                                              // zero padding (Conv2d padding=1)
        if (h >= 0 && h < H && w >= 0 && w < W && c0 + c < D) v = to_f32(in[((b * H + h) * (int64_t)W + w) * in_stride + c0 + c]);
        sin[hh][ww][c] = v;
    }
    __syncthreads();
    // compute into registers (thread -> channel tid % CB, tokens tid / CB + 16 k), then park the results in the same
    // shared memory (the halo tile is dead by then) for the token-contiguous stores
    constexpr int PER = DW_CB * DW_TH * DW_TW / 256;
    float acc[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = tid + 256 * k;
        const int c = i % DW_CB, tok = i / DW_CB;
        const int hh = tok / DW_TW, ww = tok % DW_TW;
        float a = sb[c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) a = fmaf(sin[hh + ky][ww + kx][c], sw[c][ky * 3 + kx], a);
        if (SILU) a = a / (1.0f + __expf(-a));
        acc[k] = a;
    }
    __syncthreads();
    constexpr int CP = DW_TH * (DW_TW + 1) + 1;            // odd channel pitch: conflict-free channel-fastest writes
    static_assert(DW_CB * CP <= HT * WT * (DW_CB + 1), "output staging must fit in the halo tile");
    float* sout = &sin[0][0][0];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int i = tid + 256 * k;
        const int c = i % DW_CB, tok = i / DW_CB;
        sout[c * CP + (tok / DW_TW) * (DW_TW + 1) + tok % DW_TW] = acc[k];
    }
    __syncthreads();
    // row-major plane: runs of DW_TW tokens
    for (int i = tid; i < DW_CB * DW_TH * DW_TW; i += 256) {
        const int ww = i % DW_TW, hh = (i / DW_TW) % DW_TH, c = i / (DW_TW * DW_TH);
        const int h = h0 + hh, w = w0 + ww;
        if (h < H && w < W && c0 + c < D) xx[((b * 2 + 0) * D + c0 + c) * L + (int64_t)h * W + w] = from_f32<T>(sout[c * CP + hh * (DW_TW + 1) + ww]);
    }
    // column-major plane: runs of DW_TH tokens
    for (int i = tid; i < DW_CB * DW_TH * DW_TW; i += 256) {
        const int hh = i % DW_TH, ww = (i / DW_TH) % DW_TW, c = i / (DW_TW * DW_TH);
        const int h = h0 + hh, w = w0 + ww;
        if (h < H && w < W && c0 + c < D) xx[((b * 2 + 1) * D + c0 + c) * L + (int64_t)w * H + h] = from_f32<T>(sout[c * CP + hh * (DW_TW + 1) + ww]);
    }
}

// ------------------------------------------------------------------------------------------ merge + norm (+ gate)
// CTA = TH x TW token tile, all D channels, fp32 tile[token][D | 1] in shared memory.
//   phase 1: tile  = y0 + y1        row-major planes, float4 = 4 consecutive w
//   phase 2: tile += y2 + y3        column-major planes, float4 = 4 consecutive h
//   phase 3: warp per token: two-pass LayerNorm over the D channels, affine, optional gate, channel-last store
template <typename TO, int TH, int TW>
__global__ void __launch_bounds__(256) ss2d_merge_norm_kernel(const float* __restrict__ ys, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const TO* __restrict__ zact,
                                                              TO* __restrict__ out, int D, int H, int W, int tiles_w,
                                                              int tiles_h, float eps) {
    extern __shared__ __align__(16) float tile[];
    const int P = D | 1;                                   // odd pitch: conflict-free token-major writes
    const int64_t L = (int64_t)H * W;
    int64_t t = blockIdx.x;
    const int tw = (int)(t % tiles_w); t /= tiles_w;
    const int th = (int)(t % tiles_h); t /= tiles_h;
    const int64_t b = t;
    const int h0 = th * TH, w0 = tw * TW;
    const int tid = threadIdx.x;
    const float* y0 = ys + (b * 4 + 0) * D * L;
    const float* y1 = ys + (b * 4 + 1) * D * L;
    const float* y2 = ys + (b * 4 + 2) * D * L;
    const float* y3 = ys + (b * 4 + 3) * D * L;
    constexpr int QW = TW / 4, QH = TH / 4;
    // phase 1
    for (int i = tid; i < D * TH * QW; i += 256) {
        const int q = i % QW, hh = (i / QW) % TH, c = i / (QW * TH);
        const int h = h0 + hh, w = w0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h < H && w < W) {                               // W % 4 == 0: quads are all-or-nothing
            const int64_t o = (int64_t)c * L + (int64_t)h * W + w;
            const float4 a = __ldg(reinterpret_cast<const float4*>(y0 + o));
            const float4 r = __ldg(reinterpret_cast<const float4*>(y1 + o));
            v = make_float4(a.x + r.x, a.y + r.y, a.z + r.z, a.w + r.w);
        }
        float* dst = tile + (hh * TW + 4 * q) * P + c;
        dst[0] = v.x; dst[P] = v.y; dst[2 * P] = v.z; dst[3 * P] = v.w;
    }
    __syncthreads();
    // phase 2
    for (int i = tid; i < D * TW * QH; i += 256) {
        const int q = i % QH, ww = (i / QH) % TW, c = i / (QH * TW);
        const int w = w0 + ww, h = h0 + 4 * q;
        if (h < H && w < W) {
            const int64_t o = (int64_t)c * L + (int64_t)w * H + h;
            const float4 a = __ldg(reinterpret_cast<const float4*>(y2 + o));
            const float4 r = __ldg(reinterpret_cast<const float4*>(y3 + o));
            float* dst = tile + ((4 * q) * TW + ww) * P + c;
            dst[0] += a.x + r.x; dst[TW * P] += a.y + r.y; dst[2 * TW * P] += a.z + r.z; dst[3 * TW * P] += a.w + r.w;
        }
    }
    __syncthreads();
    // phase 3
    const int warp = tid >> 5, lane = tid & 31;
    for (int tok = warp; tok < TH * TW; tok += 8) {
        const int h = h0 + tok / TW, w = w0 + tok % TW;
        if (h >= H || w >= W) continue;                     // warp-uniform
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

template <typename TO, int TH, int TW>
static int merge_norm_launch(const float* ys, const float* gamma, const float* beta, const void* zact, void* out, int64_t B,
                             int64_t D, int64_t H, int64_t W, float eps, cudaStream_t st) {
    const int tiles_w = (int)ceil_div(W, TW), tiles_h = (int)ceil_div(H, TH);
    const int smem = TH * TW * (int)(D | 1) * 4;
    auto kern = ss2d_merge_norm_kernel<TO, TH, TW>;
    XP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(unsigned)(B * tiles_h * tiles_w), 256, smem, st>>>(ys, gamma, beta, (const TO*)zact, (TO*)out, (int)D, (int)H, (int)W,
                                                                 tiles_w, tiles_h, eps);
    XP_LAUNCH_CHECK("ss2d_merge_norm_kernel");
    return XP_OK;
}

template <typename TO>
static int merge_norm_dispatch(const float* ys, const float* gamma, const float* beta, const void* zact, void* out, int64_t B,
                               int64_t D, int64_t H, int64_t W, float eps, cudaStream_t st) {
    // 8x8 tokens while the fp32 tile fits comfortably (two CTAs per SM up to D = 384), 4x8 / 4x4 for wide blocks
    if (D <= 768) return merge_norm_launch<TO, 8, 8>(ys, gamma, beta, zact, out, B, D, H, W, eps, st);
    if (D <= 1536) return merge_norm_launch<TO, 4, 8>(ys, gamma, beta, zact, out, B, D, H, W, eps, st);
    return merge_norm_launch<TO, 4, 4>(ys, gamma, beta, zact, out, B, D, H, W, eps, st);
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
#define XP_DW(T, S) ss2d_dwconv_pack_kernel<T, S><<<(unsigned)grid, 256, 0, st>>>((const T*)in, weight, bias, (T*)xx, D, (int)H, \
                                                                                   (int)W, in_token_stride, tiles_w, tiles_h, cbs)
    if (dtype == XP_F32) { if (silu) XP_DW(float, true); else XP_DW(float, false); }
    else if (dtype == XP_F16) { if (silu) XP_DW(__half, true); else XP_DW(__half, false); }
    else { if (silu) XP_DW(__nv_bfloat16, true); else XP_DW(__nv_bfloat16, false); }
#undef XP_DW
    XP_LAUNCH_CHECK("ss2d_dwconv_pack_kernel");
    return XP_OK;
}

extern "C" int xp_ss2d_merge_norm(const float* ys, const float* gamma, const float* beta, const void* zact, void* out,
                                  int64_t B, int64_t D, int64_t H, int64_t W, int32_t out_dtype, float eps,
                                  xp_stream_t stream) {
    XP_REQUIRE(ys && gamma && beta && out, "xp_ss2d_merge_norm: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && D > 0 && H > 0 && W > 0, "xp_ss2d_merge_norm: bad shape");
    XP_REQUIRE(H % 4 == 0 && W % 4 == 0, "xp_ss2d_merge_norm: H and W must be multiples of 4 (got %lld x %lld)", (long long)H,
               (long long)W);
    XP_REQUIRE(D <= 3072, "xp_ss2d_merge_norm: D must be <= 3072 (got %lld)", (long long)D);
    XP_REQUIRE(out_dtype >= XP_F32 && out_dtype <= XP_BF16, "xp_ss2d_merge_norm: unsupported dtype %d", out_dtype);
    XP_REQUIRE((reinterpret_cast<uintptr_t>(ys) & 15) == 0, "xp_ss2d_merge_norm: ys must be 16-byte aligned");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (out_dtype) {
        case XP_F32: return merge_norm_dispatch<float>(ys, gamma, beta, zact, out, B, D, H, W, eps, st);
        case XP_F16: return merge_norm_dispatch<__half>(ys, gamma, beta, zact, out, B, D, H, W, eps, st);
        default: return merge_norm_dispatch<__nv_bfloat16>(ys, gamma, beta, zact, out, B, D, H, W, eps, st);
    }
}
