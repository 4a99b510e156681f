// heads.cu -- channel-last forms of the head post-processing and the encoder tail that feeds the heads.
//
// The two head branches (XPoint.py:112-138, 348-371) end in 1x1 convolutions, which the drop-in model runs as GEMMs on
// channel-last activations.  These kernels consume the GEMM outputs where they lie, so the (B, C, Hc, Wc) permute copies
// of the reference layout never hit HBM:
//   xp_detector_post_cl  Softmax2d -> [:, :-1] -> PixelShuffle(r)  (XPoint.py:356-357) from (cells, ld) rows
//   xp_l2_normalize_cl   F.normalize(dim=1)                       (XPoint.py:365-366) from (cells, C) rows
//   xp_encoder_tail      x + branch -> permute -> depth_to_space(4) (VMamba.py:1500-1505,1521-1523) -> clone
//                        (XPoint.py:309) -> ReflectionPad2d(1) -> 16-bit channels-last (the heads' first convolution input)
// All HBM-bound, small tensors: one coalesced pass each.
#include "common.cuh"

namespace xp {

// thread = one cell: the r*r+1 logits of a cell are contiguous
template <typename T, int R>
__global__ void __launch_bounds__(128) detector_post_cl_kernel(const T* __restrict__ logits, float* __restrict__ prob,
                                                               int64_t cells, int Hc, int Wc, int64_t ld) {
    constexpr int CN = R * R + 1;
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= cells) return;
    const int64_t HW = (int64_t)Hc * Wc;
    const int64_t b = cell / HW;
    const int hw = (int)(cell % HW), h = hw / Wc, w = hw % Wc;
    const T* lg = logits + cell * ld;
    float e[CN];
    float m = -INFINITY;
    constexpr int PER = 16 / (int)sizeof(T);
    if (ld % PER == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {   // 16-byte row loads
#pragma unroll
        for (int v = 0; v * PER < CN; ++v) {
            float f[PER];
            VecIO<T, PER>::load(lg + v * PER, f);
#pragma unroll
            for (int i = 0; i < PER; ++i)
                if (v * PER + i < CN) e[v * PER + i] = f[i];
        }
    } else {
#pragma unroll
        for (int c = 0; c < CN; ++c) e[c] = to_f32(lg[c]);
    }
#pragma unroll
    for (int c = 0; c < CN; ++c) m = fmaxf(m, e[c]);
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CN; ++c) { e[c] = expf(e[c] - m); s += e[c]; }
    const int64_t W = (int64_t)Wc * R;
    float* out = prob + (b * Hc * R + (int64_t)h * R) * W + (int64_t)w * R;
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int j = 0; j < R; ++j) out[i * W + j] = e[i * R + j] / s;
    }
}

// CTA = 32 cells x all channels (tile in smem); coalesced channel-last reads, both output layouts optional
template <typename T>
__global__ void __launch_bounds__(256) l2_normalize_cl_kernel(const T* __restrict__ x, float* __restrict__ out_cf,
                                                              float* __restrict__ out_cl, int C, int64_t HW) {
    extern __shared__ float tile[];                 // [C][33]
    __shared__ float inv_norm[32];
    const int64_t b = blockIdx.y, p0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int px = ty; px < 32; px += 8) {           // warp ty: cells ty, ty+8, ...; lanes stride over channels
        const int64_t p = p0 + px;
        float s = 0.0f;
        for (int c = tx; c < C; c += 32) {
            const float v = p < HW ? to_f32(x[(b * HW + p) * C + c]) : 0.0f;
            tile[c * 33 + px] = v;
            s += v * v;
        }
        s = warp_sum(s);
        const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
        if (tx == 0) inv_norm[px] = inv;
        if (out_cl && p < HW)
            for (int c = tx; c < C; c += 32) out_cl[(b * HW + p) * C + c] = tile[c * 33 + px] * inv;
    }
    if (!out_cf) return;
    __syncthreads();
    for (int c = ty; c < C; c += 8) {
        const int64_t p = p0 + tx;
        if (p < HW) out_cf[(b * C + c) * HW + p] = tile[c * 33 + tx] * inv_norm[tx];
    }
}

// thread = one output pixel (Y, X) of the depth-to-space image, all CO channels
template <typename XT, typename PT, typename OT, int CO>
__global__ void __launch_bounds__(128) encoder_tail_kernel(const XT* __restrict__ x, const PT* __restrict__ pend,
                                                           float* __restrict__ enc, OT* __restrict__ padded, int64_t B, int H,
                                                           int W, int bs) {
    const int Ho = H * bs, Wo = W * bs;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= B * Ho * Wo) return;
    const int X = (int)(pix % Wo), Y = (int)((pix / Wo) % Ho);
    const int64_t b = pix / ((int64_t)Wo * Ho);
    const int Cin = CO * bs * bs;
    // out[b, c, Y, X] = in[b, Y / bs, X / bs, ((Y % bs) * bs + X % bs) * CO + c]
    const int64_t src = ((b * H + Y / bs) * W + X / bs) * Cin + ((Y % bs) * bs + X % bs) * CO;
    float v[CO];
#pragma unroll
    for (int c = 0; c < CO; c += 4) {
        float f[4];
        VecIO<float, 4>::load(reinterpret_cast<const float*>(x) + src + c, f);     // XT == float (checked on the host)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[c + i] = f[i];
    }
    if (pend) {
#pragma unroll
        for (int c = 0; c < CO; ++c) v[c] += to_f32(pend[src + c]);
    }
    if (enc) {
#pragma unroll
        for (int c = 0; c < CO; ++c) enc[((b * CO + c) * Ho + Y) * Wo + X] = v[c];
    }
    if (padded) {
        // ReflectionPad2d(1): padded row py holds source row |py - 1| mirrored at both ends (0 <- 1, Ho + 1 <- Ho - 2)
        const int Hp = Ho + 2, Wp = Wo + 2;
        OT o[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c) o[c] = from_f32<OT>(v[c]);
        int pys[2] = {Y + 1, -1}, pxs[2] = {X + 1, -1};
        if (Y == 1) pys[1] = 0;
        if (Y == Ho - 2) pys[1] = Hp - 1;
        if (X == 1) pxs[1] = 0;
        if (X == Wo - 2) pxs[1] = Wp - 1;
        for (int a = 0; a < 2; ++a) {
            if (pys[a] < 0) continue;
            for (int q = 0; q < 2; ++q) {
                if (pxs[q] < 0) continue;
                OT* dst = padded + ((b * Hp + pys[a]) * Wp + pxs[q]) * CO;
#pragma unroll
                for (int c = 0; c < CO; ++c) dst[c] = o[c];
            }
        }
    }
}

template <typename T> static int det_cl_launch(const void* logits, float* prob, int64_t B, int64_t Hc, int64_t Wc, int r, int64_t ld,
                                               cudaStream_t st) {
    const int64_t cells = B * Hc * Wc;
    const unsigned grid = (unsigned)ceil_div(cells, 128);
    switch (r) {
        case 8: detector_post_cl_kernel<T, 8><<<grid, 128, 0, st>>>((const T*)logits, prob, cells, (int)Hc, (int)Wc, ld); break;
        case 4: detector_post_cl_kernel<T, 4><<<grid, 128, 0, st>>>((const T*)logits, prob, cells, (int)Hc, (int)Wc, ld); break;
        case 2: detector_post_cl_kernel<T, 2><<<grid, 128, 0, st>>>((const T*)logits, prob, cells, (int)Hc, (int)Wc, ld); break;
        default: set_error("xp_detector_post_cl: r must be 2, 4 or 8 (got %d)", r); return XP_ERR_INVALID_ARG;
    }
    XP_LAUNCH_CHECK("detector_post_cl_kernel");
    return XP_OK;
}

template <typename T> static int l2_cl_launch(const void* x, float* out_cf, float* out_cl, int64_t B, int64_t C, int64_t HW,
                                              cudaStream_t st) {
    const size_t smem = (size_t)C * 33 * sizeof(float);
    XP_CUDA_OK(cudaFuncSetAttribute(l2_normalize_cl_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)B);
    l2_normalize_cl_kernel<T><<<grid, 256, smem, st>>>((const T*)x, out_cf, out_cl, (int)C, HW);
    XP_LAUNCH_CHECK("l2_normalize_cl_kernel");
    return XP_OK;
}

template <typename PT, typename OT>
static int enc_tail_launch(const void* x, const void* pend, float* enc, void* padded, int64_t B, int64_t H, int64_t W, int64_t CO,
                           int64_t bs, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(B * H * bs * W * bs, 128);
#define XP_ENC_TAIL(N)                                                                                                   \
    encoder_tail_kernel<float, PT, OT, N><<<grid, 128, 0, st>>>((const float*)x, (const PT*)pend, enc, (OT*)padded, B, (int)H, \
                                                                (int)W, (int)bs)
    switch (CO) {
        case 8: XP_ENC_TAIL(8); break;
        case 16: XP_ENC_TAIL(16); break;
        case 32: XP_ENC_TAIL(32); break;
        case 48: XP_ENC_TAIL(48); break;
        case 64: XP_ENC_TAIL(64); break;
        default:
            set_error("xp_encoder_tail: out channels must be 8, 16, 32, 48 or 64 (got %lld)", (long long)CO);
            return XP_ERR_UNSUPPORTED;
    }
#undef XP_ENC_TAIL
    XP_LAUNCH_CHECK("encoder_tail_kernel");
    return XP_OK;
}

}  // namespace xp

using namespace xp;

extern "C" int xp_detector_post_cl(const void* logits, float* prob, int64_t B, int64_t Hc, int64_t Wc, int32_t r, int64_t ld,
                                   int32_t dtype, xp_stream_t stream) {
    XP_REQUIRE(logits && prob, "xp_detector_post_cl: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && Hc > 0 && Wc > 0 && ld >= (int64_t)r * r + 1, "xp_detector_post_cl: bad shape / row stride");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case XP_F32: return det_cl_launch<float>(logits, prob, B, Hc, Wc, r, ld, st);
        case XP_F16: return det_cl_launch<__half>(logits, prob, B, Hc, Wc, r, ld, st);
        case XP_BF16: return det_cl_launch<__nv_bfloat16>(logits, prob, B, Hc, Wc, r, ld, st);
        default: set_error("xp_detector_post_cl: unsupported dtype %d", dtype); return XP_ERR_INVALID_ARG;
    }
}

extern "C" int xp_l2_normalize_cl(const void* x, float* out_cf, float* out_cl, int64_t B, int64_t C, int64_t HW, int32_t dtype,
                                  xp_stream_t stream) {
    XP_REQUIRE(x && (out_cf || out_cl), "xp_l2_normalize_cl: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && B <= 65535 && C > 0 && C <= 1024 && HW > 0, "xp_l2_normalize_cl: bad shape (C <= 1024, B <= 65535)");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case XP_F32: return l2_cl_launch<float>(x, out_cf, out_cl, B, C, HW, st);
        case XP_F16: return l2_cl_launch<__half>(x, out_cf, out_cl, B, C, HW, st);
        case XP_BF16: return l2_cl_launch<__nv_bfloat16>(x, out_cf, out_cl, B, C, HW, st);
        default: set_error("xp_l2_normalize_cl: unsupported dtype %d", dtype); return XP_ERR_INVALID_ARG;
    }
}

extern "C" int xp_encoder_tail(const void* x, const void* pend, float* enc_out, void* padded, int64_t B, int64_t H, int64_t W,
                               int64_t C_out, int64_t bs, int32_t x_dtype, int32_t pend_dtype, int32_t pad_dtype,
                               xp_stream_t stream) {
    XP_REQUIRE(x && (enc_out || padded), "xp_encoder_tail: NULL tensor pointer");
    XP_REQUIRE(x_dtype == XP_F32, "xp_encoder_tail: the residual stream x must be fp32");
    XP_REQUIRE(B >= 0 && H > 0 && W > 0 && bs >= 1 && bs <= 8 && H * bs >= 3 && W * bs >= 3, "xp_encoder_tail: bad shape");
    XP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "xp_encoder_tail: x must be 16-byte aligned");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int key = (pend ? pend_dtype : XP_F32) * 4 + (padded ? pad_dtype : XP_F32);
    switch (key) {
        case XP_F32 * 4 + XP_F32: return enc_tail_launch<float, float>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        case XP_F32 * 4 + XP_F16: return enc_tail_launch<float, __half>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        case XP_F32 * 4 + XP_BF16: return enc_tail_launch<float, __nv_bfloat16>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        case XP_F16 * 4 + XP_F16: return enc_tail_launch<__half, __half>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        case XP_F16 * 4 + XP_F32: return enc_tail_launch<__half, float>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        case XP_BF16 * 4 + XP_BF16: return enc_tail_launch<__nv_bfloat16, __nv_bfloat16>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        case XP_BF16 * 4 + XP_F32: return enc_tail_launch<__nv_bfloat16, float>(x, pend, enc_out, padded, B, H, W, C_out, bs, st);
        default: break;
    }
    set_error("xp_encoder_tail: unsupported dtype combination pend=%d padded=%d", pend_dtype, pad_dtype);
    return XP_ERR_INVALID_ARG;
}
