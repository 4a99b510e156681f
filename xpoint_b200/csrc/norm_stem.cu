// norm_stem.cu -- the glue kernels between the library GEMMs / convolutions of the VMamba encoder.
//
//  * xp_add_layer_norm: [residual add] + [per-channel bias] + LayerNorm in one pass over channel-last rows.
//      VSSBlock computes x = x + branch; n = LayerNorm(x) twice per block (VMamba.py:1222-1234) and the patch-embed /
//      downsample convolutions are followed by bias-add, permute and LayerNorm (VMamba.py:1405-1440): each of those is
//      2-3 full passes over the activations in the reference; here it is one (read branch + residual, write the new
//      residual and its normalised copy).
//  * xp_patch_embed_stem: Conv2d(Cin -> C1, 3x3, stride 2, pad 1) + bias + LayerNorm(C1) + GELU, channel-last output
//      (the first half of VMamba's patch-embed v2, VMamba.py:1405-1413), reading the image directly.
#include "common.cuh"

namespace xp {

__device__ __forceinline__ float ld_as_f32(const void* p, int dt, int64_t i) {
    if (dt == XP_F32) return reinterpret_cast<const float*>(p)[i];
    if (dt == XP_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_from_f32(void* p, int dt, int64_t i, float v) {
    if (dt == XP_F32) reinterpret_cast<float*>(p)[i] = v;
    else if (dt == XP_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

struct AddLnParams {
    const void* x; const void* res; const float* pre_bias; const float* gamma; const float* beta;
    void* y; void* sum_out;
    int64_t rows; int C; float eps;
    int x_dt, res_dt, y_dt, sum_dt;
};

__device__ __forceinline__ float4 ld4_as_f32(const void* p, int dt, int64_t i4) {     // elements 4*i4 .. 4*i4+3
    if (dt == XP_F32) return __ldg(reinterpret_cast<const float4*>(p) + i4);
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p) + i4);
    if (dt == XP_F16) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                       __uint_as_float(raw.y & 0xffff0000u));
}
__device__ __forceinline__ void st4_from_f32(void* p, int dt, int64_t i4, float4 v) {
    if (dt == XP_F32) { reinterpret_cast<float4*>(p)[i4] = v; return; }
    uint2 raw;
    if (dt == XP_F16) {
        *reinterpret_cast<__half2*>(&raw.x) = __floats2half2_rn(v.x, v.y);
        *reinterpret_cast<__half2*>(&raw.y) = __floats2half2_rn(v.z, v.w);
    } else {
        *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(v.z, v.w);
    }
    reinterpret_cast<uint2*>(p)[i4] = raw;
}

// LPR lanes per row (power of two), each lane owns CH interleaved 4-element chunks held in registers; a warp covers
// 32 / LPR rows, so short rows (C = 48 .. 192) still move 8-16 bytes per lane per access.  Two-pass variance (torch).
template <int CH>
__global__ void __launch_bounds__(256) add_layer_norm_kernel(const AddLnParams p, int lpr) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & (lpr - 1), rows_per_warp = 32 / lpr;
    const int64_t row = ((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5)) * rows_per_warp + lane / lpr;
    const bool live = row < p.rows;
    const int nch = p.C >> 2;                       // chunks per row (C % 4 == 0)
    const int64_t o4 = row * nch;
    float4 v[CH];
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        const int c4 = q * lpr + sub;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && c4 < nch) {
            t = ld4_as_f32(p.x, p.x_dt, o4 + c4);
            if (p.res) { const float4 r = ld4_as_f32(p.res, p.res_dt, o4 + c4); t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
            if (p.pre_bias) { const float4 r = __ldg(reinterpret_cast<const float4*>(p.pre_bias) + c4); t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
            if (p.sum_out) st4_from_f32(p.sum_out, p.sum_dt, o4 + c4, t);
        }
        v[q] = t;
        s += (t.x + t.y) + (t.z + t.w);
    }
    if (!p.y) return;
    for (int o = lpr >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)p.C;
    float ss = 0.0f;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        if (q * lpr + sub < nch) {
            v[q].x -= mean; v[q].y -= mean; v[q].z -= mean; v[q].w -= mean;
            ss = fmaf(v[q].x, v[q].x, fmaf(v[q].y, v[q].y, fmaf(v[q].z, v[q].z, fmaf(v[q].w, v[q].w, ss))));
        }
    }
    for (int o = lpr >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)p.C + p.eps);
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        const int c4 = q * lpr + sub;
        if (live && c4 < nch) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(p.beta) + c4);
            st4_from_f32(p.y, p.y_dt, o4 + c4, make_float4(fmaf(v[q].x * rstd, g.x, b.x), fmaf(v[q].y * rstd, g.y, b.y),
                                                            fmaf(v[q].z * rstd, g.z, b.z), fmaf(v[q].w * rstd, g.w, b.w)));
        }
    }
}


// ---- compile-time specialisation of the kernel above for the combinations the XPoint encoder issues (18 launches per
// step): dtypes, lanes per row and the presence of every optional operand are template parameters, chunk indices are 32-bit.
// The generic kernel spends 33 instructions per element on run-time dtype dispatch, 64-bit indexing and rolled shuffle
// loops and is issue-bound (76 % busy, ncu); this one is bound by its four memory streams.
template <int DT> __device__ __forceinline__ float4 ld4_t(const void* p, int i4) {
    if constexpr (DT == XP_F32) return __ldg(reinterpret_cast<const float4*>(p) + i4);
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p) + i4);
    if constexpr (DT == XP_F16) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u), __uint_as_float(raw.y << 16),
                       __uint_as_float(raw.y & 0xffff0000u));
}
template <int DT> __device__ __forceinline__ void st4_t(void* p, int i4, float4 v) {
    if constexpr (DT == XP_F32) { reinterpret_cast<float4*>(p)[i4] = v; return; }
    uint2 raw;
    if constexpr (DT == XP_F16) {
        *reinterpret_cast<__half2*>(&raw.x) = __floats2half2_rn(v.x, v.y);
        *reinterpret_cast<__half2*>(&raw.y) = __floats2half2_rn(v.z, v.w);
    } else {
        *reinterpret_cast<__nv_bfloat162*>(&raw.x) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162*>(&raw.y) = __floats2bfloat162_rn(v.z, v.w);
    }
    reinterpret_cast<uint2*>(p)[i4] = raw;
}

// XDT: dtype of x; RDT / SDT / YDT: dtype of res / sum_out / y or -1 when absent; BIAS: pre_bias present
template <int CH, int LPR, int XDT, int RDT, int SDT, int YDT, bool BIAS>
__global__ void __launch_bounds__(256) add_layer_norm_fast_kernel(const AddLnParams p) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane & (LPR - 1);
    const int row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;      // rows * C / 4 < 2^31 (host-checked)
    const bool live = row < (int)p.rows;
    const int nch = p.C >> 2;
    const int o4 = row * nch;
    float4 v[CH];
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        const int c4 = q * LPR + sub;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live && c4 < nch) {
            t = ld4_t<XDT>(p.x, o4 + c4);
            if constexpr (RDT >= 0) { const float4 r = ld4_t<RDT>(p.res, o4 + c4); t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
            if constexpr (BIAS) { const float4 r = __ldg(reinterpret_cast<const float4*>(p.pre_bias) + c4); t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w; }
            if constexpr (SDT >= 0) st4_t<SDT>(p.sum_out, o4 + c4, t);
        }
        v[q] = t;
        s += (t.x + t.y) + (t.z + t.w);
    }
    if constexpr (YDT >= 0) {
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)p.C;
    float ss = 0.0f;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        if (q * LPR + sub < nch) {
            v[q].x -= mean; v[q].y -= mean; v[q].z -= mean; v[q].w -= mean;
            ss = fmaf(v[q].x, v[q].x, fmaf(v[q].y, v[q].y, fmaf(v[q].z, v[q].z, fmaf(v[q].w, v[q].w, ss))));
        }
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss / (float)p.C + p.eps);
#pragma unroll
    for (int q = 0; q < CH; ++q) {
        const int c4 = q * LPR + sub;
        if (live && c4 < nch) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(p.beta) + c4);
            st4_t<YDT>(p.y, o4 + c4, make_float4(fmaf(v[q].x * rstd, g.x, b.x), fmaf(v[q].y * rstd, g.y, b.y),
                                                                       fmaf(v[q].z * rstd, g.z, b.z), fmaf(v[q].w * rstd, g.w, b.w)));
        }
    }
    }
}

// launches the specialised kernel if (ch, lpr, dtypes, operands) is one of the encoder's combinations; false otherwise
template <int XDT, int RDT, int SDT, int YDT, bool BIAS>
static bool add_ln_fast_shape(const AddLnParams& p, int ch, int lpr, unsigned grid, cudaStream_t st) {
    if (ch == 3 && lpr == 8) add_layer_norm_fast_kernel<3, 8, XDT, RDT, SDT, YDT, BIAS><<<grid, 256, 0, st>>>(p);
    else if (ch == 3 && lpr == 16) add_layer_norm_fast_kernel<3, 16, XDT, RDT, SDT, YDT, BIAS><<<grid, 256, 0, st>>>(p);
    else if (ch == 3 && lpr == 32) add_layer_norm_fast_kernel<3, 32, XDT, RDT, SDT, YDT, BIAS><<<grid, 256, 0, st>>>(p);
    else if (ch == 6 && lpr == 32) add_layer_norm_fast_kernel<6, 32, XDT, RDT, SDT, YDT, BIAS><<<grid, 256, 0, st>>>(p);
    else return false;
    return true;
}
template <int LO>   // LO: the 16-bit dtype of the activations (XP_F16 | XP_BF16), or XP_F32 for the fp32 model
static bool add_ln_fast(const AddLnParams& p, int ch, int lpr, unsigned grid, cudaStream_t st) {
    const bool has_res = p.res != nullptr, has_sum = p.sum_out != nullptr, has_y = p.y != nullptr, has_bias = p.pre_bias != nullptr;
    if (p.x_dt != LO) return false;
    if constexpr (LO == XP_F32) {
        // LayerNorm of the fp32 residual stream into the activation dtype: the first norm of a stage
        if (!has_res && !has_sum && has_y && !has_bias && p.y_dt == XP_F16) return add_ln_fast_shape<XP_F32, -1, -1, XP_F16, false>(p, ch, lpr, grid, st);
        if (!has_res && !has_sum && has_y && !has_bias && p.y_dt == XP_BF16) return add_ln_fast_shape<XP_F32, -1, -1, XP_BF16, false>(p, ch, lpr, grid, st);
    }
    // x + res -> sum (fp32) and LayerNorm -> y (activation dtype): every VSSBlock branch
    if (has_res && p.res_dt == XP_F32 && has_sum && p.sum_dt == XP_F32 && has_y && p.y_dt == LO && !has_bias)
        return add_ln_fast_shape<LO, XP_F32, XP_F32, LO, false>(p, ch, lpr, grid, st);
    // conv output + bias -> LayerNorm -> fp32 residual stream: patch-embed / downsample tails
    if (!has_res && !has_sum && has_y && p.y_dt == XP_F32 && has_bias)
        return add_ln_fast_shape<LO, -1, -1, XP_F32, true>(p, ch, lpr, grid, st);
    // x + res -> sum in the activation dtype, no norm: the residual sum in front of a downsample convolution
    if (has_res && p.res_dt == XP_F32 && has_sum && p.sum_dt == LO && !has_y && !has_bias)
        return add_ln_fast_shape<LO, XP_F32, LO, -1, false>(p, ch, lpr, grid, st);
    return false;
}

// scalar fallback (C % 4 != 0 or unaligned pointers): warp per row
template <int PER>
__global__ void __launch_bounds__(256) add_layer_norm_scalar_kernel(const AddLnParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= p.rows) return;
    const int C = p.C;
    const int64_t o = row * C;
    float v[PER];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        float t = 0.0f;
        if (c < C) {
            t = ld_as_f32(p.x, p.x_dt, o + c);
            if (p.res) t += ld_as_f32(p.res, p.res_dt, o + c);
            if (p.pre_bias) t += p.pre_bias[c];
            if (p.sum_out) st_from_f32(p.sum_out, p.sum_dt, o + c, t);
        }
        v[i] = t;
        s += t;
    }
    if (!p.y) return;
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const float d = (lane + 32 * i < C) ? v[i] - mean : 0.0f;
        v[i] = d;
        ss = fmaf(d, d, ss);
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + p.eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = lane + 32 * i;
        if (c < C) st_from_f32(p.y, p.y_dt, o + c, fmaf(v[i] * rstd, p.gamma[c], p.beta[c]));
    }
}

// ------------------------------------------------------------------------------------------ patch-embed stem
// One thread = one output token, all C1 (<= 64) channels in registers; weights broadcast from shared memory.
constexpr int STEM_MAXC = 64;

// CT: compile-time channel count (48 for XPoint: registers and loops sized exactly), 0 = runtime C1 <= STEM_MAXC
template <int CIN, int CT>
__global__ void __launch_bounds__(128) patch_embed_stem_kernel(const float* __restrict__ img, const float* __restrict__ wgt,
                                                               const float* __restrict__ bias, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, void* __restrict__ out, int B,
                                                               int H, int W, int C1, float eps, int out_dt, int gelu) {
    __shared__ __align__(16) float sw[CIN * 9 * STEM_MAXC];     // [cin][tap][channel], channels padded to STEM_MAXC
    __shared__ float sb[STEM_MAXC], sg[STEM_MAXC], sbe[STEM_MAXC];
    constexpr int CC = CT ? CT : STEM_MAXC;                     // accumulators per thread
    if (CT) C1 = CT;
    __shared__ uint4 stage[128 * (CC / 8)];                     // 16-bit outputs leave through a per-warp transpose
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;               // k3 s2 p1
    for (int i = threadIdx.x; i < CIN * 9 * STEM_MAXC; i += blockDim.x) {
        const int c = i % STEM_MAXC, tap = (i / STEM_MAXC) % 9, ci = i / (STEM_MAXC * 9);
        sw[i] = c < C1 ? wgt[((int64_t)c * CIN + ci) * 9 + tap] : 0.0f;
    }
    for (int c = threadIdx.x; c < STEM_MAXC; c += blockDim.x) {
        sb[c] = (c < C1 && bias) ? bias[c] : 0.0f;
        sg[c] = c < C1 ? gamma[c] : 0.0f;
        sbe[c] = c < C1 ? beta[c] : 0.0f;
    }
    __syncthreads();
    const int64_t total = (int64_t)B * Ho * Wo;
    const int64_t tok_raw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = tok_raw < total;
    const int64_t tok = live ? tok_raw : total - 1;             // dead threads of the last CTA redo the last token (no store)
    const int wo = (int)(tok % Wo), ho = (int)((tok / Wo) % Ho);
    const int64_t b = tok / ((int64_t)Wo * Ho);
    float acc[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) acc[c] = sb[c];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
        const float* plane = img + (b * CIN + ci) * (int64_t)H * W;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int h = 2 * ho - 1 + ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int w = 2 * wo - 1 + kx;
                const float px = (h >= 0 && h < H && w >= 0 && w < W) ? __ldg(plane + (int64_t)h * W + w) : 0.0f;
                const float4* wv = reinterpret_cast<const float4*>(sw + (ci * 9 + ky * 3 + kx) * STEM_MAXC);
#pragma unroll
                for (int q = 0; q < CC / 4; ++q) {
                    if (4 * q < C1) {                      // uniform: skips the padded channel quads
                        const float4 w4 = wv[q];           // two packed FFMA2 per quad: half the issue slots of four FFMA
                        const float2 lo = fma2(make_float2(px, px), make_float2(w4.x, w4.y), make_float2(acc[4 * q], acc[4 * q + 1]));
                        const float2 hi = fma2(make_float2(px, px), make_float2(w4.z, w4.w), make_float2(acc[4 * q + 2], acc[4 * q + 3]));
                        acc[4 * q] = lo.x; acc[4 * q + 1] = lo.y; acc[4 * q + 2] = hi.x; acc[4 * q + 3] = hi.y;
                    }
                }
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < CC; ++c) s += c < C1 ? acc[c] : 0.0f;
    const float mean = s / (float)C1;
    float ss = 0.0f;
#pragma unroll
    for (int c = 0; c < CC; ++c) {
        const float d = c < C1 ? acc[c] - mean : 0.0f;
        acc[c] = d;
        ss = fmaf(d, d, ss);
    }
    const float rstd = rsqrtf(ss / (float)C1 + eps);
    const int64_t o = tok * C1;
#pragma unroll
    for (int c = 0; c < CC; ++c) {
        if (c < C1) {
            float v = fmaf(acc[c] * rstd, sg[c], sbe[c]);
            if (gelu) v = gelu_erf_f(v);                                     // exact (erf) GELU, nn.GELU default
            acc[c] = v;
        }
    }
    if (out_dt == XP_F32) {
        float* dst = reinterpret_cast<float*>(out) + o;
        if (live) {
#pragma unroll
            for (int c = 0; c < CC; ++c) if (c < C1) dst[c] = acc[c];
        }
    } else if (C1 % 8 == 0) {
        // the warp's 32 tokens are one contiguous run of 32 * C1 * 2 bytes: pack to 16-byte chunks, transpose through shared
        // memory and store whole 512-byte lines per instruction (a thread-per-token store touches 32 half-used sectors)
        const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nq = C1 / 8;
        uint4* wst = stage + wrp * 32 * (CC / 8);
#pragma unroll
        for (int q = 0; q < CC / 8; ++q) {
            if (8 * q < C1) {
                uint4 raw;
                if (out_dt == XP_F16) {
                    __half2* h = reinterpret_cast<__half2*>(&raw);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(acc[8 * q + 2 * j], acc[8 * q + 2 * j + 1]);
                } else {
                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(acc[8 * q + 2 * j], acc[8 * q + 2 * j + 1]);
                }
                wst[lane * nq + q] = raw;
            }
        }
        __syncwarp();
        const int64_t tok0 = (int64_t)blockIdx.x * blockDim.x + wrp * 32;           // first token of the warp
        const int64_t nvalid = min((int64_t)32, total - tok0);                    // <= 0 for fully dead warps
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned short*>(out) + tok0 * C1);
        for (int i = lane; i < nvalid * nq; i += 32) dst[i] = wst[i];
    } else if (live) {
#pragma unroll
        for (int c = 0; c < CC; ++c) if (c < C1) st_from_f32(out, out_dt, o + c, acc[c]);
    }
}

}  // namespace xp

using namespace xp;

extern "C" int xp_add_layer_norm(const void* x, const void* res, const float* pre_bias, const float* gamma, const float* beta,
                                 void* y, void* sum_out, int64_t rows, int64_t C, int32_t x_dtype, int32_t res_dtype,
                                 int32_t y_dtype, int32_t sum_dtype, float eps, xp_stream_t stream) {
    XP_REQUIRE(x, "xp_add_layer_norm: x is NULL");
    XP_REQUIRE(y || sum_out, "xp_add_layer_norm: at least one of y / sum_out must be given");
    XP_REQUIRE(!y || (gamma && beta), "xp_add_layer_norm: gamma / beta are required when y is given");
    XP_REQUIRE(rows >= 0 && C > 0 && C <= 1536, "xp_add_layer_norm: need 0 < C <= 1536 (got %lld)", (long long)C);
    for (int dt : {x_dtype, res ? res_dtype : XP_F32, y ? y_dtype : XP_F32, sum_out ? sum_dtype : XP_F32})
        XP_REQUIRE(dt >= XP_F32 && dt <= XP_BF16, "xp_add_layer_norm: unsupported dtype %d", dt);
    if (rows == 0) return XP_OK;
    AddLnParams p{x, res, pre_bias, gamma, beta, y, sum_out, rows, (int)C, eps, x_dtype, res_dtype, y_dtype, sum_dtype};
    cudaStream_t st = (cudaStream_t)stream;
    bool vec = C % 4 == 0;
    for (const void* q : {x, res, (const void*)pre_bias, (const void*)gamma, (const void*)beta, (const void*)y, (const void*)sum_out})
        vec = vec && (reinterpret_cast<uintptr_t>(q) & 15) == 0;
    if (vec) {
        const int nch = (int)(C / 4);
        int lpr = 2;
        while (lpr < 32 && lpr * 3 < nch) lpr <<= 1;          // <= 3 chunks per lane until the warp is one row wide
        const int ch = (nch + lpr - 1) / lpr;
        const unsigned grid = (unsigned)ceil_div(rows, 8 * (32 / lpr));
        if (rows * nch < ((int64_t)1 << 31) - 8 * 32 * nch) {              // 32-bit chunk indices (incl. the last CTA's dead rows)
            bool done = x_dtype == XP_F16 ? add_ln_fast<XP_F16>(p, ch, lpr, grid, st)
                      : x_dtype == XP_BF16 ? add_ln_fast<XP_BF16>(p, ch, lpr, grid, st) : add_ln_fast<XP_F32>(p, ch, lpr, grid, st);
            if (done) {
                XP_LAUNCH_CHECK("add_layer_norm_fast_kernel");
                return XP_OK;
            }
        }
        if (ch <= 1) add_layer_norm_kernel<1><<<grid, 256, 0, st>>>(p, lpr);
        else if (ch <= 2) add_layer_norm_kernel<2><<<grid, 256, 0, st>>>(p, lpr);
        else if (ch <= 3) add_layer_norm_kernel<3><<<grid, 256, 0, st>>>(p, lpr);
        else if (ch <= 6) add_layer_norm_kernel<6><<<grid, 256, 0, st>>>(p, lpr);
        else add_layer_norm_kernel<12><<<grid, 256, 0, st>>>(p, lpr);
    } else {
        const unsigned grid = (unsigned)ceil_div(rows, 8);
        const int per = (int)ceil_div(C, 32);
        if (per <= 3) add_layer_norm_scalar_kernel<3><<<grid, 256, 0, st>>>(p);
        else if (per <= 12) add_layer_norm_scalar_kernel<12><<<grid, 256, 0, st>>>(p);
        else add_layer_norm_scalar_kernel<48><<<grid, 256, 0, st>>>(p);
    }
    XP_LAUNCH_CHECK("add_layer_norm_kernel");
    return XP_OK;
}

extern "C" int xp_patch_embed_stem(const float* img, const float* weight, const float* bias, const float* gamma,
                                   const float* beta, void* out, int64_t B, int64_t Cin, int64_t H, int64_t W, int64_t C1,
                                   float eps, int32_t out_dtype, int32_t gelu, xp_stream_t stream) {
    XP_REQUIRE(img && weight && gamma && beta && out, "xp_patch_embed_stem: NULL tensor pointer");
    XP_REQUIRE(Cin == 1 || Cin == 3, "xp_patch_embed_stem: Cin must be 1 or 3 (got %lld)", (long long)Cin);
    XP_REQUIRE(C1 > 0 && C1 <= STEM_MAXC && C1 % 4 == 0, "xp_patch_embed_stem: C1 must be a multiple of 4, <= %d (got %lld)",
               STEM_MAXC, (long long)C1);
    XP_REQUIRE(B >= 0 && H > 0 && W > 0, "xp_patch_embed_stem: bad shape");
    XP_REQUIRE(out_dtype >= XP_F32 && out_dtype <= XP_BF16, "xp_patch_embed_stem: unsupported dtype %d", out_dtype);
    if (B == 0) return XP_OK;
    const int64_t total = B * ((H + 1) / 2) * ((W + 1) / 2);
    const unsigned grid = (unsigned)ceil_div(total, 128);
    cudaStream_t st = (cudaStream_t)stream;
#define XP_STEM(CIN_, CT_) patch_embed_stem_kernel<CIN_, CT_><<<grid, 128, 0, st>>>(img, weight, bias, gamma, beta, out, (int)B, (int)H, \
                                                                               (int)W, (int)C1, eps, out_dtype, gelu)
    if (Cin == 1) { if (C1 == 48) XP_STEM(1, 48); else XP_STEM(1, 0); }
    else { if (C1 == 48) XP_STEM(3, 48); else XP_STEM(3, 0); }
#undef XP_STEM
    XP_LAUNCH_CHECK("patch_embed_stem_kernel");
    return XP_OK;
}
