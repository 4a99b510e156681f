// cross_scan.cu -- CrossScan / CrossMerge and the fused merge + out_norm (+ gate) tail of SS2D.
//
// Replaces csm_triton.py:22-179 (torch paths) and :278-517 (the Triton kernel triton_cross_scan_flex and
// its autograd wrappers), plus VMamba.py:632-646 / :364-372 (cross_merge -> LayerNorm [-> * z]).
//
//  * cross_scan_tiled_kernel / cross_merge_tiled_kernel: the hot channel-first, scans=0 case.  One CTA moves a
//    32x32 (h, w) tile of one (b, c) plane; the two column-major directions go through a padded shared-memory
//    transpose so every global access is a full 128-byte row.
//  * *_generic_kernel: every other flag combination (channel-last, one_by_one, scans 1|2), one thread per
//    output element, index routing only.
//  * merge_norm_gate: two tiled merge passes into a channel-last fp32 buffer (association (y0+y2')+(y1'+y3'),
//    exactly the torch path) followed by a warp-per-token LayerNorm (+ gate).
#include "common.cuh"

namespace xp {

// position of token (h, w) in the sequence of direction k
__device__ __forceinline__ int64_t scan_pos(int k, int scans, int64_t h, int64_t w, int64_t H, int64_t W) {
    const int64_t L = H * W, l0 = h * W + w;
    if (scans == 0) {
        const int64_t l1 = w * H + h;
        return k == 0 ? l0 : k == 1 ? l1 : k == 2 ? L - 1 - l0 : L - 1 - l1;
    }
    if (scans == 1) return l0;
    return k < 2 ? l0 : L - 1 - l0;
}

// ---------------------------------------------------------------------------------- generic scan
template <typename T>
__global__ void cross_scan_generic_kernel(const T* __restrict__ x, T* __restrict__ xs, int64_t B, int64_t C, int64_t H,
                                          int64_t W, int in_cf, int out_cf, int one_by_one, int scans) {
    const int64_t L = H * W, total = B * 4 * C * L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // decode the SOURCE-ordered index (b, k, c, l0) so that reads of channel-first inputs coalesce
        int64_t t = i;
        int64_t c, l0, k, b;
        if (in_cf) { l0 = t % L; t /= L; c = t % C; t /= C; k = t % 4; b = t / 4; }
        else       { c = t % C; t /= C; k = t % 4; t /= 4; l0 = t % L; b = t / L; }
        const int64_t h = l0 / W, w = l0 % W;
        int64_t src;
        if (in_cf) src = one_by_one ? ((b * 4 + k) * C + c) * L + l0 : (b * C + c) * L + l0;
        else       src = one_by_one ? ((b * L + l0) * 4 + k) * C + c : (b * L + l0) * C + c;
        const int64_t pos = scan_pos((int)k, scans, h, w, H, W);
        const int64_t dst = out_cf ? ((b * 4 + k) * C + c) * L + pos : ((b * L + pos) * 4 + k) * C + c;
        xs[dst] = x[src];
    }
}

// ---------------------------------------------------------------------------------- tiled scan (cf -> cf, scans 0)
template <typename T>
__global__ void __launch_bounds__(256) cross_scan_tiled_kernel(const T* __restrict__ x, T* __restrict__ xs, int64_t C,
                                                               int H, int W) {
    __shared__ T tile[32][33];
    const int64_t L = (int64_t)H * W;
    const int64_t bc = blockIdx.z;                 // b * C + c
    const int64_t b = bc / C, c = bc % C;
    const int h0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const T* src = x + bc * L;
    T* d0 = xs + ((b * 4 + 0) * C + c) * L;
    T* d1 = xs + ((b * 4 + 1) * C + c) * L;
    T* d2 = xs + ((b * 4 + 2) * C + c) * L;
    T* d3 = xs + ((b * 4 + 3) * C + c) * L;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const T v = src[(int64_t)h * W + w];
            tile[r][tx] = v;
            const int64_t l0 = (int64_t)h * W + w;
            d0[l0] = v;
            d2[L - 1 - l0] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;          // lanes run along h: contiguous in the column-major walk
        if (h < H && w < W) {
            const T v = tile[tx][r];
            const int64_t l1 = (int64_t)w * H + h;
            d1[l1] = v;
            d3[L - 1 - l1] = v;
        }
    }
}

// ---------------------------------------------------------------------------------- generic merge
template <typename T>
__global__ void cross_merge_generic_kernel(const T* __restrict__ ys, T* __restrict__ y, int64_t B, int64_t C, int64_t H,
                                           int64_t W, int in_cf, int out_cf, int one_by_one, int scans) {
    // flag names follow the reference: `out_channel_first` describes ys (the merge INPUT), `in_channel_first`
    // the merged result (csm_triton.py:56-85).  Here in_cf/out_cf already mean input/output of this kernel.
    const int64_t L = H * W;
    const int64_t total = (one_by_one ? 4 : 1) * B * C * L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = i, c, l0, b, k = 0;
        if (out_cf) { l0 = t % L; t /= L; c = t % C; t /= C; if (one_by_one) { k = t % 4; t /= 4; } b = t; }
        else        { c = t % C; t /= C; if (one_by_one) { k = t % 4; t /= 4; } l0 = t % L; b = t / L; }
        const int64_t h = l0 / W, w = l0 % W;
        auto at = [&](int kk) -> float {
            const int64_t pos = scan_pos(kk, scans, h, w, H, W);
            const int64_t s = in_cf ? ((b * 4 + kk) * C + c) * L + pos : ((b * L + pos) * 4 + kk) * C + c;
            return to_f32(ys[s]);
        };
        float v;
        if (one_by_one) v = at((int)k);
        else if (scans == 1) v = ((at(0) + at(1)) + at(2)) + at(3);
        else v = (at(0) + at(2)) + (at(1) + at(3));
        int64_t dst;
        if (one_by_one) dst = out_cf ? ((b * 4 + k) * C + c) * L + l0 : ((b * L + l0) * 4 + k) * C + c;
        else dst = out_cf ? (b * C + c) * L + l0 : (b * L + l0) * C + c;
        y[dst] = from_f32<T>(v);
    }
}

// ---------------------------------------------------------------------------------- tiled merge (cf -> cf, scans 0)
template <typename T>
__global__ void __launch_bounds__(256) cross_merge_tiled_kernel(const T* __restrict__ ys, T* __restrict__ y, int64_t C,
                                                                int H, int W) {
    __shared__ float tile[32][33];
    const int64_t L = (int64_t)H * W;
    const int64_t bc = blockIdx.z;
    const int64_t b = bc / C, c = bc % C;
    const int h0 = blockIdx.y * 32, w0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* s0 = ys + ((b * 4 + 0) * C + c) * L;
    const T* s1 = ys + ((b * 4 + 1) * C + c) * L;
    const T* s2 = ys + ((b * 4 + 2) * C + c) * L;
    const T* s3 = ys + ((b * 4 + 3) * C + c) * L;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int w = w0 + r, h = h0 + tx;
        if (h < H && w < W) {
            const int64_t l1 = (int64_t)w * H + h;
            tile[tx][r] = to_f32(s1[l1]) + to_f32(s3[L - 1 - l1]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, w = w0 + tx;
        if (h < H && w < W) {
            const int64_t l0 = (int64_t)h * W + w;
            y[bc * L + l0] = from_f32<T>((to_f32(s0[l0]) + to_f32(s2[L - 1 - l0])) + tile[r][tx]);
        }
    }
}

// ---------------------------------------------------------------------------------- merge + LayerNorm + gate
// pass 1: tmp[b, l0, c] = ys0[c][l0] + ys2[c][L-1-l0]          (32 tokens x 32 channels through smem)
// pass 2: tmp[b, l0(l1), c] += ys1[c][l1] + ys3[c][L-1-l1]      (same tiling over the column-major order)
// pass 3: out[b, l0, :] = LayerNorm(tmp[b, l0, :]) * gamma + beta [* zact]   (warp per token)
template <typename T, int PASS>
__global__ void __launch_bounds__(256) merge_pass_kernel(const T* __restrict__ ys, float* __restrict__ tmp, int64_t C,
                                                         int64_t H, int64_t W) {
    __shared__ float tile[32][33];   // [channel][token]
    const int64_t L = H * W;
    const int64_t b = blockIdx.z;
    const int64_t c0 = (int64_t)blockIdx.y * 32, p0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* sa = ys + ((b * 4 + (PASS == 1 ? 0 : 1)) * C) * L;
    const T* sb = ys + ((b * 4 + (PASS == 1 ? 2 : 3)) * C) * L;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t c = c0 + r, pos = p0 + tx;
        if (c < C && pos < L) tile[r][tx] = to_f32(sa[c * L + pos]) + to_f32(sb[c * L + (L - 1 - pos)]);
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int64_t pos = p0 + r, c = c0 + tx;
        if (c < C && pos < L) {
            const int64_t l0 = PASS == 1 ? pos : (pos % H) * W + pos / H;   // l1 = w*H + h  ->  l0 = h*W + w
            float* dst = tmp + (b * L + l0) * C + c;
            if (PASS == 1) *dst = tile[tx][r];
            else *dst = *dst + tile[tx][r];
        }
    }
}

template <typename TO>
__global__ void __launch_bounds__(256) norm_gate_kernel(const float* __restrict__ tmp, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const TO* __restrict__ zact,
                                                        TO* __restrict__ out, int64_t tokens, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t tok = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= tokens) return;
    const float* row = tmp + tok * C;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += row[c];
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.0f;
    for (int c = lane; c < C; c += 32) { const float d = row[c] - mean; ss += d * d; }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    for (int c = lane; c < C; c += 32) {
        float v = (row[c] - mean) * rstd * gamma[c] + beta[c];
        if (zact) v *= to_f32(zact[tok * C + c]);
        out[tok * C + c] = from_f32<TO>(v);
    }
}

template <typename T> static int scan_launch(const void* x, void* xs, int64_t B, int64_t C, int64_t H, int64_t W, int in_cf,
                                             int out_cf, int obo, int scans, cudaStream_t st) {
    if (in_cf && out_cf && !obo && scans == 0 && B * C <= 65535) {   // gridDim.z limit; larger goes generic
        dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 32), (unsigned)(B * C));
        cross_scan_tiled_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (T*)xs, C, (int)H, (int)W);
        XP_LAUNCH_CHECK("cross_scan_tiled_kernel");
        return XP_OK;
    }
    const int64_t total = B * 4 * C * H * W;
    const unsigned grid = (unsigned)(ceil_div(total, 256) < (int64_t)num_sms() * 32 ? ceil_div(total, 256) : (int64_t)num_sms() * 32);
    cross_scan_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)x, (T*)xs, B, C, H, W, in_cf, out_cf, obo, scans);
    XP_LAUNCH_CHECK("cross_scan_generic_kernel");
    return XP_OK;
}

template <typename T> static int merge_launch(const void* ys, void* y, int64_t B, int64_t C, int64_t H, int64_t W, int in_cf,
                                              int out_cf, int obo, int scans, cudaStream_t st) {
    if (in_cf && out_cf && !obo && scans == 0 && B * C <= 65535) {
        dim3 grid((unsigned)ceil_div(W, 32), (unsigned)ceil_div(H, 32), (unsigned)(B * C));
        cross_merge_tiled_kernel<T><<<grid, 256, 0, st>>>((const T*)ys, (T*)y, C, (int)H, (int)W);
        XP_LAUNCH_CHECK("cross_merge_tiled_kernel");
        return XP_OK;
    }
    const int64_t total = (obo ? 4 : 1) * B * C * H * W;
    const unsigned grid = (unsigned)(ceil_div(total, 256) < (int64_t)num_sms() * 32 ? ceil_div(total, 256) : (int64_t)num_sms() * 32);
    cross_merge_generic_kernel<T><<<grid, 256, 0, st>>>((const T*)ys, (T*)y, B, C, H, W, in_cf, out_cf, obo, scans);
    XP_LAUNCH_CHECK("cross_merge_generic_kernel");
    return XP_OK;
}

}  // namespace xp

using namespace xp;

static int check_cs(const void* a, const void* b, int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int scans,
                    const char* who) {
    XP_REQUIRE(a && b, "%s: NULL tensor pointer", who);
    XP_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "%s: bad shape (%lld,%lld,%lld,%lld)", who, (long long)B, (long long)C,
               (long long)H, (long long)W);
    XP_REQUIRE(dtype >= XP_F32 && dtype <= XP_BF16, "%s: unsupported dtype %d", who, dtype);
    XP_REQUIRE(scans >= 0 && scans <= 2, "%s: scans must be 0, 1 or 2 (got %d)", who, scans);
    return XP_OK;
}

extern "C" int xp_cross_scan(const void* x, void* xs, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype,
                             int32_t in_cf, int32_t out_cf, int32_t one_by_one, int32_t scans, xp_stream_t stream) {
    int rc = check_cs(x, xs, B, C, H, W, dtype, scans, "xp_cross_scan");
    if (rc) return rc;
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == XP_F32) return scan_launch<float>(x, xs, B, C, H, W, in_cf, out_cf, one_by_one, scans, st);
    // 16-bit types are moved as raw 16-bit words
    return scan_launch<unsigned short>(x, xs, B, C, H, W, in_cf, out_cf, one_by_one, scans, st);
}

extern "C" int xp_cross_merge(const void* ys, void* y, int64_t B, int64_t C, int64_t H, int64_t W, int32_t dtype,
                              int32_t in_cf, int32_t out_cf, int32_t one_by_one, int32_t scans, xp_stream_t stream) {
    int rc = check_cs(ys, y, B, C, H, W, dtype, scans, "xp_cross_merge");
    if (rc) return rc;
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case XP_F32: return merge_launch<float>(ys, y, B, C, H, W, in_cf, out_cf, one_by_one, scans, st);
        case XP_F16: return merge_launch<__half>(ys, y, B, C, H, W, in_cf, out_cf, one_by_one, scans, st);
        default: return merge_launch<__nv_bfloat16>(ys, y, B, C, H, W, in_cf, out_cf, one_by_one, scans, st);
    }
}

template <typename T, typename TO>
static int mng_launch(const void* ys, const float* gamma, const float* beta, const void* zact, void* out, float* tmp,
                      int64_t B, int64_t C, int64_t H, int64_t W, float eps, cudaStream_t st) {
    const int64_t L = H * W;
    dim3 grid((unsigned)ceil_div(L, 32), (unsigned)ceil_div(C, 32), (unsigned)B);
    merge_pass_kernel<T, 1><<<grid, 256, 0, st>>>((const T*)ys, tmp, C, H, W);
    XP_LAUNCH_CHECK("merge_pass_kernel<1>");
    merge_pass_kernel<T, 2><<<grid, 256, 0, st>>>((const T*)ys, tmp, C, H, W);
    XP_LAUNCH_CHECK("merge_pass_kernel<2>");
    norm_gate_kernel<TO><<<(unsigned)ceil_div(B * L, 8), 256, 0, st>>>(tmp, gamma, beta, (const TO*)zact, (TO*)out, B * L,
                                                                         (int)C, eps);
    XP_LAUNCH_CHECK("norm_gate_kernel");
    return XP_OK;
}

extern "C" int64_t xp_merge_norm_gate_workspace_bytes(int64_t B, int64_t C, int64_t H, int64_t W) {
    return B * C * H * W * (int64_t)sizeof(float);
}

extern "C" int xp_merge_norm_gate(const void* ys, const float* gamma, const float* beta, const void* zact, void* out,
                                  int64_t B, int64_t C, int64_t H, int64_t W, int32_t ys_dtype, int32_t out_dtype,
                                  float eps, void* workspace, int64_t workspace_bytes, xp_stream_t stream) {
    XP_REQUIRE(ys && gamma && beta && out, "xp_merge_norm_gate: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "xp_merge_norm_gate: bad shape");
    XP_REQUIRE(B <= 65535, "xp_merge_norm_gate: batch > 65535 not supported");
    if (workspace_bytes < xp_merge_norm_gate_workspace_bytes(B, C, H, W) || !workspace) {
        set_error("xp_merge_norm_gate: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
                  (long long)xp_merge_norm_gate_workspace_bytes(B, C, H, W));
        return XP_ERR_WORKSPACE;
    }
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float* tmp = (float*)workspace;
    const int key = ys_dtype * 4 + out_dtype;
    switch (key) {
        case XP_F32 * 4 + XP_F32: return mng_launch<float, float>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        case XP_F32 * 4 + XP_F16: return mng_launch<float, __half>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        case XP_F32 * 4 + XP_BF16: return mng_launch<float, __nv_bfloat16>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        case XP_F16 * 4 + XP_F16: return mng_launch<__half, __half>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        case XP_F16 * 4 + XP_F32: return mng_launch<__half, float>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        case XP_BF16 * 4 + XP_BF16: return mng_launch<__nv_bfloat16, __nv_bfloat16>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        case XP_BF16 * 4 + XP_F32: return mng_launch<__nv_bfloat16, float>(ys, gamma, beta, zact, out, tmp, B, C, H, W, eps, st);
        default: break;
    }
    set_error("xp_merge_norm_gate: unsupported dtype combination ys=%d out=%d", ys_dtype, out_dtype);
    return XP_ERR_INVALID_ARG;
}

// ================================================================================================
// Channel-last LayerNorm over the last dimension (rows x C), fp32 statistics.
// Replaces the nn.LayerNorm calls of VSSBlock / patch-embed / downsample (VMamba.py:1222-1234, :1405-1440):
// token rows are short (C = 48..1536) and there are millions of them, so one warp per row with the row held
// in registers (two-pass variance, exactly torch's definition) beats a block-per-row library kernel.
// ================================================================================================
namespace xp {
template <typename TI, typename TO, int PER_LANE>
__global__ void __launch_bounds__(256) layer_norm_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, TO* __restrict__ y, int64_t rows,
                                                         int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const TI* xr = x + row * C;
    float v[PER_LANE];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < C ? to_f32(xr[c]) : 0.0f;
        s += v[i];
    }
    const float mean = warp_sum(s) / (float)C;
    float ss = 0.0f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        const int c = lane + 32 * i;
        const float d = c < C ? v[i] - mean : 0.0f;
        ss += d * d;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)C + eps);
    TO* yr = y + row * C;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        const int c = lane + 32 * i;
        if (c < C) yr[c] = from_f32<TO>((v[i] - mean) * rstd * gamma[c] + beta[c]);
    }
}

template <typename TI, typename TO> static int ln_launch(const void* x, const float* g, const float* b, void* y, int64_t rows,
                                                         int C, float eps, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(rows, 8);
    const int per = (C + 31) / 32;
#define XP_LN_CASE(P) layer_norm_kernel<TI, TO, P><<<grid, 256, 0, st>>>((const TI*)x, g, b, (TO*)y, rows, C, eps)
    if (per <= 2) XP_LN_CASE(2);
    else if (per <= 3) XP_LN_CASE(3);
    else if (per <= 6) XP_LN_CASE(6);
    else if (per <= 12) XP_LN_CASE(12);
    else if (per <= 24) XP_LN_CASE(24);
    else if (per <= 48) XP_LN_CASE(48);
    else { set_error("xp_layer_norm: C must be <= 1536 (got %d)", C); return XP_ERR_INVALID_ARG; }
#undef XP_LN_CASE
    XP_LAUNCH_CHECK("layer_norm_kernel");
    return XP_OK;
}
}  // namespace xp

extern "C" int xp_layer_norm(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int64_t C,
                             int32_t in_dtype, int32_t out_dtype, float eps, xp_stream_t stream) {
    XP_REQUIRE(x && gamma && beta && y, "xp_layer_norm: NULL tensor pointer");
    XP_REQUIRE(rows >= 0 && C > 0, "xp_layer_norm: bad shape");
    if (rows == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (in_dtype * 4 + out_dtype) {
        case XP_F32 * 4 + XP_F32: return xp::ln_launch<float, float>(x, gamma, beta, y, rows, (int)C, eps, st);
        case XP_F32 * 4 + XP_F16: return xp::ln_launch<float, __half>(x, gamma, beta, y, rows, (int)C, eps, st);
        case XP_F32 * 4 + XP_BF16: return xp::ln_launch<float, __nv_bfloat16>(x, gamma, beta, y, rows, (int)C, eps, st);
        case XP_F16 * 4 + XP_F16: return xp::ln_launch<__half, __half>(x, gamma, beta, y, rows, (int)C, eps, st);
        case XP_F16 * 4 + XP_F32: return xp::ln_launch<__half, float>(x, gamma, beta, y, rows, (int)C, eps, st);
        case XP_BF16 * 4 + XP_BF16: return xp::ln_launch<__nv_bfloat16, __nv_bfloat16>(x, gamma, beta, y, rows, (int)C, eps, st);
        case XP_BF16 * 4 + XP_F32: return xp::ln_launch<__nv_bfloat16, float>(x, gamma, beta, y, rows, (int)C, eps, st);
        default: xp::set_error("xp_layer_norm: unsupported dtype combination %d -> %d", in_dtype, out_dtype); return XP_ERR_INVALID_ARG;
    }
}
