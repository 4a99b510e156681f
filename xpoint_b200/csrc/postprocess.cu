// postprocess.cu -- XPoint post-processing tail on sm_100a:
//   detector softmax + depth-to-space      (XPoint.py:356-357)
//   descriptor-map L2 normalisation        (XPoint.py:365-366)
//   greedy box NMS + top-k + compaction    (xpoint/utils/utils.py:148-192, evaluation.py:281-282)
//   bilinear descriptor sampling + L2 norm (xpoint/utils/utils.py:229-238)
// All kernels are HBM/L2-bound integer/byte/fp32 work: coalesced 128-byte rows, warp shuffles and ballots,
// no tensor cores.
#include <stdlib.h>

#include "common.cuh"

namespace xp {

// ================================================================================================
// detector post: one thread per low-res cell; lanes run along w so every logit plane read is a
// coalesced row, and each thread writes r contiguous floats per output row.
// ================================================================================================
template <typename T, int R>
__global__ void __launch_bounds__(128) detector_post_kernel(const T* __restrict__ logits, float* __restrict__ prob,
                                                            int64_t B, int Hc, int Wc) {
    constexpr int CN = R * R + 1;
    const int64_t HW = (int64_t)Hc * Wc;
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= B * HW) return;
    const int64_t b = cell / HW;
    const int hw = (int)(cell % HW), h = hw / Wc, w = hw % Wc;
    const T* lg = logits + b * CN * HW + hw;
    float e[CN];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < CN; ++c) { e[c] = to_f32(lg[(int64_t)c * HW]); m = fmaxf(m, e[c]); }
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

// ================================================================================================
// L2 normalise over channels; optional channel-last copy for the sampler.
// CTA = 32 pixels x all channels (tile in smem, transposed write).
// ================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) l2_normalize_kernel(const T* __restrict__ x, float* __restrict__ out_cf,
                                                           float* __restrict__ out_cl, int C, int64_t HW) {
    extern __shared__ float tile[];                 // [C][33]
    __shared__ float inv_norm[32];
    const int64_t b = blockIdx.y, p0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const T* src = x + b * C * HW;
    for (int c = ty; c < C; c += 8) {
        const int64_t p = p0 + tx;
        tile[c * 33 + tx] = p < HW ? to_f32(src[(int64_t)c * HW + p]) : 0.0f;
    }
    __syncthreads();
    // warp ty handles pixels ty, ty+8, ...: lanes stride over channels
    for (int px = ty; px < 32; px += 8) {
        float s = 0.0f;
        for (int c = tx; c < C; c += 32) { const float v = tile[c * 33 + px]; s += v * v; }
        s = warp_sum(s);
        if (tx == 0) inv_norm[px] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();
    if (out_cf) {
        for (int c = ty; c < C; c += 8) {
            const int64_t p = p0 + tx;
            if (p < HW) out_cf[(b * C + c) * HW + p] = tile[c * 33 + tx] * inv_norm[tx];
        }
    }
    if (out_cl) {
        for (int px = ty; px < 32; px += 8) {
            const int64_t p = p0 + px;
            if (p < HW)
                for (int c = tx; c < C; c += 32) out_cl[(b * HW + p) * C + c] = tile[c * 33 + px] * inv_norm[px];
        }
    }
}

// ================================================================================================
// Greedy box NMS, data-parallel fixed point (bit-exact with sequential greedy NMS):
//   repeat { every undecided candidate with no undecided higher-priority candidate inside its suppression
//            footprint is KEPT; every undecided candidate inside a newly kept footprint is SUPPRESSED }
// Priority = (score desc, flat index asc).  A candidate that is the best undecided one in its footprint can no
// longer be suppressed by anything greedy NMS would process before it, so both algorithms keep the same set.
// One CTA per image; state bytes live in the caller's workspace (L2-resident).
// ================================================================================================
constexpr int NMS_THREADS = 1024;
constexpr int NMS_MAX_OFFS = 31 * 31;

enum : uint8_t { ST_NONE = 0, ST_ALIVE = 1, ST_KEPT = 2, ST_NEW = 3 };

struct NmsParams {
    const float* prob; float* out; uint8_t* state; int32_t* kp; int32_t* kp_count;
    int H, W; float size, min_prob, iou, kp_thr; int64_t topk, kp_cap;
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
    // 1024 threads: inclusive warp scan + scan of 32 warp totals
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    __syncthreads();
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int ws = warp_sums[lane];
        int winc = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        warp_sums[lane] = winc - ws;            // exclusive warp offsets
        if (lane == 31) warp_sums[32] = winc;   // block total
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[wid] + inc - v;
}

// higher priority = processed earlier by greedy NMS: larger score, ties -> lower flat index
__device__ __forceinline__ bool nms_better(float t, int r, float s, int q) { return t > s || (t == s && r < q); }

__global__ void __launch_bounds__(NMS_THREADS) box_nms_kernel(const NmsParams p) {
    __shared__ int8_t off_dy[NMS_MAX_OFFS + 32], off_dx[NMS_MAX_OFFS + 32];
    __shared__ int n_offs_s, alive_s;
    __shared__ int scan_ws[33];
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_remaining;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int H = p.H, W = p.W, HW = H * W;
    const float* prob = p.prob + (int64_t)b * HW;
    uint8_t* st = p.state + (int64_t)b * HW;

    // suppression footprint: offsets whose boxes overlap with IoU > iou (fp32 arithmetic as torchvision), nearest
    // first so that the cooperative scan below usually exits after its first 32 offsets
    if (tid == 0) {
        int n = 0;
        const int R = (int)ceilf(p.size) - 1;
        const float area2 = 2.0f * p.size * p.size;
        for (int ring = 1; ring <= R; ++ring)
            for (int dy = -ring; dy <= ring; ++dy)
                for (int dx = -ring; dx <= ring; ++dx) {
                    if (max(abs(dy), abs(dx)) != ring) continue;
                    const float iw = p.size - fabsf((float)dx), ih = p.size - fabsf((float)dy);
                    if (iw <= 0.0f || ih <= 0.0f) continue;
                    const float inter = iw * ih;
                    if ((double)(inter / (area2 - inter)) > (double)p.iou) { off_dy[n] = (int8_t)dy; off_dx[n] = (int8_t)dx; ++n; }
                }
        n_offs_s = n;
        alive_s = 0;
    }
    __syncthreads();
    const int n_offs = n_offs_s;
    const int HWr = (HW + 31) & ~31;          // warp-uniform loop bound for the ballot-based phases

    int local_alive = 0;
    for (int q = tid; q < HW; q += NMS_THREADS) {
        const bool cand = prob[q] > p.min_prob;
        st[q] = cand ? ST_ALIVE : ST_NONE;
        local_alive += cand;
    }
    if (local_alive) atomicAdd(&alive_s, local_alive);
    __syncthreads();

    while (true) {
        const int alive = alive_s;
        __syncthreads();
        if (alive == 0) break;
        if (tid == 0) alive_s = 0;
        // ---- phase A: an undecided candidate with no undecided higher-priority candidate in its footprint is kept.
        //  A1 (lane per pixel): cheap necessary test on the 8-neighbourhood (always inside the footprint for s >= 2);
        //  A2 (warp per surviving pixel): the 32 lanes test 32 footprint offsets at a time, exit on the first hit.
        for (int q0 = tid - lane; q0 < HWr; q0 += NMS_THREADS) {
            const int q = q0 + lane;
            bool pre = false;
            float s = 0.0f;
            int y = 0, x = 0;
            if (q < HW && st[q] == ST_ALIVE) {
                s = prob[q];
                y = q / W; x = q - y * W;
                pre = true;
                const int nn = n_offs < 8 ? n_offs : 8;
                for (int o = 0; o < nn; ++o) {           // ring 1 comes first in the offset list
                    const int yy = y + off_dy[o], xx = x + off_dx[o];
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const int r = yy * W + xx;
                    const uint8_t sr = st[r];
                    if ((sr == ST_ALIVE || sr == ST_NEW) && nms_better(prob[r], r, s, q)) pre = false;
                }
            }
            unsigned todo = __ballot_sync(0xffffffffu, pre);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int cq = q0 + src;
                const float cs = __shfl_sync(0xffffffffu, s, src);
                const int cy = __shfl_sync(0xffffffffu, y, src), cx = __shfl_sync(0xffffffffu, x, src);
                bool beaten = false;
                for (int o0 = 8; o0 < n_offs && !beaten; o0 += 32) {
                    const int o = o0 + lane;
                    bool hit = false;
                    if (o < n_offs) {
                        const int yy = cy + off_dy[o], xx = cx + off_dx[o];
                        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                            const int r = yy * W + xx;
                            const uint8_t sr = st[r];
                            hit = (sr == ST_ALIVE || sr == ST_NEW) && nms_better(prob[r], r, cs, cq);
                        }
                    }
                    beaten = __any_sync(0xffffffffu, hit);
                }
                if (!beaten && lane == src) st[cq] = ST_NEW;   // NEW is treated like ALIVE by concurrent readers
            }
        }
        __syncthreads();
        // ---- phase B (cooperative scatter): every NEW pixel suppresses the undecided candidates in its footprint and
        // becomes KEPT.  Two NEW pixels are never inside each other's footprint, so NEW is never overwritten.
        for (int q0 = tid - lane; q0 < HWr; q0 += NMS_THREADS) {
            const int q = q0 + lane;
            const bool isnew = q < HW && st[q] == ST_NEW;
            unsigned todo = __ballot_sync(0xffffffffu, isnew);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int cq = q0 + src, cy = cq / W, cx = cq - cy * W;
                for (int o = lane; o < n_offs; o += 32) {
                    const int yy = cy + off_dy[o], xx = cx + off_dx[o];
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    const int r = yy * W + xx;
                    if (st[r] == ST_ALIVE) st[r] = ST_NONE;
                }
            }
            if (isnew) st[q] = ST_KEPT;
        }
        __syncthreads();
        int still = 0;
        for (int q = tid; q < HW; q += NMS_THREADS) still += st[q] == ST_ALIVE;
        if (still) atomicAdd(&alive_s, still);
        __syncthreads();
    }

    // ---- top-k by score among kept (ties at the threshold: lower flat index first) ----
    unsigned thr_bits = 0;      // keep score bits > thr_bits, plus the first `need_eq` with == thr_bits
    int need_eq = -1;           // -1: keep everything
    if (p.topk > 0) {
        int kept_local = 0;
        for (int q = tid; q < HW; q += NMS_THREADS) kept_local += st[q] == ST_KEPT;
        int total;
        block_exclusive_scan(kept_local, scan_ws, total);
        if (total > p.topk) {
            // radix select (8 bits x 4) of the topk-th largest score; positive floats order like their bits
            if (tid == 0) { sel_prefix = 0; sel_remaining = (unsigned)p.topk; }
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                const unsigned prefix = sel_prefix;
                const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
                for (int q = tid; q < HW; q += NMS_THREADS)
                    if (st[q] == ST_KEPT) {
                        const unsigned bits = __float_as_uint(prob[q]);
                        if ((bits & himask) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
                    }
                __syncthreads();
                if (tid == 0) {
                    unsigned rem = sel_remaining;
                    int d = 255;
                    for (; d > 0; --d) { if (hist[d] >= rem) break; rem -= hist[d]; }
                    sel_prefix = prefix | ((unsigned)d << shift);
                    sel_remaining = rem;
                }
                __syncthreads();
            }
            thr_bits = sel_prefix;
            need_eq = (int)sel_remaining;   // how many of the == threshold scores are still wanted
        }
    }

    // ---- selection + raster-order compaction.  Warp w owns the contiguous pixel range [w*span, (w+1)*span): lanes
    //      stride inside it (coalesced), ranks come from ballots, one block scan per quantity orders the warps. ----
    float* out = p.out ? p.out + (int64_t)b * HW : nullptr;
    int32_t* kp = p.kp ? p.kp + (int64_t)b * p.kp_cap * 2 : nullptr;
    const int wid = tid >> 5;
    const int span = (((HW + 31) / 32 + 31) / 32) * 32;          // pixels per warp, multiple of 32
    const int w0 = min(wid * span, HW), w1 = min(w0 + span, HW);
    const unsigned lt_mask = (1u << lane) - 1u;
    int eq_base = 0;
    if (need_eq >= 0) {
        int eq_warp = 0;
        for (int q0 = w0; q0 < w1; q0 += 32) {
            const int q = q0 + lane;
            const bool eq = q < w1 && st[q] == ST_KEPT && __float_as_uint(prob[q]) == thr_bits;
            eq_warp += __popc(__ballot_sync(0xffffffffu, eq));
        }
        int tot;
        const int ex = block_exclusive_scan(lane == 0 ? eq_warp : 0, scan_ws, tot);
        eq_base = __shfl_sync(0xffffffffu, ex, 0);
    }
    int kp_warp = 0;
    for (int q0 = w0; q0 < w1; q0 += 32) {
        const int q = q0 + lane;
        bool kept = false, eq = false;
        float s = 0.0f;
        if (q < w1 && st[q] == ST_KEPT) {
            s = prob[q];
            const unsigned bits = __float_as_uint(s);
            if (need_eq < 0 || bits > thr_bits) kept = true;
            else if (bits == thr_bits) eq = true;
        }
        if (need_eq >= 0) {
            const unsigned em = __ballot_sync(0xffffffffu, eq);
            if (eq && eq_base + __popc(em & lt_mask) < need_eq) kept = true;
            eq_base += __popc(em);
        }
        if (q < w1) st[q] = kept ? ST_KEPT : ST_NONE;
        kp_warp += __popc(__ballot_sync(0xffffffffu, kept && s > p.kp_thr));
    }
    int kp_total = 0, kp_off = 0;
    if (kp || p.kp_count) {
        const int ex = block_exclusive_scan(lane == 0 ? kp_warp : 0, scan_ws, kp_total);
        kp_off = __shfl_sync(0xffffffffu, ex, 0);
    }
    if (kp) {
        for (int q0 = w0; q0 < w1; q0 += 32) {
            const int q = q0 + lane;
            const bool iskp = q < w1 && st[q] == ST_KEPT && prob[q] > p.kp_thr;
            const unsigned m = __ballot_sync(0xffffffffu, iskp);
            const int pos = kp_off + __popc(m & lt_mask);
            if (iskp && pos < p.kp_cap) { kp[2 * pos] = q / W; kp[2 * pos + 1] = q % W; }
            kp_off += __popc(m);
        }
    }
    if (p.kp_count && tid == 0) p.kp_count[b] = kp_total;
    __syncthreads();
    if (out)
        for (int q = tid; q < HW; q += NMS_THREADS) out[q] = st[q] == ST_KEPT ? prob[q] : 0.0f;   // coalesced
}

// ================================================================================================
// descriptor sampling: warp per keypoint, lanes over channels, 4 corner gathers, warp-reduced norm.
// grid_sample(bilinear, zeros, align_corners=True) arithmetic, then F.normalize (utils.py:229-238).
// ================================================================================================
template <bool CL>
__global__ void __launch_bounds__(256) sample_desc_kernel(const int32_t* __restrict__ kp, const int32_t* __restrict__ kp_count,
                                                          int64_t kp_stride, const float* __restrict__ desc, int C, int Hc,
                                                          int Wc, int H, int W, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t b = blockIdx.y;
    if (i >= kp_stride) return;
    float* o = out + (b * kp_stride + i) * C;
    const int n = kp_count ? kp_count[b] : (int)kp_stride;
    if (i >= n) {
        for (int c = lane; c < C; c += 32) o[c] = 0.0f;
        return;
    }
    const int ky = kp[(b * kp_stride + i) * 2], kx = kp[(b * kp_stride + i) * 2 + 1];
    const float gy = (float)ky / ((float)H * 0.5f) - 1.0f;
    const float gx = (float)kx / ((float)W * 0.5f) - 1.0f;
    const float ix = ((gx + 1.0f) / 2.0f) * (float)(Wc - 1);
    const float iy = ((gy + 1.0f) / 2.0f) * (float)(Hc - 1);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((float)x1 - ix) * ((float)y1 - iy), wne = (ix - (float)x0) * ((float)y1 - iy);
    const float wsw = ((float)x1 - ix) * (iy - (float)y0), wse = (ix - (float)x0) * (iy - (float)y0);
    const bool inx0 = x0 >= 0 && x0 < Wc, inx1 = x1 >= 0 && x1 < Wc, iny0 = y0 >= 0 && y0 < Hc, iny1 = y1 >= 0 && y1 < Hc;
    const float* d = desc + b * (int64_t)C * Hc * Wc;
    const int64_t HWc = (int64_t)Hc * Wc;
    float ss = 0.0f;
    for (int c = lane; c < C; c += 32) {
        auto at = [&](int y, int x) -> float {
            return CL ? d[((int64_t)y * Wc + x) * C + c] : d[(int64_t)c * HWc + (int64_t)y * Wc + x];
        };
        float v = 0.0f;
        if (iny0 && inx0) v += at(y0, x0) * wnw;
        if (iny0 && inx1) v += at(y0, x1) * wne;
        if (iny1 && inx0) v += at(y1, x0) * wsw;
        if (iny1 && inx1) v += at(y1, x1) * wse;
        o[c] = v;
        ss += v * v;
    }
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < C; c += 32) o[c] = o[c] / nrm;
}

}  // namespace xp

using namespace xp;

template <typename T> static int det_launch(const void* logits, float* prob, int64_t B, int64_t Hc, int64_t Wc, int r,
                                            cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(B * Hc * Wc, 128);
    switch (r) {
        case 8: detector_post_kernel<T, 8><<<grid, 128, 0, st>>>((const T*)logits, prob, B, (int)Hc, (int)Wc); break;
        case 4: detector_post_kernel<T, 4><<<grid, 128, 0, st>>>((const T*)logits, prob, B, (int)Hc, (int)Wc); break;
        case 2: detector_post_kernel<T, 2><<<grid, 128, 0, st>>>((const T*)logits, prob, B, (int)Hc, (int)Wc); break;
        default: set_error("xp_detector_post: r must be 2, 4 or 8 (got %d)", r); return XP_ERR_INVALID_ARG;
    }
    XP_LAUNCH_CHECK("detector_post_kernel");
    return XP_OK;
}

extern "C" int xp_detector_post(const void* logits, float* prob, int64_t B, int64_t Hc, int64_t Wc, int32_t r, int32_t dtype,
                                xp_stream_t stream) {
    XP_REQUIRE(logits && prob, "xp_detector_post: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && Hc > 0 && Wc > 0, "xp_detector_post: bad shape");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case XP_F32: return det_launch<float>(logits, prob, B, Hc, Wc, r, st);
        case XP_F16: return det_launch<__half>(logits, prob, B, Hc, Wc, r, st);
        case XP_BF16: return det_launch<__nv_bfloat16>(logits, prob, B, Hc, Wc, r, st);
        default: set_error("xp_detector_post: unsupported dtype %d", dtype); return XP_ERR_INVALID_ARG;
    }
}

template <typename T> static int l2_launch(const void* x, float* out_cf, float* out_cl, int64_t B, int64_t C, int64_t HW,
                                           cudaStream_t st) {
    const size_t smem = (size_t)C * 33 * sizeof(float);
    XP_CUDA_OK(cudaFuncSetAttribute(l2_normalize_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)B);
    l2_normalize_kernel<T><<<grid, 256, smem, st>>>((const T*)x, out_cf, out_cl, (int)C, HW);
    XP_LAUNCH_CHECK("l2_normalize_kernel");
    return XP_OK;
}

extern "C" int xp_l2_normalize(const void* x, float* out_cf, float* out_cl, int64_t B, int64_t C, int64_t HW, int32_t dtype,
                               xp_stream_t stream) {
    XP_REQUIRE(x && (out_cf || out_cl), "xp_l2_normalize: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && B <= 65535 && C > 0 && C <= 1024 && HW > 0, "xp_l2_normalize: bad shape (C <= 1024, B <= 65535)");
    if (B == 0) return XP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case XP_F32: return l2_launch<float>(x, out_cf, out_cl, B, C, HW, st);
        case XP_F16: return l2_launch<__half>(x, out_cf, out_cl, B, C, HW, st);
        case XP_BF16: return l2_launch<__nv_bfloat16>(x, out_cf, out_cl, B, C, HW, st);
        default: set_error("xp_l2_normalize: unsupported dtype %d", dtype); return XP_ERR_INVALID_ARG;
    }
}

namespace xp {
int launch_box_nms_fast(const float* prob, float* prob_nms, int64_t B, int64_t H, int64_t W, float size, float min_prob, float iou,
                        int64_t keep_top_k, float kp_threshold, int32_t* keypoints, int32_t* kp_count, int64_t kp_capacity,
                        void* workspace, cudaStream_t st);
}

// state bytes + scan-cursor bytes (nms_fast.cu)
extern "C" int64_t xp_nms_workspace_bytes(int64_t B, int64_t H, int64_t W) { return 2 * B * H * W; }

extern "C" int xp_box_nms(const float* prob, float* prob_nms, int64_t B, int64_t H, int64_t W, float size, float min_prob,
                          float iou, int64_t keep_top_k, float kp_threshold, int32_t* keypoints, int32_t* kp_count,
                          int64_t kp_capacity, void* workspace, int64_t workspace_bytes, xp_stream_t stream) {
    XP_REQUIRE(prob, "xp_box_nms: prob is NULL");
    XP_REQUIRE(B >= 0 && H > 0 && W > 0 && H * W < (1LL << 30), "xp_box_nms: bad shape");
    XP_REQUIRE(size > 0.0f && size <= 16.0f, "xp_box_nms: box size must be in (0, 16] (got %f)", (double)size);
    XP_REQUIRE(iou > 0.0f, "xp_box_nms: iou must be > 0");
    XP_REQUIRE(min_prob >= 0.0f, "xp_box_nms: min_prob must be >= 0 (scores are probabilities)");
    XP_REQUIRE(!keypoints || kp_capacity > 0, "xp_box_nms: keypoints given but kp_capacity == 0");
    if (!workspace || workspace_bytes < xp_nms_workspace_bytes(B, H, W)) {
        set_error("xp_box_nms: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
                  (long long)xp_nms_workspace_bytes(B, H, W));
        return XP_ERR_WORKSPACE;
    }
    if (B == 0) return XP_OK;
    {   // widths that are multiples of 8 (every XPoint input size) take the vectorised kernel (nms_fast.cu)
        const uintptr_t al = reinterpret_cast<uintptr_t>(prob) | reinterpret_cast<uintptr_t>(workspace) |
                             reinterpret_cast<uintptr_t>(prob_nms);
        static const bool no_fast = getenv("XP_NMS_GENERIC") != nullptr;      // testing knob
        if (W % 8 == 0 && (al & 15) == 0 && !no_fast && size <= 8.0f)      // <= 224 footprint offsets: one-byte cursors
            return xp::launch_box_nms_fast(prob, prob_nms, B, H, W, size, min_prob, iou, keep_top_k, kp_threshold, keypoints,
                                           kp_count, kp_capacity, workspace, (cudaStream_t)stream);
    }
    NmsParams p;
    p.prob = prob; p.out = prob_nms; p.state = (uint8_t*)workspace; p.kp = keypoints; p.kp_count = kp_count;
    p.H = (int)H; p.W = (int)W; p.size = size; p.min_prob = min_prob; p.iou = iou; p.kp_thr = kp_threshold;
    p.topk = keep_top_k; p.kp_cap = kp_capacity;
    box_nms_kernel<<<(unsigned)B, NMS_THREADS, 0, (cudaStream_t)stream>>>(p);
    XP_LAUNCH_CHECK("box_nms_kernel");
    return XP_OK;
}

extern "C" int xp_sample_descriptors(const int32_t* keypoints, const int32_t* kp_count, int64_t B, int64_t kp_stride,
                                     const float* desc, int32_t channel_last, int64_t C, int64_t Hc, int64_t Wc, int64_t H,
                                     int64_t W, float* out, xp_stream_t stream) {
    XP_REQUIRE(keypoints && desc && out, "xp_sample_descriptors: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && B <= 65535 && kp_stride >= 0 && C > 0 && Hc > 0 && Wc > 0 && H > 0 && W > 0,
               "xp_sample_descriptors: bad shape");
    if (B == 0 || kp_stride == 0) return XP_OK;
    dim3 grid((unsigned)ceil_div(kp_stride, 8), (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (channel_last)
        sample_desc_kernel<true><<<grid, 256, 0, st>>>(keypoints, kp_count, kp_stride, desc, (int)C, (int)Hc, (int)Wc, (int)H,
                                                       (int)W, out);
    else
        sample_desc_kernel<false><<<grid, 256, 0, st>>>(keypoints, kp_count, kp_stride, desc, (int)C, (int)Hc, (int)Wc, (int)H,
                                                        (int)W, out);
    XP_LAUNCH_CHECK("sample_desc_kernel");
    return XP_OK;
}
