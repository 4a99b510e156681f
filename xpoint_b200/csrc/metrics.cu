// metrics.cu -- the evaluation driver's per-sample geometry on the device.
//
// Reference code replaced (all per-sample Python / numpy / OpenCV on the CPU after device->host copies):
//   * warp_keypoints            xpoint/utils/homographies.py:479-495   cv2.perspectiveTransform in float64 on (x, y), cast back
//   * filter_points             xpoint/utils/homographies.py:511-526   keep points inside the image frame
//   * compute_repeatability_for_sample   xpoint/utils/benchmark_evaluation.py:396-467
//   * the correctness / M-score part of compute_descriptor_for_sample     benchmark_evaluation.py:650-690
// Everything is batched over pairs, takes the keypoint tensors where the NMS kernel left them ((B, k, 2) int32 (y, x) + counts)
// and needs no host synchronisation.
#include "common.cuh"

namespace xp {

// H (row-major 3x3, float64) applied to pixel (y, x) as cv2.perspectiveTransform does: (x', y') = (H [x y 1]^T)[:2] / w,
// w == 0 -> 0 (OpenCV's convention).  Plain IEEE double arithmetic without FMA contraction (compiled with -fmad=false) so that
// the int truncation below is reproducible against the CPU restatement.
__device__ __forceinline__ void warp_point(const double* H, double y, double x, double& yo, double& xo) {
    const double X = H[0] * x + H[1] * y + H[2];
    const double Y = H[3] * x + H[4] * y + H[5];
    const double Wd = H[6] * x + H[7] * y + H[8];
    const double s = Wd != 0.0 ? 1.0 / Wd : 0.0;
    xo = X * s;
    yo = Y * s;
}

// kp (B, k, 2) int32 (y, x) -> out_f (B, k, 2) float64 (y', x') and / or out_i (B, k, 2) int32 = (int)(y', x') (numpy astype(int):
// truncation toward zero); inside (B, k) uint8 = filter_points on the int (or float) result; rows >= count are zeroed.
__global__ void __launch_bounds__(256) warp_keypoints_kernel(const int32_t* __restrict__ kp, const int32_t* __restrict__ cnt,
                                                            const double* __restrict__ H, int k, int height, int width,
                                                            double* __restrict__ out_f, int32_t* __restrict__ out_i,
                                                            uint8_t* __restrict__ inside, int filter_on_float) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int n = cnt ? min(cnt[b], k) : k;
    const int64_t o = ((int64_t)b * k + i) * 2;
    if (i >= n) {
        if (out_f) { out_f[o] = 0.0; out_f[o + 1] = 0.0; }
        if (out_i) { out_i[o] = 0; out_i[o + 1] = 0; }
        if (inside) inside[(int64_t)b * k + i] = 0;
        return;
    }
    double yo, xo;
    warp_point(H + b * 9, (double)kp[o], (double)kp[o + 1], yo, xo);
    if (out_f) { out_f[o] = yo; out_f[o + 1] = xo; }
    // astype(int) of a float64: truncation toward zero; values outside the int64 range are undefined in numpy as well
    const double lim = 2147483647.0;
    const int32_t yi = (int32_t)fmax(fmin(yo, lim), -lim), xi = (int32_t)fmax(fmin(xo, lim), -lim);
    if (out_i) { out_i[o] = yi; out_i[o + 1] = xi; }
    if (inside) {
        bool in;
        if (filter_on_float) in = yo >= 0.0 && xo >= 0.0 && yo < (double)height && xo < (double)width;
        else in = yi >= 0 && xi >= 0 && yi < height && xi < width;
        inside[(int64_t)b * k + i] = in ? 1 : 0;
    }
}

// Repeatability counts of one direction (benchmark_evaluation.py:432-461): for every warped point of set A that lies inside the
// image, the distance to the nearest keypoint of set B (integer pixel coordinates on both sides); counts[b][t] = number of such
// points with sqrt(d2) <= thr[t], n_inside[b] = number of points inside.  One CTA per pair; set B staged through shared memory.
constexpr int RP_TILE = 2048;
constexpr int RP_MAX_THR = 8;
__global__ void __launch_bounds__(256) repeat_count_kernel(const int32_t* __restrict__ warped, const uint8_t* __restrict__ inside,
                                                          const int32_t* __restrict__ cnt_a, const int32_t* __restrict__ kp_b,
                                                          const int32_t* __restrict__ cnt_b, int k, const double* __restrict__ thr,
                                                          int nthr, int32_t* __restrict__ counts, int32_t* __restrict__ n_inside) {
    __shared__ int2 tile[RP_TILE];
    __shared__ int s_counts[RP_MAX_THR + 1];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid <= RP_MAX_THR) s_counts[tid] = 0;
    const int na = min(cnt_a[b], k), nb = min(cnt_b[b], k);
    const int32_t* wa = warped + (int64_t)b * k * 2;
    const uint8_t* ia = inside + (int64_t)b * k;
    const int32_t* kb = kp_b + (int64_t)b * k * 2;
    __syncthreads();
    for (int base = 0; base < na; base += 256) {           // 256 points of A per sweep over B
        const int i = base + tid;
        const bool live = i < na && ia[i];
        const int ay = live ? wa[2 * i] : 0, ax = live ? wa[2 * i + 1] : 0;
        long long best = 0x7fffffffffffffffLL;
        for (int t0 = 0; t0 < nb; t0 += RP_TILE) {
            const int m = min(RP_TILE, nb - t0);
            __syncthreads();
            for (int j = tid; j < m; j += 256) tile[j] = make_int2(kb[2 * (t0 + j)], kb[2 * (t0 + j) + 1]);
            __syncthreads();
            if (live) {
                for (int j = 0; j < m; ++j) {
                    const long long dy = (long long)ay - tile[j].x, dx = (long long)ax - tile[j].y;
                    const long long d2 = dy * dy + dx * dx;
                    best = d2 < best ? d2 : best;
                }
            }
        }
        if (live) {
            atomicAdd(&s_counts[RP_MAX_THR], 1);
            if (nb > 0) {
                const double d = sqrt((double)best);         // numpy: np.linalg.norm on integer arrays -> float64
                for (int t = 0; t < nthr; ++t)
                    if (d <= thr[t]) atomicAdd(&s_counts[t], 1);
            }
        }
    }
    __syncthreads();
    if (tid < nthr) counts[b * nthr + tid] = s_counts[tid];
    if (tid == 0) n_inside[b] = s_counts[RP_MAX_THR];
}

// Match correctness / M-score of one direction (benchmark_evaluation.py:650-690): query keypoints warped by the ground-truth
// homography in float64 (warp_keypoints(..., float)); a match (q, t) is correct iff || float32(warped_q - kp_t) ||_2 <= thr
// (torch.norm(dist.float(), dim=-1) <= th); n_correct[b][t], n_possible[b] = warped query points inside the image (filter_points on
// the float coordinates), n_gt[b][t] = query keypoints with at least one train keypoint within thr.
__global__ void __launch_bounds__(256) match_score_kernel(const double* __restrict__ warped_f, const uint8_t* __restrict__ inside,
                                                         const int32_t* __restrict__ cnt_q, const int32_t* __restrict__ kp_t,
                                                         const int32_t* __restrict__ cnt_t, const int32_t* __restrict__ match_idx,
                                                         int k, const double* __restrict__ thr, int nthr,
                                                         int32_t* __restrict__ n_correct, int32_t* __restrict__ n_gt,
                                                         int32_t* __restrict__ n_possible, int32_t* __restrict__ n_matches) {
    __shared__ int2 tile[RP_TILE];
    __shared__ int s_correct[RP_MAX_THR], s_gt[RP_MAX_THR], s_misc[2];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid < RP_MAX_THR) { s_correct[tid] = 0; s_gt[tid] = 0; }
    if (tid < 2) s_misc[tid] = 0;
    const int nq = min(cnt_q[b], k), nt = min(cnt_t[b], k);
    const double* wq = warped_f + (int64_t)b * k * 2;
    const int32_t* kt = kp_t + (int64_t)b * k * 2;
    const int32_t* mi = match_idx + (int64_t)b * k;
    __syncthreads();
    for (int base = 0; base < nq; base += 256) {
        const int i = base + tid;
        const bool live = i < nq;
        const double qy = live ? wq[2 * i] : 0.0, qx = live ? wq[2 * i + 1] : 0.0;
        float best = INFINITY;
        for (int t0 = 0; t0 < nt; t0 += RP_TILE) {
            const int m = min(RP_TILE, nt - t0);
            __syncthreads();
            for (int j = tid; j < m; j += 256) tile[j] = make_int2(kt[2 * (t0 + j)], kt[2 * (t0 + j) + 1]);
            __syncthreads();
            if (live) {
                for (int j = 0; j < m; ++j) {
                    const float dy = (float)(qy - (double)tile[j].x), dx = (float)(qx - (double)tile[j].y);
                    const float d = sqrtf(dy * dy + dx * dx);
                    best = d < best ? d : best;
                }
            }
        }
        if (live) {
            if (inside[(int64_t)b * k + i]) atomicAdd(&s_misc[0], 1);
            for (int t = 0; t < nthr; ++t)
                if (nt > 0 && (double)best <= thr[t]) atomicAdd(&s_gt[t], 1);
            const int j = mi[i];
            if (j >= 0 && j < nt) {
                atomicAdd(&s_misc[1], 1);
                const float dy = (float)(qy - (double)kt[2 * j]), dx = (float)(qx - (double)kt[2 * j + 1]);
                const float d = sqrtf(dy * dy + dx * dx);
                for (int t = 0; t < nthr; ++t)
                    if ((double)d <= thr[t]) atomicAdd(&s_correct[t], 1);
            }
        }
    }
    __syncthreads();
    if (tid < nthr) { n_correct[b * nthr + tid] = s_correct[tid]; n_gt[b * nthr + tid] = s_gt[tid]; }
    if (tid == 0) { n_possible[b] = s_misc[0]; n_matches[b] = s_misc[1]; }
}

}  // namespace xp

using namespace xp;

extern "C" int xp_warp_keypoints(const int32_t* keypoints, const int32_t* count, const double* H, int64_t B, int64_t k,
                                 int64_t height, int64_t width, double* out_float, int32_t* out_int, uint8_t* inside,
                                 int32_t filter_on_float, xp_stream_t stream) {
    XP_REQUIRE(keypoints && H, "xp_warp_keypoints: NULL keypoints / homography");
    XP_REQUIRE(B >= 0 && B <= 65535 && k >= 0 && height > 0 && width > 0, "xp_warp_keypoints: bad shape");
    XP_REQUIRE(out_float || out_int || inside, "xp_warp_keypoints: no output requested");
    if (B == 0 || k == 0) return XP_OK;
    dim3 grid((unsigned)ceil_div(k, 256), (unsigned)B);
    warp_keypoints_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(keypoints, count, H, (int)k, (int)height, (int)width, out_float,
                                                                  out_int, inside, filter_on_float);
    XP_LAUNCH_CHECK("warp_keypoints_kernel");
    return XP_OK;
}

extern "C" int xp_repeatability_counts(const int32_t* warped_a, const uint8_t* inside_a, const int32_t* count_a,
                                       const int32_t* keypoints_b, const int32_t* count_b, int64_t B, int64_t k,
                                       const double* thresholds, int64_t n_thresholds, int32_t* counts, int32_t* n_inside,
                                       xp_stream_t stream) {
    XP_REQUIRE(warped_a && inside_a && count_a && keypoints_b && count_b && thresholds && counts && n_inside,
               "xp_repeatability_counts: NULL pointer");
    XP_REQUIRE(B >= 0 && k >= 0 && n_thresholds >= 1 && n_thresholds <= RP_MAX_THR, "xp_repeatability_counts: bad shape (1..8 thresholds)");
    if (B == 0) return XP_OK;
    repeat_count_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(warped_a, inside_a, count_a, keypoints_b, count_b, (int)k,
                                                                       thresholds, (int)n_thresholds, counts, n_inside);
    XP_LAUNCH_CHECK("repeat_count_kernel");
    return XP_OK;
}

extern "C" int xp_match_score_counts(const double* warped_q, const uint8_t* inside_q, const int32_t* count_q,
                                     const int32_t* keypoints_t, const int32_t* count_t, const int32_t* match_idx, int64_t B,
                                     int64_t k, const double* thresholds, int64_t n_thresholds, int32_t* n_correct, int32_t* n_gt,
                                     int32_t* n_possible, int32_t* n_matches, xp_stream_t stream) {
    XP_REQUIRE(warped_q && inside_q && count_q && keypoints_t && count_t && match_idx && thresholds && n_correct && n_gt &&
                   n_possible && n_matches, "xp_match_score_counts: NULL pointer");
    XP_REQUIRE(B >= 0 && k >= 0 && n_thresholds >= 1 && n_thresholds <= RP_MAX_THR, "xp_match_score_counts: bad shape (1..8 thresholds)");
    if (B == 0) return XP_OK;
    match_score_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(warped_q, inside_q, count_q, keypoints_t, count_t, match_idx,
                                                                      (int)k, thresholds, (int)n_thresholds, n_correct, n_gt,
                                                                      n_possible, n_matches);
    XP_LAUNCH_CHECK("match_score_kernel");
    return XP_OK;
}
