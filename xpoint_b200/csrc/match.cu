// match.cu -- mutual-nearest-neighbour descriptor matching (xp_mnn_match).
//
// Replaces get_matches(d1, d2, 'bfmatcher', crossCheck=True) = cv2.BFMatcher(NORM_L2, crossCheck=True).match
// (xpoint/utils/matching.py:4-36) and the in-repo NNMatcher (matching.py:38-75), which both run on the CPU
// after a device->host copy of the descriptors.
//
//   nn12[i] = argmin_j ||d1_i - d2_j||^2 = argmin_j (|d2_j|^2 - 2 d1_i.d2_j)      (first minimum on ties)
//   nn21[j] = argmin_i (|d1_i|^2 - 2 d1_i.d2_j)
//   match i <-> nn12[i]  iff  nn21[nn12[i]] == i ;  distance recomputed exactly as sqrt(sum (a-b)^2) in fp32.
//
// Two implementations of the similarity GEMM + fused arg-min:
//   * tensor-core path (match_tc.cu): tcgen05.mma kind::tf32 with a 3xTF32 operand split, accumulators in TMEM,
//     row/column arg-min fused in the epilogue -- the similarity matrix is never written.
//   * exact fp32 CUDA-core path (this file): classic smem-tiled SGEMM with the same fused arg-min; used as the
//     on-device self-check of the tensor-core path and for shapes the tensor-core path does not take.
#include "common.cuh"

namespace xp {

int mnn_argmin_tc(const float* X, const float* Y, const int32_t* nx, const int32_t* ny, int64_t P, int64_t x_stride,
                  int64_t y_stride, int64_t C, const float* xnorm, const float* ynorm, int32_t* nn_x, int32_t* nn_y,
                  void* split_ws, cudaStream_t st);
int64_t mnn_tc_workspace_bytes(int64_t P, int64_t x_stride, int64_t y_stride, int64_t C);

// ---------------------------------------------------------------------------------- row norms
__global__ void __launch_bounds__(256) row_norm2_kernel(const float* __restrict__ X, float* __restrict__ out, int64_t rows,
                                                        int C) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) { const float v = X[r * C + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
}

// ---------------------------------------------------------------------------------- exact fp32 arg-min
// CTA: 64 X-rows, loops over all Y rows in tiles of 64; K chunks of 16.  128 threads, 8x4 register tile.
constexpr int MM_BM = 64, MM_BN = 64, MM_BK = 16;

__global__ void __launch_bounds__(128) nn_argmin_fp32_kernel(const float* __restrict__ X, const float* __restrict__ Y,
                                                             const float* __restrict__ ynorm, const int32_t* __restrict__ nx,
                                                             const int32_t* __restrict__ ny, int64_t x_stride, int64_t y_stride,
                                                             int C, int32_t* __restrict__ nn) {
    __shared__ float Xs[MM_BK][MM_BM + 4];
    __shared__ float Ys[MM_BK][MM_BN + 4];
    __shared__ float red_key[MM_BM][16];
    __shared__ int red_idx[MM_BM][16];
    const int pair = blockIdx.y;
    const int n_x = nx ? nx[pair] : (int)x_stride, n_y = ny ? ny[pair] : (int)y_stride;
    const int i0 = blockIdx.x * MM_BM;
    if (i0 >= n_x) return;
    const float* Xp = X + (int64_t)pair * x_stride * C;
    const float* Yp = Y + (int64_t)pair * y_stride * C;
    const float* yn = ynorm + (int64_t)pair * y_stride;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;      // 8 x 16 threads; thread owns rows ty*8..+8, cols tx*4..+4
    float best[8];
    int besti[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { best[r] = INFINITY; besti[r] = 0x7fffffff; }

    for (int j0 = 0; j0 < n_y; j0 += MM_BN) {
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
        for (int k0 = 0; k0 < C; k0 += MM_BK) {
            // load 64x16 of X and Y (row-major, K contiguous): 1024 floats each, 8 per thread
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int e = tid + t * 128;          // 0..1023
                const int row = e >> 4, k = e & 15;
                const int xi = i0 + row, yj = j0 + row;
                Xs[k][row] = (xi < n_x && k0 + k < C) ? Xp[(int64_t)xi * C + k0 + k] : 0.0f;
                Ys[k][row] = (yj < n_y && k0 + k < C) ? Yp[(int64_t)yj * C + k0 + k] : 0.0f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < MM_BK; ++k) {
                float a[8], b[4];
#pragma unroll
                for (int r = 0; r < 8; ++r) a[r] = Xs[k][ty * 8 + r];
#pragma unroll
                for (int c = 0; c < 4; ++c) b[c] = Ys[k][tx * 4 + c];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx * 4 + c;
            if (j < n_y) {
                const float ynj = yn[j];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float key = fmaf(-2.0f, acc[r][c], ynj);
                    if (key < best[r]) { best[r] = key; besti[r] = j; }   // j ascending: first minimum wins
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) { red_key[ty * 8 + r][tx] = best[r]; red_idx[ty * 8 + r][tx] = besti[r]; }
    __syncthreads();
    if (tid < MM_BM && i0 + tid < n_x) {
        float bk = INFINITY; int bi = 0x7fffffff;
        for (int t = 0; t < 16; ++t) {
            const float k = red_key[tid][t]; const int i = red_idx[tid][t];
            if (k < bk || (k == bk && i < bi)) { bk = k; bi = i; }
        }
        nn[(int64_t)pair * x_stride + i0 + tid] = bi;
    }
}

// ---------------------------------------------------------------------------------- mutual check + exact distance
__global__ void __launch_bounds__(256) mutual_kernel(const float* __restrict__ d1, const float* __restrict__ d2,
                                                     const int32_t* __restrict__ n1, const int32_t* __restrict__ nn12,
                                                     const int32_t* __restrict__ nn21, int64_t n1_stride, int64_t n2_stride,
                                                     int C, int32_t* __restrict__ match_idx, float* __restrict__ match_dist,
                                                     int32_t* __restrict__ match_count) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t pair = blockIdx.y;
    if (i >= n1_stride) return;
    const int n = n1 ? n1[pair] : (int)n1_stride;
    int j = -1;
    if (i < n) {
        j = nn12[pair * n1_stride + i];
        if (j < 0 || j >= n2_stride || nn21[pair * n2_stride + j] != (int32_t)i) j = -1;
    }
    float dist = 0.0f;
    if (j >= 0 && match_dist) {
        const float* a = d1 + (pair * n1_stride + i) * C;
        const float* b = d2 + (pair * n2_stride + j) * C;
        float s = 0.0f;
        for (int c = lane; c < C; c += 32) { const float t = a[c] - b[c]; s = fmaf(t, t, s); }
        dist = sqrtf(warp_sum(s));
    }
    if (lane == 0) {
        if (match_idx) match_idx[pair * n1_stride + i] = j;
        if (match_dist) match_dist[pair * n1_stride + i] = dist;
        if (j >= 0 && match_count) atomicAdd(&match_count[pair], 1);
    }
}

}  // namespace xp

using namespace xp;

// workspace layout: [xnorm P*n1][ynorm P*n2][nn12 P*n1][nn21 P*n2][tf32 hi/lo copies of d1 and d2 (tensor-core path)]
extern "C" int64_t xp_match_workspace_bytes(int64_t P, int64_t n1_stride, int64_t n2_stride, int64_t C) {
    const int64_t a = P * n1_stride, b = P * n2_stride;
    return (a + b) * 4 * 2 + 256 + mnn_tc_workspace_bytes(P, n1_stride, n2_stride, C);
}

extern "C" int xp_mnn_match(const float* d1, const float* d2, const int32_t* n1, const int32_t* n2, int64_t P,
                            int64_t n1_stride, int64_t n2_stride, int64_t C, int32_t* nn12, int32_t* nn21, int32_t* match_idx,
                            float* match_dist, int32_t* match_count, int32_t use_tensor_cores, void* workspace,
                            int64_t workspace_bytes, xp_stream_t stream) {
    XP_REQUIRE(d1 && d2, "xp_mnn_match: NULL descriptor pointer");
    XP_REQUIRE(P >= 0 && P <= 65535 && n1_stride >= 0 && n2_stride >= 0 && C > 0, "xp_mnn_match: bad shape");
    XP_REQUIRE(n1_stride < (1LL << 24) && n2_stride < (1LL << 24), "xp_mnn_match: at most 2^24 descriptors per image");
    if (!workspace || workspace_bytes < xp_match_workspace_bytes(P, n1_stride, n2_stride, C)) {
        set_error("xp_mnn_match: workspace too small (%lld < %lld bytes)", (long long)workspace_bytes,
                  (long long)xp_match_workspace_bytes(P, n1_stride, n2_stride, C));
        return XP_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (match_count && P > 0) XP_CUDA_OK(cudaMemsetAsync(match_count, 0, sizeof(int32_t) * P, st));
    if (P == 0 || n1_stride == 0) return XP_OK;
    const int64_t a = P * n1_stride, b = P * n2_stride;
    uint8_t* ws = (uint8_t*)workspace;
    ws = (uint8_t*)(((uintptr_t)ws + 15) & ~(uintptr_t)15);
    float* xnorm = (float*)ws;
    float* ynorm = xnorm + a;
    int32_t* w12 = (int32_t*)(ynorm + b);
    int32_t* w21 = w12 + a;
    void* split_ws = (void*)(w21 + b);
    int32_t* o12 = nn12 ? nn12 : w12;
    int32_t* o21 = nn21 ? nn21 : w21;
    if (n2_stride == 0) {
        // no train descriptors: nothing can match (reference: get_matches returns [] , matching.py:32-33)
        XP_CUDA_OK(cudaMemsetAsync(o12, 0xff, sizeof(int32_t) * a, st));
        if (match_idx) XP_CUDA_OK(cudaMemsetAsync(match_idx, 0xff, sizeof(int32_t) * a, st));
        if (match_dist) XP_CUDA_OK(cudaMemsetAsync(match_dist, 0, sizeof(float) * a, st));
        return XP_OK;
    }
    row_norm2_kernel<<<(unsigned)ceil_div(a, 8), 256, 0, st>>>(d1, xnorm, a, (int)C);
    XP_LAUNCH_CHECK("row_norm2_kernel");
    row_norm2_kernel<<<(unsigned)ceil_div(b, 8), 256, 0, st>>>(d2, ynorm, b, (int)C);
    XP_LAUNCH_CHECK("row_norm2_kernel");
    // The tensor-core kernel needs C % 32 == 0, 32 <= C <= 1024 and 16-byte-aligned descriptors (TMA boxes); any other
    // descriptor size the reference accepts (get_matches has no such restriction, matching.py:4-36) takes the exact
    // fp32 CUDA-core kernels instead of failing.
    const bool tc_ok = C % 32 == 0 && C >= 32 && C <= 1024 && ((reinterpret_cast<uintptr_t>(d1) | reinterpret_cast<uintptr_t>(d2)) & 15) == 0;
    if (use_tensor_cores && tc_ok) {
        int rc = mnn_argmin_tc(d1, d2, n1, n2, P, n1_stride, n2_stride, C, xnorm, ynorm, o12, o21, split_ws, st);
        if (rc) return rc;
    } else {
        // rows with index >= n (per pair) are never written: pre-fill with -1
        XP_CUDA_OK(cudaMemsetAsync(o12, 0xff, sizeof(int32_t) * a, st));
        XP_CUDA_OK(cudaMemsetAsync(o21, 0xff, sizeof(int32_t) * b, st));
        dim3 g12((unsigned)ceil_div(n1_stride, MM_BM), (unsigned)P), g21((unsigned)ceil_div(n2_stride, MM_BM), (unsigned)P);
        nn_argmin_fp32_kernel<<<g12, 128, 0, st>>>(d1, d2, ynorm, n1, n2, n1_stride, n2_stride, (int)C, o12);
        XP_LAUNCH_CHECK("nn_argmin_fp32_kernel");
        nn_argmin_fp32_kernel<<<g21, 128, 0, st>>>(d2, d1, xnorm, n2, n1, n2_stride, n1_stride, (int)C, o21);
        XP_LAUNCH_CHECK("nn_argmin_fp32_kernel");
    }
    if (match_idx || match_dist || match_count) {
        dim3 grid((unsigned)ceil_div(n1_stride, 8), (unsigned)P);
        mutual_kernel<<<grid, 256, 0, st>>>(d1, d2, n1, o12, o21, n1_stride, n2_stride, (int)C, match_idx, match_dist,
                                            match_count);
        XP_LAUNCH_CHECK("mutual_kernel");
    }
    return XP_OK;
}
