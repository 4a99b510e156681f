// homography.cu -- batched homography estimation from the mutual matches.
//
// The reference estimates one homography per pair on the CPU, after copying keypoints and matches to the host:
// cv2.findHomography(optical_pts, thermal_pts, USAC_MAGSAC, ransacReprojThreshold, confidence 0.9999, maxIters 10000)
// (xpoint/utils/evaluation.py:359-378, points are (x, y) = (kp[1], kp[0])).  OpenCV's USAC is a randomised, version-dependent
// estimator, so there is no bit-level contract to reproduce; what the evaluation consumes is H (corner error against the
// ground truth, evaluation.py:385-389) and the inlier mask.  This kernel keeps the whole batch on the GPU with a
// deterministic LO-RANSAC (plain IEEE fp64, compiled without FMA contraction, so a CPU restatement reproduces every decision):
//   1. matches are compacted in keypoint order; coordinates are normalised to [-1, 1] by the image size
//   2. `iters` hypotheses per pair: 4 distinct matches from a counter-based hash (same integers on CPU and GPU), the exact
//      4-point DLT (8x8 Gaussian elimination with partial pivoting, fp64), inliers = forward transfer error < threshold
//   3. the hypothesis with the most inliers wins (ties: lowest hypothesis index)
//   4. two local-optimisation rounds: least-squares DLT over the current inliers (8x8 normal equations, fp64), re-score
//   5. H is de-normalised and scaled to H[2][2] = 1 (as OpenCV returns it)
// HG_SLICES CTAs per pair score disjoint hypothesis ranges (64-bit atomicMax on a per-pair key); the last CTA of a pair to
// finish runs the local optimisation.  fp64 throughout (a B200 does the 2 G fp64 operations of a 64-pair batch in well under a millisecond, and
// CPU / GPU then agree on every inlier decision).
#include "common.cuh"

namespace xp {

constexpr int HG_THREADS = 256;
constexpr int HG_MAX_MATCHES = 11264;  // matches held in shared memory (20 B each); later ones (keypoint order) are not used
constexpr int HG_SLICES = 8;          // CTAs per pair in the hypothesis phase (the last one to finish does the refinement)

__host__ __device__ __forceinline__ uint32_t hg_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// four distinct match indices of hypothesis t of pair b; false if m < 4
__device__ __forceinline__ bool hg_sample(uint32_t seed, uint32_t b, uint32_t t, int m, int (&s)[4]) {
    int n = 0;
    for (uint32_t c = 0; c < 64 && n < 4; ++c) {
        const int idx = (int)(hg_mix(seed + 0x9E3779B9u * b + 0x85EBCA6Bu * t + 0xC2B2AE35u * c) % (uint32_t)m);
        bool dup = false;
        for (int q = 0; q < n; ++q) dup = dup || s[q] == idx;
        if (!dup) s[n++] = idx;
    }
    return n == 4;
}

// solve the 8x8 system a[r][0..7] h = a[r][8] in place (partial pivoting); false if singular
__device__ bool hg_solve8(double (&a)[8][9], double (&h)[8]) {
    for (int c = 0; c < 8; ++c) {
        int piv = c;
        double best = fabs(a[c][c]);
        for (int r = c + 1; r < 8; ++r) {
            const double v = fabs(a[r][c]);
            if (v > best) { best = v; piv = r; }
        }
        if (best < 1e-12) return false;
        if (piv != c)
            for (int k = c; k < 9; ++k) { const double tmp = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = tmp; }
        const double inv = 1.0 / a[c][c];
        for (int r = c + 1; r < 8; ++r) {
            const double f = a[r][c] * inv;
            for (int k = c; k < 9; ++k) a[r][k] -= f * a[c][k];
        }
    }
    for (int r = 7; r >= 0; --r) {
        double s = a[r][8];
        for (int k = r + 1; k < 8; ++k) s -= a[r][k] * h[k];
        h[r] = s / a[r][r];
    }
    return true;
}

// forward transfer error^2 of (x, y) -> (u, v) under h (h[8] == 1 implied); huge if the point maps to infinity
__device__ __forceinline__ double hg_err2(const double* h, double x, double y, double u, double v) {
    const double w = h[6] * x + h[7] * y + 1.0;
    if (fabs(w) < 1e-12) return 1e300;
    const double iw = 1.0 / w;
    const double du = (h[0] * x + h[1] * y + h[2]) * iw - u, dv = (h[3] * x + h[4] * y + h[5]) * iw - v;
    return du * du + dv * dv;
}

struct HomParams {
    const int32_t* kp1; const int32_t* kp2; const int32_t* n1; const int32_t* match_idx;
    double* H; uint8_t* inlier; int32_t* n_inl;
    int k, cap, iters, lo_rounds; uint32_t seed;      // cap = min(k, HG_MAX_MATCHES): capacity of the match arrays
    double cx, cy, inv_s, thr2n;       // normalisation x' = (x - cx) * inv_s; squared threshold in normalised units
};

__global__ void __launch_bounds__(HG_THREADS) homography_kernel(const HomParams p) {
    extern __shared__ __align__(16) uint8_t hg_smem[];
    float* px = reinterpret_cast<float*>(hg_smem);          // [k] normalised source / target coordinates of match j
    float* py = px + p.cap; float* qx = py + p.cap; float* qy = qx + p.cap;
    int* src = reinterpret_cast<int*>(qy + p.cap);          // [cap] keypoint index of match j
    __shared__ int warp_cnt[HG_THREADS / 32 + 1];
    __shared__ unsigned long long best_key;
    __shared__ double hcur[8];
    __shared__ int ok_s;
    __shared__ double red[HG_THREADS / 32][45];
    const int b = blockIdx.x, slice = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    // scratch inside the outputs until the last CTA overwrites them: the winning key in H[b][0], the arrival count in n_inl[b]
    unsigned long long* gkey = reinterpret_cast<unsigned long long*>(p.H + (int64_t)b * 9);
    __shared__ int last_s;
    const int32_t* kp1 = p.kp1 + (int64_t)b * p.k * 2;
    const int32_t* kp2 = p.kp2 + (int64_t)b * p.k * 2;
    const int32_t* mi = p.match_idx + (int64_t)b * p.k;
    const int n1 = min(p.n1 ? p.n1[b] : p.k, p.k);
    uint8_t* inl = p.inlier ? p.inlier + (int64_t)b * p.k : nullptr;

    // ---- 1. compact the matches in keypoint order
    int base = 0;
    for (int i0 = 0; i0 < p.k; i0 += HG_THREADS) {
        const int i = i0 + tid;
        const int j2 = (i < n1) ? mi[i] : -1;
        const bool has = j2 >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, has);
        if (lane == 0) warp_cnt[wrp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < wrp; ++w) off += warp_cnt[w];
        const int j = off + __popc(bal & ((1u << lane) - 1u));
        if (has && j < p.cap) {
            px[j] = (float)(((double)kp1[2 * i + 1] - p.cx) * p.inv_s);      // (x, y) = (kp[1], kp[0])
            py[j] = (float)(((double)kp1[2 * i] - p.cy) * p.inv_s);
            qx[j] = (float)(((double)kp2[2 * j2 + 1] - p.cx) * p.inv_s);
            qy[j] = (float)(((double)kp2[2 * j2] - p.cy) * p.inv_s);
            src[j] = i;
        }
        for (int w = 0; w < HG_THREADS / 32; ++w) base += warp_cnt[w];
        __syncthreads();
    }
    const int m = min(base, p.cap);
    if (tid == 0) { best_key = 0ull; ok_s = 0; }
    __syncthreads();
    // ---- 2./3. hypotheses of this slice (m < 4: none -- the reference returns H_est = None, evaluation.py:364-366)
    for (int t = slice * HG_THREADS + tid; t < p.iters && m >= 4; t += HG_SLICES * HG_THREADS) {
        int s[4];
        if (!hg_sample(p.seed, (uint32_t)b, (uint32_t)t, m, s)) continue;
        double a[8][9], h[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double x = px[s[q]], y = py[s[q]], u = qx[s[q]], v = qy[s[q]];
            double* r0 = a[2 * q]; double* r1 = a[2 * q + 1];
            r0[0] = x; r0[1] = y; r0[2] = 1.0; r0[3] = 0.0; r0[4] = 0.0; r0[5] = 0.0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = u;
            r1[0] = 0.0; r1[1] = 0.0; r1[2] = 0.0; r1[3] = x; r1[4] = y; r1[5] = 1.0; r1[6] = -v * x; r1[7] = -v * y; r1[8] = v;
        }
        if (!hg_solve8(a, h)) continue;
        int cnt = 0;
        for (int j = 0; j < m; ++j) cnt += hg_err2(h, px[j], py[j], qx[j], qy[j]) < p.thr2n;
        // most inliers, then lowest hypothesis index
        atomicMax(&best_key, ((unsigned long long)(uint32_t)cnt << 32) | (unsigned long long)(0xffffffffu - (uint32_t)t));
    }
    __syncthreads();
    if (tid == 0) {
        if (best_key) atomicMax(gkey, best_key);
        __threadfence();
        last_s = atomicAdd(p.n_inl + b, 1) == HG_SLICES - 1;     // the counter was zeroed by the host wrapper
    }
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    if (inl)
        for (int i = tid; i < p.k; i += HG_THREADS) inl[i] = 0;
    if (tid == 0) {
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(gkey);
        if ((key >> 32) >= 4) {
            const uint32_t t = 0xffffffffu - (uint32_t)(key & 0xffffffffull);
            int s[4];
            hg_sample(p.seed, (uint32_t)b, t, m, s);
            double a[8][9], h[8];
            for (int q = 0; q < 4; ++q) {
                const double x = px[s[q]], y = py[s[q]], u = qx[s[q]], v = qy[s[q]];
                double* r0 = a[2 * q]; double* r1 = a[2 * q + 1];
                r0[0] = x; r0[1] = y; r0[2] = 1.0; r0[3] = 0.0; r0[4] = 0.0; r0[5] = 0.0; r0[6] = -u * x; r0[7] = -u * y; r0[8] = u;
                r1[0] = 0.0; r1[1] = 0.0; r1[2] = 0.0; r1[3] = x; r1[4] = y; r1[5] = 1.0; r1[6] = -v * x; r1[7] = -v * y; r1[8] = v;
            }
            if (hg_solve8(a, h)) {
                for (int q = 0; q < 8; ++q) hcur[q] = h[q];
                ok_s = 1;
            }
        }
    }
    __syncthreads();
    if (!ok_s) {
        if (tid < 9) p.H[(int64_t)b * 9 + tid] = 0.0;
        if (tid == 0) p.n_inl[b] = -1;
        return;
    }

    // ---- 4. local optimisation: least-squares DLT over the inliers of the current H (normal equations, 36 + 8 sums + count)
    for (int round = 0; round < p.lo_rounds; ++round) {
        double acc[45];
#pragma unroll
        for (int q = 0; q < 45; ++q) acc[q] = 0.0;
        double h[8];
        for (int q = 0; q < 8; ++q) h[q] = hcur[q];
        for (int j = tid; j < m; j += HG_THREADS) {
            const double x = px[j], y = py[j], u = qx[j], v = qy[j];
            if (!(hg_err2(h, x, y, u, v) < p.thr2n)) continue;
            const double r0[8] = {x, y, 1.0, 0.0, 0.0, 0.0, -u * x, -u * y};
            const double r1[8] = {0.0, 0.0, 0.0, x, y, 1.0, -v * x, -v * y};
            int q = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int k2 = i; k2 < 8; ++k2) acc[q++] += r0[i] * r0[k2] + r1[i] * r1[k2];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[36 + i] += r0[i] * u + r1[i] * v;
            acc[44] += 1.0;
        }
#pragma unroll
        for (int q = 0; q < 45; ++q) {
            double v = acc[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[wrp][q] = v;
        }
        __syncthreads();
        if (tid == 0) {
            double a[8][9], sum[45], hn[8];
            for (int q = 0; q < 45; ++q) {
                double v = 0.0;
                for (int w = 0; w < HG_THREADS / 32; ++w) v += red[w][q];
                sum[q] = v;
            }
            if (sum[44] >= 4.0) {
                int q = 0;
                for (int i = 0; i < 8; ++i)
                    for (int k2 = i; k2 < 8; ++k2) { a[i][k2] = sum[q]; a[k2][i] = sum[q]; ++q; }
                for (int i = 0; i < 8; ++i) a[i][8] = sum[36 + i];
                if (hg_solve8(a, hn))
                    for (int i = 0; i < 8; ++i) hcur[i] = hn[i];
            }
        }
        __syncthreads();
    }

    // ---- 5. final inlier mask, de-normalised H scaled to H[2][2] = 1
    double h[8];
    for (int q = 0; q < 8; ++q) h[q] = hcur[q];
    int cnt = 0;
    for (int j = tid; j < m; j += HG_THREADS) {
        const bool in = hg_err2(h, px[j], py[j], qx[j], qy[j]) < p.thr2n;
        cnt += in;
        if (inl && in) inl[src[j]] = 1;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) warp_cnt[wrp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int total = 0;
        for (int w = 0; w < HG_THREADS / 32; ++w) total += warp_cnt[w];
        p.n_inl[b] = total;
        // H = T^-1 Hn T with T = [[s, 0, -s cx], [0, s, -s cy], [0, 0, 1]] (s = inv_s) for both images
        const double s = p.inv_s, cx = p.cx, cy = p.cy;
        const double Hn[9] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 1.0};
        double A[9];                                   // Hn T
        for (int r = 0; r < 3; ++r) {
            A[3 * r] = Hn[3 * r] * s;
            A[3 * r + 1] = Hn[3 * r + 1] * s;
            A[3 * r + 2] = Hn[3 * r + 2] - s * (Hn[3 * r] * cx + Hn[3 * r + 1] * cy);
        }
        double Hf[9];                                  // T^-1 = [[1/s, 0, cx], [0, 1/s, cy], [0, 0, 1]]
        const double is = 1.0 / s;
        for (int c = 0; c < 3; ++c) {
            Hf[c] = A[c] * is + cx * A[6 + c];
            Hf[3 + c] = A[3 + c] * is + cy * A[6 + c];
            Hf[6 + c] = A[6 + c];
        }
        const double n = fabs(Hf[8]) > 1e-300 ? 1.0 / Hf[8] : 1.0;
        for (int q = 0; q < 9; ++q) p.H[(int64_t)b * 9 + q] = Hf[q] * n;
    }
}

}  // namespace xp

using namespace xp;

extern "C" int xp_estimate_homography(const int32_t* kp1, const int32_t* kp2, const int32_t* n1, const int32_t* match_idx,
                                      int64_t B, int64_t k, int64_t height, int64_t width, int32_t iters, float reproj_threshold,
                                      int32_t lo_rounds, uint32_t seed, double* H, uint8_t* inlier_mask, int32_t* n_inliers,
                                      xp_stream_t stream) {
    XP_REQUIRE(kp1 && kp2 && match_idx && H && n_inliers, "xp_estimate_homography: NULL tensor pointer");
    XP_REQUIRE(B >= 0 && B <= 65535 && k > 0 && k <= (1 << 20), "xp_estimate_homography: need 0 <= B <= 65535, 0 < k <= 2^20");
    XP_REQUIRE(height > 0 && width > 0 && iters > 0 && iters <= (1 << 20) && reproj_threshold > 0.0f && lo_rounds >= 0 && lo_rounds <= 16,
               "xp_estimate_homography: bad parameters");
    if (B == 0) return XP_OK;
    HomParams p;
    p.kp1 = kp1; p.kp2 = kp2; p.n1 = n1; p.match_idx = match_idx; p.H = H; p.inlier = inlier_mask; p.n_inl = n_inliers;
    p.k = (int)k; p.cap = (int)(k < HG_MAX_MATCHES ? k : HG_MAX_MATCHES); p.iters = iters; p.lo_rounds = lo_rounds; p.seed = seed;
    const double s = 0.5 * (double)(height > width ? height : width);
    p.cx = 0.5 * (double)width; p.cy = 0.5 * (double)height; p.inv_s = 1.0 / s;
    p.thr2n = ((double)reproj_threshold / s) * ((double)reproj_threshold / s);
    const int smem = p.cap * 20;
    XP_CUDA_OK(cudaFuncSetAttribute(homography_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // per-pair scratch lives in the outputs (see the kernel): winning key in H[b][0], arrival counter in n_inliers[b]
    XP_CUDA_OK(cudaMemsetAsync(H, 0, sizeof(double) * 9 * B, (cudaStream_t)stream));
    XP_CUDA_OK(cudaMemsetAsync(n_inliers, 0, sizeof(int32_t) * B, (cudaStream_t)stream));
    homography_kernel<<<dim3((unsigned)B, HG_SLICES), HG_THREADS, smem, (cudaStream_t)stream>>>(p);
    XP_LAUNCH_CHECK("homography_kernel");
    return XP_OK;
}
