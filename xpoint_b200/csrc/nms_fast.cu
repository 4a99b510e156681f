// nms_fast.cu -- greedy box NMS + top-k + raster compaction for images whose width is a multiple of 8.
//
// Same algorithm and same results as box_nms_kernel (postprocess.cu: the data-parallel fixed point that equals
// sequential greedy NMS, utils/utils.py:148-192 via torchvision.ops.nms), restructured for memory-level parallelism:
// box_nms_kernel walks the image one pixel per lane and pass (320 dependent L2 round trips per thread and pass at
// 512x640, ~25 passes), which makes it latency-bound at one CTA per image.  Here a lane owns 8 consecutive pixels:
//   * every scan of the state map (alive / new / kept masks, counts, radix histograms, compaction) is ONE 8-byte load per
//     lane and step -- 40 steps per pass instead of 320, the per-pixel work happens on register masks;
//   * the footprint scan of phase A is resumable: a one-byte cursor per pixel (second plane of the workspace) remembers the
//     first offset that has not been ruled out, so that over all rounds every (pixel, offset) pair is examined once
//     (box_nms_kernel rescans the whole footprint of every undecided pixel in every round: 14 M warp instructions per
//     image, 80 % of them in later rounds); lanes scan their own pixels 16 offsets at a time, the rare deep scans are
//     finished by the whole warp 32 offsets at a time;
//   * counting the undecided pixels is folded into phase A (one pass less per round).
// The scatter per kept pixel, the tie rule (score desc, flat index asc), the radix select and the raster-order
// keypoint compaction are those of box_nms_kernel.  Needs a footprint of at most 255 offsets (box size <= 8).
#include "common.cuh"

namespace xp {

constexpr int NF_THREADS = 1024;
constexpr int NF_PX = 8;
constexpr int NF_MAX_OFFS = 31 * 31;

enum : uint8_t { NF_NONE = 0, NF_ALIVE = 1, NF_KEPT = 2, NF_NEW = 3 };

struct NmsParams {   // identical to the one in postprocess.cu
    const float* prob; float* out; uint8_t* state; int32_t* kp; int32_t* kp_count;
    int H, W; float size, min_prob, iou, kp_thr; int64_t topk, kp_cap;
};

__device__ __forceinline__ bool nf_better(float t, int r, float s, int q) { return t > s || (t == s && r < q); }

// 8 state bytes -> 8-bit mask of the bytes equal to v
__device__ __forceinline__ unsigned nf_mask(uint2 s, unsigned v) {
    const unsigned rep = v * 0x01010101u;
    const unsigned a = __vcmpeq4(s.x, rep) & 0x01010101u, b = __vcmpeq4(s.y, rep) & 0x01010101u;
    const unsigned la = (a | (a >> 7) | (a >> 14) | (a >> 21)) & 0xfu, lb = (b | (b >> 7) | (b >> 14) | (b >> 21)) & 0xfu;
    return la | (lb << 4);
}

__device__ __forceinline__ int nf_block_exclusive_scan(int v, int* warp_sums, int& total) {
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
        warp_sums[lane] = winc - ws;
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    total = warp_sums[32];
    return warp_sums[wid] + inc - v;
}

__global__ void __launch_bounds__(NF_THREADS) box_nms_fast_kernel(const NmsParams p) {
    __shared__ int8_t off_dy[NF_MAX_OFFS + 32], off_dx[NF_MAX_OFFS + 32];
    __shared__ int n_offs_s, alive_s;
    __shared__ int scan_ws[33];
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_remaining;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int H = p.H, W = p.W, HW = H * W;
    const int cpr = W / NF_PX;                    // chunks per row (W % 8 == 0)
    const int nchunks = HW / NF_PX;
    const int nchunks_r = (nchunks + 31) & ~31;   // warp-uniform bound for the ballot phases
    const float* prob = p.prob + (int64_t)b * HW;
    uint8_t* st = p.state + (int64_t)b * HW;
    const uint2* st8 = reinterpret_cast<const uint2*>(st);

    if (tid == 0) {   // suppression footprint, nearest ring first (same construction as box_nms_kernel)
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
                    if ((double)(inter / (area2 - inter)) > (double)p.iou) {
                        off_dy[n] = (int8_t)dy; off_dx[n] = (int8_t)dx; ++n;
                    }
                }
        n_offs_s = n;
        alive_s = 0;
    }
    // initial state: candidates = pixels above the threshold
    for (int c = tid; c < nchunks; c += NF_THREADS) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * NF_PX));
        const float4 d = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * NF_PX) + 1);
        uint2 s;
        s.x = (a.x > p.min_prob ? 1u : 0u) | (a.y > p.min_prob ? 1u << 8 : 0u) | (a.z > p.min_prob ? 1u << 16 : 0u) |
              (a.w > p.min_prob ? 1u << 24 : 0u);
        s.y = (d.x > p.min_prob ? 1u : 0u) | (d.y > p.min_prob ? 1u << 8 : 0u) | (d.z > p.min_prob ? 1u << 16 : 0u) |
              (d.w > p.min_prob ? 1u << 24 : 0u);
        reinterpret_cast<uint2*>(st)[c] = s;
    }
    __syncthreads();
    const int n_offs = n_offs_s;

    uint8_t* pos = p.state + (int64_t)gridDim.x * HW + (int64_t)b * HW;     // second byte plane: per-pixel scan cursor
    const uint2* pos8 = reinterpret_cast<const uint2*>(pos);
    for (int c = tid; c < nchunks; c += NF_THREADS) reinterpret_cast<uint2*>(pos)[c] = make_uint2(0u, 0u);
    __syncthreads();
    constexpr int SCAN_LIMIT = 16;                 // lane-serial steps per pixel and round before the warp takes over

    while (true) {
        // ---- phase A: an undecided candidate with no undecided higher-priority candidate in its footprint becomes NEW.
        // The footprint is scanned nearest-first and the scan is RESUMABLE: `r blocks q` needs r undecided (a state that
        // is never re-entered) and r better than q (static), so an offset that did not block q once never will.  pos[q]
        // is the first offset not yet ruled out (the current blocker while q is blocked): over ALL rounds every
        // (pixel, offset) pair is examined at most once, plus one re-check of the blocker per round.
        int local_alive = 0;
        for (int c0 = tid - lane; c0 < nchunks_r; c0 += NF_THREADS) {
            const int c = c0 + lane;
            const bool valid = c < nchunks;
            uint2 s = make_uint2(0u, 0u);
            if (valid) s = st8[c];
            unsigned am = nf_mask(s, NF_ALIVE);
            local_alive += __popc(am);
            unsigned deep = 0;                     // pixels whose scan is handed to the whole warp
            if (am) {
                const uint2 pv = pos8[c];
                const int y = c / cpr, x0 = (c - y * cpr) * NF_PX;
                while (am) {
                    const int j = __ffs(am) - 1;
                    am &= am - 1;
                    const int x = x0 + j, q = c * NF_PX + j;
                    const float sq = prob[q];
                    int o = (int)(((j < 4 ? pv.x : pv.y) >> (8 * (j & 3))) & 0xffu);
                    const int o_end = min(o + SCAN_LIMIT, n_offs);
                    bool blocked = false;
                    for (; o < o_end; ++o) {
                        const int yy = y + off_dy[o], xx = x + off_dx[o];
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        const int r = yy * W + xx;
                        const uint8_t sr = st[r];
                        if ((sr == NF_ALIVE || sr == NF_NEW) && nf_better(prob[r], r, sq, q)) { blocked = true; break; }
                    }
                    pos[q] = (uint8_t)o;
                    if (!blocked) {
                        if (o >= n_offs) st[q] = NF_NEW;        // NEW is treated like ALIVE by concurrent readers
                        else deep |= 1u << j;
                    }
                }
            }
            // deep scans: the 32 lanes test 32 offsets at a time from the pixel's cursor
            __syncwarp();                          // cursors / states written above are visible to the whole warp
#pragma unroll 1
            for (int j = 0; j < NF_PX; ++j) {
                unsigned todo = __ballot_sync(0xffffffffu, (deep >> j) & 1u);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int cc = c0 + src;
                    const int cy = cc / cpr, cx = (cc - cy * cpr) * NF_PX + j;
                    const int cq = cy * W + cx;
                    const float cs = prob[cq];
                    int first = n_offs;
                    for (int o0 = pos[cq]; o0 < n_offs; o0 += 32) {
                        const int o = o0 + lane;
                        bool hit = false;
                        if (o < n_offs) {
                            const int yy = cy + off_dy[o], xx = cx + off_dx[o];
                            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                                const int r = yy * W + xx;
                                const uint8_t sr = st[r];
                                hit = (sr == NF_ALIVE || sr == NF_NEW) && nf_better(prob[r], r, cs, cq);
                            }
                        }
                        const unsigned hm = __ballot_sync(0xffffffffu, hit);
                        if (hm) { first = o0 + __ffs(hm) - 1; break; }
                    }
                    if (lane == src) {
                        if (first >= n_offs) st[cq] = NF_NEW;
                        else pos[cq] = (uint8_t)first;
                    }
                    __syncwarp();
                }
            }
        }
        if (local_alive) atomicAdd(&alive_s, local_alive);
        __syncthreads();
        const int alive = alive_s;
        __syncthreads();
        if (alive == 0) break;
        if (tid == 0) alive_s = 0;
        // ---- phase B: every NEW pixel suppresses the undecided candidates in its footprint and becomes KEPT
        for (int c0 = tid - lane; c0 < nchunks_r; c0 += NF_THREADS) {
            const int c = c0 + lane;
            uint2 s = make_uint2(0u, 0u);
            if (c < nchunks) s = st8[c];
            const unsigned nm = nf_mask(s, NF_NEW);
#pragma unroll
            for (int j = 0; j < NF_PX; ++j) {
                unsigned todo = __ballot_sync(0xffffffffu, (nm >> j) & 1u);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int cc = c0 + src;
                    const int cy = cc / cpr, cx = (cc - cy * cpr) * NF_PX + j;
                    for (int o = lane; o < n_offs; o += 32) {
                        const int yy = cy + off_dy[o], xx = cx + off_dx[o];
                        if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                        const int r = yy * W + xx;
                        if (st[r] == NF_ALIVE) st[r] = NF_NONE;
                    }
                }
                if ((nm >> j) & 1u) st[c * NF_PX + j] = NF_KEPT;
            }
        }
        __syncthreads();
    }

    // ---- top-k by score among kept (ties at the threshold: lower flat index first) ----
    unsigned thr_bits = 0;      // keep score bits > thr_bits, plus the first `need_eq` with == thr_bits
    int need_eq = -1;           // -1: keep everything
    if (p.topk > 0) {
        int kept_local = 0;
        for (int c = tid; c < nchunks; c += NF_THREADS) kept_local += __popc(nf_mask(st8[c], NF_KEPT));
        int total;
        nf_block_exclusive_scan(kept_local, scan_ws, total);
        if (total > p.topk) {
            if (tid == 0) { sel_prefix = 0; sel_remaining = (unsigned)p.topk; }
            for (int pass = 0; pass < 4; ++pass) {
                const int shift = 24 - 8 * pass;
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                const unsigned prefix = sel_prefix;
                const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
                for (int c = tid; c < nchunks; c += NF_THREADS) {
                    unsigned km = nf_mask(st8[c], NF_KEPT);
                    while (km) {
                        const int j = __ffs(km) - 1;
                        km &= km - 1;
                        const unsigned bits = __float_as_uint(prob[c * NF_PX + j]);
                        if ((bits & himask) == prefix) atomicAdd(&hist[(bits >> shift) & 255u], 1u);
                    }
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
            need_eq = (int)sel_remaining;
        }
    }

    // ---- selection + raster-order compaction.  Warp w owns a contiguous range of chunks; within a step the lanes hold
    //      consecutive chunks, so (lane prefix, bit order) is raster order. ----
    float* out = p.out ? p.out + (int64_t)b * HW : nullptr;
    int32_t* kp = p.kp ? p.kp + (int64_t)b * p.kp_cap * 2 : nullptr;
    const int wid = tid >> 5;
    const int span = (((nchunks + 31) / 32 + 31) / 32) * 32;      // chunks per warp, multiple of 32
    const int w0 = min(wid * span, nchunks), w1 = min(w0 + span, nchunks);
    int eq_base = 0;
    if (need_eq >= 0) {
        int eq_warp = 0;
        for (int c0 = w0; c0 < w1; c0 += 32) {
            const int c = c0 + lane;
            int cnt = 0;
            if (c < w1) {
                unsigned km = nf_mask(st8[c], NF_KEPT);
                while (km) { const int j = __ffs(km) - 1; km &= km - 1; cnt += __float_as_uint(prob[c * NF_PX + j]) == thr_bits; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            eq_warp += cnt;
        }
        int tot;
        const int ex = nf_block_exclusive_scan(lane == 0 ? eq_warp : 0, scan_ws, tot);
        eq_base = __shfl_sync(0xffffffffu, ex, 0);
    }
    // pass 1: final kept mask per chunk (written back as states), number of keypoints per warp
    int kp_warp = 0;
    for (int c0 = w0; c0 < w1; c0 += 32) {
        const int c = c0 + lane;
        unsigned keepm = 0, eqm = 0, kpm = 0;
        float sc[NF_PX];
#pragma unroll
        for (int j = 0; j < NF_PX; ++j) sc[j] = 0.0f;
        uint2 s = make_uint2(0u, 0u);
        if (c < w1) {
            s = st8[c];
            const unsigned km = nf_mask(s, NF_KEPT);
            if (km) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * NF_PX));
                const float4 d = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * NF_PX) + 1);
                sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = d.x; sc[5] = d.y; sc[6] = d.z; sc[7] = d.w;
#pragma unroll
                for (int j = 0; j < NF_PX; ++j) {
                    if (!((km >> j) & 1u)) continue;
                    const unsigned bits = __float_as_uint(sc[j]);
                    if (need_eq < 0 || bits > thr_bits) keepm |= 1u << j;
                    else if (bits == thr_bits) eqm |= 1u << j;
                }
            }
        }
        if (need_eq >= 0) {
            // rank of my == threshold pixels among all of them in raster order
            int cnt = __popc(eqm), inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            int rank = eq_base + inc - cnt;
            unsigned e = eqm;
            while (e) { const int j = __ffs(e) - 1; e &= e - 1; if (rank < need_eq) keepm |= 1u << j; ++rank; }
            eq_base += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (c < w1) {
            // final state bytes: KEPT where selected, NONE elsewhere
            uint2 ns;
            ns.x = ((keepm & 1u) ? 2u : 0u) | ((keepm & 2u) ? 2u << 8 : 0u) | ((keepm & 4u) ? 2u << 16 : 0u) | ((keepm & 8u) ? 2u << 24 : 0u);
            ns.y = ((keepm & 16u) ? 2u : 0u) | ((keepm & 32u) ? 2u << 8 : 0u) | ((keepm & 64u) ? 2u << 16 : 0u) | ((keepm & 128u) ? 2u << 24 : 0u);
            reinterpret_cast<uint2*>(st)[c] = ns;
#pragma unroll
            for (int j = 0; j < NF_PX; ++j) if (((keepm >> j) & 1u) && sc[j] > p.kp_thr) kpm |= 1u << j;
        }
        int cnt = __popc(kpm);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        kp_warp += cnt;
    }
    int kp_total = 0, kp_off = 0;
    if (kp || p.kp_count) {
        const int ex = nf_block_exclusive_scan(lane == 0 ? kp_warp : 0, scan_ws, kp_total);
        kp_off = __shfl_sync(0xffffffffu, ex, 0);
    }
    __syncthreads();     // final states visible to every warp (the output map below reads other warps' chunks)
    if (kp) {
        for (int c0 = w0; c0 < w1; c0 += 32) {
            const int c = c0 + lane;
            unsigned kpm = 0;
            if (c < w1) {
                unsigned km = nf_mask(st8[c], NF_KEPT);
                while (km) { const int j = __ffs(km) - 1; km &= km - 1; if (prob[c * NF_PX + j] > p.kp_thr) kpm |= 1u << j; }
            }
            int cnt = __popc(kpm), inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            int pos = kp_off + inc - cnt;
            while (kpm) {
                const int j = __ffs(kpm) - 1;
                kpm &= kpm - 1;
                if (pos < p.kp_cap) { const int q = c * NF_PX + j; kp[2 * pos] = q / W; kp[2 * pos + 1] = q % W; }
                ++pos;
            }
            kp_off += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    if (p.kp_count && tid == 0) p.kp_count[b] = kp_total;
    if (out) {
        for (int c = tid; c < nchunks; c += NF_THREADS) {
            const unsigned km = nf_mask(st8[c], NF_KEPT);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), d = a;
            if (km) {
                const float4 pa = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * NF_PX));
                const float4 pd = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)c * NF_PX) + 1);
                a = make_float4((km & 1u) ? pa.x : 0.f, (km & 2u) ? pa.y : 0.f, (km & 4u) ? pa.z : 0.f, (km & 8u) ? pa.w : 0.f);
                d = make_float4((km & 16u) ? pd.x : 0.f, (km & 32u) ? pd.y : 0.f, (km & 64u) ? pd.z : 0.f, (km & 128u) ? pd.w : 0.f);
            }
            reinterpret_cast<float4*>(out + (int64_t)c * NF_PX)[0] = a;
            reinterpret_cast<float4*>(out + (int64_t)c * NF_PX)[1] = d;
        }
    }
}

// launched from xp_box_nms (postprocess.cu) when W % 8 == 0 and the maps are 16-byte aligned
int launch_box_nms_fast(const float* prob, float* prob_nms, int64_t B, int64_t H, int64_t W, float size, float min_prob, float iou,
                        int64_t keep_top_k, float kp_threshold, int32_t* keypoints, int32_t* kp_count, int64_t kp_capacity,
                        void* workspace, cudaStream_t st) {
    NmsParams p;
    p.prob = prob; p.out = prob_nms; p.state = (uint8_t*)workspace; p.kp = keypoints; p.kp_count = kp_count;
    p.H = (int)H; p.W = (int)W; p.size = size; p.min_prob = min_prob; p.iou = iou; p.kp_thr = kp_threshold;
    p.topk = keep_top_k; p.kp_cap = kp_capacity;
    box_nms_fast_kernel<<<(unsigned)B, NF_THREADS, 0, st>>>(p);
    XP_LAUNCH_CHECK("box_nms_fast_kernel");
    return XP_OK;
}

}  // namespace xp
