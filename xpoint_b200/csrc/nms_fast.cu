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
#include <stdlib.h>

#include "common.cuh"

namespace xp {

constexpr int NF_THREADS = 1024;
constexpr int NF_PX = 8;
constexpr int NF_MAX_OFFS = 31 * 31;

enum : uint8_t { NF_NONE = 0, NF_ALIVE = 1, NF_KEPT = 2, NF_NEW = 3 };

struct NmsFastParams {
    const float* prob; float* out; uint8_t* state; uint8_t* cursor; int32_t* kp; int32_t* kp_count;
    int H, W; float size, min_prob, iou, kp_thr; int64_t topk, kp_cap;
    int state_ready;     // 1: `state` already holds ALIVE / KEPT / NONE bytes (dense rounds ran first)
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

__global__ void __launch_bounds__(NF_THREADS) box_nms_fast_kernel(const NmsFastParams p) {
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
    // initial state: candidates = pixels above the threshold (unless the dense rounds left their state here)
    for (int c = tid; c < nchunks && !p.state_ready; c += NF_THREADS) {
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

    uint8_t* pos = p.cursor + (int64_t)b * HW;                              // second byte plane: per-pixel scan cursor
    const uint2* pos8 = reinterpret_cast<const uint2*>(pos);
    for (int c = tid; c < nchunks; c += NF_THREADS) reinterpret_cast<uint2*>(pos)[c] = make_uint2(0u, 0u);
    __syncthreads();
    constexpr int SCAN_LIMIT = 16;                 // lane-serial steps per pixel and round before the warp takes over

#ifdef XP_NMS_TIMING
    long long t_start = clock64(), t_a = 0, t_b = 0; int n_rounds = 0;
#endif
    while (true) {
#ifdef XP_NMS_TIMING
        long long t0 = clock64();
#endif
        // ---- phase A: an undecided candidate with no undecided higher-priority candidate in its footprint becomes NEW.
        // The footprint is scanned nearest-first and the scan is RESUMABLE: `r blocks q` needs r undecided (a state that
        // is never re-entered) and r better than q (static), so an offset that did not block q once never will.  pos[q]
        // is the first offset not yet ruled out (the current blocker while q is blocked): over ALL rounds every
        // (pixel, offset) pair is examined at most once, plus one re-check of the blocker per round.
        // (the state words of four chunk steps are fetched together: one L2 round trip per four steps, not per step --
        //  in phase A only the owner thread changes a pixel's state, so the early copies stay valid)
        int local_alive = 0;
        for (int cb = tid - lane; cb < nchunks_r; cb += 4 * NF_THREADS) {
          uint2 sv0, sv1, sv2, sv3;
          { const int c = cb + lane; sv0 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
          { const int c = cb + NF_THREADS + lane; sv1 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
          { const int c = cb + 2 * NF_THREADS + lane; sv2 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
          { const int c = cb + 3 * NF_THREADS + lane; sv3 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
#pragma unroll 1
          for (int c0 = cb; c0 < min(cb + 4 * NF_THREADS, nchunks_r); c0 += NF_THREADS) {
            const int c = c0 + lane;
            const uint2 s = sv0;
            sv0 = sv1; sv1 = sv2; sv2 = sv3;
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
            if (!__any_sync(0xffffffffu, deep != 0)) continue;
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
        }
        if (local_alive) atomicAdd(&alive_s, local_alive);
        __syncthreads();
        const int alive = alive_s;
        __syncthreads();
#ifdef XP_NMS_TIMING
        long long t1 = clock64(); t_a += t1 - t0; ++n_rounds;
        if (b == 0 && tid == 0) printf("round %d alive %d phaseA %lld\n", n_rounds, alive, t1 - t0);
#endif
        if (alive == 0) break;
        if (tid == 0) alive_s = 0;
        // ---- phase B: every NEW pixel suppresses the undecided candidates in its footprint and becomes KEPT
        // (NEW flags are only cleared by their owner, so the four-step prefetch is safe here as well)
        for (int cb = tid - lane; cb < nchunks_r; cb += 4 * NF_THREADS) {
          uint2 sv0, sv1, sv2, sv3;
          { const int c = cb + lane; sv0 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
          { const int c = cb + NF_THREADS + lane; sv1 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
          { const int c = cb + 2 * NF_THREADS + lane; sv2 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
          { const int c = cb + 3 * NF_THREADS + lane; sv3 = c < nchunks ? st8[c] : make_uint2(0u, 0u); }
#pragma unroll 1
          for (int c0 = cb; c0 < min(cb + 4 * NF_THREADS, nchunks_r); c0 += NF_THREADS) {
            const int c = c0 + lane;
            const uint2 s = sv0;
            sv0 = sv1; sv1 = sv2; sv2 = sv3;
            const unsigned nm = nf_mask(s, NF_NEW);
            if (!__any_sync(0xffffffffu, nm != 0)) continue;
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
        }
        __syncthreads();
#ifdef XP_NMS_TIMING
        t_b += clock64() - t1;
#endif
    }
#ifdef XP_NMS_TIMING
    long long t_loop = clock64();
#endif

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

#ifdef XP_NMS_TIMING
    long long t_topk = clock64();
#endif
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
#ifdef XP_NMS_TIMING
    if (b == 0 && tid == 0)
        printf("init %lld  A %lld  B %lld  topk %lld  compaction %lld cycles\n", 0LL, t_a, t_b, t_topk - t_loop, clock64() - t_topk);
    (void)t_start;
#endif
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

// ================================================================================================================
// Dense rounds.  On dense score maps (XPoint heat maps put ~half of all pixels above the threshold) the first rounds of the
// fixed point touch every pixel, and a per-candidate footprint walk (136 offsets, divergent) is the wrong shape for that.
// One dense round = the same round as a tiled, state-free image filter on all SMs:
//   NEW(q)  = alive(q) and score(q) beats every alive pixel of its footprint      -> a max filter over the footprint
//   alive'  = alive minus NEW minus every pixel with a NEW pixel in its footprint -> a dilation of the NEW mask
// The footprint {(dy, dx): |dx| <= wx[|dy|]} is a stack of centred row segments, so both filters are separable: running
// maxima per distinct half-width along x (4 arrays in shared memory), 2*ry+1 lookups along y; the NEW mask lives in
// 32-bit row words and is dilated with funnel shifts.  Ties follow greedy NMS exactly (score desc, flat index asc): the
// half of the footprint that precedes q in raster order must be beaten strictly, the other half weakly.
// A CTA owns a 32 x 64 tile (+ 2*ry halo of keys, ry halo of NEW flags, recomputed identically by the neighbours) and the
// rounds ping-pong between the two byte planes of the workspace.  After a fixed number of dense rounds the few undecided
// pixels left (<1 % on XPoint maps) are finished by box_nms_fast_kernel above, which also does top-k and compaction.
// ================================================================================================================
constexpr int ND_TX = 64, ND_TY = 32;          // tile
constexpr int ND_HX = 16;                      // key halo along x (>= 2*ry, multiple of 4: aligned 16-byte rows)
constexpr int ND_KW = ND_TX + 2 * ND_HX;       // key columns (96); key column kx <-> pixel x0 - ND_HX + kx
constexpr int ND_C0 = 8, ND_HW = 80;           // running-max / NEW columns: key columns [ND_C0, ND_C0 + ND_HW)
constexpr int ND_NWORD = 3;                    // NEW-mask words per row; bit = key column
constexpr int ND_THREADS = 512;

// compile-time footprints: half-width of the row segment at |dy| (w[0]: centre row), -1 = no such row
template <int RY, int W0, int W1, int W2, int W3, int W4, int W5, int W6> struct NdFoot {
    static constexpr int ry = RY;
    static constexpr int wx(int dy) {
        constexpr int w[7] = {W0, W1, W2, W3, W4, W5, W6};
        return w[dy];
    }
};
using NdFoot8 = NdFoot<6, 6, 6, 6, 5, 5, 4, 2>;        // box 8, IoU 0.1 (configs/cipdp.yaml:52-55): 136 offsets
using NdFoot4 = NdFoot<3, 3, 3, 2, 1, -1, -1, -1>;     // box 4, IoU 0.1: 30 offsets

template <typename F> struct NdWidths {        // distinct non-zero half-widths of rows dy >= 1, ascending
    int n = 0;
    int val[8] = {};
    int slot[8] = {};                          // slot[dy]: index into val, -1 for half-width 0
    constexpr NdWidths() {
        for (int dy = 0; dy < 8; ++dy) slot[dy] = -1;
        for (int w = 1; w <= 7; ++w) {
            bool used = false;
            for (int dy = 1; dy <= F::ry; ++dy) used = used || F::wx(dy) == w;
            if (used) {
                for (int dy = 1; dy <= F::ry; ++dy)
                    if (F::wx(dy) == w) slot[dy] = n;
                val[n++] = w;
            }
        }
    }
};

template <typename F> struct NdCfg {
    static constexpr int KH = ND_TY + 4 * F::ry;                         // key rows; key row ky <-> pixel y0 - 2*ry + ky
    static constexpr int NR = ND_TY + 2 * F::ry;                         // NEW-region rows; NEW row nr <-> key row nr + ry
    static constexpr int NWID = NdWidths<F>().n;
    static constexpr int SMEM = (KH * ND_KW + NWID * KH * ND_HW) * 4;
};

struct NmsDenseParams {
    const float* prob; const uint8_t* st_in; uint8_t* st_out;
    int H, W; float min_prob;
};

template <typename F, bool FIRST>
__global__ void __launch_bounds__(ND_THREADS, 2) nms_dense_round_kernel(const NmsDenseParams p) {
    using Cfg = NdCfg<F>;
    constexpr NdWidths<F> WD{};
    constexpr int ry = F::ry, KH = Cfg::KH, NR = Cfg::NR;
    static_assert(2 * ry <= ND_HX && ND_C0 + ND_HW + 8 <= ND_KW && ND_HX - ry >= ND_C0, "halo layout");
    extern __shared__ __align__(16) float nd_smem[];
    float* K = nd_smem;                                  // [KH][ND_KW] keys: score of alive pixels, 0 elsewhere
    float* Hm = nd_smem + KH * ND_KW;                    // [NWID][KH][ND_HW] running maxima along x, column kx - ND_C0
    __shared__ uint32_t newb[NR][ND_NWORD];              // NEW mask of the NEW region, bit = key column
    __shared__ uint32_t dil[Cfg::NWID + 1][NR][ND_NWORD];    // its dilations along x (slot NWID: undilated)
    __shared__ uint32_t supb[ND_TY][ND_NWORD];
    const int tid = threadIdx.x;
    const int H = p.H, W = p.W;
    const int64_t img = (int64_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * ND_TX, y0 = blockIdx.y * ND_TY;
    const float* prob = p.prob + img;
    const uint8_t* sin = FIRST ? nullptr : p.st_in + img;
    for (int i = tid; i < NR * ND_NWORD; i += ND_THREADS) (&newb[0][0])[i] = 0u;
    // ---- phase 1: keys, 16 bytes per thread and step (W % 8 == 0: a 4-pixel group is inside the image or outside)
    constexpr int P1 = KH * (ND_KW / 4);
#pragma unroll
    for (int it = 0; it < (P1 + ND_THREADS - 1) / ND_THREADS; ++it) {
        const int i = it * ND_THREADS + tid;
        if (i < P1) {
            const int ky = i / (ND_KW / 4), kx = (i % (ND_KW / 4)) * 4;
            const int y = y0 - 2 * ry + ky, x = x0 - ND_HX + kx;
            float4 k = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(prob + (int64_t)y * W + x));
                if (FIRST) {
                    k = make_float4(s.x > p.min_prob ? s.x : 0.f, s.y > p.min_prob ? s.y : 0.f, s.z > p.min_prob ? s.z : 0.f,
                                    s.w > p.min_prob ? s.w : 0.f);
                } else {
                    const uint32_t a = *reinterpret_cast<const uint32_t*>(sin + (int64_t)y * W + x);
                    k = make_float4((a & 0xffu) == NF_ALIVE ? s.x : 0.f, ((a >> 8) & 0xffu) == NF_ALIVE ? s.y : 0.f,
                                    ((a >> 16) & 0xffu) == NF_ALIVE ? s.z : 0.f, (a >> 24) == NF_ALIVE ? s.w : 0.f);
                }
            }
            *reinterpret_cast<float4*>(K + ky * ND_KW + kx) = k;
        }
    }
    __syncthreads();
    // ---- phase 2: running maxima along x for every distinct half-width, four columns per thread and step
    constexpr int P2 = KH * (ND_HW / 4);
#pragma unroll 2
    for (int i = tid; i < P2; i += ND_THREADS) {
        const int ky = i / (ND_HW / 4), kx = ND_C0 + (i % (ND_HW / 4)) * 4;
        float v[20];                                      // key columns kx - 8 .. kx + 11
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(K + ky * ND_KW + kx - 8 + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        float m[4] = {v[8], v[9], v[10], v[11]};
        int d = 1;
#pragma unroll
        for (int k = 0; k < WD.n; ++k) {
#pragma unroll
            for (; d <= WD.val[k]; ++d) {
#pragma unroll
                for (int j = 0; j < 4; ++j) m[j] = fmaxf(m[j], fmaxf(v[8 + j - d], v[8 + j + d]));
            }
            *reinterpret_cast<float4*>(Hm + (k * KH + ky) * ND_HW + kx - ND_C0) = make_float4(m[0], m[1], m[2], m[3]);
        }
    }
    __syncthreads();
    // ---- phase 3: NEW flags on the NEW region, four columns per thread and step
    constexpr int P3 = NR * (ND_HW / 4);
#pragma unroll 2
    for (int i = tid; i < P3; i += ND_THREADS) {
        const int nr = i / (ND_HW / 4), kx = ND_C0 + (i % (ND_HW / 4)) * 4;
        const int ky = nr + ry;
        float v[20];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(K + ky * ND_KW + kx - 8 + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        if (v[8] == 0.f && v[9] == 0.f && v[10] == 0.f && v[11] == 0.f) continue;      // nothing alive here
        float mb[4] = {0.f, 0.f, 0.f, 0.f}, ma[4] = {0.f, 0.f, 0.f, 0.f};   // maxima before / after q in raster order
#pragma unroll
        for (int d = 1; d <= F::wx(0); ++d) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { mb[j] = fmaxf(mb[j], v[8 + j - d]); ma[j] = fmaxf(ma[j], v[8 + j + d]); }
        }
#pragma unroll
        for (int dy = 1; dy <= ry; ++dy) {
            const float* up = WD.slot[dy] < 0 ? K + (ky - dy) * ND_KW + kx : Hm + (WD.slot[dy] * KH + ky - dy) * ND_HW + kx - ND_C0;
            const float* dn = WD.slot[dy] < 0 ? K + (ky + dy) * ND_KW + kx : Hm + (WD.slot[dy] * KH + ky + dy) * ND_HW + kx - ND_C0;
            const float4 u = *reinterpret_cast<const float4*>(up), w = *reinterpret_cast<const float4*>(dn);
            mb[0] = fmaxf(mb[0], u.x); mb[1] = fmaxf(mb[1], u.y); mb[2] = fmaxf(mb[2], u.z); mb[3] = fmaxf(mb[3], u.w);
            ma[0] = fmaxf(ma[0], w.x); ma[1] = fmaxf(ma[1], w.y); ma[2] = fmaxf(ma[2], w.z); ma[3] = fmaxf(ma[3], w.w);
        }
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) bits |= (v[8 + j] > 0.f && v[8 + j] > mb[j] && v[8 + j] >= ma[j]) ? 1u << j : 0u;
        if (bits) atomicOr(&newb[nr][kx >> 5], bits << (kx & 31));      // kx % 4 == 0: the nibble stays inside one word
    }
    __syncthreads();
    // ---- phase 4a: dilate the NEW mask along x by every distinct half-width (slot NWID keeps the mask itself)
    for (int item = tid; item < NR * ND_NWORD; item += ND_THREADS) {
        const int nr = item / ND_NWORD, word = item % ND_NWORD;
        const uint32_t c = newb[nr][word];
        const uint32_t l = word > 0 ? newb[nr][word - 1] : 0u, r = word + 1 < ND_NWORD ? newb[nr][word + 1] : 0u;
        uint32_t m = c;
        int d = 1;
        dil[Cfg::NWID][nr][word] = c;
#pragma unroll
        for (int k = 0; k < WD.n; ++k) {
#pragma unroll
            for (; d <= WD.val[k]; ++d) m |= __funnelshift_l(l, c, d) | __funnelshift_r(c, r, d);
            dil[k][nr][word] = m;
        }
    }
    __syncthreads();
    // ---- phase 4b: OR along y -> "a NEW pixel lies in my footprint" for the tile rows
    for (int item = tid; item < ND_TY * ND_NWORD; item += ND_THREADS) {
        const int ty = item / ND_NWORD, word = item % ND_NWORD;
        const int nr = ty + ry;
        uint32_t m = 0u;
        {   // centre row: half-width wx(0) (the centre bit itself is harmless: a NEW pixel becomes KEPT)
            const uint32_t c = newb[nr][word];
            const uint32_t l = word > 0 ? newb[nr][word - 1] : 0u, r = word + 1 < ND_NWORD ? newb[nr][word + 1] : 0u;
#pragma unroll
            for (int d = 1; d <= F::wx(0); ++d) m |= __funnelshift_l(l, c, d) | __funnelshift_r(c, r, d);
        }
#pragma unroll
        for (int dy = 1; dy <= ry; ++dy) {
            const int k = WD.slot[dy] < 0 ? Cfg::NWID : WD.slot[dy];
            m |= dil[k][nr - dy][word] | dil[k][nr + dy][word];
        }
        supb[ty][word] = m;
    }
    __syncthreads();
    // ---- phase 4c: new state of the tile pixels, four per thread
    uint8_t* sout = p.st_out + img;
#pragma unroll
    for (int it = 0; it < ND_TY * (ND_TX / 4) / ND_THREADS; ++it) {
        const int i = it * ND_THREADS + tid;
        const int ty = i / (ND_TX / 4), tx = (i % (ND_TX / 4)) * 4;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= H || x >= W) continue;
        const int kx = tx + ND_HX;
        const float4 k4 = *reinterpret_cast<const float4*>(K + (ty + 2 * ry) * ND_KW + kx);
        const float kk[4] = {k4.x, k4.y, k4.z, k4.w};
        const uint32_t nb = (newb[ty + ry][kx >> 5] >> (kx & 31)) & 0xfu, sb = (supb[ty][kx >> 5] >> (kx & 31)) & 0xfu;
        uint32_t old = 0u;
        if (!FIRST) old = *reinterpret_cast<const uint32_t*>(sin + (int64_t)y * W + x);
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t v;
            if ((nb >> j) & 1u) v = NF_KEPT;
            else if (kk[j] > 0.f) v = ((sb >> j) & 1u) ? (uint32_t)NF_NONE : (uint32_t)NF_ALIVE;
            else v = (old >> (8 * j)) & 0xffu;               // decided pixels keep their state (FIRST: NONE)
            packed |= v << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(sout + (int64_t)y * W + x) = packed;
    }
}

// Which compile-time footprint (if any) equals the one (size, iou) defines: 8 -> NdFoot8, 4 -> NdFoot4, 0 -> none.
// Same membership test as the kernels above.
static int nms_dense_footprint_class(float size, float iou) {
    const int R = (int)ceilf(size) - 1;
    if (R > 7 || R < 1) return 0;
    const float area2 = 2.0f * size * size;
    auto inside = [&](int dy, int dx) {
        const float iw = size - fabsf((float)dx), ih = size - fabsf((float)dy);
        if (iw <= 0.0f || ih <= 0.0f) return false;
        const float inter = iw * ih;
        return (double)(inter / (area2 - inter)) > (double)iou;
    };
    auto matches = [&](auto foot) {
        using F = decltype(foot);
        for (int dy = 0; dy <= R; ++dy)
            for (int dx = 0; dx <= R; ++dx) {
                if (dy == 0 && dx == 0) continue;
                const bool want = dy <= F::ry && dy < 7 && dx <= F::wx(dy);
                if (inside(dy, dx) != want) return false;
            }
        return true;
    };
    if (matches(NdFoot8{})) return 8;
    if (matches(NdFoot4{})) return 4;
    return 0;
}

template <typename F>
static int launch_nms_dense(NmsDenseParams dp, uint8_t* const plane[2], int rounds, int64_t B, int64_t H, int64_t W, cudaStream_t st) {
    XP_CUDA_OK(cudaFuncSetAttribute(nms_dense_round_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NdCfg<F>::SMEM));
    XP_CUDA_OK(cudaFuncSetAttribute(nms_dense_round_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NdCfg<F>::SMEM));
    const dim3 grid((unsigned)ceil_div(W, ND_TX), (unsigned)ceil_div(H, ND_TY), (unsigned)B);
    for (int r = 0; r < rounds; ++r) {
        dp.st_in = plane[r & 1]; dp.st_out = plane[(r + 1) & 1];
        if (r == 0) nms_dense_round_kernel<F, true><<<grid, ND_THREADS, NdCfg<F>::SMEM, st>>>(dp);
        else nms_dense_round_kernel<F, false><<<grid, ND_THREADS, NdCfg<F>::SMEM, st>>>(dp);
        XP_LAUNCH_CHECK("nms_dense_round_kernel");
    }
    return XP_OK;
}

// launched from xp_box_nms (postprocess.cu) when W % 8 == 0 and the maps are 16-byte aligned
int launch_box_nms_fast(const float* prob, float* prob_nms, int64_t B, int64_t H, int64_t W, float size, float min_prob, float iou,
                        int64_t keep_top_k, float kp_threshold, int32_t* keypoints, int32_t* kp_count, int64_t kp_capacity,
                        void* workspace, cudaStream_t st) {
    uint8_t* plane[2] = {(uint8_t*)workspace, (uint8_t*)workspace + B * H * W};
    // dense rounds first (tuning / testing knob: XP_NMS_DENSE_ROUNDS=0 disables them)
    static const char* env = getenv("XP_NMS_DENSE_ROUNDS");
    // a dense round costs ~3 us per 512x640 image and shortens the per-image finishing kernel (one CTA per image) by ~50 us in
    // total from round 3 to 5: worth it for small batches only (single pair: 1.86 -> 1.75 ms end to end)
    int rounds = env && *env ? atoi(env) : (B <= 16 ? 5 : 3);
    const int cls = rounds > 0 && H * W >= 64 * 64 && B <= 65535 ? nms_dense_footprint_class(size, iou) : 0;
    if (cls) {
        NmsDenseParams dp;
        dp.prob = prob; dp.H = (int)H; dp.W = (int)W; dp.min_prob = min_prob; dp.st_in = nullptr; dp.st_out = nullptr;
        const int rc = cls == 8 ? launch_nms_dense<NdFoot8>(dp, plane, rounds, B, H, W, st)
                                : launch_nms_dense<NdFoot4>(dp, plane, rounds, B, H, W, st);
        if (rc) return rc;
    } else {
        rounds = 0;
    }
    NmsFastParams p;
    p.prob = prob; p.out = prob_nms; p.state = plane[rounds & 1]; p.cursor = plane[(rounds + 1) & 1];
    p.kp = keypoints; p.kp_count = kp_count;
    p.H = (int)H; p.W = (int)W; p.size = size; p.min_prob = min_prob; p.iou = iou; p.kp_thr = kp_threshold;
    p.topk = keep_top_k; p.kp_cap = kp_capacity; p.state_ready = rounds > 0;
    box_nms_fast_kernel<<<(unsigned)B, NF_THREADS, 0, st>>>(p);
    XP_LAUNCH_CHECK("box_nms_fast_kernel");
    return XP_OK;
}

}  // namespace xp
