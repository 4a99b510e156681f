// Go / no-go probe for a TMA-staged ss2d_merge_norm: how fast do persistent CTAs stream the four fp32 y planes of the SS2D
// core when every TMA box has a 32-byte inner extent (8 tokens x 8 rows/columns x CCH channels)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tma_box_rate.bin scripts/tma_box_rate.cu && scripts/tma_box_rate.bin
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
    uint32_t ok;
    do { asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory"); } while (!ok);
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(s32(dst)), "l"((uint64_t)m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int TH, int TW, int CCH, int STAGES>
__global__ void __launch_bounds__(160) probe(const __grid_constant__ CUtensorMap mrow, const __grid_constant__ CUtensorMap mcol, int B, int D,
                                             int H, int W, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STAGE = 4 * CCH * TH * TW * 4;
    uint64_t* full = (uint64_t*)(smem + STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int tiles_w = (W + TW - 1) / TW, tiles_h = (H + TH - 1) / TH, ntiles = B * tiles_h * tiles_w, nch = D / CCH;
    const int mine = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (warp == 4) {
        if (lane == 0) {
            int it = 0;
            for (int i = 0; i < mine; ++i) {
                int t = blockIdx.x + i * gridDim.x;
                const int tw = t % tiles_w; t /= tiles_w;
                const int th = t % tiles_h; const int b = t / tiles_h;
                for (int c = 0; c < nch; ++c, ++it) {
                    const int s = it % STAGES;
                    mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
                    mbar_expect(&full[s], STAGE);
                    uint8_t* st = smem + s * STAGE;
                    tma4(st, &mrow, &full[s], tw * TW, th * TH, c * CCH, b * 4 + 0);
                    tma4(st + STAGE / 4, &mrow, &full[s], tw * TW, th * TH, c * CCH, b * 4 + 1);
                    tma4(st + 2 * (STAGE / 4), &mcol, &full[s], th * TH, tw * TW, c * CCH, b * 4 + 2);
                    tma4(st + 3 * (STAGE / 4), &mcol, &full[s], th * TH, tw * TW, c * CCH, b * 4 + 3);
                }
            }
        }
    } else {
        float acc = 0.f;
        int it = 0;
        for (int i = 0; i < mine; ++i)
            for (int c = 0; c < nch; ++c, ++it) {
                const int s = it % STAGES;
                mbar_wait(&full[s], (it / STAGES) & 1);
                acc += ((const float*)(smem + s * STAGE))[threadIdx.x];
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        if (acc == 12345.678f) sink[0] = acc;
    }
}

static int make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t d3, uint32_t b0, uint32_t b1, uint32_t b2) {
    static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
    if (!enc) { void* fn; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q); enc = (PFN_cuTensorMapEncodeTiled_v12000)fn; }
    cuuint64_t dims[4] = {d0, d1, d2, d3}, str[3] = {d0 * 4, d0 * d1 * 4, d0 * d1 * d2 * 4};
    cuuint32_t box[4] = {b0, b1, b2, 1}, es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

template <int TH, int TW, int CCH, int STAGES> static void run(const float* ys, float* sink, int B, int D, int H, int W, int ctas_per_sm) {
    CUtensorMap mrow, mcol;
    if (make_map(&mrow, ys, W, H, D, (uint64_t)B * 4, TW, TH, CCH) || make_map(&mcol, ys, H, W, D, (uint64_t)B * 4, TH, TW, CCH)) { printf("map failed\n"); return; }
    const int smem = STAGES * 4 * CCH * TH * TW * 4 + 2 * STAGES * 8;
    auto k = probe<TH, TW, CCH, STAGES>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) k<<<148 * ctas_per_sm, 160, smem>>>(mrow, mcol, B, D, H, W, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) k<<<148 * ctas_per_sm, 160, smem>>>(mrow, mcol, B, D, H, W, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    const double bytes = (double)B * 4 * D * H * W * 4;
    printf("tile %dx%d, %2d channels per box, %d stages (%3d KB), %d CTAs/SM: %.3f ms  %.0f GB/s   [%s]\n", TH, TW, CCH, STAGES, smem / 1024, ctas_per_sm, ms,
           bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int B = 128, D = 96, H = 128, W = 160;
    float *ys, *sink;
    cudaMalloc(&ys, (size_t)B * 4 * D * H * W * 4); cudaMalloc(&sink, 4);
    cudaMemset(ys, 0, (size_t)B * 4 * D * H * W * 4);
    run<8, 8, 16, 3>(ys, sink, B, D, H, W, 2);
    run<8, 8, 16, 3>(ys, sink, B, D, H, W, 4);
    run<8, 8, 32, 3>(ys, sink, B, D, H, W, 2);
    run<8, 8, 32, 2>(ys, sink, B, D, H, W, 3);
    run<8, 16, 16, 3>(ys, sink, B, D, H, W, 2);
    run<16, 16, 8, 3>(ys, sink, B, D, H, W, 2);
    run<16, 16, 16, 2>(ys, sink, B, D, H, W, 1);
    return 0;
}
