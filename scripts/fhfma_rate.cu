// Throughput of FHFMA (fma.rn.f32.f16, sm_100 mixed-precision FMA) against FFMA and HFMA2: 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fhfma_rate scripts/fhfma_rate.cu && /tmp/fhfma_rate
#include <cstdio>
#include <cuda_fp16.h>
template <int MODE> __global__ void k(float* out, unsigned seed, int iters) {
    float acc[8]; unsigned xa = seed + threadIdx.x, xb = seed * 3 + threadIdx.x;
    __half2 h[8];
    for (int i = 0; i < 8; ++i) { acc[i] = i; h[i] = __floats2half2_rn(i, i + 1); }
    __half2 ha = *reinterpret_cast<__half2*>(&xa), hb = *reinterpret_cast<__half2*>(&xb);
    float fa = __uint_as_float(xa | 0x3f000000u), fb = __uint_as_float(xb | 0x3f000000u);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) acc[i] = fmaf(fa, fb, acc[i]);
            if (MODE == 1) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"((unsigned short)(xa >> (16 * (i & 1)))), "h"((unsigned short)xb));
            if (MODE == 2) h[i] = __hfma2(ha, hb, h[i]);
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i] + __low2float(h[i]) + __high2float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 512 * 4);
    const char* names[3] = {"FFMA", "FHFMA", "HFMA2"};
    for (int m = 0; m < 3; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<148 * 8, 512>>>(d, 12345, iters);
            if (m == 1) k<1><<<148 * 8, 512>>>(d, 12345, iters);
            if (m == 2) k<2><<<148 * 8, 512>>>(d, 12345, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = 148.0 * 8 * 512 * 8.0 * iters;
        printf("%s: %.3f ms, %.1f lane-instr/clk/SM at 1.9 GHz\n", names[m], ms, inst / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
