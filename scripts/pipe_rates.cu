// Per-SM issue / pipe rates on B200 for the instructions the N=16 selective scan is bound by:
// MUFU.EX2 (f32, f16x2, bf16x2), FFMA, FFMA2, a Cody-Waite + degree-4 polynomial 2^x on the FMA pipe, and MUFU+FFMA mixes.
// Cycle counts come from clock64() inside the kernel (independent of the SM clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/pipe_rates.bin scripts/pipe_rates.cu && scripts/pipe_rates.bin
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2h2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2b2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
                 "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
                 "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
// 2^x for a pair, x <= 0: round-to-nearest split x = i + f, |f| <= 0.5, degree-4 polynomial, exponent add
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    x.x = fmaxf(x.x, -125.0f); x.y = fmaxf(x.y, -125.0f);
    const float2 t = add2(x, magic);
    const float2 i = add2(t, make_float2(-12582912.0f, -12582912.0f));
    const float2 f = add2(x, make_float2(-i.x, -i.y));
    float2 p = fma2(make_float2(0.0096181291f, 0.0096181291f), f, make_float2(0.0555041087f, 0.0555041087f));
    p = fma2(p, f, make_float2(0.2402265070f, 0.2402265070f));
    p = fma2(p, f, make_float2(0.6931471806f, 0.6931471806f));
    p = fma2(p, f, make_float2(1.0f, 1.0f));
    float2 r;
    r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
    r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
    return r;
}

// MODE: 0 MUFU f32, 1 MUFU f16x2, 2 MUFU bf16x2, 3 FFMA, 4 FFMA2, 5 poly-exp pair, 6 MUFU f32 + 4 FFMA each,
//       7 MUFU f32 + 2 FFMA2 each, 8 MUFU f16x2 + 2 cvt + 4 FFMA, 9: 3 MUFU f32 : 1 poly pair (offload mix)
template <int MODE> __global__ void k(float* out, long long* cyc, float seed, int iters) {
    float a[8];
    float2 b[8];
    unsigned h[8];
    for (int i = 0; i < 8; ++i) { a[i] = -0.01f * (i + 1) - seed * threadIdx.x * 1e-6f; b[i] = make_float2(a[i], a[i] * 0.5f); h[i] = 0xb800b400u + i; }
    const float c0 = 0.999f + seed * 1e-9f, c1 = -1e-3f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2f(a[i]) - 1.5f * 0 + c1;   // (the add folds away; keeps a dependent chain of MUFU + FADD)
            if (MODE == 1) h[i] = ex2h2(h[i]) ^ 0x80008000u;
            if (MODE == 2) h[i] = ex2b2(h[i]) ^ 0x80008000u;
            if (MODE == 3) a[i] = fmaf(a[i], c0, c1);
            if (MODE == 4) b[i] = fma2(b[i], make_float2(c0, c0), make_float2(c1, c1));
            if (MODE == 5) { b[i] = ex2_poly2(b[i]); b[i].x -= 1.001f; b[i].y -= 1.002f; }
            if (MODE == 6) { float e = ex2f(a[i]); a[i] = fmaf(e, c0, c1); a[i] = fmaf(a[i], c0, c1); a[i] = fmaf(a[i], c0, c1); a[i] = fmaf(a[i], c0, c1); }
            if (MODE == 7) { float e = ex2f(b[i].x); b[i] = fma2(b[i], make_float2(e, c0), make_float2(c1, c1)); b[i] = fma2(b[i], make_float2(c0, c0), make_float2(c1, c1)); }
            if (MODE == 8) {
                unsigned e = ex2h2(h[i]);
                const float2 f = __half22float2(*reinterpret_cast<__half2*>(&e));
                a[i] = fmaf(a[i], f.x, c1); a[i] = fmaf(a[i], f.y, c1); a[i] = fmaf(a[i], c0, c1); a[i] = fmaf(a[i], c0, c1);
                h[i] = (h[i] + 1) | 0x80008000u;
            }
            if (MODE == 10) {                                   // 1 FFMA2 + 1 FFMA per unit (does the scalar one ride the idle half?)
                b[i] = fma2(b[i], make_float2(c0, c0), make_float2(c1, c1));
                a[i] = fmaf(a[i], c0, c1);
            }
            if (MODE == 11) {                                   // 1 FFMA2 + 2 FFMA per unit
                b[i] = fma2(b[i], make_float2(c0, c0), make_float2(c1, c1));
                a[i] = fmaf(a[i], c0, c1); a[i] = fmaf(a[i], c1, c0);
            }
            if (MODE == 12) {                                   // 2 FFMA2 + 1 MUFU.RCP per unit (the GELU epilogue's mix is 12 : 2)
                b[i] = fma2(b[i], make_float2(c0, c0), make_float2(c1, c1));
                b[i] = fma2(b[i], make_float2(c1, c1), make_float2(c0, c0));
                float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r + c0;
            }
            if (MODE == 9) {
                if (i % 4 == 3) { b[i] = ex2_poly2(b[i]); b[i].x -= 1.001f; b[i].y -= 1.002f; }
                else a[i] = ex2f(a[i]) + c1;
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + b[i].x + b[i].y + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    const int threads = 512, blocks = 148 * 2;      // 1024 threads / SM = 8 warps per SMSP
    float* d; long long* c;
    cudaMalloc(&d, blocks * threads * 4); cudaMalloc(&c, blocks * 8);
    const char* names[13] = {"MUFU.EX2 f32", "MUFU.EX2 f16x2 (per instr)", "MUFU.EX2 bf16x2 (per instr)", "FFMA", "FFMA2 (per instr)",
                             "poly 2^x pair (per pair)", "1 MUFU + 4 FFMA (per group)", "1 MUFU + 2 FFMA2 (per group)",
                             "1 MUFU.f16x2 + 2 cvt + 4 FFMA (per group)", "3 MUFU : 1 poly pair (per 4 slots)",
                             "1 FFMA2 + 1 FFMA (per unit)", "1 FFMA2 + 2 FFMA (per unit)", "2 FFMA2 + 1 MUFU.RCP + FADD (per unit)"};
    const int iters = 4000;
    for (int m = 0; m < 13; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (m) {
                case 0: k<0><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 1: k<1><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 2: k<2><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 3: k<3><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 4: k<4><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 5: k<5><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 6: k<6><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 7: k<7><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 8: k<8><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 9: k<9><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 10: k<10><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 11: k<11><<<blocks, threads>>>(d, c, 1.f, iters); break;
                case 12: k<12><<<blocks, threads>>>(d, c, 1.f, iters); break;
            }
            cudaDeviceSynchronize();
        }
        long long hc[blocks];
        cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < blocks; ++i) mean += hc[i]; mean /= blocks;
        // per SM: 1024 threads each doing 8 * iters "units"
        const double units = 1024.0 * 8 * iters;
        printf("%-46s %8.0f cycles  -> %6.2f lane-units/clk/SM  (%.3f clk per warp-unit per SMSP)\n", names[m], mean, units / mean,
               mean / (8.0 * 8 * iters));
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
