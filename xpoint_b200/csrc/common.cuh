// common.cuh -- shared device/host helpers for the sm_100a kernels of xpoint_b200.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/xpoint_b200.h"

namespace xp {

// ------------------------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);

#define XP_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            ::xp::set_error(__VA_ARGS__);    \
            return XP_ERR_INVALID_ARG;       \
        }                                    \
    } while (0)

#define XP_CUDA_OK(expr)                                             \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) return ::xp::cuda_fail(_e, #expr);    \
    } while (0)

#define XP_LAUNCH_CHECK(name)                                        \
    do {                                                             \
        cudaError_t _e = cudaGetLastError();                         \
        if (_e != cudaSuccess) return ::xp::cuda_fail(_e, name);     \
    } while (0)

int num_sms();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int dtype_size(int dt) { return dt == XP_F32 ? 4 : 2; }

// ------------------------------------------------------------------------------------ dtype conversion
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Load V consecutive elements (V * sizeof(T) == 16 or 32 bytes for the vector paths) and widen to fp32.
// p must be 16-byte aligned.
template <typename T, int V> struct VecIO;

template <> struct VecIO<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&r)[4]) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    }
    static __device__ __forceinline__ void load_stream(const float* p, float (&r)[4]) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&r)[4]) {
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
    }
};

template <> struct VecIO<__half, 8> {
    static __device__ __forceinline__ void widen(const uint4& v, float (&r)[8]) {
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); r[2 * i] = f.x; r[2 * i + 1] = f.y; }
    }
    static __device__ __forceinline__ void load(const __half* p, float (&r)[8]) {
        widen(__ldg(reinterpret_cast<const uint4*>(p)), r);
    }
    static __device__ __forceinline__ void load_stream(const __half* p, float (&r)[8]) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        widen(v, r);
    }
    static __device__ __forceinline__ void store(__half* p, const float (&r)[8]) {
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(r[2 * i], r[2 * i + 1]);
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
};

template <> struct VecIO<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void widen(const uint4& v, float (&r)[8]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
            r[2 * i] = __uint_as_float(w[i] << 16);
            r[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&r)[8]) {
        widen(__ldg(reinterpret_cast<const uint4*>(p)), r);
    }
    static __device__ __forceinline__ void load_stream(const __nv_bfloat16* p, float (&r)[8]) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
        widen(v, r);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&r)[8]) {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r[2 * i], r[2 * i + 1]);
        asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                     :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
};

// ------------------------------------------------------------------------------------ math
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed FP32 pairs (Blackwell FFMA2 / FMUL2: two IEEE fp32 operations per issue slot, same rounding as the scalar ops)
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)),
        "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

// softplus(beta=1, threshold=20) as torch / the reference kernel define it (csms6s.py:49-50,
// selective_scan_fwd_kernel_oflex.cuh:124-127): x > 20 ? x : log1p(exp(x)).  Branch-free, 2 MUFU + ~10 FP32 ops
// (libdevice log1pf costs ~35 instructions and two branches):  e = exp(x);
//   e <  1/16: log1p(e) = e - e^2/2 + e^3/3 - e^4/4 + e^5/5       (truncation e^5/6 < 1.6e-7 relative)
//   e >= 1/16: log1p(e) = ln2 * lg2(1 + e)                          (lg2.approx abs error 2^-22 on >= 0.087)
// The series branch keeps the small-delta regime (dt in [1e-3, 0.1], VMamba.py:181-186) at fp32 accuracy.
__device__ __forceinline__ float softplus_f(float x) {
    const float e = ex2_approx(x * kLog2e);
    const float small = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
    const float big = lg2_approx(1.0f + e) * kLn2;
    const float r = e < 0.0625f ? small : big;
    return x > 20.0f ? x : r;
}

// The same function with ONE MUFU, for kernels that are bound by the XU pipe (the N >= 4 scan spends 16 ex2 per token and row on
// its decay factors; softplus's ex2 + lg2 were another 11 % of that pipe):  softplus(x) = max(x, 0) + log1p(t), t = e^-|x| in (0, 1],
// log1p(t) = t * Q(t) with Q the degree-8 minimax polynomial of log1p(t) / t on [0, 1] (max relative error 2.0e-7 evaluated in
// fp32, i.e. the small-delta regime keeps fp32 accuracy; for x > 20 the correction is below half an ulp of x, so the result is x
// exactly as torch's threshold branch returns it).
__device__ __forceinline__ float softplus1_f(float x) {
    const float t = ex2_approx(-fabsf(x) * kLog2e);
    float q = fmaf(t, 0.005253457929939032f, -0.02958850748836994f);
    q = fmaf(q, t, 0.07836166769266129f);
    q = fmaf(q, t, -0.13674770295619965f);
    q = fmaf(q, t, 0.19111430644989014f);
    q = fmaf(q, t, -0.24844369292259216f);
    q = fmaf(q, t, 0.33319270610809326f);
    q = fmaf(q, t, -0.49999502301216125f);
    q = fmaf(q, t, 1.0f);
    return fmaf(q, t, fmaxf(x, 0.0f));
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// exact (erf) GELU: erfc(x) = (1 + a1 x + .. + a6 x^6)^-16, x >= 0 (Abramowitz-Stegun 7.1.28, |err| <= 3e-7), evaluated in |v|
// (the 1/sqrt(2) of x = |v| / sqrt(2) folded into the coefficients) and GELU(v) = relu(v) - 0.5 |v| erfc(|v| / sqrt(2)):
// one MUFU.RCP, one FMNMX and 12 FP32 ops instead of libdevice erff's ~25 instructions and branches (gelu_erf2 in tcgen05.cuh is
// the packed form of the same arithmetic).
__device__ __forceinline__ float gelu_erf_f(float v) {
    const float av = fabsf(v);
    float p = fmaf(av, 5.3829750000e-06f, 4.8890635643e-05f);
    p = fmaf(p, av, 3.8003575000e-05f);
    p = fmaf(p, av, 3.2776263241e-03f);
    p = fmaf(p, av, 2.1141006150e-02f);
    p = fmaf(p, av, 4.9867346967e-02f);
    p = fmaf(p, av, 1.0f);
    p *= p; p *= p; p *= p; p *= p;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    return fmaf(av * r, -0.5f, fmaxf(v, 0.0f));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------ mbarrier / TMA (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {   // non-blocking poll
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait sleeps in hardware; the bound turns a lost TMA completion into a trap instead of a hung GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
// The same with a suspend-time hint: the thread stays parked until the phase completes (or ~20 us pass) instead of re-polling
// every few dozen cycles.  For waits that are EXPECTED to block in a kernel whose other warps are bound by instruction issue
// (mlp_res_ln_tc: a quarter of all issued instructions were epilogue warps polling h_empty).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (++spins > (1u << 22)) __trap();
    } while (!ok);
}

// 32-bit shared-window variants (keep hot loops on LDS/STS/SYNCS with no generic-address arithmetic)
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t saddr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();   // a lost completion becomes a trap, not a hung GPU
    } while (!ok);
}

// 4-D tiled TMA load: global (tensor map, coords c0 innermost) -> shared, completion on mbarrier.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
           "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// 4-D tiled TMA store: shared -> global.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// 1-D bulk async copy global -> shared (UBLKCP), completion on mbarrier.  dst/src/bytes: multiples of 16.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ------------------------------------------------------------------------------------ host: tensor maps
// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda).
// dims/strides innermost-first; strides_bytes has rank-1 entries (dim 1..rank-1).  swizzle: 0 none, 1 128B, 2 64B, 3 32B.
// Returns XP_OK or error.
int make_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

}  // namespace xp
