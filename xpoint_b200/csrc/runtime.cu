// runtime.cu -- library-level host helpers: error text, device query, TMA tensor-map encoding.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace xp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
    set_error("CUDA error at %s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return XP_ERR_CUDA;
}

int num_sms() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

int make_tensor_map(CUtensorMap* map, int dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            set_error("cannot resolve cuTensorMapEncodeTiled from the CUDA driver (%s)", cudaGetErrorString(e));
            return XP_ERR_CUDA;
        }
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    const CUtensorMapDataType dt = dtype == XP_F32    ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : dtype == XP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                     : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint64_t gdims[5], gstr[4];
    cuuint32_t gbox[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    const CUresult r = encode(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu, box %u x %u)", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 1), box[0], rank > 1 ? box[1] : 1);
        return XP_ERR_CUDA;
    }
    return XP_OK;
}

}  // namespace xp

extern "C" int xp_abi_version(void) { return XP_ABI_VERSION; }

extern "C" const char* xp_last_error(void) { return xp::g_err; }

extern "C" int xp_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return xp::cuda_fail(e, "cudaGetDevice");
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        xp::set_error("xpoint_b200 is built for sm_100a only; device %d has compute capability %d.%d", dev, major, minor);
        return XP_ERR_UNSUPPORTED;
    }
    return XP_OK;
}
