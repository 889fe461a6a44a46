// Host runtime helpers: last-error string, TMA tensor-map encoding through the driver
// entry point (no link-time dependency on libcuda), device capability probe.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "ops.h"

namespace pp {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

bool pdl_enabled() {
    static const bool on = [] {
        // opt-in: measured on the C2 step it does not pay (20.5 ms with programmatic edges vs 20.2 ms
        // without — the persistent kernels own a whole SM each, so a successor CTA can only become
        // resident when a predecessor CTA has already exited)
        const char* e = getenv("PP_B200_PDL");
        return e && e[0] == '1';
    }();
    return on;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
    return make_tmap_bf16_sw(out, base, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0);
}

int make_tmap_bf16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
        return PP_ERR_CUDA;
    }
    PP_REQUIRE(rank >= 1 && rank <= 5, "tensor map rank %d out of range", rank);
    PP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16-byte aligned", base);
    cuuint64_t gd[5];
    cuuint64_t gs[4];
    cuuint32_t bx[5];
    cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        PP_REQUIRE(dims[i] >= 1, "tensor map dim %d is zero", i);
        PP_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map box[%d]=%u out of range", i, box[i]);
    }
    for (int i = 0; i + 1 < rank; ++i) {
        gs[i] = strides_bytes[i];
        PP_REQUIRE((strides_bytes[i] & 15) == 0, "tensor map stride[%d]=%llu not a multiple of 16 bytes", i,
                   (unsigned long long)strides_bytes[i]);
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd,
                    gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                    : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu %llu)",
                       (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                       (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0),
                       (unsigned long long)(rank > 4 ? gd[4] : 0));
        return PP_ERR_CUDA;
    }
    return PP_OK;
}

}  // namespace pp

extern "C" {

const char* pp_last_error(void) { return pp::last_error(); }
int pp_abi_version(void) { return 5; }
int pp_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0, minor = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return 0;
    return (major == 10 && minor == 0) ? 1 : 0;
}

}  // extern "C"
