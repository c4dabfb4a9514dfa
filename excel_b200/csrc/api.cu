// Library-level entry points: error channel, version, device probe.
#include <stdarg.h>

#include <atomic>
#include <string.h>

#include "common.cuh"
#include "excel_b200.h"

namespace xl {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};  // kernels enqueued by this library

int check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 3;
}
}  // namespace xl

extern "C" const char* excel_last_error(void) { return xl::g_err; }
extern "C" int excel_version(void) { return 1; }
extern "C" int64_t excel_launch_count(void) { return (int64_t)xl::g_launches.load(); }
extern "C" int excel_device_arch(int device) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) {
        xl::set_error("cudaGetDeviceProperties(%d) failed", device);
        return -1;
    }
    return 10 * p.major + p.minor;
}

// ---- TMA descriptor encoding ------------------------------------------------------------------------
#include "ptx.cuh"
namespace xl {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int encode_tensor_map(CUtensorMap* tm, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                      CUtensorMapL2promotion promo) {
    EncodeTiledFn fn = get_encode_fn();
    XL_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = fn(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    XL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,..] box=[%u,%u,..]", (int)r,
               rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return 0;
}
}  // namespace xl
