// Shared helpers for the excel_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace xl {

// ---- error reporting across the C ABI (include/excel_b200.h: excel_last_error) -------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> 0 / error code (message recorded)

#define XL_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            xl::set_error(__VA_ARGS__);      \
            return 1;                        \
        }                                    \
    } while (0)

#define XL_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            xl::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                        \
        }                                                                                    \
    } while (0)

// fp32 SIMT GEMM with a two-level batch (gemm_simt.cu)
int sgemm2(const float* A, const float* B, float* C, const float* bias, const float* residual, int M, int N, int K,
           int64_t lda, int64_t ldb, int64_t ldc, int batch, int64_t sA, int64_t sB, int64_t sC, int nb2, int64_t sA2,
           int64_t sB2, int64_t sC2, float alpha, int b_is_nk, int act, cudaStream_t st);

// Function attributes (cudaFuncSetAttribute) are per DEVICE: `once` holds one bit per device ordinal; returns true the first
// time it is called with a given device current (single host thread per process).
inline bool first_use_on_device(unsigned long long& once) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) return true;   // unknown: (re)apply the attribute
    if (once & (1ull << dev)) return false;
    once |= 1ull << dev;
    return true;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide reductions through shared memory (blockDim.x multiple of 32, <= 1024)
template <typename Op>
__device__ __forceinline__ float block_reduce(float v, float* smem32, Op op, float identity) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();  // protect smem32 from a previous use
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    v = (threadIdx.x < nw) ? smem32[threadIdx.x] : identity;
    if (wid == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) smem32[0] = v;
    }
    __syncthreads();
    return smem32[0];
}
struct OpSum { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// The encoder is a chain of ~190 short dependent kernels.  Launched with the programmatic-stream-serialization attribute, a
// kernel's CTAs may become resident while the previous kernel's last CTAs are still running: launch latency, barrier init,
// tensor-memory allocation and descriptor prefetch then overlap the predecessor's tail.  Every kernel launched this way calls
// pdl_wait() before its first global-memory access (it returns once the preceding grid has completed and its writes are
// visible) and pdl_trigger() at its top (lets ITS successor be scheduled early).  Both are no-ops in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace xl
