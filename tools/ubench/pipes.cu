// Dev micro-benchmark: per-SM issue throughput of the instruction classes in the attention epilogues (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes tools/ubench/pipes.cu && build/ubench_pipes
// Each kernel runs ITER x 8 independent instances of one body per thread; 148 CTAs x 512 threads (16 warps / SM, like the
// 16 epilogue warps of attn_pv_kernel).  Prints thread-instructions per clock per SM (at the nominal clock).
#include <cuda_fp16.h>
#include <cstdio>
constexpr int ITER = 2048;
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned pack(float a, float b) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float2 unpack(unsigned h) { __half2 x = *reinterpret_cast<__half2*>(&h); return __half22float2(x); }
struct Mufu { static __device__ __forceinline__ void step(float2& v) { v.x = ex2(v.x); v.y = ex2(v.y); } };
struct F2fp { static __device__ __forceinline__ void step(float2& v) { unsigned h = pack(v.x, v.y); v.x = __uint_as_float(h | 0x3c003c00u); } };
struct Unpack { static __device__ __forceinline__ void step(float2& v) { float2 f = unpack(__float_as_uint(v.x)); v.x = f.x + 1.f; v.y = f.y; } };
struct Ffma2 { static __device__ __forceinline__ void step(float2& v) { v = __ffma2_rn(v, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f)); } };
struct Fadd2 { static __device__ __forceinline__ void step(float2& v) { v = __fadd2_rn(v, make_float2(0.5f, 0.25f)); } };
struct Ffma { static __device__ __forceinline__ void step(float2& v) { v.x = fmaf(v.x, 1.0001f, 0.5f); v.y = fmaf(v.y, 0.9999f, 0.25f); } };
struct Mix {   // the attn_pv chain per pair: 2 MUFU, F2FP, 2 HADD2.F32, FADD2, F2FP
    static __device__ __forceinline__ void step(float2& v) {
        float2 e = make_float2(ex2(v.x), ex2(v.y));
        unsigned h = pack(e.x, e.y);
        float2 f = unpack(h);
        float2 lo = __fadd2_rn(e, make_float2(-f.x, -f.y));
        unsigned l = pack(lo.x, lo.y);
        v.x = __uint_as_float((h ^ l) & 0x3fffffffu) * 1e-9f; v.y = lo.y;
    }
};
struct MixInt {   // same with an integer fp16 -> fp32 unpack (normal range only)
    static __device__ __forceinline__ void step(float2& v) {
        float2 e = make_float2(ex2(v.x), ex2(v.y));
        unsigned h = pack(e.x, e.y);
        float fx = __uint_as_float(((h & 0x7fffu) << 13) + 0x38000000u), fy = __uint_as_float(((h >> 3) & 0x0fffe000u) + 0x38000000u);
        float2 lo = __fadd2_rn(e, make_float2(-fx, -fy));
        unsigned l = pack(lo.x, lo.y);
        v.x = __uint_as_float((h ^ l) & 0x3fffffffu) * 1e-9f; v.y = lo.y;
    }
};
struct MixNoLo {   // 2 MUFU + F2FP only
    static __device__ __forceinline__ void step(float2& v) {
        float2 e = make_float2(ex2(v.x), ex2(v.y));
        unsigned h = pack(e.x, e.y);
        v.x = __uint_as_float(h & 0x3fffffffu) * 1e-9f; v.y = e.y * 0.5f;
    }
};
template <typename T>
__global__ void __launch_bounds__(512) kern(float* out, float seed) {
    float2 v[8];
    for (int i = 0; i < 8; ++i) v[i] = make_float2(seed + threadIdx.x * 1e-3f + i, seed + i * 0.5f);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) T::step(v[i]);
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    if (s == 123.456f) out[threadIdx.x] = s;
}
template <typename T>
void run(const char* name, double per_body) {
    float* out; cudaMalloc(&out, 4096);
    kern<T><<<148, 512>>>(out, 0.3f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) kern<T><<<148, 512>>>(out, 0.3f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double bodies = (double)ITER * 8 * 512;                     // per SM
    printf("%-10s %8.3f ms  clk/body/SM %7.3f  thread-instr/clk/SM @%d MHz nominal: %6.1f  (x%.0f instr per body)\n", name, ms,
           ms * 1e-3 * khz * 1e3 / bodies, khz / 1000, bodies * per_body / (ms * 1e-3) / (khz * 1e3), per_body);
    cudaFree(out);
}
int main() {
    run<Mufu>("mufu.ex2", 2); run<F2fp>("f2fp", 1); run<Unpack>("unpack", 2); run<Ffma2>("ffma2", 1); run<Fadd2>("fadd2", 1);
    run<Ffma>("ffma", 2); run<Mix>("mix", 1); run<MixInt>("mix_int", 1); run<MixNoLo>("mix_nolo", 1);
    return 0;
}
