// Dev micro-benchmark: SM -> L2 store throughput for the access patterns of the attention-map flush (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_stores tools/ubench/stores.cu && build/ubench_stores
// 148 CTAs x 512 threads; each CTA writes a [128 rows x 128 floats] tile (64 KB) per iteration into its own region of a
// [B*N, Npad] fp32 matrix (row pitch 4112 B like the attention maps), ITER times.  Patterns: a warp-wide 16 B store covers
//   seg64 : 8 rows x 64 B (the current flush)    seg128: 4 rows x 128 B    seg512: 1 row x 512 B    red64: seg64 with red.add.v4
#include <cstdio>
#include <cstdint>
constexpr int ITER = 64, NPAD = 1028;
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* base = out + (size_t)blockIdx.x * 128 * NPAD;
    const float4 v = make_float4(lane, warp, 1.f, 2.f);
    for (int it = 0; it < ITER; ++it) {
        const int col0 = (it % 8) * 128;                       // walk 8 key blocks like the kernel
        if (MODE == 0 || MODE == 3) {                         // warp = 32 rows x 16 cols block; 2 chunks; 4 instr of 8 rows x 64 B
            const int lg = warp & 3, cq = warp >> 2;          // cq 0..3: 32-col slice
            for (int c = 0; c < 2; ++c)
                for (int i = 0; i < 4; ++i) {
                    float* p = base + (size_t)(lg * 32 + (lane >> 2) + 8 * i) * NPAD + col0 + cq * 32 + c * 16 + (lane & 3) * 4;
                    if (MODE == 0) *reinterpret_cast<float4*>(p) = v;
                    else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
        } else if (MODE == 1) {                               // 4 rows x 128 B per instr: warp covers 32 rows x 32 cols in 8 instr
            const int lg = warp & 3, cq = warp >> 2;
            for (int i = 0; i < 8; ++i) {
                float* p = base + (size_t)(lg * 32 + (lane >> 3) + 4 * i) * NPAD + col0 + cq * 32 + (lane & 7) * 4;
                *reinterpret_cast<float4*>(p) = v;
            }
        } else {                                              // 1 row x 512 B per instr: warp covers 8 rows x 128 cols
            for (int i = 0; i < 8; ++i) {
                float* p = base + (size_t)(warp * 8 + i) * NPAD + col0 + lane * 4;
                *reinterpret_cast<float4*>(p) = v;
            }
        }
    }
}
template <int MODE>
void run(const char* name, float* out) {
    k<MODE><<<148, 512>>>(out);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) k<MODE><<<148, 512>>>(out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double bytes = (double)ITER * 65536;               // per SM
    printf("%-8s %7.3f ms   %6.1f B/clk/SM (nominal %d MHz)   %6.0f clk per 64 KB tile   chip %5.2f TB/s\n", name, ms,
           bytes / (ms * 1e-3 * khz * 1e3), khz / 1000, ms * 1e-3 * khz * 1e3 / ITER, bytes * 148 / (ms * 1e-3) / 1e12);
}
int main() {
    float* out; cudaMalloc(&out, (size_t)148 * 128 * NPAD * 4 + 4096);
    cudaMemset(out, 0, (size_t)148 * 128 * NPAD * 4);
    run<0>("seg64", out); run<1>("seg128", out); run<2>("seg512", out); run<3>("red64", out);
    return 0;
}
