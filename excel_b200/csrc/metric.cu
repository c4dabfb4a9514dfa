// On-device confusion histogram (reference: utils/evaluate.py:9-15 _fast_hist), the quantity the per-rank
// shards reduce with ONE NCCL all-reduce at the end of a run.
#include "common.cuh"
#include "excel_b200.h"

namespace xl {
// hist[nc*t + p] += 1 for every pixel with 0 <= t < nc; block-private shared histogram, then global atomics
__global__ void __launch_bounds__(256)
confusion_hist_kernel(const int64_t* __restrict__ truth, const int64_t* __restrict__ pred, int64_t n, int nc,
                      unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned int sh[];
    const int bins = nc * nc;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = truth[i], p = pred[i];
        if (t >= 0 && t < nc && p >= 0 && p < nc) atomicAdd(&sh[(int)t * nc + (int)p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}
}  // namespace xl

extern "C" int excel_confusion_hist(const int64_t* label_true, const int64_t* label_pred, int64_t n, int num_classes,
                                    int64_t* hist, void* stream) {
    XL_REQUIRE(n >= 0 && num_classes >= 1 && num_classes <= 181, "confusion_hist: bad arguments (num_classes=%d)", num_classes);
    if (n == 0) return 0;
    const size_t smem = (size_t)num_classes * num_classes * sizeof(unsigned int);
    XL_CUDA(cudaFuncSetAttribute(xl::confusion_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (int)(xl::ceil_div64(n, 256 * 16) < 4 * xl::kNumSMs ? xl::ceil_div64(n, 256 * 16) : 4 * xl::kNumSMs);
    xl::confusion_hist_kernel<<<blocks > 0 ? blocks : 1, 256, smem, (cudaStream_t)stream>>>(
        label_true, label_pred, n, num_classes, reinterpret_cast<unsigned long long*>(hist));
    return xl::check_launch("confusion_hist_kernel");
}
