// fp32 SIMT GEMM for the small / oddly shaped contractions of the path (patch x text-bank similarity with
// T = 45..103 columns, the ln_post projection) where exact fp32 products matter more than tensor-core
// throughput.  The large encoder GEMMs run on tcgen05 (gemm_tc.cu).
//
//   C[b] = act(alpha * A[b] * op(B[b]) + bias) + residual[b]
//   A [M,K] row-major (lda); B is [N,K] row-major (ldb) when b_is_nk (the nn.Linear weight layout,
//   y = x W^T) or [K,N] row-major otherwise; C/residual [M,N] (ldc).
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool B_NK>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
             const float* __restrict__ bias, const float* __restrict__ residual, int M, int N, int K, int64_t lda,
             int64_t ldb, int64_t ldc, int64_t sA, int64_t sB, int64_t sC, int nb2, int64_t sA2, int64_t sB2,
             int64_t sC2, float alpha, int act) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int bz = blockIdx.z / nb2, b2 = blockIdx.z % nb2;  // two-level batch (image, head)
    A += (int64_t)bz * sA + (int64_t)b2 * sA2;
    Bm += (int64_t)bz * sB + (int64_t)b2 * sB2;
    C += (int64_t)bz * sC + (int64_t)b2 * sC2;
    if (residual) residual += (int64_t)bz * sC + (int64_t)b2 * sC2;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4x4 outputs each
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: 64 rows x 16 k -> each thread loads 4 elements (k fastest in memory)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256, r = e >> 4, kk = e & 15;
            const int gm = m0 + r, gk = k0 + kk;
            As[kk][r] = (gm < M && gk < K) ? __ldg(A + (int64_t)gm * lda + gk) : 0.f;
        }
        if (B_NK) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * 256, r = e >> 4, kk = e & 15;
                const int gn = n0 + r, gk = k0 + kk;
                Bs[kk][r] = (gn < N && gk < K) ? __ldg(Bm + (int64_t)gn * ldb + gk) : 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int e = tid + i * 256, kk = e >> 6, c = e & 63;
                const int gn = n0 + c, gk = k0 + kk;
                Bs[kk][c] = (gn < N && gk < K) ? __ldg(Bm + (int64_t)gk * ldb + gn) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = alpha * acc[i][j];
            if (bias) v += bias[gn];
            if (act == 1) v = v * (1.f / (1.f + expf(-1.702f * v)));  // QuickGELU x*sigmoid(1.702x)
            else if (act == 2) v = fmaxf(v, 0.f);                       // ReLU
            if (residual) v += residual[(int64_t)gm * ldc + gn];
            C[(int64_t)gm * ldc + gn] = v;
        }
    }
}

}  // namespace xl

using namespace xl;

namespace xl {
int sgemm2(const float* A, const float* B, float* C, const float* bias, const float* residual, int M, int N, int K,
           int64_t lda, int64_t ldb, int64_t ldc, int batch, int64_t sA, int64_t sB, int64_t sC, int nb2, int64_t sA2,
           int64_t sB2, int64_t sC2, float alpha, int b_is_nk, int act, cudaStream_t st) {
    XL_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0 && nb2 >= 1, "sgemm: bad dimension");
    if (M == 0 || N == 0 || batch == 0) return 0;
    XL_REQUIRE((int64_t)batch * nb2 <= 65535 && ceil_div(M, BM) <= 65535, "sgemm: grid too large");
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM), batch * nb2);
    if (b_is_nk)
        sgemm_kernel<true><<<grid, 256, 0, st>>>(A, B, C, bias, residual, M, N, K, lda, ldb, ldc, sA, sB, sC, nb2, sA2, sB2,
                                                 sC2, alpha, act);
    else
        sgemm_kernel<false><<<grid, 256, 0, st>>>(A, B, C, bias, residual, M, N, K, lda, ldb, ldc, sA, sB, sC, nb2, sA2, sB2,
                                                  sC2, alpha, act);
    return check_launch("sgemm_kernel");
}
}  // namespace xl

extern "C" int excel_sgemm(const float* A, const float* B, float* C, const float* bias, const float* residual, int M,
                           int N, int K, int64_t lda, int64_t ldb, int64_t ldc, int batch, int64_t strideA,
                           int64_t strideB, int64_t strideC, float alpha, int b_is_nk, int act, void* stream) {
    return xl::sgemm2(A, B, C, bias, residual, M, N, K, lda, ldb, ldc, batch, strideA, strideB, strideC, 1, 0, 0, 0, alpha,
                      b_is_nk, act, (cudaStream_t)stream);
}
