// LVC bias of the surgery attention (reference: clip/clip_surgery_model.py:127-141, Attention.forward with ex_feats) and
// the small row-wise helpers of utils/attrutils.py.
//
//   ex_attn[b] = softmax_j( mask_{<0 -> -inf}( (cos_sim(f_b)[i,j] - mean_over_batch(cos_sim)) * gamma ) ),
//   f_b = decoder features [C, n_p] L2-normalised over C per position (F.normalize, eps 1e-12).
// The reference adds ex_attn to EVERY head's patch block of the new-path attention and then sums the heads (:139-146), so
// the encoder adds H * ex_attn to the head-summed map (vit.cu).  The similarity is an exact fp32 SIMT GEMM (gemm_simt.cu);
// the batch-wide mean is accumulated in double in a fixed order (the threshold at 0 right next to it is discontinuous).
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

// qt[b, m, c] = f[b, c, m] / max(||f[b, :, m]||, 1e-12): normalise over channels and transpose to [n_p, C].
// A block of 32 x 8 threads owns 32 positions and the channel slice blockIdx.z (of gridDim.z): channel norms over ALL channels
// with coalesced reads along m (fixed summation order: channel c in slice c % 8, slices combined 0..7 -- every z-slice computes
// the same norm), then its 32 x 32 tiles transposed through shared memory so that the writes are coalesced along c.  The channel
// split exists for small batches: one image of 1024 positions is 32 blocks without it.
__global__ void __launch_bounds__(256)
lvc_normalize_t_kernel(const float* __restrict__ f, int C, int np, float* __restrict__ qt) {
    __shared__ float part[8][33], tile[32][33], inv[32];
    const int tx = threadIdx.x, ty = threadIdx.y, m0 = blockIdx.x * 32, b = blockIdx.y;
    const float* src = f + (int64_t)b * C * np;
    float ss = 0.f;
    if (m0 + tx < np) {
#pragma unroll 8
        for (int c = ty; c < C; c += 8) {
            const float v = src[(int64_t)c * np + m0 + tx];
            ss = fmaf(v, v, ss);
        }
    }
    part[ty][tx] = ss;
    __syncthreads();
    if (ty == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][tx];
        inv[tx] = 1.f / fmaxf(sqrtf(t), 1e-12f);
    }
    __syncthreads();
    const int cper = ((C + 31) / 32 + gridDim.z - 1) / gridDim.z * 32;   // channels per z-slice (multiple of 32)
    const int cbeg = blockIdx.z * cper, cend = min(C, cbeg + cper);
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + ty + 8 * k;
            tile[ty + 8 * k][tx] = (c < C && m0 + tx < np) ? src[(int64_t)c * np + m0 + tx] * inv[tx] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int m = m0 + ty + 8 * k;
            if (m < np && c0 + tx < C) qt[((int64_t)b * np + m) * C + c0 + tx] = tile[tx][ty + 8 * k];
        }
        __syncthreads();
    }
}

// rowsum[r] = sum_j x[r, j] in double, one warp per row
__global__ void __launch_bounds__(256)
row_sum_f64_kernel(const float* __restrict__ x, int64_t rows, int cols, double* __restrict__ rowsum) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    double s = 0.0;
    for (int j = lane; j < cols; j += 32) s += (double)x[row * cols + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rowsum[row] = s;
}

// mean[0] = sum(rowsum) / count, one block, fixed order
__global__ void __launch_bounds__(1024)
total_mean_kernel(const double* __restrict__ rowsum, int64_t rows, double count, float* __restrict__ mean) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 1024) s += rowsum[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) mean[0] = (float)(s / count);
    }
}

// out[r, :] = softmax_j(t(x[r, :])), t(v) = v (plain) or ((v - mean*beta) * gamma, negatives -> -inf); one warp per row.
// A row whose entries are all masked gives NaN like torch.softmax of an all -inf row.
template <bool MASKED>
__global__ void __launch_bounds__(256)
row_softmax_kernel(const float* __restrict__ x, int64_t rows, int cols, const float* __restrict__ mean, float beta, float gamma,
                   float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * cols;
    float* yr = out + row * cols;
    const float sub = MASKED ? mean[0] * beta : 0.f;
    auto tr = [&](float v) {
        if (!MASKED) return v;
        const float t = (v - sub) * gamma;
        return t < 0.f ? -INFINITY : t;
    };
    float mx = -INFINITY;
    for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, tr(xr[j]));
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < cols; j += 32) s += expf(tr(xr[j]) - mx);
    s = warp_sum(s);
    for (int j = lane; j < cols; j += 32) yr[j] = expf(tr(xr[j]) - mx) / s;
}

__global__ void __launch_bounds__(256)
row_l2_normalize_kernel(const float* __restrict__ x, int64_t rows, int cols, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float ss = 0.f;
    for (int j = lane; j < cols; j += 32) ss = fmaf(x[row * cols + j], x[row * cols + j], ss);
    const float n = sqrtf(warp_sum(ss));
    for (int j = lane; j < cols; j += 32) out[row * cols + j] = x[row * cols + j] / n;
}

// x <- sigmoid((x - mean * beta) * gamma), elementwise (model/model_excel.py:75-76)
__global__ void centred_sigmoid_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ mean, float beta, float gamma) {
    const float sub = mean[0] * beta;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = 1.f / (1.f + expf(-(x[i] - sub) * gamma));
}

// normalise + similarity + batch mean, shared by the LVC bias and attn_pred: sim [B,np,np] and mean_ws[0]
static int cosine_similarity_and_mean(const float* feats, int B, int C, int np, float* qt_ws, double* rowsum_ws, float* mean_ws,
                                      float* sim, cudaStream_t st) {
    // channel slices so that a small batch still fills the chip (>= ~2 blocks per SM), at most one slice per 32 channels
    int zs = 1;
    while (zs < 8 && ceil_div(np, 32) * B * zs < 2 * kNumSMs && zs * 2 * 32 <= C) zs *= 2;
    lvc_normalize_t_kernel<<<dim3(ceil_div(np, 32), B, zs), dim3(32, 8), 0, st>>>(feats, C, np, qt_ws);
    if (int e = check_launch("lvc_normalize_t_kernel")) return e;
    // sim[b] = qt[b] qt[b]^T  (exact fp32)
    if (int e = sgemm2(qt_ws, qt_ws, sim, nullptr, nullptr, np, np, C, C, C, np, B, (int64_t)np * C, (int64_t)np * C,
                       (int64_t)np * np, 1, 0, 0, 0, 1.f, 1, 0, st)) return e;
    const int64_t rows = (int64_t)B * np;
    row_sum_f64_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, st>>>(sim, rows, np, rowsum_ws);
    if (int e = check_launch("row_sum_f64_kernel")) return e;
    total_mean_kernel<<<1, 1024, 0, st>>>(rowsum_ws, rows, (double)rows * np, mean_ws);
    return check_launch("total_mean_kernel");
}

}  // namespace xl

using namespace xl;

extern "C" int excel_attn_pred(const float* feats, int B, int C, int np, float beta, float gamma, float* qt_ws,
                               double* rowsum_ws, float* mean_ws, float* attn_pred, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && C >= 1 && np >= 1 && B <= 65535, "attn_pred: bad shape B=%d C=%d np=%d", B, C, np);
    XL_REQUIRE(feats && qt_ws && rowsum_ws && mean_ws && attn_pred, "attn_pred: missing buffers");
    if (B == 0) return 0;
    if (int e = cosine_similarity_and_mean(feats, B, C, np, qt_ws, rowsum_ws, mean_ws, attn_pred, st)) return e;
    const int64_t n = (int64_t)B * np * np;
    centred_sigmoid_kernel<<<(unsigned)(ceil_div64(n, 256) < 148 * 16 ? ceil_div64(n, 256) : 148 * 16), 256, 0, st>>>(attn_pred, n, mean_ws,
                                                                                                              beta, gamma);
    return check_launch("centred_sigmoid_kernel");
}

extern "C" int excel_row_softmax(const float* x, int rows, int cols, float* out, void* stream) {
    XL_REQUIRE(rows >= 0 && cols >= 1, "row_softmax: bad shape");
    if (rows == 0) return 0;
    row_softmax_kernel<false><<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, nullptr, 0.f, 1.f, out);
    return check_launch("row_softmax_kernel");
}

extern "C" int excel_row_l2_normalize(const float* x, int rows, int cols, float* out, void* stream) {
    XL_REQUIRE(rows >= 0 && cols >= 1, "row_l2_normalize: bad shape");
    if (rows == 0) return 0;
    row_l2_normalize_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, out);
    return check_launch("row_l2_normalize_kernel");
}

extern "C" int excel_lvc_attention(const float* ex_feats, int B, int C, int np, float beta, float gamma, float* qt_ws,
                                   double* rowsum_ws, float* mean_ws, float* ex_attn, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && C >= 1 && np >= 1 && B <= 65535, "lvc_attention: bad shape B=%d C=%d np=%d", B, C, np);
    XL_REQUIRE(ex_feats && qt_ws && rowsum_ws && mean_ws && ex_attn, "lvc_attention: missing buffers");
    if (B == 0) return 0;
    // sim[b] = cosine similarity, written into ex_attn and transformed in place row by row
    if (int e = cosine_similarity_and_mean(ex_feats, B, C, np, qt_ws, rowsum_ws, mean_ws, ex_attn, st)) return e;
    const int64_t rows = (int64_t)B * np;
    row_softmax_kernel<true><<<(unsigned)ceil_div64(rows, 8), 256, 0, st>>>(ex_attn, rows, np, mean_ws, beta, gamma, ex_attn);
    return check_launch("row_softmax_kernel<masked>");
}
