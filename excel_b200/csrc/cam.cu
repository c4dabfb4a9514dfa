// Patch x text-bank CAM (reference: clip/clip.py:288-310 clip_feature_surgery, clip/clip.py:353 token norm).
//
// The reference materialises feats[B,N,T,E] = F (x) T (1.5 GB at 512^2 x16) only to reduce it again; by
// linearity  sim[n,t] = w_t S[n,t] - mean_t'(w_t' S[n,t'])  with S = F T^T, so the path is one GEMM
// (excel_sgemm, exact fp32) plus two small epilogue kernels:
//   row kernel : w = softmax_t(2 S[b,0,:]) / mean(..)  (clip.py:295-297); sim row (clip.py:301-306)
//   col kernel : per (b,t) min / max over ALL N tokens incl. CLS, (sim-min)/(max-min), no epsilon (clip.py:308)
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

// ---- image_features / image_features.norm(dim=1)  (norm over the TOKEN axis, clip/clip.py:353) -------
__global__ void __launch_bounds__(1024)
token_sumsq_kernel(const float* __restrict__ tok, int N, int E, float* __restrict__ norm) {
    __shared__ float red[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y;
    float s = 0.f;
    if (e < E)
        for (int n = threadIdx.y; n < N; n += 32) {
            const float v = tok[((int64_t)b * N + n) * E + e];
            s = fmaf(v, v, s);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < E) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) t += red[r][threadIdx.x];
        norm[(int64_t)b * E + e] = sqrtf(t);
    }
}

__global__ void token_div_kernel(const float* __restrict__ tok, const float* __restrict__ norm, float* __restrict__ out,
                                 int N, int E, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = (int)(i % E);
    const int64_t b = i / ((int64_t)N * E);
    out[i] = tok[i] / norm[b * E + e];
}

// ---- sim[b,n,:] from S[b,n,:]; one warp per token row ------------------------------------------------
__global__ void __launch_bounds__(256)
cam_row_kernel(const float* __restrict__ S, float* __restrict__ sim, int N, int T) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y;
    if (n >= N) return;
    const float* cls = S + (int64_t)b * N * T;  // row 0 = CLS token
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) mx = fmaxf(mx, 2.f * cls[t]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int t = lane; t < T; t += 32) se += expf(2.f * cls[t] - mx);
    se = warp_sum(se);
    // w_t = p_t / mean(p), p = softmax;  mean(p) = (sum p)/T
    float sp = 0.f;
    for (int t = lane; t < T; t += 32) sp += expf(2.f * cls[t] - mx) / se;
    sp = warp_sum(sp);
    const float pmean = sp / (float)T;
    const float* row = S + ((int64_t)b * N + n) * T;
    float acc = 0.f;
    for (int t = lane; t < T; t += 32) acc += row[t] * ((expf(2.f * cls[t] - mx) / se) / pmean);
    acc = warp_sum(acc);
    const float mean = acc / (float)T;
    float* o = sim + ((int64_t)b * N + n) * T;
    for (int t = lane; t < T; t += 32) o[t] = row[t] * ((expf(2.f * cls[t] - mx) / se) / pmean) - mean;
}

// ---- per (b,t): min-max over the N tokens, in place --------------------------------------------------
__global__ void __launch_bounds__(256)
cam_col_kernel(float* __restrict__ sim, int N, int T) {
    __shared__ float red[32];
    const int t = blockIdx.x, b = blockIdx.y;
    float* col = sim + (int64_t)b * N * T + t;
    float lo = INFINITY, hi = -INFINITY;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float v = col[(int64_t)n * T];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    lo = block_reduce(lo, red, OpMin(), INFINITY);
    hi = block_reduce(hi, red, OpMax(), -INFINITY);
    const float rng = hi - lo;
    for (int n = threadIdx.x; n < N; n += blockDim.x) col[(int64_t)n * T] = (col[(int64_t)n * T] - lo) / rng;
}

}  // namespace xl

using namespace xl;

extern "C" int excel_token_normalize(const float* tok, int B, int N, int E, float* norm_ws, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && N >= 1 && E >= 1 && B <= 65535, "token_normalize: bad shape");
    if (B == 0) return 0;
    dim3 grid(ceil_div(E, 32), B), block(32, 32);
    token_sumsq_kernel<<<grid, block, 0, st>>>(tok, N, E, norm_ws);
    if (int e = check_launch("token_sumsq_kernel")) return e;
    const int64_t total = (int64_t)B * N * E;
    token_div_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(tok, norm_ws, out, N, E, total);
    return check_launch("token_div_kernel");
}

// utils/camutils.py:19-26 (cure_attr_map_flip): out[b,p,k] = (m - min_p m) / (max_p (m - min_p m) + 1e-5),
// m[b,p,k] = max(x[b,p,k], x[B+b, flip_x(p), k]); x [2B, gh*gw, K].  One block per (k, b).
__global__ void __launch_bounds__(256)
flip_merge_kernel(const float* __restrict__ x, int B, int gh, int gw, int K, float* __restrict__ out) {
    __shared__ float red[32];
    const int k = blockIdx.x, b = blockIdx.y, np = gh * gw;
    const float* xa = x + (int64_t)b * np * K + k;
    const float* xb = x + (int64_t)(B + b) * np * K + k;
    float mn = INFINITY;
    for (int p = threadIdx.x; p < np; p += 256) {
        const int py = p / gw, px = p - py * gw;
        mn = fminf(mn, fmaxf(xa[(int64_t)p * K], xb[(int64_t)(py * gw + gw - 1 - px) * K]));
    }
    mn = block_reduce(mn, red, OpMin(), INFINITY);
    float mx = -INFINITY;
    for (int p = threadIdx.x; p < np; p += 256) {
        const int py = p / gw, px = p - py * gw;
        mx = fmaxf(mx, fmaxf(xa[(int64_t)p * K], xb[(int64_t)(py * gw + gw - 1 - px) * K]) - mn);
    }
    mx = block_reduce(mx, red, OpMax(), -INFINITY);
    const float den = mx + 1e-5f;
    for (int p = threadIdx.x; p < np; p += 256) {
        const int py = p / gw, px = p - py * gw;
        out[((int64_t)b * np + p) * K + k] = (fmaxf(xa[(int64_t)p * K], xb[(int64_t)(py * gw + gw - 1 - px) * K]) - mn) / den;
    }
}

extern "C" int excel_flip_merge(const float* attr_2b, int B, int gh, int gw, int K, float* out, void* stream) {
    XL_REQUIRE(B >= 0 && gh >= 1 && gw >= 1 && K >= 1 && B <= 65535, "flip_merge: bad shape");
    if (B == 0) return 0;
    flip_merge_kernel<<<dim3(K, B), 256, 0, (cudaStream_t)stream>>>(attr_2b, B, gh, gw, K, out);
    return check_launch("flip_merge_kernel");
}

extern "C" int excel_cam_surgery(const float* feats, const float* text, int B, int N, int E, int T, float* S_ws,
                                 float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && N >= 1 && E >= 1 && T >= 1 && B <= 65535 && T <= 65535, "cam_surgery: bad shape");
    if (B == 0) return 0;
    // S[B*N, T] = feats[B*N, E] * text[T, E]^T
    if (int e = excel_sgemm(feats, text, S_ws, nullptr, nullptr, B * N, T, E, E, E, T, 1, 0, 0, 0, 1.f, 1, 0, stream)) return e;
    dim3 g1(ceil_div(N, 8), B);
    cam_row_kernel<<<g1, 256, 0, st>>>(S_ws, out, N, T);
    if (int e = check_launch("cam_row_kernel")) return e;
    dim3 g2(T, B);
    cam_col_kernel<<<g2, 256, 0, st>>>(out, N, T);
    return check_launch("cam_col_kernel");
}
