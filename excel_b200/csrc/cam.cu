// Patch x text-bank CAM (reference: clip/clip.py:288-310 clip_feature_surgery, clip/clip.py:353 token norm).
//
// The reference materialises feats[B,N,T,E] = F (x) T (1.5 GB at 512^2 x16) only to reduce it again; by
// linearity  sim[n,t] = w_t S[n,t] - mean_t'(w_t' S[n,t'])  with S = F T^T, so the path is ONE GEMM on the tcgen05
// engine (gemm_tc.cu: split-fp16 operands, 3 MMA passes, fp32 accumulation in TMEM, TMA-fed) plus two epilogue kernels:
//   sim kernel : w = softmax_t(2 S[b,0,:]) / mean(..) once per block (clip.py:295-297); sim rows (clip.py:301-306),
//                one warp per token row (coalesced along t), per-block column min / max partials;
//   norm kernel: per (b,t) min / max over ALL N tokens incl. CLS from the partials, (sim-min)/(max-min), no epsilon
//                (clip.py:308), coalesced along t -- no strided column walk.
// Operand range: each operand is multiplied by a power of two chosen ON THE DEVICE from its max-abs (amax kernel) so that
// hi / lo stay inside fp16's normal range whatever the caller passes (token-normalised features are ~1/sqrt(N): their lo
// halves would be fp16 subnormals); the exact inverse factor is applied by the sim kernel.
#include <cuda_fp16.h>

#include "common.cuh"
#include "excel_b200.h"
#include "gemm_tc.cuh"

namespace xl {

// ---- image_features / image_features.norm(dim=1)  (norm over the TOKEN axis, clip/clip.py:353) -------
__global__ void __launch_bounds__(1024)
token_sumsq_kernel(const float* __restrict__ tok, int N, int E, float* __restrict__ norm) {
    pdl_trigger();
    pdl_wait();
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
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = (int)(i % E);
    const int64_t b = i / ((int64_t)N * E);
    out[i] = tok[i] / norm[b * E + e];
}

// ---- operand pre-scale: amax -> power of two (device side, no host sync) --------------------------------
__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ amax) {
    pdl_trigger();
    pdl_wait();
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));   // m >= 0: bit order == value order
}
// 2^k with amax * 2^k in [2^13, 2^14): hi = fp16(x) keeps 11 bits, lo ~ 2^-11 x stays a NORMAL fp16 down to |x| = amax * 2^-16
__device__ __forceinline__ float pow2_scale(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    return exp2f((float)(13 - ilogbf(amax)));
}
__global__ void __launch_bounds__(256)
split_scaled_kernel(const float* __restrict__ x, int64_t ldx, int cols, int Kp, __half* __restrict__ out, const float* __restrict__ amax) {
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (c >= Kp) return;
    const float v = c < cols ? x[r * ldx + c] * pow2_scale(*amax) : 0.f;
    const __half h = __float2half_rn(v);
    out[r * 2 * Kp + c] = h;
    out[r * 2 * Kp + Kp + c] = __float2half_rn(v - __half2float(h));
}

// ---- sim[b,n,:] from S[b,n,:] (pitch Tp); one warp per token row, kCamRows rows per block -------------------
constexpr int kCamRows = 64, kCamSlots = 16;   // T <= 32 * kCamSlots
__global__ void __launch_bounds__(256)
cam_sim_kernel(const float* __restrict__ S, int Tp, const float* __restrict__ amax2, float* __restrict__ sim, float* __restrict__ part,
               int N, int T, int nblk) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sh[];            // w[T], then [8][T] min, [8][T] max
    __shared__ float red[32];
    float* w = sh;
    float* smin = sh + T;
    float* smax = smin + 8 * T;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, b = blockIdx.y;
    const float inv = 1.f / (pow2_scale(amax2[0]) * pow2_scale(amax2[1]));   // exact: powers of two
    const float* cls = S + (int64_t)b * N * Tp;  // row 0 = CLS token
    // w_t = softmax_t(2 S[b,0,t]) / mean_t(softmax)   (clip.py:295-297)
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < T; t += 256) mx = fmaxf(mx, 2.f * cls[t] * inv);
    mx = block_reduce(mx, red, OpMax(), -INFINITY);
    float se = 0.f;
    for (int t = threadIdx.x; t < T; t += 256) se += expf(2.f * cls[t] * inv - mx);
    se = block_reduce(se, red, OpSum(), 0.f);
    float sp = 0.f;
    for (int t = threadIdx.x; t < T; t += 256) sp += expf(2.f * cls[t] * inv - mx) / se;
    sp = block_reduce(sp, red, OpSum(), 0.f);
    const float pmean = sp / (float)T;
    for (int t = threadIdx.x; t < T; t += 256) w[t] = (expf(2.f * cls[t] * inv - mx) / se) / pmean;
    __syncthreads();
    float lo[kCamSlots], hi[kCamSlots];
#pragma unroll
    for (int i = 0; i < kCamSlots; ++i) { lo[i] = INFINITY; hi[i] = -INFINITY; }
    const int n0 = blockIdx.x * kCamRows;
    for (int n = n0 + wid; n < min(n0 + kCamRows, N); n += 8) {
        const float* row = S + ((int64_t)b * N + n) * Tp;
        float v[kCamSlots], acc = 0.f;
#pragma unroll
        for (int i = 0; i < kCamSlots; ++i) {
            const int t = lane + 32 * i;
            v[i] = t < T ? row[t] * inv * w[t] : 0.f;     // sum_c f*t*w (clip.py:301-302)
            acc += v[i];
        }
        const float mean = warp_sum(acc) / (float)T;      // redundant term (clip.py:303-304)
        float* o = sim + ((int64_t)b * N + n) * T;
#pragma unroll
        for (int i = 0; i < kCamSlots; ++i) {
            const int t = lane + 32 * i;
            if (t < T) {
                const float sv = v[i] - mean;
                o[t] = sv;
                lo[i] = fminf(lo[i], sv);
                hi[i] = fmaxf(hi[i], sv);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < kCamSlots; ++i) {
        const int t = lane + 32 * i;
        if (t < T) { smin[wid * T + t] = lo[i]; smax[wid * T + t] = hi[i]; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += 256) {
        float a = smin[t], c = smax[t];
#pragma unroll
        for (int k = 1; k < 8; ++k) { a = fminf(a, smin[k * T + t]); c = fmaxf(c, smax[k * T + t]); }
        part[(((int64_t)b * nblk + blockIdx.x) * 2) * T + t] = a;
        part[(((int64_t)b * nblk + blockIdx.x) * 2 + 1) * T + t] = c;
    }
}

// ---- (sim - min_n) / (max_n - min_n) per (b,t), in place; min / max over ALL N tokens from the block partials ----
__global__ void __launch_bounds__(256)
cam_norm_kernel(float* __restrict__ sim, const float* __restrict__ part, int N, int T, int nblk) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sh[];            // lo[T], rng[T]
    const int b = blockIdx.y;
    for (int t = threadIdx.x; t < T; t += 256) {
        float a = INFINITY, c = -INFINITY;
        for (int k = 0; k < nblk; ++k) {
            a = fminf(a, part[(((int64_t)b * nblk + k) * 2) * T + t]);
            c = fmaxf(c, part[(((int64_t)b * nblk + k) * 2 + 1) * T + t]);
        }
        sh[t] = a;
        sh[T + t] = c - a;
    }
    __syncthreads();
    const int64_t total = (int64_t)N * T;
    float* o = sim + (int64_t)b * total;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int t = (int)(i % T);
        o[i] = (o[i] - sh[t]) / sh[T + t];
    }
}

}  // namespace xl

using namespace xl;

extern "C" int excel_token_normalize(const float* tok, int B, int N, int E, float* norm_ws, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && N >= 1 && E >= 1 && B <= 65535, "token_normalize: bad shape");
    if (B == 0) return 0;
    dim3 grid(ceil_div(E, 32), B), block(32, 32);
    XL_CUDA(launch_pdl(token_sumsq_kernel, dim3(grid), dim3(block), 0, st, tok, N, E, norm_ws));
    if (int e = check_launch("token_sumsq_kernel")) return e;
    const int64_t total = (int64_t)B * N * E;
    XL_CUDA(launch_pdl(token_div_kernel, dim3((unsigned)ceil_div64(total, 256)), dim3(256), 0, st, tok, norm_ws, out, N, E, total));
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

extern "C" int64_t excel_cam_workspace_bytes(int B, int N, int E, int T) {
    const int64_t Ep = (E + 63) & ~63, Tp = (T + 3) & ~3, nblk = ceil_div(N, kCamRows);
    auto al = [](int64_t b) { return (b + 255) & ~int64_t(255); };
    return al(16) + al((int64_t)B * N * 2 * Ep * 2) + al((int64_t)T * 2 * Ep * 2) + al((int64_t)B * N * Tp * 4) + al((int64_t)B * nblk * 2 * T * 4);
}

extern "C" int excel_cam_surgery(const float* feats, const float* text, int B, int N, int E, int T, void* workspace,
                                 int64_t workspace_bytes, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && N >= 1 && E >= 1 && T >= 1 && B <= 65535 && T <= 32 * kCamSlots, "cam_surgery: bad shape (T <= %d)", 32 * kCamSlots);
    XL_REQUIRE((int64_t)B * N < (1ll << 31), "cam_surgery: too many tokens");
    if (B == 0) return 0;
    XL_REQUIRE(workspace && workspace_bytes >= excel_cam_workspace_bytes(B, N, E, T) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "cam_surgery: workspace too small or not 256 B-aligned (excel_cam_workspace_bytes)");
    const int Ep = (E + 63) & ~63, Tp = (T + 3) & ~3, nblk = ceil_div(N, kCamRows);
    const int64_t M = (int64_t)B * N;
    uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
    auto take = [&](int64_t bytes) { uint8_t* r = p; p += (bytes + 255) & ~int64_t(255); return r; };
    float* amax = (float*)take(16);
    __half* Fs = (__half*)take(M * 2 * Ep * 2);
    __half* Ts = (__half*)take((int64_t)T * 2 * Ep * 2);
    float* S = (float*)take(M * Tp * 4);
    float* part = (float*)take((int64_t)B * nblk * 2 * T * 4);
    XL_CUDA(cudaMemsetAsync(amax, 0, 16, st));
    XL_CUDA(launch_pdl(amax_kernel, dim3((unsigned)(ceil_div64(M * E, 256 * 8) < 4 * kNumSMs ? ceil_div64(M * E, 256 * 8) : 4 * kNumSMs)), dim3(256), 0, st, feats, M * E, amax));
    if (int e = check_launch("amax_kernel")) return e;
    XL_CUDA(launch_pdl(amax_kernel, dim3((unsigned)(ceil_div64((int64_t)T * E, 256) < kNumSMs ? ceil_div64((int64_t)T * E, 256) : kNumSMs)), dim3(256), 0, st, text, (int64_t)T * E, amax + 1));
    if (int e = check_launch("amax_kernel")) return e;
    for (int64_t r0 = 0; r0 < M; r0 += 65535) {
        const int nr = (int)(M - r0 < 65535 ? M - r0 : 65535);
        XL_CUDA(launch_pdl(split_scaled_kernel, dim3(dim3(ceil_div(Ep, 256), nr)), dim3(256), 0, st, feats + r0 * E, E, E, Ep, Fs + r0 * 2 * Ep, amax));
        if (int e = check_launch("split_scaled_kernel")) return e;
    }
    XL_CUDA(launch_pdl(split_scaled_kernel, dim3(dim3(ceil_div(Ep, 256), T)), dim3(256), 0, st, text, E, E, Ep, Ts, amax + 1));
    if (int e = check_launch("split_scaled_kernel")) return e;
    // S[B*N, T] (pitch Tp) = F_s T_s^T on the tensor cores: M128 x N64/128 tiles, K = Ep
    CUtensorMap tmA, tmB;
    const int bn = T <= 64 ? 64 : 128;
    if (int e = make_operand_map(&tmA, Fs, M, 2 * Ep, 2 * Ep, 128)) return e;
    if (int e = make_operand_map(&tmB, Ts, T, 2 * Ep, 2 * Ep, bn == 64 ? 64 : 128)) return e;
    TcParams q = {};
    q.M = (int)M; q.N = T; q.kblocks = Ep / 64; q.a_lo_off = Ep; q.b_lo_off = Ep; q.nb2 = 1;
    q.C = S; q.ldc = Tp; q.alpha = 1.f;
    if (int e = tc_gemm(tmA, tmB, q, 1, bn, st)) return e;
    const size_t sm1 = (size_t)(17 * T) * sizeof(float), sm2 = (size_t)(2 * T) * sizeof(float);
    XL_CUDA(launch_pdl(cam_sim_kernel, dim3(dim3(nblk, B)), dim3(256), sm1, st, S, Tp, amax, out, part, N, T, nblk));
    if (int e = check_launch("cam_sim_kernel")) return e;
    const int gx = (int)(ceil_div64((int64_t)N * T, 256 * 4) < 64 ? ceil_div64((int64_t)N * T, 256 * 4) : 64);
    XL_CUDA(launch_pdl(cam_norm_kernel, dim3(dim3(gx, B)), dim3(256), sm2, st, out, part, N, T, nblk));
    return check_launch("cam_norm_kernel");
}
