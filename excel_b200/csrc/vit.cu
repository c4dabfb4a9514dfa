// CLIP-surgery ViT forward (reference: clip/clip_surgery_model.py:76-159, 285-371, 418-448) for a batch.
//
// Token layout is [B, N, D] (batch-major; the reference runs LND, which only permutes the same numbers).
// Outputs follow clip.generate_clip_fts (clip/clip.py:348-358) BEFORE the token-axis normalisation:
//   tokens [B,N,E]; attn [L,B,N,N] (blocks before the surgery: head-MEAN of softmax(q k^T/sqrt(dh)),
//   surgery blocks: head-SUM); feats [L,B,N,D] with the reference's view-aliasing reproduced
//   (SURVEY.md §8 a5): feats[first-1] = final new-path x with the CLS row of the final x_ori,
//   feats[l] (first <= l < L-1) = x_ori_l + x_ori_res_{l+1}, feats[L-1] = x_ori_{L-1}.
// The aliasing falls out of the buffer plan: every block writes its state straight into feats[l], and the
// surgery blocks update feats[l-1] / feats[first-1] in place exactly where the reference's in-place `+=`
// mutates the views it had already appended.
#include <cuda_fp16.h>

#include "common.cuh"
#include "excel_b200.h"
#include "attn_tc.cuh"
#include "gemm_tc.cuh"

namespace xl {

// Probabilities (<= 1, mostly ~1/N) are scaled by 2^10 before the fp16 split so that hi/lo stay clear of fp16's
// subnormal range (absolute resolution 6e-8); the exact factor 2^-10 is folded into the GEMM's alpha.
constexpr float kProbScale = 1024.f;

// hi saturates at fp16's largest finite value, so |v| up to 2 x 65504 still splits into finite halves (no inf - inf = NaN)
__device__ __forceinline__ void split_store(__half* hi, __half* lo, float v) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    *hi = h;
    *lo = __float2half_rn(v - __half2float(h));
}

// ---- patch embedding: im2col (conv1 16x16/16, no bias == GEMM; clip_surgery_model.py:421), split-fp16 ----
__global__ void im2col_kernel(const float* __restrict__ img, int64_t sb, int64_t sc, int64_t sy, int S, int P, int g,
                              __half* __restrict__ col, int KKp) {
    // col[(b*g*g + py*g + px), c*P*P + iy*P + ix] = img[b, c, py*P+iy, px*P+ix]   (hi | lo halves, KKp apart)
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;  // column in [0, KKp)
    const int p = blockIdx.y, b = blockIdx.z;
    const int KK = 3 * P * P;
    if (kk >= KKp) return;
    float v = 0.f;
    if (kk < KK) {
        const int c = kk / (P * P), r = kk - c * P * P, iy = r / P, ix = r - iy * P;
        const int py = p / g, px = p - py * g;
        v = img[(int64_t)b * sb + (int64_t)c * sc + (int64_t)(py * P + iy) * sy + px * P + ix];
    }
    __half* row = col + ((int64_t)b * g * g + p) * 2 * KKp;
    split_store(row + kk, row + KKp + kk, v);
}

// ---- positional embedding, bilinear align_corners=False from g0 x g0 to g x g (:426-435) -------------
__global__ void pos_resize_kernel(const float* __restrict__ pos, int g0, int g, int D, float* __restrict__ out) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;  // n in [0, 1+g*g)
    if (d >= D) return;
    if (n == 0) { out[d] = pos[d]; return; }
    const int p = n - 1, oy = p / g, ox = p - oy * g;
    const float scale = (float)g0 / (float)g;
    float fy = scale * (oy + 0.5f) - 0.5f, fx = scale * (ox + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < g0 - 1 ? 1 : 0), x1 = x0 + (x0 < g0 - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* base = pos + D + d;  // grid part, [g0*g0, D]
    const float v = hy * (hx * base[(int64_t)(y0 * g0 + x0) * D] + lx * base[(int64_t)(y0 * g0 + x1) * D]) +
                    ly * (hx * base[(int64_t)(y1 * g0 + x0) * D] + lx * base[(int64_t)(y1 * g0 + x1) * D]);
    out[(int64_t)n * D + d] = v;
}

// ---- LayerNorm (fp32, eps 1e-5; clip_surgery_model.py:271-277), one warp per row -----------------------
// 16 B loads (a lane owns 4 consecutive channels per 128-channel group), 16 B fp32 stores, 8 B stores of the split-fp16
// halves.  Optional prologue used for the embedding: row n==0 of every image takes `cls`, and `pos[n]` is added
// before normalising (x = ln_pre(cat(cls, patches) + pos), :424-438).
template <bool EMBED>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bvec,
                 float* __restrict__ y, __half* __restrict__ ys, int64_t rows, int D, int N,
                 const float* __restrict__ cls, const float* __restrict__ pos) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * D;
    float* yr = y ? y + row * D : nullptr;
    __half* sr = ys ? ys + row * 2 * D : nullptr;  // split-fp16 copy (operand of the next GEMM): hi | lo, D apart
    const int n = EMBED ? (int)(row % N) : 0;
    constexpr int MAXV = 8;  // D <= 1024, D % 4 == 0
    float4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int d = 4 * (lane + 32 * i);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d < D) {
            t = *reinterpret_cast<const float4*>(xr + d);
            if (EMBED) {
                if (n == 0) t = *reinterpret_cast<const float4*>(cls + d);
                const float4 pp = *reinterpret_cast<const float4*>(pos + (int64_t)n * D + d);
                t.x += pp.x; t.y += pp.y; t.z += pp.z; t.w += pp.w;
            }
        }
        v[i] = t;
        s += (t.x + t.y) + (t.z + t.w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        if (4 * (lane + 32 * i) < D) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
            q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(e, e, q);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int d = 4 * (lane + 32 * i);
        if (d < D) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(w + d)), bb = __ldg(reinterpret_cast<const float4*>(bvec + d));
            const float4 o = make_float4((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y,
                                         (v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
            if (yr) *reinterpret_cast<float4*>(yr + d) = o;
            if (sr) {
                // hi saturates at fp16's largest finite value (|v| up to 2 x 65504 stays finite), lo = v - hi
                const float c0 = fminf(fmaxf(o.x, -65504.f), 65504.f), c1 = fminf(fmaxf(o.y, -65504.f), 65504.f);
                const float c2 = fminf(fmaxf(o.z, -65504.f), 65504.f), c3 = fminf(fmaxf(o.w, -65504.f), 65504.f);
                __align__(8) __half2 h[2], l[2];
                h[0] = __floats2half2_rn(c0, c1);
                h[1] = __floats2half2_rn(c2, c3);
                const float2 f0 = __half22float2(h[0]), f1 = __half22float2(h[1]);
                l[0] = __floats2half2_rn(o.x - f0.x, o.y - f0.y);
                l[1] = __floats2half2_rn(o.z - f1.x, o.w - f1.y);
                *reinterpret_cast<uint2*>(sr + d) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(sr + D + d) = *reinterpret_cast<const uint2*>(l);
            }
        }
    }
}

// pnew[b, 1+i, 1+j] += coef * ex_attn[b, i, j]  (row-padded map, pitch npad)
__global__ void lvc_add_kernel(const float* __restrict__ ex_attn, int np, float coef, float* __restrict__ pnew, int npad, int N) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= np) return;
    pnew[((int64_t)b * N + 1 + i) * npad + 1 + j] += coef * ex_attn[((int64_t)b * np + i) * np + j];
}

__global__ void copy_cls_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t stride_b, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (d < D) dst[(int64_t)b * stride_b + d] = src[(int64_t)b * stride_b + d];
}

struct Ws {  // workspace carve-up
    float *pos, *m, *pnew, *part, *x0;
    __half *col, *h, *qkv, *pn, *o, *o2, *u;
};

static size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

// partial head-sum maps of a split attention launch (small batches): [ngrp][B,N,Npad]
static size_t part_bytes(int B, int N, int H) {
    int hpi, gsplit;
    attn_pv_plan(B, H, N, &hpi, &gsplit);
    return gsplit ? (size_t)((H + hpi - 1) / hpi) * B * N * ((N + 3) & ~3) * 4 : 0;
}

static size_t ws_bytes(int B, int N, int D, int H, int KKp) {
    const size_t BN = (size_t)B * N, np = (size_t)((N + 63) & ~63);
    size_t t = 0;
    t += align256((size_t)N * D * 4);                 // pos
    t += align256((size_t)3 * B * H * N * 4);         // softmax row statistic (up to 3 score sets)
    t += align256((size_t)B * N * ((N + 3) & ~3) * 4); // pnew (row-padded new-path map: LVC branch only)
    t += align256(part_bytes(B, N, H));               // part
    t += align256(BN * D * 4);                        // x0
    t += align256((size_t)B * (N - 1) * 2 * KKp * 2); // col
    t += align256(BN * 2 * D * 2) + align256(2 * BN * 2 * D * 2);   // h; o2 | o
    t += align256(BN * 6 * D * 2);                    // qkv
    t += align256(BN * 2 * np * 2);                   // pn
    t += align256(BN * 8 * D * 2);                    // u
    return t + 256;
}

struct Maps {  // tensor maps of the activation operands (built once per forward)
    CUtensorMap col, h, qkv_a, qkv_v, pn, o, o2, oo, u;   // oo: o2 | o as one [2 BN, 2 D] matrix (merged out_proj)
};

struct Ctx {
    int B, N, D, H, dh, np, L;
    int64_t BN;
    Ws w;
    Maps m;
    cudaStream_t st;
};

static int layernorm(const Ctx& c, const float* x, const float* w, const float* b, __half* ys) {
    XL_CUDA(launch_pdl(layernorm_kernel<false>, dim3((unsigned)ceil_div64(c.BN, 8)), dim3(256), 0, c.st, x, w, b, nullptr, ys, c.BN, c.D, 1,
                       nullptr, nullptr));
    return check_launch("layernorm_kernel");
}

// A weight matrix in the engine's operand format: split fp16 [out, 2 K] (hi | lo) of  wscale * W  (wscale: the power of two
// chosen at load time that puts max|W| at 2^13..2^14, clear of fp16's subnormal range; folded back through alpha).
struct Wt { const void* ws; float scale; };

// y = act(x W^T + bias) (+ residual): x split [M rows per batch, 2K] (map ma), W split [Nout, 2K].
// batch = 2 runs two activations (row blocks a_row1 apart in ma) against the SAME weights into outputs c1 floats apart.
// residual == y (an in-place update) leaves as TMA reduce-adds: the epilogue then never waits for residual loads (a quarter of
// the stall samples of the residual GEMMs) and the add happens once, in fp32, in L2 -- the same single rounding.
static int linear(const Ctx& c, const CUtensorMap& ma, Wt w, int K, int Nout, const float* bias, int act,
                  const float* residual, float* y, __half* ys, int batch = 1, int64_t c1 = 0) {
    TcParams p = {};
    p.M = (int)c.BN; p.N = Nout; p.kblocks = K / 64; p.a_lo_off = K; p.b_lo_off = K; p.nb2 = 1;
    const bool in_place = residual != nullptr && residual == y && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && Nout % 4 == 0 &&
                          (batch == 1 || c1 % 4 == 0);
    if (in_place) { residual = nullptr; p.c_add = 1; }
    p.C = y; p.ldc = Nout; p.bias = bias; p.residual = residual; p.alpha = 1.f / w.scale; p.act = act;
    p.Cs = ys; p.lds = 2 * Nout; p.cs_lo_off = Nout;
    if (batch > 1) { p.a_row1 = (int)c.BN; p.c1 = c1; }
    const int bn = tc_pick_bn(c.BN, Nout, batch);
    CUtensorMap mw;
    if (int e = make_operand_map(&mw, w.ws, Nout, 2 * K, 2 * K, bn == 64 ? 64 : 128)) return e;
    return tc_gemm(ma, mw, p, batch, bn, c.st);
}

// Row statistics of softmax(scale * X_t,h Y_t,h^T) for `ntypes` score sets (column blocks xo[t], yo[t] of qkv_s) and, unless
// stats_only, the row-padded map  out[b] = coef * sum_t sum_h softmax(...)  -- the scores are never materialised (attn_tc.cu).
static int scores(const Ctx& c, int ntypes, const int* xo, const int* yo, float scale, float* out_padded, __half* out_split,
                  float coef, bool stats_only) {
    AttnParams p = {};
    p.B = c.B; p.H = c.H; p.N = c.N; p.ntypes = ntypes; p.lo_off = 3 * c.D;
    for (int t = 0; t < ntypes; ++t) { p.xo[t] = xo[t]; p.yo[t] = yo[t]; }
    p.alpha = scale * 1.4426950408889634f;  // exp2 domain
    p.m = c.w.m; p.out = out_padded; p.coef = coef; p.out_split = out_split; p.np = c.np;
    return attn_scores(c.m.qkv_a, p, c.st, stats_only);
}

// Original-path attention: stats pass, then ONE fused kernel (attn_pv.cu): out = coef * sum_h softmax(q_h k_h^T), written
// straight into the API's attention tensor [B,N,Npad], and o_s[b, :, h*dh..] = softmax(q_h k_h^T) V[b,h] -- the per-head
// probabilities stay in tensor memory.
static int attention_qk(const Ctx& c, float scale, float* out, float coef) {
    const int qx[1] = {0}, ky[1] = {c.D};
    if (int e = scores(c, 1, qx, ky, scale, nullptr, nullptr, 0.f, true)) return e;
    AttnPvParams q = {};
    q.B = c.B; q.H = c.H; q.N = c.N; q.D = c.D; q.xo = 0; q.yo = c.D; q.vo = 2 * c.D; q.lo_off = 3 * c.D;
    q.alpha = scale * 1.4426950408889634f; q.ml = c.w.m; q.out = out; q.coef = coef; q.o = c.w.o;
    attn_pv_plan(c.B, c.H, c.N, &q.hpi, &q.gsplit);
    q.part = c.w.part;
    return attn_pv(c.m.qkv_a, q, c.st);
}

// ln_1 -> in_proj -> split qkv (V is consumed in place, MN-major, by the attention kernel and the new-path GEMM)
static int qkv_stage(const Ctx& c, const float* src, const ExcelVitLayer& Lw) {
    if (int e = layernorm(c, src, Lw.ln1_w, Lw.ln1_b, c.w.h)) return e;
    return linear(c, c.m.h, Wt{Lw.in_ws, Lw.in_scale}, c.D, 3 * c.D, Lw.in_b, 0, nullptr, nullptr, c.w.qkv);
}

// feat = mid + c_proj(QuickGELU(c_fc(ln_2(mid))))
static int mlp_stage(const Ctx& c, const float* mid, const ExcelVitLayer& Lw, float* feat) {
    if (int e = layernorm(c, mid, Lw.ln2_w, Lw.ln2_b, c.w.h)) return e;
    if (int e = linear(c, c.m.h, Wt{Lw.fc_ws, Lw.fc_scale}, c.D, 4 * c.D, Lw.fc_b, 1, nullptr, nullptr, c.w.u)) return e;
    return linear(c, c.m.u, Wt{Lw.proj_ws, Lw.proj_scale}, 4 * c.D, c.D, Lw.proj_b, 0, mid, feat, nullptr);
}

}  // namespace xl

using namespace xl;

extern "C" int64_t excel_vit_workspace_bytes(int B, int S, int patch, int D, int heads) {
    const int g = S / patch, N = g * g + 1;
    return (int64_t)ws_bytes(B, N, D, heads, (3 * patch * patch + 63) & ~63);
}

extern "C" int excel_split_f16(const float* x, int64_t ldx, int rows, int cols, int Kp, float scale, void* out, void* stream) {
    XL_REQUIRE(scale > 0.f, "split_f16: scale must be positive");
    return split_f16(x, ldx, rows, cols, Kp, reinterpret_cast<__half*>(out), (cudaStream_t)stream, scale);
}

extern "C" int excel_vit_forward(const ExcelVitWeights* Wt, const float* img, int64_t img_stride_b, int64_t img_stride_c,
                                 int64_t img_stride_y, int B, int S, float* workspace, int64_t workspace_bytes,
                                 float* tokens, float* attn, int64_t attn_row_pitch, float* feats, const float* lvc_attn,
                                 void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(Wt != nullptr, "vit_forward: null weights");
    const int L = Wt->layers, D = Wt->width, H = Wt->heads, P = Wt->patch, E = Wt->embed, g0 = Wt->grid0,
              nsur = Wt->n_surgery;
    XL_REQUIRE(L >= 1 && D >= 64 && D <= 1024 && D % 64 == 0 && H >= 1 && D / H == 64 && P >= 1 && E >= 1 && g0 >= 1,
               "vit_forward: unsupported geometry L=%d D=%d H=%d P=%d (head dim must be 64)", L, D, H, P);
    XL_REQUIRE(nsur >= 1 && nsur < L, "vit_forward: n_surgery=%d must be in [1, L-1]", nsur);
    XL_REQUIRE(B >= 0 && S >= P && S % P == 0, "vit_forward: image size %d is not a multiple of the patch size %d", S, P);
    XL_REQUIRE(Wt->conv1_s && Wt->proj_t_s && Wt->conv1_scale > 0.f && Wt->proj_t_scale > 0.f,
               "vit_forward: split weights / scales missing (excel_split_f16 at load time)");
    if (B == 0) return 0;
    const int g = S / P, npatch = g * g, N = npatch + 1, KK = 3 * P * P, KKp = (KK + 63) & ~63, dh = D / H, first = L - nsur;
    const int np = (N + 63) & ~63, Npad = (N + 3) & ~3;
    const int64_t BN = (int64_t)B * N, ND = (int64_t)N * D;
    XL_REQUIRE(attn_row_pitch == Npad && (reinterpret_cast<uintptr_t>(attn) & 15) == 0,
               "vit_forward: attn must be [L,B,N,%d] (row pitch round_up(N,4): TMA store target), 16 B-aligned", Npad);
    XL_REQUIRE(workspace_bytes >= (int64_t)ws_bytes(B, N, D, H, KKp), "vit_forward: workspace too small");
    XL_REQUIRE(npatch <= 65535 && B <= 65535 && (int64_t)B * H * N < (1ll << 31) && BN * 6 * D < (1ll << 31),
               "vit_forward: problem too large");
    XL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vit_forward: workspace must be 256 B-aligned");
    const float scale = 1.f / sqrtf((float)dh);

    Ctx c;
    c.B = B; c.N = N; c.D = D; c.H = H; c.dh = dh; c.np = np; c.L = L; c.BN = BN; c.st = st;
    {
        uint8_t* p = reinterpret_cast<uint8_t*>(workspace);
        auto take = [&](size_t bytes) { uint8_t* r = p; p += align256(bytes); return r; };
        c.w.pos = (float*)take((size_t)N * D * 4);
        c.w.m = (float*)take((size_t)3 * B * H * N * 4);
        c.w.pnew = (float*)take((size_t)B * N * Npad * 4);
        c.w.part = (float*)take(part_bytes(B, N, H));
        c.w.x0 = (float*)take(BN * D * 4);
        c.w.col = (__half*)take((size_t)B * npatch * 2 * KKp * 2);
        c.w.h = (__half*)take(BN * 2 * D * 2);
        c.w.o2 = (__half*)take(2 * BN * 2 * D * 2);   // o2 | o contiguous: the merged out_proj reads them as one [2 BN, 2 D] matrix
        c.w.o = c.w.o2 + BN * 2 * D;
        c.w.qkv = (__half*)take(BN * 6 * D * 2);
        c.w.pn = (__half*)take(BN * 2 * np * 2);
        c.w.u = (__half*)take(BN * 8 * D * 2);
    }
    {
        int e = 0;
        e |= make_operand_map(&c.m.col, c.w.col, (int64_t)B * npatch, 2 * KKp, 2 * KKp, 128);
        e |= make_operand_map(&c.m.h, c.w.h, BN, 2 * D, 2 * D, 128);
        e |= make_operand_map(&c.m.qkv_a, c.w.qkv, BN, 6 * D, 6 * D, 128);
        e |= make_operand_map(&c.m.qkv_v, c.w.qkv, BN, 6 * D, 6 * D, 64);   // 64 x 64 boxes: MN-major B operand (V)
        e |= make_operand_map(&c.m.pn, c.w.pn, BN, 2 * np, 2 * np, 128);
        e |= make_operand_map(&c.m.o, c.w.o, BN, 2 * D, 2 * D, 128);
        e |= make_operand_map(&c.m.o2, c.w.o2, BN, 2 * D, 2 * D, 128);
        e |= make_operand_map(&c.m.oo, c.w.o2, 2 * BN, 2 * D, 2 * D, 128);
        e |= make_operand_map(&c.m.u, c.w.u, BN, 8 * D, 8 * D, 128);
        if (e) return e;
    }

    // ---- embedding: conv1 as GEMM, + cls, + pos, ln_pre (:421-438)
    {
        dim3 grid(ceil_div(KKp, 256), npatch, B);
        im2col_kernel<<<grid, 256, 0, st>>>(img, img_stride_b, img_stride_c, img_stride_y, S, P, g, c.w.col, KKp);
        if (int e = check_launch("im2col_kernel")) return e;
        CUtensorMap m_conv;
        if (int e = make_operand_map(&m_conv, Wt->conv1_s, D, 2 * KKp, 2 * KKp, 128)) return e;
        TcParams p = {};  // patches of image b land in rows 1..npatch of x0[b]
        p.M = npatch; p.N = D; p.kblocks = KKp / 64; p.a_lo_off = KKp; p.b_lo_off = KKp; p.nb2 = 1; p.a_row1 = npatch;
        p.C = c.w.x0 + D; p.ldc = D; p.c1 = ND; p.alpha = 1.f / Wt->conv1_scale;
        if (int e = tc_gemm(c.m.col, m_conv, p, B, 128, st)) return e;
        const float* pos = Wt->pos;
        if (g != g0) {
            dim3 gp(ceil_div(D, 256), N);
            pos_resize_kernel<<<gp, 256, 0, st>>>(Wt->pos, g0, g, D, c.w.pos);
            if (int e = check_launch("pos_resize_kernel")) return e;
            pos = c.w.pos;
        }
        layernorm_kernel<true><<<(unsigned)ceil_div64(BN, 8), 256, 0, st>>>(c.w.x0, Wt->ln_pre_w, Wt->ln_pre_b, c.w.x0, nullptr,
                                                                            BN, D, N, Wt->cls, pos);
        if (int e = check_launch("layernorm_kernel<embed>")) return e;
    }

    const int self_xy[3] = {0, D, 2 * D};  // column blocks of qkv: q, k, v
    const float* x = c.w.x0;  // current single-path state (blocks before the surgery)
    for (int l = 0; l < L; ++l) {
        const ExcelVitLayer& Lw = Wt->blocks[l];
        XL_REQUIRE(Lw.in_ws && Lw.out_ws && Lw.fc_ws && Lw.proj_ws && Lw.in_scale > 0.f && Lw.out_scale > 0.f && Lw.fc_scale > 0.f &&
                   Lw.proj_scale > 0.f, "vit_forward: split weights / scales of block %d missing", l);
        const struct Wt w_out = {Lw.out_ws, Lw.out_scale};
        float* attn_l = attn + (int64_t)l * B * N * Npad;
        float* feat_l = feats + (int64_t)l * BN * D;
        if (l < first) {  // ---- standard block (:332-337)
            if (int e = qkv_stage(c, x, Lw)) return e;
            if (int e = attention_qk(c, scale, attn_l, 1.f / H)) return e;    // need_weights: head mean; o = attn @ v
            if (int e = linear(c, c.m.o, w_out, D, D, Lw.out_b, 0, x, feat_l, nullptr)) return e;         // x + attn, into this block's slot
            if (int e = mlp_stage(c, feat_l, Lw, feat_l)) return e;                                      // += MLP, in place
            x = feat_l;
        } else {  // ---- surgery block (:309-330, Attention.forward :95-159)
            float* xnew = feats + (int64_t)(first - 1) * BN * D;              // new path, accumulates x_res in place
            float* src = feats + (int64_t)(l - 1) * BN * D;                   // X_{first-1} or previous x_ori
            if (int e = qkv_stage(c, src, Lw)) return e;
            // new path: (softmax(qq^T) + softmax(kk^T) + softmax(vv^T))/3 summed over heads (:119-125,146).  The map pass
            // writes the 2^10-scaled map straight into the split-fp16 A operand of the P V GEMM below; only the LVC branch
            // (which adds ex_attn to the fp32 map first) takes the fp32 map + split pass.
            if (!lvc_attn) {
                if (int e = scores(c, 3, self_xy, self_xy, scale, nullptr, c.w.pn, 1.f / 3.f, false)) return e;
            } else {   // LVC: + ex_attn on every head's patch block, then summed over heads (:139-146) == + H * ex_attn
                if (int e = scores(c, 3, self_xy, self_xy, scale, c.w.pnew, nullptr, 1.f / 3.f, false)) return e;
                dim3 grid(ceil_div(N - 1, 256), N - 1, B);
                lvc_add_kernel<<<grid, 256, 0, st>>>(lvc_attn, N - 1, (float)H, c.w.pnew, Npad, N);
                if (int e = check_launch("lvc_add_kernel")) return e;
                if (int e = split_f16(c.w.pnew, Npad, (int)BN, N, np, c.w.pn, st, kProbScale)) return e;
            }
            {   // x = attn @ v with the head-summed map applied to every head's v (:149): [N,N] x [N,D] per image
                TcParams p = {};
                // B operand = V [keys, D] in place inside the split qkv matrix (MN-major): no transpose pass
                p.M = N; p.N = D; p.kblocks = np / 64; p.a_lo_off = np; p.nb2 = 1;
                p.a_row1 = N; p.b_mn = 1; p.b_row1 = N; p.b_col0 = 2 * D; p.b_lo_off = 3 * D; p.alpha = 1.f / kProbScale;
                p.Cs = c.w.o2; p.lds = 2 * D; p.cs1 = (int64_t)N * 2 * D; p.cs_lo_off = D;
                if (int e = tc_gemm(c.m.pn, c.m.qkv_v, p, B, tc_pick_bn(N, D, B), st)) return e;
            }
            // original path: softmax(q k^T); returned attention = head SUM (:101-102,154)
            if (int e = attention_qk(c, scale, attn_l, 1.f)) return e;          // x_ori = attn_ori @ v
            // mid = src + proj(x_ori): a separate buffer for the first surgery block, in place afterwards
            // (the reference's `x_ori += x_ori_res` mutates the view it stored in all_feats[l-1], :317)
            float* mid = (l == first) ? feat_l : src;   // first surgery block: into its own slot, then the MLP updates it in place
            if (l == first) {   // src IS xnew here: the two products read / update the same rows, so they stay two launches
                if (int e = linear(c, c.m.o, w_out, D, D, Lw.out_b, 0, src, mid, nullptr)) return e;
                if (int e = linear(c, c.m.o2, w_out, D, D, Lw.out_b, 0, xnew, xnew, nullptr)) return e;   // x += x_res (:319,329)
            } else {            // one launch: [o2 | o] x out_proj^T -> xnew += .., src += ..  (in place, disjoint buffers)
                if (int e = linear(c, c.m.oo, w_out, D, D, Lw.out_b, 0, xnew, xnew, nullptr, 2, src - xnew)) return e;
            }
            if (int e = mlp_stage(c, mid, Lw, feat_l)) return e;                                         // x_ori
        }
    }
    // x[0] = x_ori[0] (:442), ln_post, @ proj (:445-446)
    float* xnew = feats + (int64_t)(first - 1) * BN * D;
    {
        dim3 grid(ceil_div(D, 256), B);
        copy_cls_kernel<<<grid, 256, 0, st>>>(feats + (int64_t)(L - 1) * BN * D, xnew, ND, D);
        if (int e = check_launch("copy_cls_kernel")) return e;
    }
    if (int e = layernorm(c, xnew, Wt->ln_post_w, Wt->ln_post_b, c.w.h)) return e;
    CUtensorMap m_pt;
    if (int e = make_operand_map(&m_pt, Wt->proj_t_s, E, 2 * D, 2 * D, E <= 64 ? 64 : 128)) return e;
    TcParams p = {};
    p.M = (int)BN; p.N = E; p.kblocks = D / 64; p.a_lo_off = D; p.b_lo_off = D; p.nb2 = 1;
    p.C = tokens; p.ldc = E; p.alpha = 1.f / Wt->proj_t_scale;
    return tc_gemm(c.m.h, m_pt, p, 1, E <= 64 ? 64 : 128, st);
}
