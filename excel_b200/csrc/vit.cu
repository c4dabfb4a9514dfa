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
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

// ---- patch embedding: im2col (conv1 16x16/16, no bias == GEMM; clip_surgery_model.py:421) ------------
__global__ void im2col_kernel(const float* __restrict__ img, int64_t sb, int64_t sc, int64_t sy, int S, int P, int g,
                              float* __restrict__ col) {
    // col[(b*g*g + py*g + px), c*P*P + iy*P + ix] = img[b, c, py*P+iy, px*P+ix]
    const int kk = blockIdx.x * blockDim.x + threadIdx.x;  // column in [0, 3*P*P)
    const int p = blockIdx.y, b = blockIdx.z;
    const int KK = 3 * P * P;
    if (kk >= KK) return;
    const int c = kk / (P * P), r = kk - c * P * P, iy = r / P, ix = r - iy * P;
    const int py = p / g, px = p - py * g;
    col[((int64_t)b * g * g + p) * KK + kk] = img[(int64_t)b * sb + (int64_t)c * sc + (int64_t)(py * P + iy) * sy + px * P + ix];
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
// Optional prologue used for the embedding: row n==0 of every image takes `cls`, and `pos[n]` is added
// before normalising (x = ln_pre(cat(cls, patches) + pos), :424-438).
template <bool EMBED>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bvec,
                 float* __restrict__ y, int64_t rows, int D, int N, const float* __restrict__ cls,
                 const float* __restrict__ pos) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* xr = x + row * D;
    float* yr = y + row * D;
    const int n = EMBED ? (int)(row % N) : 0;
    constexpr int MAXV = 32;  // D <= 1024
    float v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int d = lane + 32 * i;
        float t = 0.f;
        if (d < D) {
            t = xr[d];
            if (EMBED) t = (n == 0 ? cls[d] : t) + pos[(int64_t)n * D + d];
        }
        v[i] = t;
        s += t;
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int d = lane + 32 * i;
        const float t = d < D ? v[i] - mean : 0.f;
        q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int d = lane + 32 * i;
        if (d < D) yr[d] = (v[i] - mean) * rstd * w[d] + bvec[d];
    }
}

// ---- row softmax in place (rows of length n); one 128-thread block per row ----------------------------
__global__ void __launch_bounds__(128)
softmax_rows_kernel(float* __restrict__ S, int n) {
    __shared__ float red[32];
    float* row = S + (int64_t)blockIdx.x * n;
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < n; j += 128) mx = fmaxf(mx, row[j]);
    mx = block_reduce(mx, red, OpMax(), -INFINITY);
    float sum = 0.f;
    for (int j = threadIdx.x; j < n; j += 128) {
        const float e = expf(row[j] - mx);
        row[j] = e;
        sum += e;
    }
    sum = block_reduce(sum, red, OpSum(), 0.f);
    for (int j = threadIdx.x; j < n; j += 128) row[j] = row[j] / sum;
}

// ---- out[b,i,j] (+)= coef * sum_h P[b,h,i,j] ------------------------------------------------------------
__global__ void head_reduce_kernel(const float* __restrict__ P, float* __restrict__ out, int H, int64_t nn, float coef,
                                   int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= nn) return;
    const float* p = P + (int64_t)b * H * nn + i;
    float s = 0.f;
    for (int h = 0; h < H; ++h) s += p[(int64_t)h * nn];
    s *= coef;
    float* o = out + (int64_t)b * nn + i;
    *o = accumulate ? *o + s : s;
}

__global__ void copy_cls_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t stride_b, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (d < D) dst[(int64_t)b * stride_b + d] = src[(int64_t)b * stride_b + d];
}

struct Ws {  // workspace carve-up (floats)
    float *col, *pos, *h, *qkv, *S, *pnew, *o, *o2, *mid, *u, *x0;
};

static size_t ws_floats(int B, int N, int D, int H, int KK) {
    const size_t BN = (size_t)B * N;
    return (size_t)B * (N - 1) * KK + (size_t)N * D + BN * D + BN * 3 * D + (size_t)B * H * N * N + (size_t)B * N * N +
           BN * D * 4 + BN * 4 * D + 64;
}

static int layernorm(const float* x, const float* w, const float* b, float* y, int64_t rows, int D, cudaStream_t st) {
    layernorm_kernel<false><<<(unsigned)ceil_div64(rows, 8), 256, 0, st>>>(x, w, b, y, rows, D, 1, nullptr, nullptr);
    return check_launch("layernorm_kernel");
}

// y = act(x W^T + bias) + residual   (x [M,K], W [Nout,K])
static int linear(const float* x, const float* W, const float* bias, const float* residual, float* y, int M, int Nout,
                  int K, int act, cudaStream_t st) {
    return sgemm2(x, W, y, bias, residual, M, Nout, K, K, K, Nout, 1, 0, 0, 0, 1, 0, 0, 0, 1.f, 1, act, st);
}

// S[b,h] = softmax(scale * X_h Y_h^T) for X, Y column blocks of qkv (offsets xo, yo into the 3D row)
static int scores(const float* qkv, int xo, int yo, float* S, int B, int N, int D, int H, float scale, cudaStream_t st) {
    const int dh = D / H;
    if (int e = sgemm2(qkv + xo, qkv + yo, S, nullptr, nullptr, N, N, dh, 3 * D, 3 * D, N, B, (int64_t)N * 3 * D,
                       (int64_t)N * 3 * D, (int64_t)H * N * N, H, dh, dh, (int64_t)N * N, scale, 1, 0, st)) return e;
    softmax_rows_kernel<<<(unsigned)((int64_t)B * H * N), 128, 0, st>>>(S, N);
    return check_launch("softmax_rows_kernel");
}

static int head_reduce(const float* P, float* out, int B, int H, int N, float coef, int accumulate, cudaStream_t st) {
    const int64_t nn = (int64_t)N * N;
    dim3 grid((unsigned)ceil_div64(nn, 256), B);
    head_reduce_kernel<<<grid, 256, 0, st>>>(P, out, H, nn, coef, accumulate);
    return check_launch("head_reduce_kernel");
}

}  // namespace xl

using namespace xl;

extern "C" int64_t excel_vit_workspace_bytes(int B, int S, int patch, int D, int heads) {
    const int g = S / patch, N = g * g + 1;
    return (int64_t)(ws_floats(B, N, D, heads, 3 * patch * patch) * sizeof(float));
}

extern "C" int excel_vit_forward(const ExcelVitWeights* Wt, const float* img, int64_t img_stride_b, int64_t img_stride_c,
                                 int64_t img_stride_y, int B, int S, float* workspace, int64_t workspace_bytes,
                                 float* tokens, float* attn, float* feats, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(Wt != nullptr, "vit_forward: null weights");
    const int L = Wt->layers, D = Wt->width, H = Wt->heads, P = Wt->patch, E = Wt->embed, g0 = Wt->grid0,
              nsur = Wt->n_surgery;
    XL_REQUIRE(L >= 1 && D >= 64 && D <= 1024 && H >= 1 && D % H == 0 && P >= 1 && E >= 1 && g0 >= 1,
               "vit_forward: unsupported geometry L=%d D=%d H=%d P=%d", L, D, H, P);
    XL_REQUIRE(nsur >= 1 && nsur < L, "vit_forward: n_surgery=%d must be in [1, L-1]", nsur);
    XL_REQUIRE(B >= 0 && S >= P && S % P == 0, "vit_forward: image size %d is not a multiple of the patch size %d", S, P);
    if (B == 0) return 0;
    const int g = S / P, np = g * g, N = np + 1, KK = 3 * P * P, dh = D / H, first = L - nsur;
    const int64_t BN = (int64_t)B * N, ND = (int64_t)N * D;
    XL_REQUIRE(workspace_bytes >= (int64_t)(ws_floats(B, N, D, H, KK) * sizeof(float)), "vit_forward: workspace too small");
    XL_REQUIRE(np <= 65535 && B <= 65535 && (int64_t)B * H * N < (1ll << 31), "vit_forward: problem too large");
    const float scale = 1.f / sqrtf((float)dh);

    Ws w;
    float* p = workspace;
    w.col = p;  p += (size_t)B * np * KK;
    w.pos = p;  p += (size_t)N * D;
    w.h = p;    p += BN * D;
    w.qkv = p;  p += BN * 3 * D;
    w.S = p;    p += (size_t)B * H * N * N;
    w.pnew = p; p += (size_t)B * N * N;
    w.o = p;    p += BN * D;
    w.o2 = p;   p += BN * D;
    w.mid = p;  p += BN * D;
    w.x0 = p;   p += BN * D;
    w.u = p;    p += BN * 4 * D;

    // ---- embedding: conv1 as GEMM, + cls, + pos, ln_pre (:421-438)
    {
        dim3 grid(ceil_div(KK, 256), np, B);
        im2col_kernel<<<grid, 256, 0, st>>>(img, img_stride_b, img_stride_c, img_stride_y, S, P, g, w.col);
        if (int e = check_launch("im2col_kernel")) return e;
        // patches of image b land in rows 1..np of x0[b]
        if (int e = sgemm2(w.col, Wt->conv1, w.x0 + D, nullptr, nullptr, np, D, KK, KK, KK, D, B, (int64_t)np * KK, 0, ND, 1,
                           0, 0, 0, 1.f, 1, 0, st)) return e;
        const float* pos = Wt->pos;
        if (g != g0) {
            dim3 gp(ceil_div(D, 256), N);
            pos_resize_kernel<<<gp, 256, 0, st>>>(Wt->pos, g0, g, D, w.pos);
            if (int e = check_launch("pos_resize_kernel")) return e;
            pos = w.pos;
        }
        layernorm_kernel<true><<<(unsigned)ceil_div64(BN, 8), 256, 0, st>>>(w.x0, Wt->ln_pre_w, Wt->ln_pre_b, w.x0, BN, D, N,
                                                                            Wt->cls, pos);
        if (int e = check_launch("layernorm_kernel<embed>")) return e;
    }

    const float* x = w.x0;  // current single-path state (blocks before the surgery)
    for (int l = 0; l < L; ++l) {
        const ExcelVitLayer& Lw = Wt->blocks[l];
        float* attn_l = attn + (int64_t)l * B * N * N;
        float* feat_l = feats + (int64_t)l * BN * D;
        if (l < first) {  // ---- standard block (:332-337)
            if (int e = layernorm(x, Lw.ln1_w, Lw.ln1_b, w.h, BN, D, st)) return e;
            if (int e = linear(w.h, Lw.in_w, Lw.in_b, nullptr, w.qkv, (int)BN, 3 * D, D, 0, st)) return e;
            if (int e = scores(w.qkv, 0, D, w.S, B, N, D, H, scale, st)) return e;
            if (int e = head_reduce(w.S, attn_l, B, H, N, 1.f / H, 0, st)) return e;  // need_weights: head mean
            if (int e = sgemm2(w.S, w.qkv + 2 * D, w.o, nullptr, nullptr, N, dh, N, N, 3 * D, D, B, (int64_t)H * N * N,
                               (int64_t)N * 3 * D, ND, H, (int64_t)N * N, dh, dh, 1.f, 0, 0, st)) return e;
            if (int e = linear(w.o, Lw.out_w, Lw.out_b, x, w.mid, (int)BN, D, D, 0, st)) return e;       // x + attn
            if (int e = layernorm(w.mid, Lw.ln2_w, Lw.ln2_b, w.h, BN, D, st)) return e;
            if (int e = linear(w.h, Lw.fc_w, Lw.fc_b, nullptr, w.u, (int)BN, 4 * D, D, 1, st)) return e; // QuickGELU
            if (int e = linear(w.u, Lw.proj_w, Lw.proj_b, w.mid, feat_l, (int)BN, D, 4 * D, 0, st)) return e;
            x = feat_l;
        } else {  // ---- surgery block (:309-330, Attention.forward :95-159)
            float* xnew = feats + (int64_t)(first - 1) * BN * D;              // new path, accumulates x_res in place
            float* src = feats + (int64_t)(l - 1) * BN * D;                   // X_{first-1} or previous x_ori
            if (int e = layernorm(src, Lw.ln1_w, Lw.ln1_b, w.h, BN, D, st)) return e;
            if (int e = linear(w.h, Lw.in_w, Lw.in_b, nullptr, w.qkv, (int)BN, 3 * D, D, 0, st)) return e;
            // new path: (softmax(qq^T) + softmax(kk^T) + softmax(vv^T))/3 summed over heads (:119-125,146)
            for (int t = 0; t < 3; ++t) {
                if (int e = scores(w.qkv, t * D, t * D, w.S, B, N, D, H, scale, st)) return e;
                if (int e = head_reduce(w.S, w.pnew, B, H, N, 1.f / 3.f, t > 0, st)) return e;
            }
            // original path: softmax(q k^T); returned attention = head SUM (:101-102,154)
            if (int e = scores(w.qkv, 0, D, w.S, B, N, D, H, scale, st)) return e;
            if (int e = head_reduce(w.S, attn_l, B, H, N, 1.f, 0, st)) return e;
            if (int e = sgemm2(w.S, w.qkv + 2 * D, w.o, nullptr, nullptr, N, dh, N, N, 3 * D, D, B, (int64_t)H * N * N,
                               (int64_t)N * 3 * D, ND, H, (int64_t)N * N, dh, dh, 1.f, 0, 0, st)) return e;  // x_ori = attn_ori @ v
            if (int e = sgemm2(w.pnew, w.qkv + 2 * D, w.o2, nullptr, nullptr, N, D, N, N, 3 * D, D, B, (int64_t)N * N,
                               (int64_t)N * 3 * D, ND, 1, 0, 0, 0, 1.f, 0, 0, st)) return e;                 // x = attn @ v (all heads)
            // mid = src + proj(x_ori): a separate buffer for the first surgery block, in place afterwards
            // (the reference's `x_ori += x_ori_res` mutates the view it stored in all_feats[l-1], :317)
            float* mid = (l == first) ? w.mid : src;
            if (int e = linear(w.o, Lw.out_w, Lw.out_b, src, mid, (int)BN, D, D, 0, st)) return e;
            if (int e = linear(w.o2, Lw.out_w, Lw.out_b, xnew, xnew, (int)BN, D, D, 0, st)) return e;        // x += x_res (:319,329)
            if (int e = layernorm(mid, Lw.ln2_w, Lw.ln2_b, w.h, BN, D, st)) return e;
            if (int e = linear(w.h, Lw.fc_w, Lw.fc_b, nullptr, w.u, (int)BN, 4 * D, D, 1, st)) return e;
            if (int e = linear(w.u, Lw.proj_w, Lw.proj_b, mid, feat_l, (int)BN, D, 4 * D, 0, st)) return e;  // x_ori
        }
    }
    // x[0] = x_ori[0] (:442), ln_post, @ proj (:445-446)
    float* xnew = feats + (int64_t)(first - 1) * BN * D;
    {
        dim3 grid(ceil_div(D, 256), B);
        copy_cls_kernel<<<grid, 256, 0, st>>>(feats + (int64_t)(L - 1) * BN * D, xnew, ND, D);
        if (int e = check_launch("copy_cls_kernel")) return e;
    }
    if (int e = layernorm(xnew, Wt->ln_post_w, Wt->ln_post_b, w.h, BN, D, st)) return e;
    return sgemm2(w.h, Wt->proj, tokens, nullptr, nullptr, (int)BN, E, D, D, E, E, 1, 0, 0, 0, 1, 0, 0, 0, 1.f, 0, 0, st);
}
