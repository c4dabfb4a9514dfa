// SVC -- attention-affinity propagation of the CAMs (reference: utils/affutils.py) for a batch of images.
//
// The reference builds, per image, T = compute_trans_mat(mean of the last 6 attention maps) with three
// Sinkhorn rounds, a symmetrisation and one dense squaring (affutils.py:8-24), then per present class
// refined = ((T@T) * box_mask[None,:]) @ cam (affutils.py:206-221).  Restated here without ever forming
// T or T@T:
//   * Sinkhorn: T_sink = diag(r) A diag(c) with c = 1/(A^T r), r = 1/(A c), starting from r = 1 and
//     repeated 3 times -- exactly the column-then-row normalisations of affutils.py:11-16;
//   * T = (T_sink + T_sink^T)/2 is applied as T v = (r*(A (c*v)) + c*(A^T (r*v)))/2;
//   * (T@T * mask) @ cam = T (T (mask*cam)): two applications of T per class (associativity), i.e.
//     mat-vec passes over the 4 MB matrix A (L2-resident) instead of a 2*n_p^3 GEMM per image.
// The only HBM-heavy step is the 6-layer mean (reads 6 * n_p^2 floats per image).
//
// Box mask (affutils.py:26-53,209-212): uint8(cam*255) truncation, thr = int(caa*max), strict >, the
// union of the bounding boxes of the 8-connected components, box end clipped to size-1 and used as an
// exclusive bound -- identical to cv2.findContours + boundingRect (oracle/check_port.py: 0 mismatches).
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

// ---- A[b] = mean_l attn[l0+l, b, 1:, 1:]  (affutils.py:180,197) ----------------------------------
// attn: [L,B,N,N]; A: [B,n_p,n_p], n_p = N-1.  One thread per output element, x fastest.
__global__ void svc_mean_kernel(const float* __restrict__ attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int N,
                                int l0, int nl, float* __restrict__ A) {
    pdl_trigger();
    pdl_wait();
    const int np = N - 1;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= np) return;
    const float* p = attn + (int64_t)l0 * stride_l + (int64_t)b * stride_b + (int64_t)(i + 1) * stride_r + (j + 1);
    float s = 0.f;
    for (int l = 0; l < nl; ++l) s += __ldcs(p + (int64_t)l * stride_l);
    A[((int64_t)b * np + i) * np + j] = s / (float)nl;
}

// ---- seg_attn branch (affutils.py:182-195): keep the layers whose sum(seg_attn - A_l) is <= the mean over layers ----
// d[b,l] = sum_ij (seg[b,i,j] - attn[l0+l, b, 1+i, 1+j]); one block per (l, b), fixed summation order
__global__ void __launch_bounds__(1024)
svc_layer_diff_kernel(const float* __restrict__ attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int N, int l0,
                      const float* __restrict__ seg, float* __restrict__ d) {
    __shared__ float red[32];
    const int np = N - 1, l = blockIdx.x, b = blockIdx.y, nl = gridDim.x;
    const float* a = attn + (int64_t)(l0 + l) * stride_l + (int64_t)b * stride_b;
    const float* s = seg + (int64_t)b * np * np;
    float acc = 0.f;
    for (int i = threadIdx.x >> 5; i < np; i += 32) {
        float r = 0.f;
        for (int j = threadIdx.x & 31; j < np; j += 32) r += s[(int64_t)i * np + j] - a[(int64_t)(i + 1) * stride_r + j + 1];
        acc += r;
    }
    acc = block_reduce(acc, red, OpSum(), 0.f);
    if (threadIdx.x == 0) d[b * nl + l] = acc;
}

// A[b] = (sum_l keep_l A_l) / (sum_l keep_l + 1e-5) * seg[b],  keep_l = d[b,l] <= mean_l d[b,:]
__global__ void svc_seg_mean_kernel(const float* __restrict__ attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int N, int l0, int nl,
                                    const float* __restrict__ seg, const float* __restrict__ d, float* __restrict__ A) {
    const int np = N - 1;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= np) return;
    float mean = 0.f;
    for (int l = 0; l < nl; ++l) mean += d[b * nl + l];
    mean /= (float)nl;
    const float* p = attn + (int64_t)l0 * stride_l + (int64_t)b * stride_b + (int64_t)(i + 1) * stride_r + (j + 1);
    float s = 0.f, cnt = 0.f;
    for (int l = 0; l < nl; ++l)
        if (d[b * nl + l] <= mean) { s += __ldcs(p + (int64_t)l * stride_l); cnt += 1.f; }
    const int64_t o = ((int64_t)b * np + i) * np + j;
    A[o] = s / (cnt + 1e-5f) * seg[o];
}

// ---- batched mat-vec passes over A -----------------------------------------------------------------
// Vector slot q belongs to image img_of[q] (nullptr: q itself).  All vectors are [Q, n_p].
//   row pass: y_i = so_i * sum_j A_ij (si_j x_j)       col pass: y_j = so_j * sum_i A_ij (si_i x_i)
// x == nullptr means x = 1; si/so == nullptr mean 1 (both indexed by IMAGE: [B, n_p]).
// mode 0: y = value; 1: y = 1/value (Sinkhorn); 2: y = 0.5*(add[q] + value) (second half of T v).
__global__ void __launch_bounds__(256)
svc_rowpass_kernel(const float* __restrict__ A, const int* __restrict__ img_of, const float* __restrict__ x,
                   const float* __restrict__ si, const float* __restrict__ so, const float* __restrict__ add,
                   float* __restrict__ y, int np, int mode) {
    pdl_trigger();
    pdl_wait();
    const int q = blockIdx.y, b = img_of ? img_of[q] : q;
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= np) return;
    const float* row = A + ((int64_t)b * np + i) * np;
    const float* xq = x ? x + (int64_t)q * np : nullptr;
    const float* sib = si ? si + (int64_t)b * np : nullptr;
    float s = 0.f;
    if ((np & 3) == 0 && ((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(xq) | reinterpret_cast<uintptr_t>(sib)) & 15) == 0) {
        // 16 B loads, every load of the row in flight at once (n_p = 1024: eight per lane); lane-local order j, j+1, j+2, j+3
        const float4* r4 = reinterpret_cast<const float4*>(row);
        const float4* x4 = reinterpret_cast<const float4*>(xq);
        const float4* s4 = reinterpret_cast<const float4*>(sib);
#pragma unroll 8
        for (int j = lane; j < np / 4; j += 32) {
            const float4 a = __ldg(r4 + j);
            float4 v = x4 ? x4[j] : make_float4(1.f, 1.f, 1.f, 1.f);
            if (s4) { const float4 w = s4[j]; v.x *= w.x; v.y *= w.y; v.z *= w.z; v.w *= w.w; }
            s = fmaf(a.x, v.x, s); s = fmaf(a.y, v.y, s); s = fmaf(a.z, v.z, s); s = fmaf(a.w, v.w, s);
        }
    } else {
        for (int j = lane; j < np; j += 32) {
            float v = xq ? xq[j] : 1.f;
            if (sib) v *= sib[j];
            s = fmaf(__ldg(row + j), v, s);
        }
    }
    s = warp_sum(s);
    if (lane == 0) {
        if (so) s *= so[(int64_t)b * np + i];
        if (mode == 1) s = 1.f / s;
        else if (mode == 2) s = 0.5f * (add[(int64_t)q * np + i] + s);
        y[(int64_t)q * np + i] = s;
    }
}

// 32 columns per block; 32 row-lanes stride over the rows, reduced through shared memory (deterministic).
__global__ void __launch_bounds__(1024)
svc_colpass_kernel(const float* __restrict__ A, const int* __restrict__ img_of, const float* __restrict__ x,
                   const float* __restrict__ si, const float* __restrict__ so, const float* __restrict__ add,
                   float* __restrict__ y, int np, int mode) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32][33];
    const int q = blockIdx.y, b = img_of ? img_of[q] : q;
    const int cx = threadIdx.x, ry = threadIdx.y;
    const int j = blockIdx.x * 32 + cx;
    const float* Ab = A + (int64_t)b * np * np;
    const float* xq = x ? x + (int64_t)q * np : nullptr;
    const float* sib = si ? si + (int64_t)b * np : nullptr;
    float s = 0.f;
    if (j < np) {
#pragma unroll 8
        for (int i = ry; i < np; i += 32) {
            float v = xq ? xq[i] : 1.f;
            if (sib) v *= sib[i];
            s = fmaf(__ldg(Ab + (int64_t)i * np + j), v, s);
        }
    }
    red[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && j < np) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) t += red[r][cx];
        if (so) t *= so[(int64_t)b * np + j];
        if (mode == 1) t = 1.f / t;
        else if (mode == 2) t = 0.5f * (add[(int64_t)q * np + j] + t);
        y[(int64_t)q * np + j] = t;
    }
}

// ---- box mask + masked CAM vector --------------------------------------------------------------------
// One block per vector slot q = (image, present class).  cam = attr[img, :, cls] as a gh x gw map.
// Writes v[q] = mask * cam (the vector T is applied to) and, optionally, the mask itself.
__global__ void __launch_bounds__(1024)
svc_boxmask_kernel(const float* __restrict__ attr, int64_t attr_stride_b, int64_t attr_stride_p, const int* __restrict__ img_of,
                   const int* __restrict__ cls_of, int gh, int gw, double caa_thre, float* __restrict__ v,
                   float* __restrict__ mask_out) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ int sh[];
    const int n = gh * gw;
    int* label = sh;                 // [n] component label (min cell index) or -1
    int* bx0 = sh + n;               // per-root bounding box
    int* by0 = bx0 + n;
    int* bx1 = by0 + n;
    int* by1 = bx1 + n;
    unsigned char* msk = reinterpret_cast<unsigned char*>(by1 + n);  // [n]
    __shared__ int s_max, s_changed;
    const int q = blockIdx.x, tid = threadIdx.x;
    const float* cam = attr + (int64_t)img_of[q] * attr_stride_b + cls_of[q];
    if (tid == 0) s_max = 0;
    __syncthreads();
    // uint8(cam*255): truncation toward zero like numpy's astype(np.uint8) (affutils.py:28)
    int lmax = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        const int u = (int)(cam[(int64_t)i * attr_stride_p] * 255.f) & 255;
        label[i] = u;  // stash the 8-bit image
        lmax = max(lmax, u);
    }
    atomicMax(&s_max, lmax);
    __syncthreads();
    const int thr = (int)(caa_thre * (double)s_max);  // int(threshold * np.max(img)), affutils.py:31
    for (int i = tid; i < n; i += blockDim.x) {
        label[i] = label[i] > thr ? i : -1;
        bx0[i] = gw; by0[i] = gh; bx1[i] = -1; by1[i] = -1;
        msk[i] = 0;
    }
    __syncthreads();
    // 8-connected components by iterated min-label propagation (grids are <= 64x64 cells)
    for (;;) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < n; i += blockDim.x) {
            int l = label[i];
            if (l < 0) continue;
            const int yy = i / gw, xx = i - yy * gw;
            int m = l;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int y2 = yy + dy, x2 = xx + dx;
                    if (y2 < 0 || y2 >= gh || x2 < 0 || x2 >= gw) continue;
                    const int l2 = label[y2 * gw + x2];
                    if (l2 >= 0 && l2 < m) m = l2;
                }
            if (m < l) {
                atomicMin(&label[i], m);
                s_changed = 1;
            }
        }
        __syncthreads();
        const int ch = s_changed;
        __syncthreads();
        if (!ch) break;
    }
    for (int i = tid; i < n; i += blockDim.x) {
        const int l = label[i];
        if (l < 0) continue;
        const int yy = i / gw, xx = i - yy * gw;
        atomicMin(&bx0[l], xx); atomicMin(&by0[l], yy);
        atomicMax(&bx1[l], xx); atomicMax(&by1[l], yy);
    }
    __syncthreads();
    // boundingRect gives [x, x+w) ; the reference clips the end to size-1 and slices [y0:y1, x0:x1)
    for (int i = tid; i < n; i += blockDim.x) {
        if (label[i] != i) continue;  // roots paint their box
        const int xe = min(bx1[i] + 1, gw - 1), ye = min(by1[i] + 1, gh - 1);
        for (int yy = by0[i]; yy < ye; ++yy)
            for (int xx = bx0[i]; xx < xe; ++xx) msk[yy * gw + xx] = 1;
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const float m = msk[i] ? 1.f : 0.f;
        v[(int64_t)q * n + i] = m * cam[(int64_t)i * attr_stride_p];
        if (mask_out) mask_out[(int64_t)q * n + i] = m;
    }
}

// ---- per-class min-max, bilinear up-sampling, background channel (affutils.py:55-78,161-166) ----------
__global__ void __launch_bounds__(256)
svc_minmax_kernel(const float* __restrict__ x, int n, float* __restrict__ mn, float* __restrict__ mx) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    const float* p = x + (int64_t)blockIdx.x * n;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = p[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    lo = block_reduce(lo, red, OpMin(), INFINITY);
    hi = block_reduce(hi, red, OpMax(), -INFINITY);
    if (threadIdx.x == 0) {
        mn[blockIdx.x] = lo;
        mx[blockIdx.x] = hi - lo;  // max of (x - min): scale_cam_image divides by 1e-7 + max(img - min)
    }
}

// planes[plane_off[b] + 0] = 1 - max_c cams, planes[plane_off[b] + 1 + c] = resize(minmax(refined[q0_b + c]))
// cv2.resize INTER_LINEAR semantics for float images (half-pixel centres, edge clamp).
__global__ void __launch_bounds__(256)
svc_upsample_bg_kernel(const float* __restrict__ refined, const float* __restrict__ mn, const float* __restrict__ rng,
                       const int* __restrict__ plane_off, int gh, int gw, int H, int W, float* __restrict__ planes) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
    if (x >= W || y >= H) return;
    const int p0 = plane_off[b], nc = plane_off[b + 1] - p0 - 1;  // plane p0 is the background
    const int q0 = p0 - b;                                         // vector slots: one fewer per preceding image
    float fx = (float)(((double)x + 0.5) * ((double)gw / (double)W) - 0.5);
    float fy = (float)(((double)y + 0.5) * ((double)gh / (double)H) - 0.5);
    int sx = (int)floorf(fx), sy = (int)floorf(fy);
    fx -= sx; fy -= sy;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= gw - 1) { fx = 0.f; sx = gw - 1; }
    if (sy < 0) { fy = 0.f; sy = 0; }
    if (sy >= gh - 1) { fy = 0.f; sy = gh - 1; }
    const int sx1 = min(sx + 1, gw - 1), sy1 = min(sy + 1, gh - 1);
    const int64_t hw = (int64_t)H * W, o = (int64_t)y * W + x;
    float best = -INFINITY;
    for (int c = 0; c < nc; ++c) {
        const float* m = refined + (int64_t)(q0 + c) * gh * gw;
        const float lo = mn[q0 + c], inv = 1e-7f + rng[q0 + c];
        const float a00 = (m[sy * gw + sx] - lo) / inv, a01 = (m[sy * gw + sx1] - lo) / inv;
        const float a10 = (m[sy1 * gw + sx] - lo) / inv, a11 = (m[sy1 * gw + sx1] - lo) / inv;
        const float r0 = a00 * (1.f - fx) + a01 * fx, r1 = a10 * (1.f - fx) + a11 * fx;
        const float v = r0 * (1.f - fy) + r1 * fy;
        planes[(int64_t)(p0 + 1 + c) * hw + o] = v;
        best = fmaxf(best, v);
    }
    planes[(int64_t)p0 * hw + o] = 1.f - best;
}

// T = (S + S^T)/2, S = diag(r) A diag(c)  (affutils.py:17) -- only for the standalone compute_trans_mat API
__global__ void svc_build_trans_kernel(const float* __restrict__ A, const float* __restrict__ r, const float* __restrict__ c,
                                       float* __restrict__ T, int np) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= np) return;
    const float* Ab = A + (int64_t)b * np * np;
    const float* rb = r + (int64_t)b * np;
    const float* cb = c + (int64_t)b * np;
    const float sij = rb[i] * Ab[(int64_t)i * np + j] * cb[j], sji = rb[j] * Ab[(int64_t)j * np + i] * cb[i];
    T[((int64_t)b * np + i) * np + j] = (sij + sji) / 2.f;
}

}  // namespace xl

using namespace xl;

extern "C" int excel_svc_build_trans(const float* A, const float* r, const float* c, int B, int np, float* T, void* stream) {
    XL_REQUIRE(B >= 0 && np >= 1 && np <= 65535 && B <= 65535, "svc_build_trans: bad shape");
    if (B == 0) return 0;
    dim3 grid(ceil_div(np, 256), np, B);
    svc_build_trans_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, r, c, T, np);
    return check_launch("svc_build_trans_kernel");
}

extern "C" int excel_svc_mean_attention(const float* attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int L, int B, int N,
                                        int attn_layers, float* A, void* stream) {
    XL_REQUIRE(L >= 1 && B >= 0 && N >= 2 && attn_layers >= 1 && stride_r >= N, "svc_mean_attention: bad shape L=%d B=%d N=%d", L, B, N);
    if (B == 0) return 0;
    const int nl = attn_layers < L ? attn_layers : L, np = N - 1;
    XL_REQUIRE(np <= 65535 && B <= 65535, "svc_mean_attention: grid too large");
    dim3 grid(ceil_div(np, 256), np, B);
    XL_CUDA(launch_pdl(svc_mean_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, attn, stride_l, stride_b, stride_r, N, L - nl, nl, A));
    return check_launch("svc_mean_kernel");
}

extern "C" int excel_svc_seg_attention(const float* attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int L, int B, int N,
                                       int attn_layers, const float* seg_attn, float* diff_ws, float* A, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(L >= 1 && B >= 0 && N >= 2 && attn_layers >= 1 && seg_attn && diff_ws && stride_r >= N, "svc_seg_attention: bad arguments");
    if (B == 0) return 0;
    const int nl = attn_layers < L ? attn_layers : L, np = N - 1;
    XL_REQUIRE(np <= 65535 && B <= 65535, "svc_seg_attention: grid too large");
    svc_layer_diff_kernel<<<dim3(nl, B), 1024, 0, st>>>(attn, stride_l, stride_b, stride_r, N, L - nl, seg_attn, diff_ws);
    if (int e = check_launch("svc_layer_diff_kernel")) return e;
    dim3 grid(ceil_div(np, 256), np, B);
    svc_seg_mean_kernel<<<grid, 256, 0, st>>>(attn, stride_l, stride_b, stride_r, N, L - nl, nl, seg_attn, diff_ws, A);
    return check_launch("svc_seg_mean_kernel");
}

static int rowpass(const float* A, const int* img_of, const float* x, const float* si, const float* so, const float* add,
                   float* y, int Q, int np, int mode, cudaStream_t st) {
    dim3 grid(ceil_div(np, 8), Q);
    XL_CUDA(launch_pdl(svc_rowpass_kernel, dim3(grid), dim3(256), 0, st, A, img_of, x, si, so, add, y, np, mode));
    return check_launch("svc_rowpass_kernel");
}
static int colpass(const float* A, const int* img_of, const float* x, const float* si, const float* so, const float* add,
                   float* y, int Q, int np, int mode, cudaStream_t st) {
    dim3 grid(ceil_div(np, 32), Q), block(32, 32);
    XL_CUDA(launch_pdl(svc_colpass_kernel, dim3(grid), dim3(block), 0, st, A, img_of, x, si, so, add, y, np, mode));
    return check_launch("svc_colpass_kernel");
}

extern "C" int excel_svc_sinkhorn(const float* A, int B, int np, int rounds, float* r, float* c, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(B >= 0 && np >= 1 && rounds >= 1 && B <= 65535, "svc_sinkhorn: bad arguments");
    if (B == 0) return 0;
    for (int it = 0; it < rounds; ++it) {
        if (int e = colpass(A, nullptr, nullptr, it == 0 ? nullptr : r, nullptr, nullptr, c, B, np, 1, st)) return e;  // c = 1/(A^T r)
        if (int e = rowpass(A, nullptr, nullptr, c, nullptr, nullptr, r, B, np, 1, st)) return e;                      // r = 1/(A c)
    }
    return 0;
}

extern "C" int excel_svc_box_mask(const float* attr, int64_t attr_stride_b, int64_t attr_stride_p, const int* img_of_dev,
                                  const int* cls_of_dev, int Q, int gh, int gw, double caa_thre, float* v, float* mask_out,
                                  void* stream) {
    XL_REQUIRE(Q >= 0 && gh >= 1 && gw >= 1, "svc_box_mask: bad shape");
    if (Q == 0) return 0;
    const int n = gh * gw;
    const size_t smem = (size_t)n * (5 * sizeof(int) + 1);
    XL_REQUIRE(smem <= 200 * 1024, "svc_box_mask: grid %dx%d too large for the shared-memory labelling", gh, gw);
    XL_CUDA(cudaFuncSetAttribute(svc_boxmask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = n >= 1024 ? 1024 : ((n + 31) / 32) * 32;
    XL_CUDA(launch_pdl(svc_boxmask_kernel, dim3(Q), dim3(threads), smem, (cudaStream_t)stream, attr, attr_stride_b, attr_stride_p, img_of_dev,
                                                                   cls_of_dev, gh, gw, caa_thre, v, mask_out));
    return check_launch("svc_boxmask_kernel");
}

extern "C" int excel_svc_propagate(const float* A, const float* r, const float* c, const int* img_of_dev, const float* v,
                                   int Q, int np, int hops, float* tmp1, float* tmp2, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(Q >= 0 && np >= 1 && hops >= 1 && Q <= 65535, "svc_propagate: bad arguments");
    if (Q == 0) return 0;
    // x <- T x, `hops` times (hops = 2 is (T@T) v); T x = (r*(A(c*x)) + c*(A^T(r*x)))/2
    const float* x = v;
    for (int h = 0; h < hops; ++h) {
        float* dst = (h == hops - 1) ? out : tmp2;
        if (int e = rowpass(A, img_of_dev, x, c, r, nullptr, tmp1, Q, np, 0, st)) return e;
        if (int e = colpass(A, img_of_dev, x, r, c, tmp1, dst, Q, np, 2, st)) return e;
        x = dst;
        if (h + 1 < hops && h + 2 < hops) {  // more than one intermediate: ping-pong tmp2 with out is not needed for hops<=2
            XL_REQUIRE(false, "svc_propagate: hops > 2 not supported");
        }
    }
    return 0;
}

extern "C" int excel_svc_cams_to_planes(const float* refined, int Q, int gh, int gw, const int* plane_off_dev, int B,
                                        int H, int W, float* minmax_ws, float* planes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    XL_REQUIRE(Q >= 0 && B >= 0 && gh >= 1 && gw >= 1 && H >= 1 && W >= 1 && B <= 65535, "svc_cams_to_planes: bad shape");
    if (B == 0) return 0;
    if (Q > 0) {
        XL_CUDA(launch_pdl(svc_minmax_kernel, dim3(Q), dim3(256), 0, st, refined, gh * gw, minmax_ws, minmax_ws + Q));
        if (int e = check_launch("svc_minmax_kernel")) return e;
    }
    dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B), block(32, 8);
    XL_CUDA(launch_pdl(svc_upsample_bg_kernel, dim3(grid), dim3(block), 0, st, refined, minmax_ws, minmax_ws + Q, plane_off_dev, gh, gw, H, W, planes));
    return check_launch("svc_upsample_bg_kernel");
}
