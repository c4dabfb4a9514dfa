// Multi-scale + flip merge of the segmentation logits (reference: tools/infer_seg_voc.py:56-88).
//
// Per scale the reference runs model(cat[x, flip(x)])[0] -> segs [2,C,g,g], up-samples both maps bilinearly
// (align_corners=False) to the image size (h,w), takes  segs[0]  at the base scale (:71: the flipped half is dropped there)
// and  (segs[0] + flip_x(segs[1])) / 2  at the others (:80), averages the scales (:83), resizes to the label size and takes
// the argmax (:85-86).  Here one kernel per scale accumulates the up-sampled (and un-flipped) maps straight into the running
// sum -- the [2,C,h,w] up-sampled tensors and the stack are never materialised -- and one kernel does resize + argmax.
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

// torch upsample_bilinear2d, align_corners=False: source coordinate max((dst + 0.5) * in/out - 0.5, 0)
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_of(int dst, int in, int out) {
    const float scale = (float)in / (float)out;
    float f = scale * ((float)dst + 0.5f) - 0.5f;
    f = f < 0.f ? 0.f : f;
    Lerp r;
    r.i0 = (int)f;
    r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
    r.l1 = f - (float)r.i0;
    r.l0 = 1.f - r.l1;
    return r;
}
__device__ __forceinline__ float bilerp(const float* __restrict__ m, int gw, const Lerp& ly, const Lerp& lx) {
    return ly.l0 * (lx.l0 * m[ly.i0 * gw + lx.i0] + lx.l1 * m[ly.i0 * gw + lx.i1]) +
           ly.l1 * (lx.l0 * m[ly.i1 * gw + lx.i0] + lx.l1 * m[ly.i1 * gw + lx.i1]);
}

// acc[c,y,x] = ((first ? 0 : acc) + s) * out_scale,  s = up(seg[0,c])[y,x]  or  (up(seg[0,c])[y,x] + up(seg[1,c])[y,w-1-x]) / 2
__global__ void __launch_bounds__(256)
seg_accumulate_kernel(const float* __restrict__ seg, int C, int gh, int gw, int flip_merge, float* __restrict__ acc, int h, int w,
                      int first, float out_scale) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const Lerp ly = lerp_of(y, gh, h), lx = lerp_of(x, gw, w), lf = lerp_of(w - 1 - x, gw, w);
    const int64_t hw = (int64_t)h * w, o = (int64_t)y * w + x, gg = (int64_t)gh * gw;
    for (int c = 0; c < C; ++c) {
        float s = bilerp(seg + c * gg, gw, ly, lx);
        if (flip_merge) s = (s + bilerp(seg + (C + c) * gg, gw, ly, lf)) / 2.f;
        const float a = first ? s : acc[c * hw + o] + s;
        acc[c * hw + o] = a * out_scale;
    }
}

// labels[Y,X] = argmax_c resize(acc)[c,Y,X]  (first maximum wins, NaN is a maximum: torch.argmax)
__global__ void __launch_bounds__(256)
seg_argmax_kernel(const float* __restrict__ acc, int C, int h, int w, int H, int W, int64_t* __restrict__ labels) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W || y >= H) return;
    const bool same = H == h && W == w;
    const Lerp ly = lerp_of(y, h, H), lx = lerp_of(x, w, W);
    const int64_t hw = (int64_t)h * w;
    float best = 0.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        const float v = same ? acc[c * hw + (int64_t)y * w + x] : bilerp(acc + c * hw, w, ly, lx);
        if (c == 0 || v > best || (v != v && best == best)) { best = v; arg = c; }
    }
    labels[(int64_t)y * W + x] = arg;
}

}  // namespace xl

using namespace xl;

extern "C" int excel_seg_accumulate(const float* seg, int C, int gh, int gw, int flip_merge, float* acc, int h, int w, int first,
                                    float out_scale, void* stream) {
    XL_REQUIRE(seg && acc && C >= 1 && gh >= 1 && gw >= 1 && h >= 1 && w >= 1, "seg_accumulate: bad arguments");
    seg_accumulate_kernel<<<dim3(ceil_div(w, 32), ceil_div(h, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(seg, C, gh, gw, flip_merge, acc,
                                                                                                        h, w, first, out_scale);
    return check_launch("seg_accumulate_kernel");
}

extern "C" int excel_seg_argmax(const float* acc, int C, int h, int w, int H, int W, int64_t* labels, void* stream) {
    XL_REQUIRE(acc && labels && C >= 1 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "seg_argmax: bad arguments");
    seg_argmax_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 8)), dim3(32, 8), 0, (cudaStream_t)stream>>>(acc, C, h, w, H, W, labels);
    return check_launch("seg_argmax_kernel");
}
