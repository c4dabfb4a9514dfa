// Training-side label utilities that consume the pseudo labels (reference: utils/camutils.py:123-143, 438-476;
// SURVEY.md §8 f2).  The reference builds these with Python loops on the host every iteration
// (get_mask_by_radius is O(n_p * r^2) interpreted code, scripts/train_voc.py:208).
#include "common.cuh"
#include "excel_b200.h"

namespace xl {

// get_mask_by_radius (camutils.py:459-476): mask[i,j] = 1 iff |y_i-y_j| <= r and |x_i-x_j| <= r
__global__ void radius_mask_kernel(int h, int w, int radius, float* __restrict__ mask) {
    const int n = h * w;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    const int yi = i / w, xi = i - yi * w, yj = j / w, xj = j - yj * w;
    mask[(int64_t)i * n + j] = (abs(yi - yj) <= radius && abs(xi - xj) <= radius) ? 1.f : 0.f;
}

// cams_to_affinity_label (camutils.py:438-457): nearest-neighbour down-sampling of the label map by the patch
// size, pairwise equality, ignore where either label is ignore_index or the radius mask is 0
__global__ void affinity_label_kernel(const int64_t* __restrict__ label, int H, int W, int gh, int gw,
                                      const float* __restrict__ mask, int64_t ignore, int64_t* __restrict__ out) {
    const int n = gh * gw;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, b = blockIdx.z;
    if (j >= n) return;
    // F.interpolate(mode="nearest"): src = floor(dst * in/out)
    const int yi = (int)floorf((i / gw) * ((float)H / gh)), xi = (int)floorf((i % gw) * ((float)W / gw));
    const int yj = (int)floorf((j / gw) * ((float)H / gh)), xj = (int)floorf((j % gw) * ((float)W / gw));
    const int64_t li = label[((int64_t)b * H + min(yi, H - 1)) * W + min(xi, W - 1)];
    const int64_t lj = label[((int64_t)b * H + min(yj, H - 1)) * W + min(xj, W - 1)];
    int64_t v = li == lj ? 1 : 0;
    if (mask && mask[(int64_t)i * n + j] == 0.f) v = ignore;
    if (li == ignore || lj == ignore) v = ignore;
    out[((int64_t)b * n + i) * n + j] = v;
}

// lam_to_label (camutils.py:123-143, img_box=None): valid_cam = cls * cam; label = argmax_c + 1, thresholded
__global__ void lam_to_label_kernel(const float* __restrict__ cam, const float* __restrict__ cls, int C, int64_t hw,
                                    float bkg_thre, float high_thre, float low_thre, int ignore_mid, int64_t ignore,
                                    float* __restrict__ valid_cam, int64_t* __restrict__ label) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (p >= hw) return;
    float best = 0.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        const float v = cls[b * C + c] * cam[((int64_t)b * C + c) * hw + p];
        valid_cam[((int64_t)b * C + c) * hw + p] = v;
        if (c == 0 || v > best || (v != v && !(best != best))) { best = v; arg = c; }
    }
    int64_t l = arg + 1;
    if (ignore_mid) {
        if (best <= high_thre) l = ignore;
        if (best <= low_thre) l = 0;
    } else if (best <= bkg_thre) {
        l = 0;
    }
    label[(int64_t)b * hw + p] = l;
}

}  // namespace xl

using namespace xl;

extern "C" int excel_radius_mask(int h, int w, int radius, float* mask, void* stream) {
    XL_REQUIRE(h >= 1 && w >= 1 && radius >= 0 && h * w <= 65535, "radius_mask: bad shape");
    const int n = h * w;
    radius_mask_kernel<<<dim3(ceil_div(n, 256), n), 256, 0, (cudaStream_t)stream>>>(h, w, radius, mask);
    return check_launch("radius_mask_kernel");
}

extern "C" int excel_affinity_label(const int64_t* label, int B, int H, int W, int gh, int gw, const float* mask,
                                    int64_t ignore_index, int64_t* out, void* stream) {
    XL_REQUIRE(B >= 0 && H >= 1 && W >= 1 && gh >= 1 && gw >= 1 && gh * gw <= 65535 && B <= 65535, "affinity_label: bad shape");
    if (B == 0) return 0;
    const int n = gh * gw;
    affinity_label_kernel<<<dim3(ceil_div(n, 256), n, B), 256, 0, (cudaStream_t)stream>>>(label, H, W, gh, gw, mask, ignore_index, out);
    return check_launch("affinity_label_kernel");
}

extern "C" int excel_lam_to_label(const float* cam, const float* cls_label, int B, int C, int H, int W, float bkg_thre,
                                  float high_thre, float low_thre, int ignore_mid, int64_t ignore_index, float* valid_cam,
                                  int64_t* label, void* stream) {
    XL_REQUIRE(B >= 0 && C >= 1 && H >= 1 && W >= 1 && B <= 65535, "lam_to_label: bad shape");
    if (B == 0) return 0;
    const int64_t hw = (int64_t)H * W;
    lam_to_label_kernel<<<dim3((unsigned)ceil_div64(hw, 256), B), 256, 0, (cudaStream_t)stream>>>(
        cam, cls_label, C, hw, bkg_thre, high_thre, low_thre, ignore_mid, ignore_index, valid_cam, label);
    return check_launch("lam_to_label_kernel");
}
