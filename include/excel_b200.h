/* excel_b200 -- C ABI of the B200 (sm_100a) CAM -> SVC -> PAR hot path of ExCEL.
 *
 * The reference (zwyang6/ExCEL) has no FFI: its boundary is a Python call surface
 * (SURVEY.md §8b).  Each entry point below replaces the arithmetic of the cited reference
 * function; the Python shims in excel_b200/ keep the reference signatures and call these through
 * ctypes.  Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are fp32, planar,
 *     innermost dimension contiguous; the library never allocates: outputs and workspaces are
 *     caller-owned (PyTorch tensors' data_ptr()), so lifetime follows the caching allocator;
 *   - `stream` is a cudaStream_t (0 = legacy default stream); work is only enqueued, never synced;
 *   - return value 0 = ok; non-zero = error, message via excel_last_error() (thread-local);
 *   - NaN/Inf propagate IEEE-style like the reference (no hidden epsilons).
 */
#ifndef EXCEL_B200_H
#define EXCEL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XL_PAR_MAX_DIL 8

/* library / diagnostics */
const char* excel_last_error(void);
int excel_version(void);              /* 100*major + minor */
int excel_device_arch(int device);    /* 10*major + minor of `device` (100 on B200), <0 on error */

/* ---------------------------------------------------------------- PAR (utils/PAR.py) ---------- */

/* utils/PAR.py:64-92 (PAR.forward) for a batch of B images.
 *   img [B,3,hi,wi], element strides (stride_b, stride_c, stride_y, 1); resized bilinearly with
 *   align_corners=True (:67) to HxW when the sizes differ -- resize_ws [B,3,H,W] is then required.
 *   Affinity (:69-86): K = 8*n_dil planes per image,
 *     aff_k = softmax_k(-mean_c((|I_k-I_0|/(std_c+1e-8)/w1)^2)) + w2*softmax_k(-(pos_k/(std(pos)+1e-8)/w1)^2),
 *     neighbour k = dilation-major, taps in the order of get_kernel (:10-24), replicate padding.
 *   Propagation (:88-90): num_iter steps over the packed mask planes [P,H,W]; plane_off_dev [B+1]
 *     (int32, device) gives image b's planes, total_planes = P, max_c = max planes of one image.  The result lands in
 *     planes_out; planes_tmp [P,H,W] is the ping-pong buffer (needed when num_iter > 1).
 *   Images are processed in launch groups of `group` (<=0: all B): affinity of the group, then all
 *   its steps, so aff_ws only needs [group,K,H,Wp] floats, Wp = round_up(W,4) (internal layout: the
 *   steps stream it with TMA, whose row stride must be a multiple of 16 B), and for one 512^2 image
 *   (50 MB) it stays in the 126 MB L2 across the steps.
 *   planes_out == NULL or num_iter == 0: affinity only, aff_ws must then hold [B,K,H,Wp]. */
int excel_par_forward(const float* img, int64_t stride_b, int64_t stride_c, int64_t stride_y, int B,
                      int hi, int wi, int H, int W, const int* dilations_host, int n_dil, float w1, float w2,
                      int num_iter, int group, float* resize_ws, float* aff_ws, const float* planes_in,
                      float* planes_out, float* planes_tmp, const int* plane_off_dev, int total_planes, int max_c,
                      void* stream);

/* utils/affutils.py:86-87 (_refine_cams): labels[b] = plane_key[argmax_c planes of image b]
 * (first maximum wins, NaN is a maximum); labels [B,H,W] int64, plane_key_dev [P] int64. */
int excel_par_labels(const float* planes, const int* plane_off_dev, const int64_t* plane_key_dev,
                     int64_t* labels, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EXCEL_B200_H */
