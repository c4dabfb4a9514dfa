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
int64_t excel_launch_count(void);     /* kernels this library has enqueued so far in this process */

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
 *   steps stream it with TMA, whose row stride must be a multiple of 16 B).  (group = 1 keeps one 512^2 image's
 *   50 MB of affinities in the 126 MB L2 across the steps but under-fills the GPU: measured 1.8x slower than
 *   whole-batch launches, which stream the affinities from HBM at 67-95 % of its peak.)
 *   planes_out == NULL or num_iter == 0: affinity only, aff_ws must then hold [B,K,H,Wp].
 *   img_index_dev: NULL, or int32 [B] (device): slot b of this call (its planes are plane_off_dev[b]..) reads image
 *   img_index_dev[b] of `img` -- lets a caller process the images of a batch in another order (e.g. sorted by plane count)
 *   without gathering them. */
int excel_par_forward(const float* img, int64_t stride_b, int64_t stride_c, int64_t stride_y, int B,
                      int hi, int wi, int H, int W, const int* dilations_host, int n_dil, float w1, float w2,
                      int num_iter, int group, float* resize_ws, float* aff_ws, const float* planes_in,
                      float* planes_out, float* planes_tmp, const int* plane_off_dev, int total_planes, int max_c,
                      const int* img_index_dev, void* stream);

/* utils/affutils.py:86-87 (_refine_cams): labels[o(b)] = plane_key[argmax_c planes of slot b]
 * (first maximum wins, NaN is a maximum); labels [B,H,W] int64, plane_key_dev [P] int64; o(b) = out_index_dev[b]
 * (int32 [B], device) or b when NULL. */
int excel_par_labels(const float* planes, const int* plane_off_dev, const int64_t* plane_key_dev,
                     int64_t* labels, int B, int H, int W, const int* out_index_dev, void* stream);

/* ---------------------------------------------------------------- SVC (utils/affutils.py) ------ */

/* utils/affutils.py:180,197 (training-free branch of refine_cams_with_aff): A[b] = mean over the last
 * `attn_layers` of attn[l, b, 1:, 1:].  attn [L,B,N,N] with element strides (stride_l, stride_b, stride_r, 1) -- the
 * encoder returns a row pitch of round_up(N,4) (excel_vit_forward); A [B, N-1, N-1]. */
int excel_svc_mean_attention(const float* attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int L, int B, int N,
                             int attn_layers, float* A, void* stream);

/* utils/affutils.py:182-195 (the seg_attn / LVC branch of refine_cams_with_aff): per image keep the layers l of the
 * last `attn_layers` whose sum(seg_attn - attn_l[1:,1:]) is <= the mean over those layers;
 * A = (sum of kept layers) / (number kept + 1e-5) * seg_attn.  seg_attn [B,N-1,N-1]; diff_ws [B*attn_layers]. */
int excel_svc_seg_attention(const float* attn, int64_t stride_l, int64_t stride_b, int64_t stride_r, int L, int B, int N, int attn_layers,
                            const float* seg_attn, float* diff_ws, float* A, void* stream);

/* utils/affutils.py:11-16 (compute_trans_mat, the 1 + 2 rounds of column / row normalisation) in scaling
 * form: trans = diag(r) A diag(c).  A [B,np,np]; r, c [B,np] outputs. rounds = 3 in the reference. */
int excel_svc_sinkhorn(const float* A, int B, int np, int rounds, float* r, float* c, void* stream);

/* utils/affutils.py:17: T = (S + S^T)/2 with S = diag(r) A diag(c), materialised [B,np,np] (only the
 * standalone compute_trans_mat API needs it; the squaring of :20 is then excel_sgemm). */
int excel_svc_build_trans(const float* A, const float* r, const float* c, int B, int np, float* T, void* stream);

/* utils/affutils.py:26-53 (scoremap2bbox, multi_contour_eval=True) + :207-215 for Q (image, class) pairs:
 * cam = attr[img_of[q], :, cls_of[q]] viewed gh x gw (element strides attr_stride_b, attr_stride_p, 1);
 * uint8(cam*255) truncation, thr = int(caa_thre*max), strict >, union of the clipped bounding boxes of the
 * 8-connected components (== cv2.findContours + boundingRect).  v[q] = mask * cam [Q, gh*gw];
 * mask_out [Q, gh*gw] optional (NULL to skip). */
int excel_svc_box_mask(const float* attr, int64_t attr_stride_b, int64_t attr_stride_p, const int* img_of_dev,
                       const int* cls_of_dev, int Q, int gh, int gw, double caa_thre, float* v, float* mask_out,
                       void* stream);

/* utils/affutils.py:17,20 + :215-221: out[q] = T^hops v[q] with T = (S + S^T)/2, S = diag(r) A diag(c)
 * of image img_of[q]; hops = 2 reproduces ((T@T) * mask) @ cam without forming T@T.  tmp1, tmp2 [Q,np]. */
int excel_svc_propagate(const float* A, const float* r, const float* c, const int* img_of_dev, const float* v, int Q,
                        int np, int hops, float* tmp1, float* tmp2, float* out, void* stream);

/* utils/affutils.py:55-78 (generate_cam_label / scale_cam_image) + :165-166: per class min-max with
 * (1e-7 + max), bilinear resize gh x gw -> H x W with cv2.resize semantics, background plane
 * 1 - max_c; writes the packed PAR planes [P,H,W] (plane_off_dev[b] = background of image b, then its
 * classes; class slot of image b starts at plane_off[b] - b).  minmax_ws [2*Q] floats. */
int excel_svc_cams_to_planes(const float* refined, int Q, int gh, int gw, const int* plane_off_dev, int B, int H,
                             int W, float* minmax_ws, float* planes, void* stream);

/* ---------------------------------------------------------------- seg inference (tools/infer_seg_voc.py) */

/* tools/infer_seg_voc.py:66-83, one scale of the multi-scale + flip loop: seg [2,C,gh,gw] = model(cat[x, flip(x)])[0].
 *   s = bilinear(seg[0] -> h x w)                                  (flip_merge = 0: the base scale, :69-72)
 *   s = (bilinear(seg[0]) + flip_x(bilinear(seg[1]))) / 2          (flip_merge = 1: the other scales, :78-80)
 *   acc [C,h,w] = ((first ? 0 : acc) + s) * out_scale              (out_scale = 1/n_scales on the last call: the mean of :83)
 * bilinear = F.interpolate(mode='bilinear', align_corners=False). */
int excel_seg_accumulate(const float* seg, int C, int gh, int gw, int flip_merge, float* acc, int h, int w, int first,
                         float out_scale, void* stream);

/* tools/infer_seg_voc.py:85-86: labels [H,W] int64 = argmax_c bilinear(acc [C,h,w] -> H x W) (first maximum wins). */
int excel_seg_argmax(const float* acc, int C, int h, int w, int H, int W, int64_t* labels, void* stream);

/* ---------------------------------------------------------------- CAM (clip/clip.py) ---------- */

/* clip/clip.py:353 (generate_clip_fts): out = tok / ||tok||_2 over the TOKEN axis, per (b, channel).
 * tok, out [B,N,E]; norm_ws [B,E]. */
int excel_token_normalize(const float* tok, int B, int N, int E, float* norm_ws, float* out, void* stream);

/* clip/clip.py:288-310 (clip_feature_surgery, redundant_feats=None): feats [B,N,E], text [T,E] ->
 * out [B,N,T], min-max normalised over all N tokens per (b,t), no epsilon.  The similarity GEMM S = feats text^T runs on the
 * tcgen05 engine (split-fp16 operands, pre-scaled on the device by a power of two of their max-abs); the w / redundant-term
 * row epilogue and the column min-max are two coalesced passes.  workspace: excel_cam_workspace_bytes(B,N,E,T) bytes,
 * 256 B-aligned.  T <= 512. */
int64_t excel_cam_workspace_bytes(int B, int N, int E, int T);
int excel_cam_surgery(const float* feats, const float* text, int B, int N, int E, int T, void* workspace,
                      int64_t workspace_bytes, float* out, void* stream);

/* utils/camutils.py:19-26 (cure_attr_map_flip): attr_2b [2B, gh*gw, K] = maps of [x, flip(x)] -> out [B, gh*gw, K]:
 * element-max of the un-flipped pair, minus its per-(b,k) spatial min, divided by (max + 1e-5). */
int excel_flip_merge(const float* attr_2b, int B, int gh, int gw, int K, float* out, void* stream);

/* ---------------------------------------------------------------- ViT (clip/clip_surgery_model.py) */

/* Weights of one ResidualAttentionBlock (clip/clip_surgery_model.py:285-337); all device pointers, fp32,
 * nn.Linear layout [out,in].  in_w/in_b = attn.in_proj (or the surgery Attention's qkv, a clone of it,
 * :396-405); out_w/out_b = attn.out_proj (or Attention.proj). */
typedef struct {
    const float *ln1_w, *ln1_b, *in_w, *in_b, *out_w, *out_b, *ln2_w, *ln2_b, *fc_w, *fc_b, *proj_w, *proj_b;
    /* split-fp16 copies of the four weight matrices (excel_split_f16 at load time): [out, 2*in] halves, hi | lo, of
     * x_scale * W; x_scale is a power of two that puts max|W| at 2^13..2^14 (fp16's normal range for both halves) and is
     * divided out again by the GEMM's alpha */
    const void *in_ws, *out_ws, *fc_ws, *proj_ws;
    float in_scale, out_scale, fc_scale, proj_scale;
} ExcelVitLayer;

/* VisionTransformer (clip/clip_surgery_model.py:374-448).  conv1 [width, 3*patch*patch]; cls [width];
 * pos [1+grid0^2, width]; proj [width, embed]; blocks: HOST array of `layers` entries.  The last
 * n_surgery blocks run the dual-path surgery attention (reload_self_attn(layers=6) -> 5, :399). */
typedef struct {
    int layers, width, heads, patch, embed, grid0, n_surgery;
    const float *conv1, *cls, *pos, *ln_pre_w, *ln_pre_b, *ln_post_w, *ln_post_b, *proj;
    const void *conv1_s, *proj_t_s;   /* split-fp16 conv1 [width, 2*Kp] and proj^T [embed, 2*width], pre-scaled like the blocks' */
    float conv1_scale, proj_t_scale;
    const ExcelVitLayer* blocks;
} ExcelVitWeights;

int64_t excel_vit_workspace_bytes(int B, int S, int patch, int D, int heads);

/* fp32 [rows, cols] (row pitch ldx) -> split fp16 [rows, 2*Kp] (hi | lo, zero padded) of scale * x, Kp % 64 == 0: the
 * operand format of the tcgen05 GEMM engine (x = hi + lo keeps 22 significant bits while both halves are normal fp16
 * numbers: 6.1e-5 <= |lo|, |hi| <= 65504 -- pick `scale` accordingly; hi saturates instead of overflowing to inf). */
int excel_split_f16(const float* x, int64_t ldx, int rows, int cols, int Kp, float scale, void* out, void* stream);

/* VisionTransformer.forward + Transformer.forward (clip/clip_surgery_model.py:418-448, 346-371) as called by
 * clip.generate_clip_fts (clip/clip.py:348-358), for img [B,3,S,S] (element strides b, c, y; x contiguous).
 * Outputs: tokens [B,N,embed] (BEFORE the token-axis normalisation of clip.py:353 -> excel_token_normalize),
 * attn [layers,B,N,attn_row_pitch] with attn_row_pitch = round_up(N,4) floats (the attention kernels write the maps with TMA
 * stores, whose row pitch must be a multiple of 16 B; columns >= N are padding -- callers view [.., :N]),
 * feats [layers,B,N,width] with the reference's view aliasing.
 * lvc_attn: NULL, or the LVC bias ex_attn [B,N-1,N-1] of excel_lvc_attention (Attention.forward with ex_feats,
 * clip/clip_surgery_model.py:127-141): added to every head's patch block of the surgery blocks' new-path attention. */
int excel_vit_forward(const ExcelVitWeights* w, const float* img, int64_t img_stride_b, int64_t img_stride_c,
                      int64_t img_stride_y, int B, int S, float* workspace, int64_t workspace_bytes, float* tokens,
                      float* attn, int64_t attn_row_pitch, float* feats, const float* lvc_attn, void* stream);

/* clip/clip_surgery_model.py:127-137: ex_feats [B,C,np] (decoder features, np = h*w positions) -> ex_attn [B,np,np] =
 * softmax_j of ((cosine similarity - mean over the WHOLE batch * beta) * gamma) with negatives set to -inf.
 * Workspaces: qt_ws [B*np*C] floats, rowsum_ws [B*np] doubles, mean_ws [1] float. */
int excel_lvc_attention(const float* ex_feats, int B, int C, int np, float beta, float gamma, float* qt_ws,
                        double* rowsum_ws, float* mean_ws, float* ex_attn, void* stream);

/* model/model_excel.py:71-76 (attn_pred): feats [B,C,np] -> sigmoid((cosine similarity - batch mean * beta) * gamma)
 * [B,np,np]; same workspaces as excel_lvc_attention. */
int excel_attn_pred(const float* feats, int B, int C, int np, float beta, float gamma, float* qt_ws, double* rowsum_ws,
                    float* mean_ws, float* attn_pred, void* stream);

/* utils/attrutils.py helpers: out[r,:] = softmax(x[r,:]) / x[r,:] / ||x[r,:]||_2 for x [rows, cols]. */
int excel_row_softmax(const float* x, int rows, int cols, float* out, void* stream);
int excel_row_l2_normalize(const float* x, int rows, int cols, float* out, void* stream);

/* ---------------------------------------------------------------- metric (utils/evaluate.py) --- */

/* utils/evaluate.py:9-15 (_fast_hist), accumulated on the device: hist[nc*t + p] += 1 over n pixels with
 * 0 <= t < nc (other labels, e.g. 255 = ignore, are skipped).  hist: int64 [nc*nc], NOT cleared here. */
int excel_confusion_hist(const int64_t* label_true, const int64_t* label_pred, int64_t n, int num_classes,
                         int64_t* hist, void* stream);

/* ---------------------------------------------------------------- label utilities (utils/camutils.py) */

/* utils/camutils.py:459-476 (get_mask_by_radius): mask [h*w, h*w] fp32, 1 where both |dy|,|dx| <= radius. */
int excel_radius_mask(int h, int w, int radius, float* mask, void* stream);

/* utils/camutils.py:438-457 (cams_to_affinity_label): label [B,H,W] int64 -> out [B, gh*gw, gh*gw] int64
 * (1 same label, 0 different, ignore_index where a label is ignore_index or mask[i,j] == 0; mask may be NULL);
 * the label map is down-sampled to gh x gw with nearest-neighbour (F.interpolate) semantics. */
int excel_affinity_label(const int64_t* label, int B, int H, int W, int gh, int gw, const float* mask,
                         int64_t ignore_index, int64_t* out, void* stream);

/* utils/camutils.py:123-143 (lam_to_label, img_box=None): valid_cam = cls*cam [B,C,H,W]; label [B,H,W] int64 =
 * argmax_c + 1 with the bkg / high / low thresholds applied. */
int excel_lam_to_label(const float* cam, const float* cls_label, int B, int C, int H, int W, float bkg_thre,
                       float high_thre, float low_thre, int ignore_mid, int64_t ignore_index, float* valid_cam,
                       int64_t* label, void* stream);

/* ---------------------------------------------------------------- dense fp32 GEMM ------------- */

/* C[b] = act(alpha * A[b] * op(B[b]) + bias) + residual[b]; exact fp32 (SIMT).  A [M,K] (lda); B [N,K] (ldb)
 * if b_is_nk (nn.Linear weight layout) else [K,N]; C, residual [M,N] (ldc); act: 0 none, 1 QuickGELU
 * (clip/clip_surgery_model.py:280-282), 2 ReLU. */
int excel_sgemm(const float* A, const float* B, float* C, const float* bias, const float* residual, int M, int N, int K,
                int64_t lda, int64_t ldb, int64_t ldc, int batch, int64_t strideA, int64_t strideB, int64_t strideC,
                float alpha, int b_is_nk, int act, void* stream);

/* The same contraction on the tcgen05 tensor cores: C = act(alpha * A B^T + bias) + residual for fp32
 * row-major A [M,K] (lda) and B [N,K] (ldb, the nn.Linear weight layout).  Operands are split into fp16
 * hi/lo pairs (fp32-quality products, three MMA passes); ws >= 4*(M+N)*round_up(K,64) bytes. */
int excel_gemm_tc(const float* A, const float* B, float* C, const float* bias, const float* residual, int M, int N, int K,
                  int64_t lda, int64_t ldb, int64_t ldc, float alpha, int act, void* ws, int64_t ws_bytes, void* stream);

/* Batched GEMM on the tcgen05 engine with operands ALREADY in its split-fp16 format (weights split once at load time by
 * excel_split_f16, activations written in split form by a previous call) -- the decoder-side MLPs of
 * model/segformer_head.py:18-26,66-77 at inference:
 *   for z < batch:  D[z] = act(alpha * A[z] B[z]^T + bias[z])        act: 0 none, 1 QuickGELU, 2 ReLU
 *   A[z] = rows z*a_rows_z .. +M of As (row pitch lda halves; hi at column 0, lo at column a_lo_off), M x K
 *   B[z] = rows z*b_rows_z .. +N of Bs (row pitch ldb halves; hi at 0, lo at b_lo_off), N x K  (nn.Linear layout [out, in])
 *   output: fp32 C + z*c_z (row pitch ldc floats)  OR  split fp16 Cs + z*cs_z halves (row pitch lds; lo cs_lo_off after hi)
 *   -- exactly one of C / Cs; bias + z*bias_z ([N] floats each) or NULL.  K % 64 == 0. */
int excel_gemm_tc_split(const void* As, int64_t lda, int a_lo_off, int64_t a_rows_z, const void* Bs, int64_t ldb, int b_lo_off,
                        int64_t b_rows_z, float* C, int64_t ldc, int64_t c_z, void* Cs, int64_t lds, int cs_lo_off, int64_t cs_z,
                        const float* bias, int64_t bias_z, int M, int N, int K, int batch, float alpha, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EXCEL_B200_H */
