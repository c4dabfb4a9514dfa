"""The ONE label gate of the parity checks (north_star: argmax pseudo-labels bit-exact).  TEST INFRASTRUCTURE
(like everything under oracle/): imported by tests/, __graft_entry__.smoke() and bench.py's parity record only.

The GPU labels are compared pixel by pixel with the oracle's.  A mismatch is tolerated only where the oracle itself
cannot decide: its two best PAR planes at that pixel differ by a relative margin <= MARGIN (the fp32 summation order of
48 neighbours x 20 steps differs between torch-CPU and the kernel).  Everything else is a HARD mismatch and must be 0.

Stage-isolated comparisons (GPU stage fed with the oracle's inputs) use `plane_err=0`.  End-to-end comparisons, where
the GPU's PAR INPUT planes already differ from the oracle's by a measured max-abs `plane_err` (<= 1e-3, the CAM
tolerance), widen the margin by what that input difference can move a PAR output: PAR (utils/PAR.py:88-90) is a linear
map whose rows sum to 1.01 per step, so 2 * plane_err * 1.01^iters relative to the winning plane.
"""
import torch

MARGIN = 1e-5


def label_parity(ref_planes, lab_ref, lab_gpu, plane_err=0.0, iters=20):
    """ref_planes [C,H,W] (oracle PAR output), lab_ref / lab_gpu [H,W] -> (hard, total) mismatching pixels."""
    lab_ref, lab_gpu = torch.as_tensor(lab_ref), torch.as_tensor(lab_gpu)
    bad = lab_ref != lab_gpu
    if ref_planes.shape[0] < 2:
        return int(bad.sum()), int(bad.sum())
    top2 = ref_planes.topk(2, dim=0).values
    margin = (top2[0] - top2[1]) / top2[0].abs().clamp_min(1e-30)
    tol = MARGIN + 2.0 * plane_err * (1.01 ** iters) / top2[0].abs().clamp_min(1e-30)
    return int((bad & (margin > tol)).sum()), int(bad.sum())


def svc_self_noise(attr_map, attn_weights, cls_label, size, caa_thre=0.79):
    """How well is the REFERENCE itself defined on these inputs?  Max-abs difference of the PAR input planes (SVC ->
    per-class min-max -> bilinear up-sampling, utils/affutils.py:177-223,55-78) between the oracle's fp32 arithmetic and
    the same algorithm carried in float64.  With nearly uniform attention (random-init weights) the refined maps are nearly
    constant and the reference's per-class min-max amplifies its own fp32 rounding noise: this number is the floor below
    which no fp32 implementation can agree with another."""
    from . import port
    h, w = size
    gh, gw = h // 16, w // 16
    T32 = port.compute_trans_mat(port.svc_attention(attn_weights)).float()
    T64 = port.compute_trans_mat(port.svc_attention(attn_weights.double()))
    worst = 0.0
    for c in torch.where(cls_label)[0]:
        cam = attr_map[:, c].float()
        mask = torch.from_numpy(port.box_mask_cv2(cam.numpy().reshape(gh, gw), caa_thre)).reshape(1, -1)
        r32 = ((T32 * mask) @ cam.reshape(-1, 1)).reshape(gh, gw)
        r64 = ((T64 * mask.double()) @ cam.double().reshape(-1, 1)).reshape(gh, gw)
        worst = max(worst, (port.scale_cam(r32, (h, w)).double() - port.scale_cam(r64, (h, w))).abs().max().item())
    return worst
