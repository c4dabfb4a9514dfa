"""Validate oracle/port.py against the UNMODIFIED reference (build container only).

    python -B oracle/check_port.py [--size 224] [--batch 2]

Runs both on the same seeded inputs and prints max-abs differences / label mismatches
per stage.  Exit code 1 if any stage is outside its gate.  TEST INFRASTRUCTURE ONLY.
"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from oracle import port, ref_harness as H  # noqa: E402
from excel_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--dataset", default="pascal_voc")
    args = ap.parse_args()
    ref = H.load()
    torch.set_grad_enabled(False)
    ok = True

    def gate(name, diff, tol):
        nonlocal ok
        flag = "ok " if diff <= tol else "FAIL"
        ok &= diff <= tol
        print(f"[{flag}] {name:44s} max-abs {diff:.3e} (gate {tol:g})")

    enc = H.build_clip(H.VIT_B16, seed=0)
    model = H.build_model(enc, args.dataset, img_size=args.size, mode="val")
    W = port.export_visual_weights(enc.visual)
    num_fg = model.num_classes - 1
    imgs = synth.images(args.batch, args.size, seed=1)
    cls = synth.class_labels(args.batch, num_fg, seed=2, n_fixed=None)

    # ---- encoder + CAM
    seg, attn_fts, attr_ref, attn_ref, attn_pred = model(imgs)
    tok_ref, attn_ref2, feats_ref = ref.clip.generate_clip_fts(imgs, enc, return_weights=True)
    tok, attn, feats = port.generate_clip_fts(W, imgs)
    gate("generate_clip_fts.image_features", (tok - tok_ref).abs().max().item(), 2e-5)
    gate("generate_clip_fts.attn_weights", (attn - attn_ref2).abs().max().item(), 2e-5)
    for l in range(feats.shape[0]):
        gate(f"generate_clip_fts.all_feats[{l}]", (feats[l] - feats_ref[l]).abs().max().item(), 2e-4)
    text_t = model.text_attr.permute(1, 0).contiguous()
    attr = port.clip_feature_surgery(tok, text_t)[:, 1:, :num_fg]
    gate("clip_feature_surgery (attr_maps_raw)", (attr - attr_ref).abs().max().item(), 5e-5)

    # ---- LVC branch: ExCEL_model.forward(img, ex_feats=attn_fts) -> attr_maps_raw only (model_excel.py:50-53)
    attr_lvc_ref = model(imgs, ex_feats=attn_fts)
    tok_l, attn_l, feats_l = port.generate_clip_fts(W, imgs, ex_feats=attn_fts)
    attr_lvc = port.clip_feature_surgery(tok_l, text_t)[:, 1:, :num_fg]
    gate("LVC: model(img, ex_feats) attr_maps_raw", (attr_lvc - attr_lvc_ref).abs().max().item(), 5e-5)
    tok_lr, attn_lr, feats_lr = ref.clip.generate_clip_fts(imgs, enc, return_weights=True, ex_feats=attn_fts)
    gate("LVC: generate_clip_fts.image_features", (tok_l - tok_lr).abs().max().item(), 2e-5)
    gate("LVC: generate_clip_fts.all_feats", (feats_l - feats_lr).abs().max().item(), 2e-4)
    print(f"       LVC bias changes attr_maps_raw by up to {(attr_lvc_ref - attr_ref).abs().max().item():.3e}")

    # ---- decoder-side inference (f4): SegFormerHead + attn_pred on the reference's own all_feats
    Wd = {k: v.detach() for k, v in model.decoder_fts_fuse.state_dict().items()}
    g_ = args.size // 16
    x_all = feats_ref[:, :, 1:].permute(0, 1, 3, 2).reshape(feats_ref.shape[0], args.batch, feats_ref.shape[-1], g_, g_)
    gate("SegFormerHead.forward (attn_fts)", (port.segformer_head(Wd, x_all) - attn_fts).abs().max().item(), 1e-5)
    gate("attn_pred", (port.attn_pred(attn_fts) - attn_pred).abs().max().item(), 1e-6)

    # ---- attrutils (a10): no live caller in the reference, checked function by function
    gA = torch.Generator().manual_seed(9)
    flag = (torch.rand(20, 112, generator=gA) > 0.8).float()
    amap = torch.rand(2, 49, 112, generator=gA)
    gate("attrutils.attrmap2clsmap", (ref.attrutils.attrmap2clsmap(flag, amap) - port.attrmap2clsmap(flag, amap)).abs().max().item(), 1e-6)
    tf = torch.randn(20, 64, generator=gA)
    bank = torch.randn(64, 112, generator=gA)
    gate("attrutils.attr2cls_embedings (no bg rows)",
         (ref.attrutils.attr2cls_embedings(tf, bank, 20) - port.attr2cls_embedings(tf, bank, 20)).abs().max().item(), 1e-6)

    # ---- SVC + PAR per image, on the REFERENCE's encoder outputs (stage isolation)
    par = ref.PAR(num_iter=20, dilations=list(port.PAR_DILATIONS))
    mism = 0
    for i in range(args.batch):
        r_list, r_cls = ref.affutils.refine_cams_with_aff(attr_ref[i], attn_ref[:, i], cls[i], imgs.shape[2:], caa_thre=0.79)
        p_list, p_cls = port.refine_cams_with_aff(attr_ref[i], attn_ref[:, i], cls[i], imgs.shape[2:], caa_thre=0.79)
        q_list, _ = port.refine_cams_with_aff(attr_ref[i], attn_ref[:, i], cls[i], imgs.shape[2:], caa_thre=0.79, use_cv2=False)
        assert torch.equal(r_cls, p_cls)
        gate(f"refine_cams_with_aff[{i}] (cv2 boxes)", max((a - b).abs().max().item() for a, b in zip(r_list, p_list)), 1e-6)
        gate(f"refine_cams_with_aff[{i}] (cc boxes)", max((a - b).abs().max().item() for a, b in zip(r_list, q_list)), 1e-6)
        r_lab, r_cams = ref.affutils.refine_cams_with_bkg_weclip(r_list, imgs[i], r_cls, par, imgs.shape[-2:])
        p_lab, p_cams, _ = port.refine_cams_with_bkg_weclip(r_list, imgs[i], r_cls, tuple(imgs.shape[-2:]))
        q_lab, q_cams, _ = port.refine_cams_with_bkg_weclip(r_list, imgs[i], r_cls, tuple(imgs.shape[-2:]), use_cv2=False)
        gate(f"refine_cams_with_bkg_weclip[{i}].cams (cv2)", (r_cams - p_cams).abs().max().item(), 1e-6)
        gate(f"refine_cams_with_bkg_weclip[{i}].cams (torch)", (r_cams - q_cams).abs().max().item(), 5e-6)
        m1 = (r_lab != p_lab).sum().item()
        m2 = (r_lab != q_lab).sum().item()
        mism += m1
        print(f"       labels[{i}]: mismatches vs reference: cv2-path {m1}, torch-resize path {m2} of {r_lab.numel()}")
    ok &= mism == 0

    # ---- PAR alone, non-square, with image resize
    g = torch.Generator().manual_seed(5)
    im = torch.rand(2, 3, 40, 56, generator=g)
    mk = torch.softmax(torch.randn(2, 4, 61, 83, generator=g), 1)
    gate("PAR.forward (resize, non-square)", (par(im, mk) - port.par_forward(im, mk)).abs().max().item(), 1e-6)

    # ---- trans mat, box masks on random maps
    A = torch.rand(196, 196, generator=g) + 0.01
    gate("compute_trans_mat", (ref.affutils.compute_trans_mat(A) - port.compute_trans_mat(A)).abs().max().item(), 1e-8)
    rng = np.random.default_rng(0)
    bad = 0
    for t in range(500):
        gsz = int(rng.integers(4, 33))
        cam = rng.random((gsz, gsz)).astype(np.float32)
        if t % 3 == 0:
            cam = (cam > 0.7).astype(np.float32) * cam
        thr = float(rng.choice([0.75, 0.79, 0.88]))
        bad += int(not np.array_equal(port.box_mask_cv2(cam, thr), port.box_mask_cc(cam, thr)))
    print(f"[{'ok ' if bad == 0 else 'FAIL'}] box_mask_cc vs cv2 contours: {bad}/500 mismatching maps")
    ok &= bad == 0
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
