"""CPU: the oracle port (oracle/port.py) replayed against fixtures frozen from the UNMODIFIED
reference (oracle/make_golden.py).  This is what pins the oracle (the reference
ships no golden vectors of its own)."""
import numpy as np
import torch

from excel_b200 import synth
from oracle import port
from oracle.make_golden_cfg import TINY, checksum

t = torch.from_numpy


def test_par_golden(golden):
    G = golden("par")
    im_a = synth.images(2, 48, seed=11)
    assert abs(checksum(im_a) - float(G["chk_a"])) < 1e-3
    out = port.par_forward(im_a, t(G["mk_a"]), num_iter=20)
    assert (out - t(G["out_a"])).abs().max() < 1e-6
    assert torch.equal(out.argmax(1), t(G["out_a"]).argmax(1))
    out = port.par_forward(t(G["im_b"]), t(G["mk_b"]), num_iter=3)   # image-resize branch, non-square, C=5
    assert (out - t(G["out_b"])).abs().max() < 1e-6


def test_par_rows_sum_to_1p01():
    aff = port.par_affinity(synth.images(1, 32, seed=0), (32, 32))
    assert torch.allclose(aff.sum(1), torch.full((1, 32, 32), 1.01), atol=1e-5)   # utils/PAR.py:86


def test_svc_golden(golden):
    G = golden("svc")
    A = t(G["A"])
    T = port.compute_trans_mat(A[:, 1:, 1:].mean(0))
    assert (T - t(G["T"])).abs().max() < 1e-8
    for cam, mask, thr in zip(G["maps"], G["masks"], G["thrs"]):
        assert np.array_equal(port.box_mask_cc(cam, float(thr)), mask)
        assert np.array_equal(port.box_mask_cv2(cam, float(thr)), mask)
    for use_cv2 in (True, False):
        lst, cl = port.refine_cams_with_aff(t(G["attr"]), A, t(G["cls"]), (128, 128), caa_thre=0.79, use_cv2=use_cv2)
        assert np.array_equal(cl.numpy(), G["cls_lst"])
        assert (torch.stack(lst) - t(G["refined"])).abs().max() < 1e-7
        lst_s, _ = port.refine_cams_with_aff(t(G["attr"]), A, t(G["cls"]), (128, 128), caa_thre=0.75,
                                             seg_attn=t(G["seg_attn"]), use_cv2=use_cv2)
        assert (torch.stack(lst_s) - t(G["refined_seg"])).abs().max() < 1e-7
    img = synth.images(1, 128, seed=12)[0]
    assert abs(checksum(img) - float(G["chk_img"])) < 1e-3
    lab, cams, _ = port.refine_cams_with_bkg_weclip(list(t(G["refined"])), img, t(G["cls_lst"]), (96, 112))
    assert (cams - t(G["cams"])).abs().max() < 1e-6
    assert np.array_equal(lab.numpy().astype(np.int16), G["labels"])
    lab2, cams2, _ = port.refine_cams_with_bkg_weclip(list(t(G["refined"])), img, t(G["cls_lst"]), (96, 112), use_cv2=False)
    assert (cams2 - t(G["cams"])).abs().max() < 5e-6
    assert (lab2.numpy() != G["labels"]).mean() < 1e-4


def test_cam_golden(golden):
    G = golden("cam")
    cam = port.clip_feature_surgery(t(G["F"]), t(G["T"]))
    assert (cam - t(G["cam"])).abs().max() < 5e-6
    assert (port.cure_attr_map_flip_post(t(G["lam2b"]), 6) - t(G["merged"])).abs().max() < 1e-6


def test_vit_tiny_golden(golden):
    G = golden("vit_tiny")
    W = port.random_visual_weights(seed=3, **TINY)
    imgs = synth.images(2, 96, seed=13)
    text = synth.text_bank(45, TINY["embed"], seed=6)
    assert abs(checksum(*[v for k, v in W.items() if k != "meta"]) - float(G["chk_w"])) < 1e-2
    assert abs(checksum(imgs) - float(G["chk_img"])) < 1e-3 and abs(checksum(text) - float(G["chk_text"])) < 1e-4
    tok, attn, feats = port.generate_clip_fts(W, imgs)
    assert (tok - t(G["tok"])).abs().max() < 1e-5
    assert (attn - t(G["attn"])).abs().max() < 1e-5
    assert (feats - t(G["feats"])).abs().max() < 1e-4          # incl. the aliasing quirk of rows 6..10 (here 1..5)
    attr = port.clip_feature_surgery(tok, text)[:, 1:, :20]
    assert (attr - t(G["attr_maps"])).abs().max() < 1e-4
    # stage-isolated tail: reference attr/attn in -> identical labels out
    cls2 = synth.class_labels(2, 20, seed=14, n_fixed=2)
    for i in range(2):
        lst, cl = port.refine_cams_with_aff(t(G["attr_maps"])[i], t(G["attn"])[:, i], cls2[i], (96, 96), caa_thre=0.79)
        lab, cams, _ = port.refine_cams_with_bkg_weclip(lst, imgs[i], cl, (96, 96))
        assert (cams - t(G["cams"][i])).abs().max() < 1e-6
        assert np.array_equal(lab.numpy().astype(np.int16), G["labels"][i])


def test_label_utils_golden(golden):
    G = golden("labels")
    assert np.array_equal(port.get_mask_by_radius(6, 7, 2), G["mask"])
    assert np.array_equal(port.cams_to_affinity_label(t(G["lab"]), G["mask"], 255).numpy(), G["aff"])
    v, l = port.lam_to_label(t(G["cam"]), t(G["cls"]), 0.45, 0.6, 0.3, True, 255)
    assert np.array_equal(l.numpy(), G["l_mid"]) and np.array_equal(v.numpy(), G["valid"])
    assert np.array_equal(port.lam_to_label(t(G["cam"]), t(G["cls"]), 0.45)[1].numpy(), G["l_bkg"])


def test_lvc_branch_and_attrutils_oracle_vs_reference_golden(golden):
    """oracle/port.py, LVC branch of the surgery attention (clip/clip_surgery_model.py:127-141) and utils/attrutils.py,
    against outputs of the unmodified reference (oracle/make_golden_lvc.py)."""
    import torch
    from excel_b200 import synth
    from oracle import port
    from oracle.make_golden_cfg import TINY, checksum
    G = golden("lvc")
    t = torch.from_numpy
    W = port.random_visual_weights(seed=3, **TINY)
    imgs = synth.images(2, 96, seed=13)
    assert abs(checksum(*[v for k, v in W.items() if k != "meta"]) - float(G["chk_w"])) < 1e-6 * float(G["chk_w"])
    tok, attn, feats = port.generate_clip_fts(W, imgs, ex_feats=t(G["ex"]))
    assert (tok - t(G["tok"])).abs().max() < 2e-5
    assert (attn - t(G["attn"])).abs().max() < 2e-5
    assert (feats - t(G["feats"])).abs().max() < 2e-4
    assert (port.attrmap2clsmap(t(G["flag"]), t(G["amap"])) - t(G["clsmap"])).abs().max() < 1e-5
    assert (port.attr2cls_embedings(t(G["tf"]), t(G["bank"]), 20) - t(G["agg"])).abs().max() < 1e-6


def test_decoder_inference_oracle_vs_reference_golden(golden):
    """oracle/port.py segformer_head / attn_pred against the reference's module outputs."""
    import torch
    from oracle import port
    from oracle.make_golden_cfg import TINY
    G, GL = golden("decoder"), golden("lvc")
    t = torch.from_numpy
    Wd = {k[5:]: t(v) for k, v in G.items() if k.startswith("head.")}
    feats = t(GL["feats"])
    L, B, N, D = feats.shape
    x_all = feats[:, :, 1:].permute(0, 1, 3, 2).reshape(L, B, D, 6, 6)
    assert L == TINY["layers"] and (port.segformer_head(Wd, x_all) - t(G["fts"])).abs().max() < 1e-6
    assert (port.attn_pred(t(G["fts"])) - t(G["attn_pred"])).abs().max() < 1e-6
