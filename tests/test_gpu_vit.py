"""GPU parity: CLIP-surgery encoder (clip/clip_surgery_model.py) + whole hot path vs the oracle and the fixtures."""
import pytest
import torch

from _parity import label_parity
from excel_b200 import synth
from oracle import port
from oracle.make_golden_cfg import TINY

pytestmark = pytest.mark.gpu
t = torch.from_numpy


def test_vit_tiny_golden(golden):
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    from excel_b200.clip import clip_feature_surgery
    G = golden("vit_tiny")
    enc = SurgeryViT(port.random_visual_weights(seed=3, **TINY))
    imgs = synth.images(2, 96, seed=13).cuda()
    tok, attn, feats = generate_clip_fts(imgs, enc)
    assert (tok.cpu() - t(G["tok"])).abs().max() < 2e-5
    assert (attn.cpu() - t(G["attn"])).abs().max() < 2e-5
    d = (feats.cpu() - t(G["feats"])).abs().amax(dim=(1, 2, 3))
    assert d.max() < 2e-4, d                                      # rows first-1 .. L-2 carry the aliasing quirk
    text = synth.text_bank(45, TINY["embed"], seed=6).cuda()
    attr = clip_feature_surgery(tok, text)[:, 1:, :20].cpu()
    assert (attr - t(G["attr_maps"])).abs().max() < 1e-3           # north_star: fp32 CAM values within 1e-3


def test_vit_b16_vs_oracle():
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    W = port.random_visual_weights(seed=1)                         # ViT-B/16, 12 layers, grid0 = 14
    enc = SurgeryViT(W)
    imgs = synth.images(2, 224, seed=21)
    tok, attn, feats = generate_clip_fts(imgs.cuda(), enc)
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs)
    assert (attn.cpu() - attn_r).abs().max() < 5e-5
    assert (attn.cpu()[:7].sum(-1) - 1).abs().max() < 1e-4 and (attn.cpu()[7:].sum(-1) - 12).abs().max() < 1e-3
    assert ((feats.cpu() - feats_r).abs().amax(dim=(1, 2, 3)) / feats_r.abs().amax(dim=(1, 2, 3))).max() < 1e-4
    assert (tok.cpu() - tok_r).abs().max() < 1e-4
    # positional-embedding resize path (grid 14 -> 20) and batch of 1
    imgs = synth.images(1, 320, seed=22)
    tok, attn, feats = generate_clip_fts(imgs.cuda(), enc)
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs)
    assert (attn.cpu() - attn_r).abs().max() < 5e-5 and (tok.cpu() - tok_r).abs().max() < 1e-4


def test_hot_path_end_to_end_vs_oracle():
    """Whole path on the GPU vs the whole oracle.  CAMs within 1e-3; labels identical except near-ties of the PAR argmax
    (tests/_parity.py).  The box masks are a discontinuous function of the CAMs (uint8 truncation + threshold): where the
    GPU's and the oracle's masks differ, the tail (SVC + PAR + argmax) is compared on the ORACLE's CAMs instead."""
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath
    from excel_b200 import affutils
    W = port.random_visual_weights(seed=2)
    text = synth.text_bank(45, 512, seed=3)
    imgs = synth.images(2, 224, seed=23)
    cls = synth.class_labels(2, 20, seed=24, n_fixed=2)
    hp = ExCELHotPath(SurgeryViT(W), text, 20)
    attr, attn, _ = hp.cams(imgs.cuda())
    ref = port.hot_path(W, text, imgs, cls, 20)
    assert (attr.cpu() - ref["attr_maps_raw"]).abs().max() < 1e-3
    labels = hp(imgs.cuda(), cls.cuda()).cpu()
    assert labels.shape == (2, 224, 224) and labels.dtype == torch.int64
    lists = affutils._class_lists(cls)
    m_gpu = affutils.box_masks(attr, lists, 14, 14, 0.79).cpu()
    m_ref = affutils.box_masks(ref["attr_maps_raw"].cuda(), lists, 14, 14, 0.79).cpu()
    lab_iso, planes_iso, off, _ = affutils.refine_batch(ref["attr_maps_raw"].cuda(), ref["attn_weights"].cuda(), cls, imgs.cuda(),
                                                        hp.par, return_cams=True)
    _, planes_e2e, _, _ = affutils.refine_batch(attr, attn, cls, imgs.cuda(), hp.par, return_cams=True)
    off = off.cpu().tolist()
    q = 0
    for b in range(2):
        same = torch.equal(m_gpu[q:q + 2], m_ref[q:q + 2])
        q += 2
        lst, cl = port.refine_cams_with_aff(ref["attr_maps_raw"][b], ref["attn_weights"][:, b], cls[b], (224, 224))
        lab_r, cams_r, ref_planes = port.refine_cams_with_bkg_weclip(lst, imgs[b], cl, (224, 224))
        assert torch.equal(lab_r, ref["labels"][b])
        planes = planes_e2e if same else planes_iso
        err = (planes[off[b]:off[b + 1]].cpu() - cams_r).abs().max().item()
        assert err < 1e-3, (b, same, err)
        hard, total = label_parity(ref_planes, lab_r[0], (labels if same else lab_iso.cpu())[b], plane_err=err)
        assert hard == 0 and total < 2e-3 * 224 * 224, (b, same, hard, total)
