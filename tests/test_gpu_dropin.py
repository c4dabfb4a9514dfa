"""The drop-in call surface on the GPU: the reference's per-image loop (tools/infer_lam.py:70-94) through
excel_b200.install() on the stand-in module tree tests/dropin_tree (the reference itself is not on the GPU box), checked
against the batched public API on the same images and against the oracle."""
import os
import sys

import pytest
import torch

from excel_b200 import synth
from oracle import port
from oracle.make_golden_cfg import TINY

TREE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin_tree")
_TOP = ("utils", "model", "clip", "_stub")


def _foreign_modules_loaded():
    return any(m in sys.modules and TREE not in (getattr(sys.modules[m], "__file__", "") or "") for m in _TOP)


@pytest.fixture
def dropin():
    """install() on the stand-in tree; everything it imported is removed again afterwards."""
    if _foreign_modules_loaded():
        pytest.skip("another `utils` / `model` / `clip` package (the real reference) is already imported in this process")
    from excel_b200 import install as inst
    sys.path.insert(0, TREE)
    state = {}

    def start(graph=False):
        state["orig"] = inst.install(graph=graph)
        return inst
    yield start
    if "orig" in state:
        inst.uninstall(state["orig"])
    inst.install.__globals__["_CLASS_PATCHES"].clear()
    from excel_b200 import encoder
    encoder.ENGINE_OPTS["graph"] = False
    encoder._ENGINES.clear()
    sys.path.remove(TREE)
    for m in list(sys.modules):
        if m in _TOP or m.split(".")[0] in _TOP:
            del sys.modules[m]


def test_install_patches_every_stub(dropin):
    """Not a GPU test: every hot-path symbol of the tree is a raising stub before install() and ours after."""
    sys.path.insert(0, TREE)
    import utils.affutils as aff
    with pytest.raises(NotImplementedError):
        aff.refine_cams_with_aff()
    sys.path.remove(TREE)
    dropin()
    import clip
    import utils.PAR
    import utils.camutils as cam
    import utils.attrutils as attr
    from model.model_excel import ExCEL_model
    from model.segformer_head import SegFormerHead
    for fn in (aff.refine_cams_with_aff, aff.refine_cams_with_bkg_weclip, aff.compute_trans_mat, utils.PAR.PAR, cam.cure_attr_map,
               cam.cure_attr_map_flip, cam.lam_to_label, cam.cams_to_affinity_label, cam.get_mask_by_radius, attr.attrmap2clsmap,
               attr.attr2cls_embedings, clip.generate_clip_fts, clip.clip_feature_surgery, clip.clip.generate_clip_fts):
        assert fn.__module__.startswith("excel_b200."), fn
    assert ExCEL_model.forward.__module__ == "excel_b200.install" and SegFormerHead.forward.__module__ == "excel_b200.install"


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_dropin_loop_equals_batched_api_and_oracle(dropin, graph):
    dropin(graph=graph)
    from model.model_excel import ExCEL_model
    from utils.affutils import refine_cams_with_aff, refine_cams_with_bkg_weclip
    from utils.PAR import PAR
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath
    from oracle.parity import label_parity
    S, B = 96, 3
    W = synth.random_visual_weights(seed=3, **TINY)
    text = synth.text_bank(45, TINY["embed"], seed=6)
    imgs = synth.images(B, S, seed=13)
    cls = synth.class_labels(B, 20, seed=14, n_fixed=None)
    model = ExCEL_model(W, text.t().contiguous(), 21, embedding_dim=64).cuda().eval()
    par = PAR(num_iter=20, dilations=[1, 2, 4, 8, 12, 24]).cuda()
    got, cams_raw = [], []
    with torch.no_grad():
        for k in range(B):                                              # tools/infer_lam.py:70-94, batch 1
            inputs = imgs[k:k + 1].cuda()
            cls_labels = cls[k:k + 1].cuda()
            seg, ex_feats, attr_maps_raw, attn_weights, attn_pred = model(inputs)
            assert seg.shape == (1, 21, S // 16, S // 16) and attn_pred.shape == (1, (S // 16) ** 2, (S // 16) ** 2)
            for i, attr_map in enumerate(attr_maps_raw):
                refined, cls_lst = refine_cams_with_aff(attr_map, attn_weights[:, i, ...], cls_labels[i], size=inputs.shape[2:],
                                                        seg_attn=None, caa_thre=0.79)
                labels, normed = refine_cams_with_bkg_weclip(refined, inputs[i], cls_lst, par, inputs.shape[-2:])
            assert not cls_lst.is_cuda and labels.shape == (1, S, S)
            got.append(labels[0].cpu())
            cams_raw.append(attr_maps_raw[0].cpu().clone())
    # same images through the batched public API: identical labels (same kernels, other launch grouping)
    hp = ExCELHotPath(SurgeryViT(W), text, 20)
    batched = hp(imgs.cuda(), cls).cpu()
    for k in range(B):
        assert torch.equal(batched[k], got[k]), k
    # and against the oracle: CAMs within 1e-3; labels through the one gate, stage-isolated where the box masks could differ
    with torch.no_grad():
        ref = port.hot_path(W, text, imgs, cls, 20)
    for k in range(B):
        assert (cams_raw[k] - ref["attr_maps_raw"][k]).abs().max() < 1e-3
    from excel_b200 import affutils
    lab_iso, planes, off, _ = affutils.refine_batch(ref["attr_maps_raw"].cuda(), ref["attn_weights"].cuda(), cls, imgs.cuda(), hp.par,
                                                    return_cams=True)
    off = off.cpu().tolist()
    for k in range(B):
        err = (planes[off[k]:off[k + 1]].cpu() - ref["cams"][k]).abs().max().item()
        hard, total = label_parity(ref["refined"][k], ref["labels"][k][0], lab_iso[k].cpu(), plane_err=err)
        assert err < 1e-3 and hard == 0, (k, err, hard, total)
