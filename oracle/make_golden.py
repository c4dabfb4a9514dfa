"""Freeze outputs of the UNMODIFIED reference into tests/golden/*.npz (build container only).

    python -B oracle/make_golden.py

The reference (pure Python, /root/reference) cannot travel to the GPU box, so its outputs on
seeded inputs are committed as small fixtures; tests/test_oracle_golden.py replays the oracle
port against them and tests/test_gpu_*.py replay the CUDA path.  Inputs are regenerated from
seeds at test time (excel_b200/synth.py, oracle/port.random_visual_weights); each fixture stores
a checksum of its inputs so generator drift is detected instead of silently mis-compared.
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
from oracle import port, ref_harness as H  # noqa: E402
from excel_b200 import synth  # noqa: E402

from oracle.make_golden_cfg import TINY, checksum  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_into_reference(visual, W):
    """Copy a weight pack into a reference VisionTransformer (before reload_self_attn)."""
    sd = visual.state_dict()
    new = {"conv1.weight": W["conv1.weight"], "class_embedding": W["class_embedding"],
           "positional_embedding": W["positional_embedding"], "proj": W["proj"]}
    for n in ("ln_pre", "ln_post"):
        new[n + ".weight"], new[n + ".bias"] = W[n + ".weight"], W[n + ".bias"]
    L = int(W["meta"][0])
    for i in range(L):
        p, o = "transformer.resblocks.%d." % i, "blocks.%d." % i
        new[p + "attn.in_proj_weight"], new[p + "attn.in_proj_bias"] = W[o + "in_proj_weight"], W[o + "in_proj_bias"]
        new[p + "attn.out_proj.weight"], new[p + "attn.out_proj.bias"] = W[o + "out_proj.weight"], W[o + "out_proj.bias"]
        for n in ("ln_1", "ln_2"):
            new[p + n + ".weight"], new[p + n + ".bias"] = W[o + n + ".weight"], W[o + n + ".bias"]
        new[p + "mlp.c_fc.weight"], new[p + "mlp.c_fc.bias"] = W[o + "c_fc.weight"], W[o + "c_fc.bias"]
        new[p + "mlp.c_proj.weight"], new[p + "mlp.c_proj.bias"] = W[o + "c_proj.weight"], W[o + "c_proj.bias"]
    assert set(new) == set(sd), set(sd) ^ set(new)
    visual.load_state_dict(new)


def main():
    ref = H.load()
    torch.set_grad_enabled(False)
    os.makedirs(OUT, exist_ok=True)
    g = torch.Generator().manual_seed(1234)

    # ---- PAR (utils/PAR.py) -- resize branch + non-square + ragged C
    par20 = ref.PAR(num_iter=20, dilations=list(port.PAR_DILATIONS))
    par3 = ref.PAR(num_iter=3, dilations=list(port.PAR_DILATIONS))
    im_a = synth.images(2, 48, seed=11)
    mk_a = torch.softmax(3 * torch.randn(2, 3, 48, 48, generator=g), 1)
    im_b = torch.rand(1, 3, 40, 56, generator=g)
    mk_b = torch.softmax(torch.randn(1, 5, 61, 83, generator=g), 1)
    np.savez(os.path.join(OUT, "par.npz"),
             mk_a=mk_a.numpy(), out_a=par20(im_a, mk_a).numpy(), chk_a=checksum(im_a),
             im_b=im_b.numpy(), mk_b=mk_b.numpy(), out_b=par3(im_b, mk_b).numpy())

    # ---- SVC pieces (utils/affutils.py)
    A = (torch.rand(6, 65, 65, generator=g) + 0.02) * torch.tensor([1., 12, 12, 12, 12, 12]).view(6, 1, 1)
    T = ref.affutils.compute_trans_mat(A[:, 1:, 1:].mean(0))
    rng = np.random.default_rng(7)
    maps, masks, thrs = [], [], []
    for t in range(64):
        cam = rng.random((8, 8)).astype(np.float32)
        if t % 2:
            cam *= (rng.random((8, 8)) > 0.6)
        thr = float([0.75, 0.79, 0.88][t % 3])
        box, cnt = ref.affutils.scoremap2bbox(cam, thr, multi_contour_eval=True)
        m = np.zeros((8, 8), np.float32)
        for k in range(cnt):
            x0, y0, x1, y1 = box[k]
            m[y0:y1, x0:x1] = 1
        maps.append(cam), masks.append(m), thrs.append(thr)
    attr = torch.rand(64, 20, generator=g)
    cls = torch.zeros(20)
    cls[[2, 7, 15]] = 1
    seg_attn = torch.rand(1, 64, 64, generator=g)
    lst, cl = ref.affutils.refine_cams_with_aff(attr, A, cls, (128, 128), caa_thre=0.79)
    lst_s, _ = ref.affutils.refine_cams_with_aff(attr, A, cls, (128, 128), caa_thre=0.75, seg_attn=seg_attn)
    img = synth.images(1, 128, seed=12)[0]
    lab, cams = ref.affutils.refine_cams_with_bkg_weclip(lst, img, cl, par20, (96, 112))
    np.savez(os.path.join(OUT, "svc.npz"), A=A.numpy(), T=T.numpy(), maps=np.stack(maps), masks=np.stack(masks),
             thrs=np.asarray(thrs, np.float32), attr=attr.numpy(), cls=cls.numpy(), seg_attn=seg_attn.numpy(),
             refined=torch.stack(lst).numpy(), refined_seg=torch.stack(lst_s).numpy(), cls_lst=cl.numpy(),
             labels=lab.numpy().astype(np.int16), cams=cams.numpy(), chk_img=checksum(img))

    # ---- CAM (clip/clip.py:288-310) + flip merge (utils/camutils.py:8-30)
    Fe = torch.randn(2, 37, 64, generator=g)
    Fe = Fe / Fe.norm(dim=1, keepdim=True)
    Te = synth.text_bank(45, 64, seed=5)
    cam = ref.clip.clip_feature_surgery(Fe, Te)

    class _M:  # cure_attr_map_flip only needs model(x)[2]
        def __call__(self, x, ex_feats=None):
            return None, None, lam2b
    lam2b = torch.rand(4, 36, 20, generator=g)
    merged = ref.camutils.cure_attr_map_flip(_M(), torch.zeros(2, 3, 96, 96), ex_fts=False, flip=True)
    np.savez(os.path.join(OUT, "cam.npz"), F=Fe.numpy(), T=Te.numpy(), cam=cam.numpy(), lam2b=lam2b.numpy(),
             merged=merged.numpy())

    # ---- tiny surgery ViT end to end (clip/clip_surgery_model.py + clip/clip.py:348-358)
    W = port.random_visual_weights(seed=3, **TINY)
    enc = ref.csm.ExCEL_CLIP(TINY["embed"], TINY["grid0"] * 16, TINY["layers"], TINY["width"], 16, 77, 49408, 64, 1, 1).float().eval()
    load_into_reference(enc.visual, W)
    enc.visual.reload_self_attn(layers=6, feat_size=6, mode="val")
    imgs = synth.images(2, 96, seed=13)
    tok, attn, feats = ref.clip.generate_clip_fts(imgs, enc, return_weights=True)
    text = synth.text_bank(45, TINY["embed"], seed=6)
    attr_maps = ref.clip.clip_feature_surgery(tok, text)[:, 1:, :20]
    cls2 = synth.class_labels(2, 20, seed=14, n_fixed=2)
    labels, cams_all = [], []
    for i in range(2):
        lst, cl = ref.affutils.refine_cams_with_aff(attr_maps[i], attn[:, i], cls2[i], imgs.shape[2:], caa_thre=0.79)
        lab, cams = ref.affutils.refine_cams_with_bkg_weclip(lst, imgs[i], cl, par20, imgs.shape[-2:])
        labels.append(lab.numpy().astype(np.int16)), cams_all.append(cams.numpy())
    np.savez(os.path.join(OUT, "vit_tiny.npz"), tok=tok.numpy(), attn=attn.numpy(), feats=feats.numpy(),
             attr_maps=attr_maps.numpy(), labels=np.stack(labels), cams=np.stack(cams_all),
             chk_w=checksum(*[v for k, v in W.items() if k != "meta"]), chk_img=checksum(imgs), chk_text=checksum(text))
    # ---- training-side label utilities (utils/camutils.py:123-143, 438-476)
    m_ref = ref.camutils.get_mask_by_radius(h=6, w=7, radius=2)
    g2 = torch.Generator().manual_seed(77)
    lab = torch.randint(0, 4, (2, 96, 112), generator=g2)
    lab[0, :20] = 255
    a_ref = ref.camutils.cams_to_affinity_label(lab.clone(), mask=m_ref, ignore_index=255)
    cam2 = torch.rand(2, 5, 24, 28, generator=g2)
    cls3 = (torch.rand(2, 5, generator=g2) > 0.4).float()
    v_ref, l_ref = ref.camutils.lam_to_label(cam2, cls3, bkg_thre=0.45, high_thre=0.6, low_thre=0.3, ignore_mid=True, ignore_index=255)
    _, l2_ref = ref.camutils.lam_to_label(cam2, cls3, bkg_thre=0.45)
    np.savez(os.path.join(OUT, "labels.npz"), mask=m_ref, lab=lab.numpy(), aff=a_ref.numpy(), cam=cam2.numpy(), cls=cls3.numpy(),
             l_mid=l_ref.numpy(), l_bkg=l2_ref.numpy(), valid=v_ref.numpy())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
