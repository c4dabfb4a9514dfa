"""Golden fixtures for the LVC branch and utils/attrutils.py (a10), from the UNMODIFIED reference.

    python -B oracle/make_golden_lvc.py        (build container only: needs /root/reference)

Writes tests/golden/lvc.npz.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
from oracle import port, ref_harness as H  # noqa: E402
from oracle.make_golden import load_into_reference  # noqa: E402
from oracle.make_golden_cfg import TINY, checksum  # noqa: E402
from excel_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    ref = H.load()
    torch.set_grad_enabled(False)
    # tiny surgery ViT, Attention.forward WITH ex_feats (clip/clip_surgery_model.py:127-141) through clip.generate_clip_fts
    W = port.random_visual_weights(seed=3, **TINY)
    enc = ref.csm.ExCEL_CLIP(TINY["embed"], TINY["grid0"] * 16, TINY["layers"], TINY["width"], 16, 77, 49408, 64, 1, 1).float().eval()
    load_into_reference(enc.visual, W)
    enc.visual.reload_self_attn(layers=6, feat_size=6, mode="val")
    imgs = synth.images(2, 96, seed=13)
    g = torch.Generator().manual_seed(99)
    ex = torch.randn(2, 16, 6, 6, generator=g)
    ex = ex + 0.5 * torch.nn.functional.avg_pool2d(torch.nn.functional.pad(ex, [1] * 4, mode="replicate"), 3, stride=1)
    tok, attn, feats = ref.clip.generate_clip_fts(imgs, enc, return_weights=True, ex_feats=ex)
    tok0, _, _ = ref.clip.generate_clip_fts(imgs, enc, return_weights=True)
    assert (tok - tok0).abs().max() > 1e-3                     # the bias matters
    # utils/attrutils.py
    flag = (torch.rand(20, 112, generator=g) > 0.8).float()
    amap = torch.rand(2, 36, 112, generator=g)
    clsmap = ref.attrutils.attrmap2clsmap(flag, amap)
    tf, bank = torch.randn(20, 64, generator=g), torch.randn(64, 112, generator=g)
    agg = ref.attrutils.attr2cls_embedings(tf, bank, 20)
    # decoder-side inference (f4): the reference's SegFormerHead (seeded) on the tiny model's all_feats, and attn_pred
    import importlib
    seg_mod = importlib.import_module("model.segformer_head")
    torch.manual_seed(123)
    head = seg_mod.SegFormerHead(in_channels=TINY["width"], embedding_dim=32, num_classes=21, index=TINY["layers"]).eval()
    x_all = feats[:, :, 1:].permute(0, 1, 3, 2).reshape(TINY["layers"], 2, TINY["width"], 6, 6)
    fts = head(x_all)
    af = torch.nn.functional.normalize(fts.reshape(2, 32, 36), dim=1)
    apred = af.transpose(2, 1).bmm(af)
    apred = torch.sigmoid((apred - torch.mean(apred) * 1.) * 3.0)                 # model/model_excel.py:71-76
    head_sd = {"head." + k: v.numpy() for k, v in head.state_dict().items()}
    np.savez(os.path.join(OUT, "decoder.npz"), fts=fts.numpy(), attn_pred=apred.numpy(), **head_sd)
    print("decoder.npz", os.path.getsize(os.path.join(OUT, "decoder.npz")) // 1024, "KiB")
    np.savez(os.path.join(OUT, "lvc.npz"), ex=ex.numpy(), tok=tok.numpy(), attn=attn.numpy(), feats=feats.numpy(),
             ex_attn=port.lvc_attention(ex).numpy(), chk_w=checksum(*[v for k, v in W.items() if k != "meta"]), chk_img=checksum(imgs),
             flag=flag.numpy(), amap=amap.numpy(), clsmap=clsmap.numpy(), tf=tf.numpy(), bank=bank.numpy(), agg=agg.numpy())
    print("lvc.npz", os.path.getsize(os.path.join(OUT, "lvc.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
