"""Dev: the drop-in per-image loop (batch 1, 512^2) for an ncu launch list: python tools/ncu_dropin.py [n_images]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "dropin_tree"))
import torch
from excel_b200 import synth, install as inst
inst.install()
from model.model_excel import ExCEL_model
from utils.affutils import refine_cams_with_aff, refine_cams_with_bkg_weclip
from utils.PAR import PAR
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
model = ExCEL_model(synth.random_visual_weights(seed=0), synth.text_bank(45, 512, seed=1).t().contiguous(), 21).cuda().eval()
par = PAR(num_iter=20, dilations=[1, 2, 4, 8, 12, 24]).cuda()
imgs = synth.images(n, S, seed=10); cls = synth.class_labels(n, 20, seed=110, n_fixed=None)
with torch.no_grad():
    for k in range(n):
        x = imgs[k:k + 1].cuda(); c = cls[k:k + 1].cuda()
        _, _, attr, attn, _ = model(x)
        ref, cl = refine_cams_with_aff(attr[0], attn[:, 0], c[0], size=x.shape[2:], caa_thre=0.79)
        lab, _ = refine_cams_with_bkg_weclip(ref, x[0], cl, par, x.shape[-2:])
        lab.cpu()
