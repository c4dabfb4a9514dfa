"""Dev: one hot-path step (cfg2 shape) for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import synth
from excel_b200.encoder import SurgeryViT
from excel_b200.pipeline import ExCELHotPath
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
hp = ExCELHotPath(SurgeryViT(synth.random_visual_weights(seed=0)), synth.text_bank(45, 512, seed=1), 20)
imgs = synth.images(B, S, seed=10).cuda(); cls = synth.class_labels(B, 20, seed=110, n_fixed=None).cuda()
for _ in range(2):
    hp(imgs, cls)
torch.cuda.synchronize()
