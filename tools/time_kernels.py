"""Dev: CUDA-event timing of one encoder forward (cfg2 shape), per run; prints ms.  Usage: python tools/time_kernels.py [B] [S]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import synth
from excel_b200.encoder import SurgeryViT
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
enc = SurgeryViT(synth.random_visual_weights(seed=0))
imgs = synth.images(B, S, seed=10).cuda()
for _ in range(2):
    enc(imgs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    enc(imgs)
b.record(); torch.cuda.synchronize()
print("encoder ms", a.elapsed_time(b) / 5)
