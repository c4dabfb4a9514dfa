"""Dev: a short PAR run for ncu captures (cfg2 shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import synth
from excel_b200.par import par_refine_planes
B, S, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
group = int(sys.argv[4]) if len(sys.argv) > 4 else 0
imgs = synth.images(B, S, seed=0).cuda()
planes = torch.softmax(torch.randn(B * C, S, S, device="cuda"), 0).contiguous()
off = torch.arange(0, (B + 1) * C, C, dtype=torch.int32, device="cuda")
for _ in range(2):
    par_refine_planes(imgs, planes, off, C, [1, 2, 4, 8, 12, 24], 4, group=group)
torch.cuda.synchronize()
