import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from excel_b200 import _lib, synth
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from excel_b200.encoder import SurgeryViT
enc = SurgeryViT(synth.random_visual_weights(seed=0))
imgs = synth.images(16, 512, seed=10).cuda()
enc(imgs); torch.cuda.synchronize()
# Variants: build private libraries with
#   nvcc <build.FLAGS> -DXL_TUNING -DXL_TC_VARIANT=<mask> -c excel_b200/csrc/attn_tc.cu  (+ link with the other objects of build/obj)
# and run each under  ncu --metrics gpu__time_duration.sum -k regex:attn_tc_kernel python tools/experiments/stats_probe.py <lib>.
# Measured (1 x B200, 16 x 512^2, one score set, shipped kernel 96 us):  no MUFU.EX2 66 us, no tcgen05.ld 90 us, no S MMAs 76 us,
# no MUFU + no ld 67 us -- the epilogue's dependent chains (max -> scale -> exp2 -> sum -> online update), not a throughput limit.
# A joint 64-column epilogue (one online update per tile, both loads in flight) spilled 148 B and ran at 115 us.
