"""Dev: A/B timing of library builds on ONE box (boxes of the pool differ by several % in sustained clocks, so only
same-box comparisons count):  python tools/ab_time.py build/base/libexcel_b200.so excel_b200/lib/libexcel_b200.so
Each library is timed in its own subprocess, interleaved A B A B; prints encoder-forward and whole-step ms (cfg2 shape)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(path, B, S):
    import torch
    from excel_b200 import _lib, synth
    _lib.LIB_PATH = os.path.abspath(path)
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath
    enc = SurgeryViT(synth.random_visual_weights(seed=0))
    hp = ExCELHotPath(enc, synth.text_bank(45, 512, seed=1), 20)
    imgs = synth.images(B, S, seed=10).cuda()
    cls = synth.class_labels(B, 20, seed=110, n_fixed=None)

    def t(fn, rep=8):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(rep):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / rep
    print(json.dumps({"lib": path, "encoder_ms": round(t(lambda: enc(imgs)), 3), "step_ms": round(t(lambda: hp(imgs, cls)), 3)}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        libs = sys.argv[1:]
        for _ in range(2):
            for lib in libs:
                subprocess.call([sys.executable, os.path.abspath(__file__), "one", lib, "16", "512"])
