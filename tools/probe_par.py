"""Dev probe: PAR throughput at the BASELINE.json shapes (cfg2 512^2 B=16, cfg5 1024^2 B=4)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import synth
from excel_b200.par import par_refine_planes, par_affinity

DIL = [1, 2, 4, 8, 12, 24]

def timeit(fn, warm=3, rep=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep

def run(B, S, C, iters, group):
    imgs = synth.images(B, S, seed=0).cuda()
    planes = torch.softmax(torch.randn(B * C, S, S, device="cuda"), 0).contiguous()
    off = torch.arange(0, (B + 1) * C, C, dtype=torch.int32, device="cuda")
    ms = timeit(lambda: par_refine_planes(imgs, planes, off, C, DIL, iters, group=group))
    ms_aff = timeit(lambda: par_affinity(imgs, (S, S), DIL)) if B * S * S * 48 * 4 < 8e9 else float("nan")
    bytes_alg = 4.0 * S * S * ((3 + 48) + iters * (48 + 2 * C)) * B
    print(json.dumps(dict(B=B, S=S, C=C, iters=iters, group=group, ms=round(ms, 3), ms_affinity_only=round(ms_aff, 3),
                          alg_GBs=round(bytes_alg / ms / 1e6, 1), img_per_s=round(B / ms * 1e3, 1))), flush=True)

if __name__ == "__main__":
    for c in (2, 3, 4, 5):
        run(16, 512, c, 20, 0)
    run(16, 512, 3, 20, 1)
    run(4, 1024, 4, 20, 0)
    run(4, 1024, 4, 1, 0); run(4, 1024, 4, 50, 0)
    run(8, 448, 3, 20, 0)
