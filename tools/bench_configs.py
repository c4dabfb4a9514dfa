"""Dev: timings at the other BASELINE.json configs (not the bench line): cfg3 per-GPU shard, cfg4 ViT-L/14@336 encoder + CAM,
cfg5 PAR iteration sweep at 1024^2.  Prints one JSON object per config."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import synth
from excel_b200.encoder import SurgeryViT, generate_clip_fts
from excel_b200.clip import clip_feature_surgery
from excel_b200.pipeline import ExCELHotPath
from excel_b200.par import par_refine_planes

DIL = [1, 2, 4, 8, 12, 24]


def timeit(fn, warm=2, rep=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / rep


def cfg3():   # COCO 448^2, batch 64 over 8 GPUs -> 8 images per GPU, T = 103 text rows, 80 fg classes, COCO class-count mix
    hp = ExCELHotPath(SurgeryViT(synth.random_visual_weights(seed=0)), synth.text_bank(103, 512, seed=1), 80)
    imgs = synth.images(8, 448, seed=3).cuda()
    cls = synth.class_labels(8, 80, seed=4, n_fixed=None, dataset="ms_coco")
    ms = timeit(lambda: hp(imgs, cls))
    print(json.dumps({"config": "cfg3 per-GPU shard: ViT-B/16 CAM+SVC+PAR, COCO 448^2, 8 images", "ms": round(ms, 2),
                      "images_per_s_per_gpu": round(8e3 / ms, 1), "classes_per_image": [int(c.sum()) for c in cls]}))


def cfg4():   # ViT-L/14@336 + 103-row text bank, batch 8
    W = synth.random_visual_weights(layers=24, width=1024, patch=14, grid0=24, embed=768, seed=4)
    enc = SurgeryViT(W)
    text = synth.text_bank(103, 768, seed=5).cuda()
    imgs = synth.images(8, 336, seed=6).cuda()

    def f():
        tok, attn, feats = generate_clip_fts(imgs, enc)
        return clip_feature_surgery(tok, text)
    ms = timeit(f)
    print(json.dumps({"config": "cfg4: ViT-L/14@336 dense forward + 103-row CAM, batch 8", "ms": round(ms, 2),
                      "images_per_s": round(8e3 / ms, 1), "tflops_fp32_equiv": round(8 * 402e9 / ms / 1e9, 1)}))


def cfg5():   # PAR sweep 1..50 iterations at 1024^2, batch 4, C = 4
    B, S, C = 4, 1024, 4
    imgs = synth.images(B, S, seed=0).cuda()
    planes = torch.softmax(torch.randn(B * C, S, S, device="cuda"), 0).contiguous()
    off = torch.arange(0, (B + 1) * C, C, dtype=torch.int32, device="cuda")
    out = {}
    for it in (1, 2, 5, 10, 20, 50):
        ms = timeit(lambda: par_refine_planes(imgs, planes, off, C, DIL, it))
        alg = 4.0 * S * S * ((3 + 48) + it * (48 + 2 * C)) * B
        out[str(it)] = {"ms": round(ms, 3), "alg_GBs": round(alg / ms / 1e6, 1)}
    print(json.dumps({"config": "cfg5: PAR 1-50 iterations @1024^2, batch 4, 4 planes (algorithmic bytes incl. the affinity set-up)", "iters": out}))


if __name__ == "__main__":
    for f in (cfg3, cfg4, cfg5):
        f()
