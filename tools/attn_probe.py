"""Dev: where does attn_pv_kernel's time go?  Builds private copies of the library with one component of the kernel
removed (-DXL_TUNING -DXL_PV_VARIANT=<mask>, see csrc/attn_pv.cu) and times one encoder forward with each.

    python tools/attn_probe.py build            # here (nvcc cross-compiles): build/dev/libxl_pv_<mask>.so
    python tools/attn_probe.py run [B] [S]      # on the GPU box: one JSON line per variant

The variants compute garbage -- they exist for timing only and never ship (the product build defines no XL_TUNING)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEV = os.path.join(ROOT, "build", "dev")
MASKS = [int(m) for m in os.environ.get("PV_MASKS", "0,1,2,4,8,16,32,64,128,256,512,1024,2048").split(",")]


def build():
    from excel_b200 import build as b
    b.build()
    os.makedirs(DEV, exist_ok=True)
    objs = [os.path.join(b.OBJ, f) for f in sorted(os.listdir(b.OBJ)) if f.endswith(".o") and f != "attn_pv.o"]
    for m in MASKS:
        obj = os.path.join(DEV, f"attn_pv_{m}.o")
        subprocess.check_call([b.NVCC] + b.FLAGS + ["-DXL_TUNING", f"-DXL_PV_VARIANT={m}", "-c", os.path.join(b.CSRC, "attn_pv.cu"), "-o", obj])
        subprocess.check_call([b.NVCC, "-shared", "-o", os.path.join(DEV, f"libxl_pv_{m}.so")] + objs + [obj] +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
        os.remove(obj)
    print("built", len(MASKS), "variants in", DEV)


def run_one(mask, B, S):
    import torch
    from excel_b200 import _lib, synth
    _lib.LIB_PATH = os.path.join(DEV, f"libxl_pv_{mask}.so")
    from excel_b200.encoder import SurgeryViT
    enc = SurgeryViT(synth.random_visual_weights(seed=0))
    imgs = synth.images(B, S, seed=10).cuda()
    for _ in range(2):
        enc(imgs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        enc(imgs)
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"mask": mask, "B": B, "S": S, "encoder_ms": round(a.elapsed_time(b) / 5, 3)}), flush=True)
    if mask & 4096:   # clock trace of the LAST attn_pv launch (CTA 0): dump for offline analysis
        import ctypes
        import numpy as np
        buf = np.zeros(3 * 4096, dtype=np.uint64)
        L = _lib.lib()
        L.excel_dev_pv_trace.argtypes, L.excel_dev_pv_trace.restype = [ctypes.c_void_p], ctypes.c_int
        assert L.excel_dev_pv_trace(buf.ctypes.data) == 0
        np.save(os.path.join(ROOT, "gpurun_out", f"pv_trace_{mask}.npy"), buf)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "one":
        run_one(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    else:
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
        S = int(sys.argv[3]) if len(sys.argv) > 3 else 512
        for m in MASKS:
            subprocess.call([sys.executable, os.path.abspath(__file__), "one", str(m), str(B), str(S)])
