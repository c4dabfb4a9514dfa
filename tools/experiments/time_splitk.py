"""Dev: residual GEMMs of the encoder (out_proj / c_proj shapes) through excel_gemm_tc with and without the split-K workspace."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import _lib

def run(M, N, K, sk):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g); B = torch.randn(N, K, device="cuda", generator=g) * 0.05
    bias = torch.randn(N, device="cuda", generator=g); res = torch.randn(M, N, device="cuda", generator=g)
    C = torch.empty(M, N, device="cuda")
    Kp = (K + 63) // 64 * 64
    ws = torch.empty(4 * (M + N) * Kp + 256 + (_lib.lib().excel_gemm_tc_splitk_bytes() if sk else 0), dtype=torch.uint8, device="cuda")
    f = lambda: _lib.call("excel_gemm_tc", _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bias), _lib.ptr(res), M, N, K, K, K, N, 1.0, 0,
                          _lib.ptr(ws), ws.numel(), _lib.stream())
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 20 * 1e3

for shape in ((16400, 768, 768), (16400, 768, 3072), (32800, 768, 768)):
    print(shape, "plain %.1f us  split-K %.1f us (incl. the two operand-split kernels)" % (run(*shape, False), run(*shape, True)), flush=True)
