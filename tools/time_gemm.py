"""Dev: the encoder's GEMM shapes through excel_gemm_tc, for an ncu launch list (python tools/time_gemm.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from excel_b200 import _lib

def tc(A, B, bias=None, residual=None, act=0):
    M, K = A.shape
    N = B.shape[0]
    ws = torch.empty(4 * (M + N) * K, dtype=torch.uint8, device=A.device)
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _lib.call("excel_gemm_tc", _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bias), _lib.ptr(residual), M, N, K, A.stride(0),
              B.stride(0), N, 1.0, act, _lib.ptr(ws), ws.numel(), _lib.stream())
    return C

M = 16400
for (N, K) in ((768, 768), (768, 3072), (2304, 768)):
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    for _ in range(2):
        tc(A, B)                 # plain
        tc(A, B, bias)           # + bias
        tc(A, B, bias, res)      # + bias + residual
torch.cuda.synchronize()
