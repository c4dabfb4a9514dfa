"""GPU: the tcgen05 split-fp16 GEMM (excel_gemm_tc) and the fp32 SIMT GEMM (excel_sgemm) vs a float64 product."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tc(A, B, bias=None, residual=None, alpha=1.0, act=0):
    from excel_b200 import _lib
    M, K = A.shape
    N = B.shape[0]
    Kp = (K + 63) // 64 * 64
    ws = torch.empty(4 * (M + N) * Kp, dtype=torch.uint8, device=A.device)
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _lib.call("excel_gemm_tc", _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bias), _lib.ptr(residual), M, N, K, A.stride(0),
              B.stride(0), N, alpha, act, _lib.ptr(ws), ws.numel(), _lib.stream())
    return C


def _simt(A, B, bias=None, residual=None, alpha=1.0, act=0):
    from excel_b200 import _lib
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    _lib.call("excel_sgemm", _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bias), _lib.ptr(residual), M, N, K, A.stride(0),
              B.stride(0), N, 1, 0, 0, 0, alpha, 1, act, _lib.stream())
    return C


@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 384, 768), (1025, 2304, 768), (300, 45, 512), (77, 200, 100),
                                   (4100, 768, 3072), (1, 8, 64)])
def test_gemm_tc_matches_fp64(shape):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g) * 3
    B = torch.randn(N, K, device="cuda", generator=g)
    ref = (A.double() @ B.double().t())
    scale = ref.abs().max().item()
    errs = {}
    for fn, tol in ((_tc, 3e-5), (_simt, 3e-6)):
        C = fn(A, B)
        errs[fn.__name__] = (C.double() - ref).abs().max().item() / scale
        assert errs[fn.__name__] < tol, (fn.__name__, shape, errs)
    print("gemm max-rel-err", shape, errs)


def test_gemm_tc_epilogue():
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(333, 192, device="cuda", generator=g)
    B = torch.randn(260, 192, device="cuda", generator=g)
    bias = torch.randn(260, device="cuda", generator=g)
    res = torch.randn(333, 260, device="cuda", generator=g)
    x = 0.5 * (A.double() @ B.double().t()) + bias.double()
    ref = x * torch.sigmoid(1.702 * x) + res.double()
    for fn in (_tc, _simt):
        C = fn(A, B, bias, res, 0.5, 1)
        assert (C.double() - ref).abs().max().item() / ref.abs().max().item() < 1e-5, fn.__name__


def test_gemm_tc_large_values_and_zeros():
    A = torch.zeros(130, 64, device="cuda")
    A[5, 3] = 1000.0
    A[7, 9] = 1e-4
    B = torch.eye(64, device="cuda")[:40].contiguous()
    C = _tc(A, B)
    # fp16 hi/lo: exact for fp16-representable values; absolute resolution 2^-24 ~ 6e-8 below fp16's normal range
    assert C[5, 3].item() == 1000.0 and abs(C[7, 9].item() - 1e-4) < 6e-8 and C.abs().sum().item() < 1000.0002
