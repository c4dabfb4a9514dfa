"""Attribute-bank utilities -- drop-ins for the reference's ``utils/attrutils.py`` on sm_100a.

``attrmap2clsmap`` (patch x attribute-bank maps against the class flags, the per-batch contraction) runs on the tcgen05
GEMM engine (``excel_gemm_tc``: TMA-fed split-fp16 operands, fp32 accumulation in TMEM); ``attr2cls_embedings`` is
init-time work on a [cls, A] matrix and runs on ``excel_sgemm`` (exact fp32) with a row-softmax kernel in between.
``load_text_attri`` is file I/O and stays PyTorch.
"""
import os

import torch

from . import _lib


def _sgemm(A, B, b_is_nk, bias=None, residual=None, alpha=1.0):
    """A [M,K] @ (B [N,K]^T if b_is_nk else B [K,N]) (+ residual) through excel_sgemm."""
    A, B = _lib.f32c(A), _lib.f32c(B)
    M, K = A.shape
    N = B.shape[0] if b_is_nk else B.shape[1]
    if (B.shape[1] if b_is_nk else B.shape[0]) != K:
        raise RuntimeError(f"attrutils: inner dimensions differ ({tuple(A.shape)} x {tuple(B.shape)})")
    C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    residual = None if residual is None else _lib.f32c(residual)
    _lib.call("excel_sgemm", _lib.ptr(A), _lib.ptr(B), _lib.ptr(C), _lib.ptr(bias), _lib.ptr(residual), M, N, K, K, B.shape[1], N,
              1, 0, 0, 0, float(alpha), 1 if b_is_nk else 0, 0, _lib.stream())
    return C


def load_text_attri(pt_path):
    """utils/attrutils.py:4-9: (text_attri [E,A] fp32, attri_flag [cls,A]) on the GPU."""
    kind = os.path.basename(pt_path).replace("_cls_include4.pth", "")
    text_attri, attri_flag = torch.load(pt_path, map_location="cpu")[kind]
    return text_attri.float().cuda(), attri_flag.cuda()


@_lib.on_tensor_device
def attrmap2clsmap(attri_flag, attr_maps):
    """utils/attrutils.py:11-17: attr_maps [B,n_p,A] @ attri_flag[cls,A]^T -> [B,n_p,cls]."""
    B, n_p, A = attr_maps.shape
    X = _lib.f32c(attr_maps).reshape(B * n_p, A)
    Fl = _lib.f32c(attri_flag)
    if Fl.shape[1] != A:
        raise RuntimeError(f"attrutils: inner dimensions differ ({tuple(attr_maps.shape)} x {tuple(attri_flag.shape)})")
    M, N = X.shape[0], Fl.shape[0]
    kp = (A + 63) // 64 * 64
    out = torch.empty((M, N), dtype=torch.float32, device=X.device)
    ws = torch.empty((4 * (M + N) * kp,), dtype=torch.uint8, device=X.device)
    _lib.call("excel_gemm_tc", _lib.ptr(X), _lib.ptr(Fl), _lib.ptr(out), None, None, M, N, A, A, A, N, 1.0, 0, _lib.ptr(ws),
              ws.numel(), _lib.stream())
    return out.view(B, n_p, -1)


@_lib.on_tensor_device
def attr2cls_embedings(text_features, text_attri, num_classes):
    """utils/attrutils.py:19-29: softmax(fg_text @ bank) @ bank^T + fg_text, background rows appended, rows
    L2-normalised, returned transposed [E, T].  (The reference adds ALL text rows at :25, which only broadcasts when
    there are no background rows; the live equivalent, model/load_attr.py:86-119, adds the foreground rows -- as here.)"""
    text_features = _lib.f32c(text_features)
    fg, bg = text_features[:num_classes], text_features[num_classes:]
    logits = _sgemm(fg, text_attri, b_is_nk=False)                           # [cls, A]
    corr = torch.empty_like(logits)
    _lib.call("excel_row_softmax", _lib.ptr(logits), logits.shape[0], logits.shape[1], _lib.ptr(corr), _lib.stream())
    agg = _sgemm(corr, text_attri, b_is_nk=True, residual=fg)                # corr @ bank^T + fg_text
    agg = torch.cat([agg, bg], dim=0)
    out = torch.empty_like(agg)
    _lib.call("excel_row_l2_normalize", _lib.ptr(agg), agg.shape[0], agg.shape[1], _lib.ptr(out), _lib.stream())
    return out.permute(1, 0)
