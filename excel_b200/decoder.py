"""Decoder-side inference pieces of ``ExCEL_model.forward`` on sm_100a:

* ``segformer_head`` -- ``SegFormerHead.forward`` (model/segformer_head.py:66-77): 12 two-layer MLPs (Linear-ReLU-Linear,
  :18-26) on the 12 ``all_feats`` levels, channel concat, 1x1 fuse conv; all of it GEMMs on the tcgen05 engine
  (``excel_gemm_tc``, fp32-quality split-fp16 products).  Dropout2d is the identity at inference.
* ``attn_pred`` -- model/model_excel.py:71-76: sigmoid((cosine similarity of the fused features - batch mean) * 3).
* ``excel_model_forward`` -- the whole ``ExCEL_model.forward`` (model/model_excel.py:48-77) for inference: encoder, CAM and the
  two pieces above on this package's kernels, the trained ``DecoderTransformer`` (``model.decoder``) as the PyTorch
  module it is.

Inference only: under autograd, or with the module in training mode (dropout), ``install`` keeps the reference forward.
"""
import torch

from . import _lib
from .clip import clip_feature_surgery
from .encoder import generate_clip_fts


def _gemm(A, W, bias, C=None, act=0):
    """C[M,N] (given, possibly a column block of a wider matrix) = act(A [M,K] @ W [N,K]^T + bias) via excel_gemm_tc."""
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K or A.stride(1) != 1 or W.stride(1) != 1:
        raise RuntimeError(f"decoder gemm: bad operands {tuple(A.shape)} x {tuple(W.shape)}")
    if C is None:
        C = torch.empty((M, N), dtype=torch.float32, device=A.device)
    kp = (K + 63) // 64 * 64
    ws = torch.empty(4 * (M + N) * kp, dtype=torch.uint8, device=A.device)
    _lib.call("excel_gemm_tc", _lib.ptr(A), _lib.ptr(W), _lib.ptr(C), _lib.ptr(bias), None, M, N, K, A.stride(0), W.stride(0),
              C.stride(0), 1.0, act, _lib.ptr(ws), ws.numel(), _lib.stream())
    return C


def _p(t):
    return _lib.f32c(t)


@_lib.on_tensor_device
def segformer_head_tokens(head, feats):
    """head: the reference's SegFormerHead module (weights read in place); feats [L, M, C] token-major rows (any M).
    Returns the fused features [M, E] (row m = the 1x1-fused embedding of token m)."""
    L, M, Cin = feats.shape
    mlps = head.linears_modulelist
    if len(mlps) != L:
        raise RuntimeError(f"SegFormerHead has {len(mlps)} levels, got {L} feature levels")
    E = mlps[0].proj.weight.shape[0]
    cat = torch.empty((M, L * E), dtype=torch.float32, device=feats.device)
    for l in range(L):
        h1 = _gemm(feats[l], _p(mlps[l].proj.weight), _p(mlps[l].proj.bias), act=2)              # Linear + ReLU (:22-24)
        _gemm(h1, _p(mlps[l].proj_2.weight), _p(mlps[l].proj_2.bias), C=cat[:, l * E:(l + 1) * E])   # Linear (:25) -> cat (:74)
    wf = _p(head.linear_fuse.weight).reshape(head.linear_fuse.weight.shape[0], L * E)
    return _gemm(cat, wf, _p(head.linear_fuse.bias))                                            # 1x1 conv (:75)


@_lib.on_tensor_device
def segformer_head(head, x_all):
    """Drop-in for SegFormerHead.forward: x_all [L, B, C, h, w] (channel-major, as model_excel.py:60-63 builds it)
    -> [B, E, h, w]."""
    L, B, C, h, w = x_all.shape
    feats = _lib.f32c(x_all).reshape(L, B, C, h * w).permute(0, 1, 3, 2).reshape(L, B * h * w, C).contiguous()
    out = segformer_head_tokens(head, feats)
    return out.reshape(B, h * w, -1).permute(0, 2, 1).reshape(B, -1, h, w).contiguous()


@_lib.on_tensor_device
def attn_pred(attn_fts, beta=1.0, gamma=3.0):
    """model/model_excel.py:71-76: attn_fts [B,C,h,w] -> sigmoid((cos-sim - batch mean) * 3) [B, h*w, h*w]."""
    f = _lib.f32c(attn_fts)
    B, C = f.shape[:2]
    f = f.reshape(B, C, -1)
    n = f.shape[2]
    dev = f.device
    qt = torch.empty((B, n, C), dtype=torch.float32, device=dev)
    rowsum = torch.empty((B * n,), dtype=torch.float64, device=dev)
    mean = torch.empty((1,), dtype=torch.float32, device=dev)
    out = torch.empty((B, n, n), dtype=torch.float32, device=dev)
    _lib.call("excel_attn_pred", _lib.ptr(f), B, C, n, float(beta), float(gamma), _lib.ptr(qt), _lib.ptr(rowsum), _lib.ptr(mean),
              _lib.ptr(out), _lib.stream())
    return out


@torch.no_grad()
def excel_model_forward(model, img, ex_feats=None):
    """ExCEL_model.forward (model/model_excel.py:48-77) at inference: same 5-tuple
    (seg, attn_fts, attr_maps_raw, attn_weights, attn_pred); with ex_feats only attr_maps_raw (:50-53)."""
    text_t = model.text_attr.permute(1, 0)
    nfg = model.num_classes - 1
    if ex_feats is not None:
        tok, _, _ = generate_clip_fts(img, model.encoder, return_weights=True, ex_feats=ex_feats)
        return clip_feature_surgery(tok, text_t)[:, 1:, :nfg]
    b, c, h, w = img.shape
    tok, attn_weights, all_feats = generate_clip_fts(img, model.encoder, return_weights=True)
    attr_maps_raw = clip_feature_surgery(tok, text_t)[:, 1:, :nfg]
    L, B, N, D = all_feats.shape
    fused = segformer_head_tokens(model.decoder_fts_fuse, all_feats.reshape(L, B * N, D))       # rows incl. the CLS tokens
    fts = fused.reshape(B, N, -1)[:, 1:].permute(0, 2, 1).reshape(B, -1, h // 16, w // 16).contiguous()   # :60-65
    attn_fts = fts.clone()
    seg, _ = model.decoder(fts)                                                                  # trained PyTorch module (:69)
    return seg, attn_fts.clone().detach(), attr_maps_raw, attn_weights, attn_pred(attn_fts)
