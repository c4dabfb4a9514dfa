"""Decoder-side inference pieces of ``ExCEL_model.forward`` on sm_100a:

* ``segformer_head`` -- ``SegFormerHead.forward`` (model/segformer_head.py:66-77): 12 two-layer MLPs (Linear-ReLU-Linear,
  :18-26) on the 12 ``all_feats`` levels, channel concat, 1x1 fuse conv; all of it GEMMs on the tcgen05 engine
  (``excel_gemm_tc_split``: fp32-quality split-fp16 products, weights split once and cached).  Dropout2d is the identity at inference.
* ``attn_pred`` -- model/model_excel.py:71-76: sigmoid((cosine similarity of the fused features - batch mean) * 3).
* ``excel_model_forward`` -- the whole ``ExCEL_model.forward`` (model/model_excel.py:48-77) for inference: encoder, CAM and the
  two pieces above on this package's kernels, the trained ``DecoderTransformer`` (``model.decoder``) as the PyTorch
  module it is.

Inference only: under autograd, or with the module in training mode (dropout), ``install`` keeps the reference forward.
"""
import math

import torch

from . import _lib
from .clip import clip_feature_surgery
from .encoder import generate_clip_fts


def _pow2_scale(t):
    """The power of two that puts max|t| at 2^13..2^14 (both split halves in fp16's normal range)."""
    amax = float(t.abs().max())
    if not math.isfinite(amax):
        raise RuntimeError("decoder: weights contain inf / nan")
    return 2.0 ** (14 - math.frexp(amax)[1]) if amax > 0 else 1.0


def _r64(n):
    return (n + 63) // 64 * 64


def _split(x, scale=1.0):
    """fp32 [rows, cols] -> split fp16 [rows, 2*Kp] (hi | lo, Kp = round_up(cols, 64), zero padded) of scale * x: the GEMM
    engine's operand format."""
    x = _lib.f32c(x)
    rows, cols = x.shape
    kp = _r64(cols)
    out = torch.empty((rows, 2 * kp), dtype=torch.float16, device=x.device)
    _lib.call("excel_split_f16", _lib.ptr(x), x.stride(0), rows, cols, kp, float(scale), _lib.ptr(out), _lib.stream())
    return out


_HEADS = {}


def _head_pack(head):
    """Split-fp16 copies of a SegFormerHead's weights, stacked per level, cached per module and weight version (the head
    is trained: a new optimizer step bumps the parameters' version counters and the pack is rebuilt)."""
    params = list(head.parameters())
    ver = tuple((p.data_ptr(), p._version) for p in params)
    hit = _HEADS.get(id(head))
    if hit is not None and hit[0] == ver:
        return hit[1]
    mlps = head.linears_modulelist
    L = len(mlps)
    W1 = torch.cat([_lib.f32c(m.proj.weight) for m in mlps], 0)            # [L*E, C]
    W2 = torch.cat([_lib.f32c(m.proj_2.weight) for m in mlps], 0)          # [L*E, E]
    Wf = _lib.f32c(head.linear_fuse.weight).reshape(head.linear_fuse.weight.shape[0], -1)   # [E, L*E] (1x1 conv)
    s1, s2, sf = _pow2_scale(W1), _pow2_scale(W2), _pow2_scale(Wf)
    pack = dict(L=L, E=mlps[0].proj.weight.shape[0], C=W1.shape[1], s1=s1, s2=s2, sf=sf,
                W1=_split(W1, s1), W2=_split(W2, s2), Wf=_split(Wf, sf),
                b1=torch.cat([_lib.f32c(m.proj.bias) for m in mlps]).contiguous(),
                b2=torch.cat([_lib.f32c(m.proj_2.bias) for m in mlps]).contiguous(), bf=_lib.f32c(head.linear_fuse.bias))
    _HEADS[id(head)] = (ver, pack)
    return pack


def _gemm_split(As, lda, a_lo, a_rows_z, Bs, ldb, b_lo, b_rows_z, M, N, K, batch, alpha, act, bias, bias_z, C=None, ldc=0, c_z=0,
                Cs=None, lds=0, cs_lo=0, cs_z=0):
    _lib.call("excel_gemm_tc_split", _lib.ptr(As), lda, a_lo, a_rows_z, _lib.ptr(Bs), ldb, b_lo, b_rows_z, _lib.ptr(C), ldc, c_z,
              _lib.ptr(Cs), lds, cs_lo, cs_z, _lib.ptr(bias), bias_z, M, N, K, batch, float(alpha), act, _lib.stream())


def _p(t):
    return _lib.f32c(t)


@_lib.on_tensor_device
def segformer_head_tokens(head, feats):
    """head: the reference's SegFormerHead module (weights read in place); feats [L, M, C] token-major rows (any M).
    Returns the fused features [M, E] (row m = the 1x1-fused embedding of token m).

    Four launches on the tcgen05 engine: split of the features, the L first-layer GEMMs (+ReLU) as ONE batched launch
    writing the hidden state in split form, the L second-layer GEMMs as one batched launch writing straight into the split
    channel-concatenated matrix, and the 1x1 fuse GEMM -- no fp32 intermediate ever touches HBM."""
    L, M, Cin = feats.shape
    pk = _head_pack(head)
    if pk["L"] != L or pk["C"] != Cin:
        raise RuntimeError(f"SegFormerHead has {pk['L']} levels of {pk['C']} channels, got features {tuple(feats.shape)}")
    E, dev = pk["E"], feats.device
    Cp, Ep, LEp = _r64(Cin), _r64(E), _r64(L * E)                    # K extents in the engine's format (zero padded to 64)
    Fs = _split(_lib.f32c(feats).reshape(L * M, Cin))
    alloc = torch.empty if E % 64 == 0 else torch.zeros              # K padding columns of the intermediates must be zero
    H1 = alloc((L * M, 2 * Ep), dtype=torch.float16, device=dev)
    _gemm_split(Fs, 2 * Cp, Cp, M, pk["W1"], 2 * Cp, Cp, E, M, E, Cp, L, 1.0 / pk["s1"], 2, pk["b1"], E,
                Cs=H1, lds=2 * Ep, cs_lo=Ep, cs_z=M * 2 * Ep)                                        # Linear + ReLU (:22-24)
    cat = (torch.empty if (L * E) % 64 == 0 else torch.zeros)((M, 2 * LEp), dtype=torch.float16, device=dev)
    _gemm_split(H1, 2 * Ep, Ep, M, pk["W2"], 2 * Ep, Ep, E, M, E, Ep, L, 1.0 / pk["s2"], 0, pk["b2"], E,
                Cs=cat, lds=2 * LEp, cs_lo=LEp, cs_z=E)                                              # Linear (:25) -> cat (:74)
    out = torch.empty((M, E), dtype=torch.float32, device=dev)
    _gemm_split(cat, 2 * LEp, LEp, 0, pk["Wf"], 2 * LEp, LEp, 0, M, E, LEp, 1, 1.0 / pk["sf"], 0, pk["bf"], 0,
                C=out, ldc=E)                                                                        # 1x1 conv (:75)
    return out


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
