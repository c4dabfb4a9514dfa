"""CLIP-surgery call surface -- drop-ins for ``clip.clip_feature_surgery`` / ``clip.generate_clip_fts``
(clip/clip.py:288-310, 348-358) on sm_100a."""
import torch

from . import _lib


@_lib.on_tensor_device
def clip_feature_surgery(image_features, text_features, redundant_feats=None, t=2):
    """clip/clip.py:288-310: image_features [B,N,E], text_features [T,E] -> [B,N,T] (detached)."""
    if redundant_feats is not None:
        raise NotImplementedError("excel_b200: clip_feature_surgery(redundant_feats=...) has no caller in ExCEL")
    F = _lib.f32c(image_features)
    T = _lib.f32c(text_features)
    B, N, E = F.shape
    if T.shape[1] != E:
        raise RuntimeError(f"clip_feature_surgery: feature dim {E} vs text dim {T.shape[1]}")
    out = torch.empty((B, N, T.shape[0]), dtype=torch.float32, device=F.device)
    nbytes = _lib.lib().excel_cam_workspace_bytes(B, N, E, T.shape[0])
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=F.device)      # split operands + S + min/max partials
    _lib.call("excel_cam_surgery", _lib.ptr(F), _lib.ptr(T), B, N, E, T.shape[0], _lib.ptr(ws), nbytes, _lib.ptr(out), _lib.stream())
    return out


@_lib.on_tensor_device
def token_normalize(tok):
    """clip/clip.py:353: tok / tok.norm(dim=1, keepdim=True) for tok [B,N,E]."""
    tok = _lib.f32c(tok)
    B, N, E = tok.shape
    ws = torch.empty((B, E), dtype=torch.float32, device=tok.device)
    out = torch.empty_like(tok)
    _lib.call("excel_token_normalize", _lib.ptr(tok), B, N, E, _lib.ptr(ws), _lib.ptr(out), _lib.stream())
    return out
