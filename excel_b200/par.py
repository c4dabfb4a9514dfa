"""PAR -- drop-in for the reference's ``utils.PAR.PAR`` (utils/PAR.py:26-92) on sm_100a.

Same constructor and ``forward(imgs, masks)`` signature; the arithmetic runs in
``excel_par_forward`` (include/excel_b200.h).  ``refine_planes`` is the batched, ragged form used by
the fused pipeline (per-image channel counts differ: background + present classes).
"""
import torch
from torch import nn

from . import _lib

W1, W2 = 0.3, 0.01  # utils/PAR.py:36-37


def _pitch(w):
    return (w + 3) & ~3


_SIDE = {}


def _side_streams(dev, n):
    """n reusable side streams of device `dev` (created once)."""
    pool = _SIDE.setdefault(dev.index if dev.index is not None else torch.cuda.current_device(), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


@_lib.on_tensor_device
def par_refine_planes(imgs, planes, plane_off, max_c, dilations, num_iter, group=0, w1=W1, w2=W2, segments=None, img_index=None):
    """imgs [B,3,hi,wi]; planes [P,H,W] packed mask planes; plane_off int32 [B+1] (device).
    segments: optional [(b0, b1, max planes per image)] runs of consecutive SLOTS launched separately (so that each
    run gets the kernel variant matching its plane count).  img_index: optional int32 [B] (device): slot b reads image
    img_index[b] (the batch processed in another order without gathering the images).
    Returns the refined planes [P,H,W] (a new tensor)."""
    _lib.ptr(imgs), _lib.ptr(planes)   # CUDA tensors only: raises for CPU inputs (no fallback)
    imgs = imgs.float()
    if imgs.stride(-1) != 1:
        imgs = imgs.contiguous()
    planes = _lib.f32c(planes)
    _, _, hi, wi = imgs.shape
    B = plane_off.numel() - 1
    P, H, W = planes.shape
    if num_iter <= 0 or P == 0:
        return planes.clone()
    K = 8 * len(dilations)
    segments = [(0, B, int(max_c))] if not segments else segments
    dev = planes.device
    out = torch.empty_like(planes)
    tmp = torch.empty_like(planes) if num_iter > 1 else None
    dil = _lib.int_array(dilations)
    # The runs touch disjoint images and planes: each gets its own affinity workspace and (beyond the first) its own
    # side stream, forked from / joined to the caller's stream with events, so that the short launches of a small run
    # fill the tail of a large one instead of queueing behind it.
    cur = torch.cuda.current_stream(dev)
    side = _side_streams(dev, len(segments) - 1)
    for i, (b0, b1, mc) in enumerate(segments):
        nb = b1 - b0
        g = nb if group <= 0 else min(group, nb)
        st = cur if i == 0 else side[i - 1]
        if st is not cur:
            st.wait_stream(cur)
        with torch.cuda.stream(st):
            aff = torch.empty((g, K, H, _pitch(W)), dtype=torch.float32, device=dev)   # internal layout: row pitch % 4 == 0
            rs = torch.empty((nb, 3, H, W), dtype=torch.float32, device=dev) if (hi, wi) != (H, W) else None
            img_ptr = _lib.ptr(imgs) + (0 if img_index is not None else b0 * imgs.stride(0) * 4)
            idx_ptr = None if img_index is None else _lib.ptr(img_index) + 4 * b0
            _lib.call("excel_par_forward", img_ptr, imgs.stride(0), imgs.stride(1),
                      imgs.stride(2), nb, hi, wi, H, W, dil, len(dilations), w1, w2, num_iter, g, _lib.ptr(rs), _lib.ptr(aff),
                      _lib.ptr(planes), _lib.ptr(out), _lib.ptr(tmp), _lib.ptr(plane_off) + 4 * b0, P, int(mc), idx_ptr, _lib.stream())
    for st in side[:len(segments) - 1]:
        cur.wait_stream(st)
    return out


@_lib.on_tensor_device
def par_affinity(imgs, size, dilations, w1=W1, w2=W2):
    """utils/PAR.py:67-86 only: imgs [B,3,hi,wi] -> aff [B,8*n_dil,H,W]."""
    imgs = imgs.float()
    if imgs.stride(-1) != 1:
        imgs = imgs.contiguous()
    B, _, hi, wi = imgs.shape
    H, W = size
    aff = torch.empty((B, 8 * len(dilations), H, _pitch(W)), dtype=torch.float32, device=imgs.device)
    rs = torch.empty((B, 3, H, W), dtype=torch.float32, device=imgs.device) if (hi, wi) != (H, W) else None
    _lib.call("excel_par_forward", _lib.ptr(imgs), imgs.stride(0), imgs.stride(1), imgs.stride(2), B, hi, wi, H, W,
              _lib.int_array(dilations), len(dilations), w1, w2, 0, B, _lib.ptr(rs), _lib.ptr(aff),
              None, None, None, None, 0, 0, None, _lib.stream())
    return aff[..., :W]


@_lib.on_tensor_device
def par_labels(planes, plane_off, plane_key, B, out_index=None):
    """utils/affutils.py:86-87: labels [B,H,W] int64 = plane_key[argmax over the planes of slot b], written to
    labels[out_index[b]] (int32 [B], device) or labels[b]."""
    P, H, W = planes.shape
    labels = torch.empty((B, H, W), dtype=torch.int64, device=planes.device)
    _lib.call("excel_par_labels", _lib.ptr(planes), _lib.ptr(plane_off), _lib.ptr(plane_key), _lib.ptr(labels),
              B, H, W, _lib.ptr(out_index), _lib.stream())
    return labels


class PAR(nn.Module):
    """Same interface as utils/PAR.py:26 -- ``PAR(dilations, num_iter)(imgs, masks)``."""

    def __init__(self, dilations, num_iter, group=0):
        super().__init__()
        self.dilations = list(dilations)
        self.num_iter = num_iter
        self.group = group          # images per launch group (L2 residency of the affinity planes)
        self.dim = 2
        self.w1 = W1
        self.w2 = W2
        # kept for state_dict compatibility with the reference module (utils/PAR.py:31-32)
        kernel = torch.zeros(8, 1, 3, 3)
        for i, (r, c) in enumerate(((0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2))):
            kernel[i, 0, r, c] = 1
        self.register_buffer("kernel", kernel)

    @torch.no_grad()
    def forward(self, imgs, masks):
        b, c, h, w = masks.shape
        if imgs.shape[0] != b:
            raise RuntimeError(f"PAR: batch mismatch imgs {tuple(imgs.shape)} vs masks {tuple(masks.shape)}")
        off = torch.arange(0, (b + 1) * c, c, dtype=torch.int32, device=masks.device)
        out = par_refine_planes(imgs, masks.reshape(b * c, h, w), off, c, self.dilations, self.num_iter,
                                self.group, self.w1, self.w2)
        return out.view(b, c, h, w)
