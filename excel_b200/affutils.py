"""SVC + pseudo-label refinement -- drop-ins for the reference's ``utils/affutils.py`` on sm_100a.

Per-image functions keep the reference signatures (``refine_cams_with_aff``,
``refine_cams_with_bkg_weclip``, ``compute_trans_mat``); ``refine_batch`` is the fused batched form the
bench and the pipeline use: one pass over the whole batch, no per-class host round trips (the reference
does one D2H sync + OpenCV call per present class, utils/affutils.py:207-208).
"""
import numpy as np
import torch

from . import _lib
from .par import par_labels, par_refine_planes


def _f32(t):
    return t.detach().to(torch.float32)


def _class_lists(cls_labels):
    """cls_labels [B,K] (any device) -> host list of per-image int64 class-index tensors
    (``torch.where(cls_label)[0]``, utils/affutils.py:203)."""
    host = cls_labels.detach().cpu()
    return [torch.where(row)[0].to(torch.int64) for row in host]


def _svc_vectors(attr_maps, attn, cls_lists, gh, gw, caa_thre, attn_layers, order=None, seg_attn=None, pair_index=None):
    """attr_maps [B,n_p,K]; attn [L,B,N,N] (arbitrary layer / image / row strides, columns contiguous: the encoder returns a
    view with row pitch round_up(N,4)).
    Returns refined [Q, n_p] for the Q = sum_b n_b (image, class) pairs, image-major (images in `order`, default
    0..B-1), classes ascending."""
    dev = attr_maps.device
    B, n_p, _ = attr_maps.shape
    L, _, N, _ = attn.shape
    if N - 1 != n_p or gh * gw != n_p:
        raise RuntimeError(f"SVC: attention has {N - 1} patches, CAM has {n_p}, grid {gh}x{gw}")
    if attn.stride(3) != 1 or attn.stride(2) < N:
        attn = attn.contiguous()
    attr_maps = _f32(attr_maps)
    if attr_maps.stride(2) != 1:
        attr_maps = attr_maps.contiguous()
    if pair_index is not None:        # (img_of, cls_of) int32 device views prepared by the caller (one pinned upload)
        img_of, cls_of = pair_index
        Q = int(img_of.numel())
    else:
        order = list(range(len(cls_lists))) if order is None else order
        io = np.asarray([b for b in order for _ in cls_lists[b]], dtype=np.int32)
        co = np.concatenate([cls_lists[b].numpy() for b in order]).astype(np.int32) if cls_lists else np.zeros(0, np.int32)
        Q = int(io.size)
        both = _lib.upload_ints(np.concatenate([io, co]), dev)
        img_of, cls_of = both[:Q], both[Q:]
    st = _lib.stream()
    A = torch.empty((B, n_p, n_p), dtype=torch.float32, device=dev)
    if seg_attn is None:
        _lib.call("excel_svc_mean_attention", _lib.ptr(attn), attn.stride(0), attn.stride(1), attn.stride(2), L, B, N, attn_layers,
                  _lib.ptr(A), st)
    else:   # LVC branch (utils/affutils.py:182-195): seg_attn [B, n_p, n_p]
        seg = _lib.f32c(seg_attn).reshape(B, n_p, n_p)
        dws = torch.empty((B * attn_layers,), dtype=torch.float32, device=dev)
        _lib.call("excel_svc_seg_attention", _lib.ptr(attn), attn.stride(0), attn.stride(1), attn.stride(2), L, B, N, attn_layers,
                  _lib.ptr(seg), _lib.ptr(dws), _lib.ptr(A), st)
    r = torch.empty((B, n_p), dtype=torch.float32, device=dev)
    c = torch.empty_like(r)
    _lib.call("excel_svc_sinkhorn", _lib.ptr(A), B, n_p, 3, _lib.ptr(r), _lib.ptr(c), st)
    v = torch.empty((Q, n_p), dtype=torch.float32, device=dev)
    _lib.call("excel_svc_box_mask", _lib.ptr(attr_maps), attr_maps.stride(0), attr_maps.stride(1), _lib.ptr(img_of),
              _lib.ptr(cls_of), Q, gh, gw, float(caa_thre), _lib.ptr(v), None, st)
    t1, t2, out = torch.empty_like(v), torch.empty_like(v), torch.empty_like(v)
    _lib.call("excel_svc_propagate", _lib.ptr(A), _lib.ptr(r), _lib.ptr(c), _lib.ptr(img_of), _lib.ptr(v), Q, n_p, 2,
              _lib.ptr(t1), _lib.ptr(t2), _lib.ptr(out), st)
    return out


@_lib.on_tensor_device
def compute_trans_mat(attn_weight):
    """utils/affutils.py:8-24 for one [n_p,n_p] attention matrix (materialises T@T; the pipeline does not)."""
    A = _lib.f32c(attn_weight)
    n_p = A.shape[0]
    st = _lib.stream()
    r = torch.empty((1, n_p), dtype=torch.float32, device=A.device)
    c = torch.empty_like(r)
    _lib.call("excel_svc_sinkhorn", _lib.ptr(A), 1, n_p, 3, _lib.ptr(r), _lib.ptr(c), st)
    T = torch.empty_like(A)
    _lib.call("excel_svc_build_trans", _lib.ptr(A), _lib.ptr(r), _lib.ptr(c), 1, n_p, _lib.ptr(T), st)
    T2 = torch.empty_like(A)
    _lib.call("excel_sgemm", _lib.ptr(T), _lib.ptr(T), _lib.ptr(T2), None, None, n_p, n_p, n_p, n_p, n_p, n_p, 1, 0, 0, 0,
              1.0, 0, 0, st)
    return T2


@_lib.on_tensor_device
def box_masks(attr_maps, cls_lists, gh, gw, caa_thre):
    """utils/affutils.py:26-53 + :209-212 on the device: [Q, gh, gw] masks for the (image, class) pairs."""
    dev = attr_maps.device
    attr_maps = _f32(attr_maps)
    if attr_maps.stride(2) != 1:
        attr_maps = attr_maps.contiguous()
    img_of = torch.tensor([b for b, c in enumerate(cls_lists) for _ in c], dtype=torch.int32, device=dev)
    cls_of = torch.cat(cls_lists).to(torch.int32).to(dev)
    Q = int(img_of.numel())
    v = torch.empty((Q, gh * gw), dtype=torch.float32, device=dev)
    m = torch.empty_like(v)
    _lib.call("excel_svc_box_mask", _lib.ptr(attr_maps), attr_maps.stride(0), attr_maps.stride(1), _lib.ptr(img_of),
              _lib.ptr(cls_of), Q, gh, gw, float(caa_thre), _lib.ptr(v), _lib.ptr(m), _lib.stream())
    return m.view(Q, gh, gw)


@_lib.on_tensor_device
def refine_cams_with_aff(attr_map, attn_weights, cls_label, size, caa_thre=0.79, attn_layers=6, seg_attn=None):
    """Same contract as utils/affutils.py:177-223: returns (list of n [h//16, w//16] CUDA tensors,
    int64 class indices on the CPU)."""
    h, w = size
    cls_lst = torch.where(cls_label)[0].detach().cpu()
    out = _svc_vectors(attr_map.unsqueeze(0), attn_weights.unsqueeze(1), [cls_lst], h // 16, w // 16, caa_thre, attn_layers,
                       seg_attn=seg_attn)
    return [o.view(h // 16, w // 16) for o in out], cls_lst


def _cams_to_planes(refined, counts, gh, gw, H, W, plane_off_dev=None):
    """refined [Q, gh*gw] -> (planes [P,H,W], plane_off int32 [B+1] device, plane_off host list)."""
    dev = refined.device
    B = len(counts)
    off = np.zeros(B + 1, dtype=np.int32)
    off[1:] = np.cumsum([c + 1 for c in counts])
    Q, P = int(sum(counts)), int(off[-1])
    plane_off = _lib.upload_ints(off, dev) if plane_off_dev is None else plane_off_dev
    planes = torch.empty((P, H, W), dtype=torch.float32, device=dev)
    ws = torch.empty((max(2 * Q, 1),), dtype=torch.float32, device=dev)
    _lib.call("excel_svc_cams_to_planes", _lib.ptr(refined), Q, gh, gw, _lib.ptr(plane_off), B, H, W, _lib.ptr(ws),
              _lib.ptr(planes), _lib.stream())
    return planes, plane_off, off


@_lib.on_tensor_device
def refine_cams_with_bkg_weclip(cam_refined_list, inputs_denorm, cls_lst, par, size):
    """Same contract as utils/affutils.py:161-174: (labels [1,H,W] int64, cams [C,H,W] fp32).
    ``size`` is (H, W) like at the reference's call sites (the reference unpacks it as ``w, h`` and swaps
    it back, :163-164)."""
    if len(cam_refined_list) == 0:
        raise RuntimeError("stack expects a non-empty TensorList")  # torch.stack([]) in the reference (:63)
    H, W = int(size[0]), int(size[1])
    gh, gw = cam_refined_list[0].shape
    refined = torch.stack([_f32(c).reshape(-1) for c in cam_refined_list], 0).contiguous()
    planes, plane_off, off = _cams_to_planes(refined, [len(cam_refined_list)], gh, gw, H, W)
    dev = planes.device
    key = _lib.upload_ints(np.concatenate([[0], cls_lst.to(torch.int64).cpu().numpy() + 1]).astype(np.int64), dev)   # :168
    out = par_refine_planes(inputs_denorm.unsqueeze(0), planes, plane_off, int(off[-1]), par.dilations, par.num_iter,
                            getattr(par, "group", 0), par.w1, par.w2)
    labels = par_labels(out, plane_off, key, 1)
    return labels, planes


def _segments(counts):
    """Runs of equal plane count (counts sorted ascending) -> [(b0, b1, planes_per_image)]; everything with more than
    4 planes shares the last run (the PAR kernel stages at most 4 planes per pass)."""
    segs, b0 = [], 0
    for b in range(1, len(counts) + 1):
        if b == len(counts) or min(counts[b], 5) != min(counts[b0], 5):
            segs.append((b0, b, max(counts[b0:b])))
            b0 = b
    return segs


@_lib.on_tensor_device
def refine_batch(attr_maps, attn_weights, cls_labels, par_imgs, par, out_size=None, caa_thre=0.79, attn_layers=6,
                 return_cams=False, cls_lists=None):
    """Fused batched SVC + PAR + argmax (tools/infer_lam.py:88-94 for the whole batch).
    attr_maps [B,n_p,K], attn_weights [L,B,N,N], cls_labels [B,K] (any device), par_imgs [B,3,h,w].
    Returns labels [B,H,W] int64 (and the packed planes, plane_off, refined CAMs when return_cams).

    Images are processed in order of their plane count so that PAR launches run the kernel variant that fits
    (2, 3 or 4 planes staged per pass); the labels come back in the caller's order."""
    B = attr_maps.shape[0]
    h, w = par_imgs.shape[-2:]
    H, W = (h, w) if out_size is None else (int(out_size[0]), int(out_size[1]))
    gh, gw = h // 16, w // 16
    # host-side class lists: pass `cls_lists` (or CPU labels) to keep the device stream free of a D2H sync here
    cls_lists = _class_lists(cls_labels) if cls_lists is None else cls_lists
    counts = [int(c.numel()) for c in cls_lists]
    if min(counts) == 0:
        raise RuntimeError("stack expects a non-empty TensorList")  # an image without classes (affutils.py:63)
    order = list(range(B)) if return_cams else sorted(range(B), key=lambda b: counts[b])
    identity = order == list(range(B))
    counts_s = [counts[b] for b in order]
    dev = attr_maps.device
    # every index array of the step in ONE pinned upload: (image, class) of the Q vector slots, plane offsets of the B
    # slots, the slot -> image order; the int64 plane keys in a second one
    io = np.asarray([b for b in order for _ in cls_lists[b]], dtype=np.int32)
    co = np.concatenate([cls_lists[b].numpy() for b in order]).astype(np.int32)
    off = np.zeros(B + 1, dtype=np.int32)
    off[1:] = np.cumsum([c + 1 for c in counts_s])
    Q = int(io.size)
    idx = _lib.upload_ints(np.concatenate([io, co, off, np.asarray(order, dtype=np.int32)]), dev)
    img_of, cls_of, plane_off, order_dev = idx[:Q], idx[Q:2 * Q], idx[2 * Q:2 * Q + B + 1], idx[2 * Q + B + 1:]
    key = _lib.upload_ints(np.concatenate([np.concatenate([[0], cls_lists[b].numpy() + 1]) for b in order]).astype(np.int64), dev)
    refined = _svc_vectors(attr_maps, attn_weights, cls_lists, gh, gw, caa_thre, attn_layers, order, pair_index=(img_of, cls_of))
    planes, plane_off, off = _cams_to_planes(refined, counts_s, gh, gw, H, W, plane_off_dev=plane_off)
    segs = _segments([c + 1 for c in counts_s])
    # the PAR runs read the images through the slot -> image index (no gathered copy of the batch) and the labels land in
    # the caller's order
    out = par_refine_planes(par_imgs, planes, plane_off, max(counts) + 1, par.dilations, par.num_iter,
                            getattr(par, "group", 0), par.w1, par.w2, segments=segs, img_index=None if identity else order_dev)
    labels = par_labels(out, plane_off, key, B, out_index=None if identity else order_dev)
    if return_cams:
        return labels, planes, plane_off, refined
    return labels
