"""Multi-scale + flip segmentation inference -- the loop body of the reference's ``tools/infer_seg_voc.py::_validate``
(:56-88; same in ``tools/infer_seg_coco.py``) on sm_100a.

``multi_scale_flip_seg`` runs the (install()-patched or stand-alone) model once per scale on ``cat[x, flip(x)]`` and merges
the per-scale logits with two small kernels (``excel_seg_accumulate`` / ``excel_seg_argmax``) instead of the reference's
up-sample / flip / stack / mean / resize / argmax chain of full-size temporaries."""
import torch
import torch.nn.functional as F

from . import _lib


@_lib.on_tensor_device
def merge_scales(seg_list, size, base_index=0):
    """seg_list: per scale the model output ``segs`` [2,C,g,g] for ``cat[x, flip(x)]``; entry ``base_index`` is the base
    scale (only its un-flipped half is used, tools/infer_seg_voc.py:69-72), the others are flip-merged (:78-80).
    Returns the mean over the scales [1,C,h,w] (:83)."""
    h, w = int(size[0]), int(size[1])
    n = len(seg_list)
    if n == 0:
        raise RuntimeError("merge_scales: no scales")
    C = seg_list[0].shape[1]
    acc = torch.empty((1, C, h, w), dtype=torch.float32, device=seg_list[0].device)
    for i, seg in enumerate(seg_list):
        seg = _lib.f32c(seg)
        if seg.dim() != 4 or seg.shape[0] != 2 or seg.shape[1] != C:
            raise RuntimeError(f"merge_scales: expected [2,{C},g,g] logits per scale, got {tuple(seg.shape)}")
        _lib.call("excel_seg_accumulate", _lib.ptr(seg), C, seg.shape[2], seg.shape[3], int(i != base_index), _lib.ptr(acc), h, w,
                  int(i == 0), 1.0 / n if i == n - 1 else 1.0, _lib.stream())
    return acc


@_lib.on_tensor_device
def seg_argmax(segs, size):
    """tools/infer_seg_voc.py:85-86: argmax over the classes of the logits [1,C,h,w] resized to ``size`` -> [1,H,W] int64."""
    segs = _lib.f32c(segs)
    _, C, h, w = segs.shape
    H, W = int(size[0]), int(size[1])
    labels = torch.empty((1, H, W), dtype=torch.int64, device=segs.device)
    _lib.call("excel_seg_argmax", _lib.ptr(segs), C, h, w, H, W, _lib.ptr(labels), _lib.stream())
    return labels


@torch.no_grad()
def multi_scale_flip_seg(model, inputs, scales=(0.7, 1.0, 1.2, 1.5), resize_size=320, label_size=None):
    """tools/infer_seg_voc.py:58-86 for one image ``inputs`` [1,3,h,w]: returns (segs [1,C,h,w] multi-scale mean logits,
    labels [1,H,W] int64).  ``model(x)[0]`` must be the segmentation logits (ExCEL_model.forward)."""
    _, _, h, w = inputs.shape
    sizes = [resize_size] + [int(resize_size * sc) for sc in scales if sc != 1.0]      # base scale first (:66, :74-76)
    seg_list = []
    for s in sizes:
        x = F.interpolate(inputs, size=[s, s], mode="bilinear", align_corners=False)
        seg_list.append(model(torch.cat([x, x.flip(-1)], dim=0))[0])
    segs = merge_scales(seg_list, (h, w), base_index=0)
    return segs, seg_argmax(segs, (h, w) if label_size is None else label_size)
