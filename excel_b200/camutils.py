"""CAM wrappers -- drop-ins for the live functions of the reference's ``utils/camutils.py``."""
import torch

from . import _lib


def cure_attr_map(model, inputs, ex_feats):
    """utils/camutils.py:93-97."""
    with torch.no_grad():
        return model(inputs, ex_feats=ex_feats)


@_lib.on_tensor_device
def merge_flipped_maps(attr_2b, b, gh, gw):
    """utils/camutils.py:19-26: element-max of the maps of x and flip(x) (un-flipped), per-(b,c) min subtracted,
    divided by (max + 1e-5).  attr_2b [2b, n_p, K] -> [b, n_p, K]."""
    x = _lib.f32c(attr_2b)
    if x.shape[0] != 2 * b or x.shape[1] != gh * gw:
        raise RuntimeError(f"merge_flipped_maps: expected [{2 * b}, {gh * gw}, K], got {tuple(x.shape)}")
    out = torch.empty((b, gh * gw, x.shape[2]), dtype=torch.float32, device=x.device)
    _lib.call("excel_flip_merge", _lib.ptr(x), b, gh, gw, x.shape[2], _lib.ptr(out), _lib.stream())
    return out


def cure_attr_map_flip(model, inputs, ex_fts=True, flip=True, raw_fts=None):
    """utils/camutils.py:8-30."""
    b, c, h, w = inputs.shape
    with torch.no_grad():
        if not flip:
            return model(inputs, ex_feats=raw_fts)
        inputs_cat = torch.cat([inputs, inputs.flip(-1)], dim=0)
        if ex_fts:
            ex_feats = model(inputs_cat)[1]
            attr = model(inputs_cat, ex_feats=ex_feats)
        else:
            attr = model(inputs_cat)[2]
        return merge_flipped_maps(attr, b, h // 16, w // 16)


def get_mask_by_radius(h=20, w=20, radius=8, device="cuda"):
    """utils/camutils.py:459-476, built on the device (the reference loops in Python and returns numpy):
    [h*w, h*w] fp32 CUDA tensor, 1 inside the (2r+1)^2 window."""
    mask = torch.empty((h * w, h * w), dtype=torch.float32, device=device)
    with torch.cuda.device(mask.device):
        _lib.call("excel_radius_mask", h, w, radius, _lib.ptr(mask), _lib.stream())
    return mask


@_lib.on_tensor_device
def cams_to_affinity_label(cam_label, mask=None, ignore_index=255):
    """utils/camutils.py:438-457: cam_label [b,h,w] -> [b, (h//16)*(w//16), (h//16)*(w//16)] int64."""
    b, h, w = cam_label.shape
    lab = cam_label.to(torch.int64).contiguous()
    if mask is not None and not torch.is_tensor(mask):
        mask = torch.as_tensor(mask)
    m = None if mask is None else mask.to(lab.device, torch.float32).contiguous()
    n = (h // 16) * (w // 16)
    out = torch.empty((b, n, n), dtype=torch.int64, device=lab.device)
    _lib.call("excel_affinity_label", _lib.ptr(lab), b, h, w, h // 16, w // 16, _lib.ptr(m), int(ignore_index), _lib.ptr(out),
              _lib.stream())
    return out


@_lib.on_tensor_device
def lam_to_label(cam, cls_label, img_box=None, bkg_thre=0.5, high_thre=None, low_thre=None, ignore_mid=False, ignore_index=None):
    """utils/camutils.py:123-143: (valid_cam [b,c,h,w], pseudo_label [b,h,w] int64)."""
    b, c, h, w = cam.shape
    cam = _lib.f32c(cam)
    cls = _lib.f32c(cls_label)
    valid = torch.empty_like(cam)
    label = torch.empty((b, h, w), dtype=torch.int64, device=cam.device)
    _lib.call("excel_lam_to_label", _lib.ptr(cam), _lib.ptr(cls), b, c, h, w, float(bkg_thre),
              float(high_thre if high_thre is not None else 0.0), float(low_thre if low_thre is not None else 0.0),
              int(bool(ignore_mid)), int(ignore_index if ignore_index is not None else 0), _lib.ptr(valid), _lib.ptr(label),
              _lib.stream())
    if img_box is not None:   # :137-140
        boxed = torch.ones_like(label) * ignore_index
        for idx, coord in enumerate(img_box):
            boxed[idx, coord[0]:coord[1], coord[2]:coord[3]] = label[idx, coord[0]:coord[1], coord[2]:coord[3]]
        label = boxed
    return valid, label
