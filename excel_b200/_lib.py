"""ctypes binding of the C-ABI library (include/excel_b200.h).

The product path is CUDA only: if libexcel_b200.so is missing or a call fails this raises -- there is
no CPU or PyTorch fallback behind any shim in this package.
"""
import ctypes
import functools
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libexcel_b200.so")

_c = ctypes
_i, _i64, _f, _p = _c.c_int, _c.c_int64, _c.c_float, _c.c_void_p

# name -> argtypes; must list every symbol declared in include/excel_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "excel_last_error": ([], _c.c_char_p),
    "excel_version": ([], _i),
    "excel_device_arch": ([_i], _i),
    "excel_launch_count": ([], _i64),
    "excel_confusion_hist": ([_p, _p, _i64, _i, _p, _p], _i),
    "excel_par_forward": ([_p, _i64, _i64, _i64, _i, _i, _i, _i, _i, _p, _i, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p], _i),
    "excel_par_labels": ([_p, _p, _p, _p, _i, _i, _i, _p, _p], _i),
    "excel_svc_mean_attention": ([_p, _i64, _i64, _i64, _i, _i, _i, _i, _p, _p], _i),
    "excel_svc_seg_attention": ([_p, _i64, _i64, _i64, _i, _i, _i, _i, _p, _p, _p, _p], _i),
    "excel_svc_sinkhorn": ([_p, _i, _i, _i, _p, _p, _p], _i),
    "excel_svc_build_trans": ([_p, _p, _p, _i, _i, _p, _p], _i),
    "excel_svc_box_mask": ([_p, _i64, _i64, _p, _p, _i, _i, _i, _c.c_double, _p, _p, _p], _i),
    "excel_svc_propagate": ([_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p], _i),
    "excel_svc_cams_to_planes": ([_p, _i, _i, _i, _p, _i, _i, _i, _p, _p, _p], _i),
    "excel_token_normalize": ([_p, _i, _i, _i, _p, _p, _p], _i),
    "excel_cam_workspace_bytes": ([_i, _i, _i, _i], _i64),
    "excel_cam_surgery": ([_p, _p, _i, _i, _i, _i, _p, _i64, _p, _p], _i),
    "excel_flip_merge": ([_p, _i, _i, _i, _i, _p, _p], _i),
    "excel_vit_workspace_bytes": ([_i, _i, _i, _i, _i], _i64),
    "excel_split_f16": ([_p, _i64, _i, _i, _i, _f, _p, _p], _i),
    "excel_vit_forward": ([_p, _p, _i64, _i64, _i64, _i, _i, _p, _i64, _p, _p, _i64, _p, _p, _p], _i),
    "excel_lvc_attention": ([_p, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p], _i),
    "excel_attn_pred": ([_p, _i, _i, _i, _f, _f, _p, _p, _p, _p, _p], _i),
    "excel_row_softmax": ([_p, _i, _i, _p, _p], _i),
    "excel_row_l2_normalize": ([_p, _i, _i, _p, _p], _i),
    "excel_gemm_tc": ([_p, _p, _p, _p, _p, _i, _i, _i, _i64, _i64, _i64, _f, _i, _p, _i64, _p], _i),
    "excel_gemm_tc_split": ([_p, _i64, _i, _i64, _p, _i64, _i, _i64, _p, _i64, _i64, _p, _i64, _i, _i64, _p, _i64, _i, _i, _i, _i, _f, _i, _p], _i),
    "excel_seg_accumulate": ([_p, _i, _i, _i, _i, _p, _i, _i, _i, _f, _p], _i),
    "excel_seg_argmax": ([_p, _i, _i, _i, _i, _i, _p, _p], _i),
    "excel_radius_mask": ([_i, _i, _i, _p, _p], _i),
    "excel_affinity_label": ([_p, _i, _i, _i, _i, _i, _p, _i64, _p, _p], _i),
    "excel_lam_to_label": ([_p, _p, _i, _i, _i, _i, _f, _f, _f, _i, _i64, _p, _p, _p], _i),
    "excel_sgemm": ([_p, _p, _p, _p, _p, _i, _i, _i, _i64, _i64, _i64, _i, _i64, _i64, _i64, _f, _i, _i, _p], _i),
}



class VitLayer(ctypes.Structure):
    _fields_ = [(n, _p) for n in ("ln1_w", "ln1_b", "in_w", "in_b", "out_w", "out_b", "ln2_w", "ln2_b", "fc_w", "fc_b",
                                  "proj_w", "proj_b", "in_ws", "out_ws", "fc_ws", "proj_ws")] + \
               [(n, _f) for n in ("in_scale", "out_scale", "fc_scale", "proj_scale")]


class VitWeights(ctypes.Structure):
    _fields_ = [(n, _i) for n in ("layers", "width", "heads", "patch", "embed", "grid0", "n_surgery")] + \
               [(n, _p) for n in ("conv1", "cls", "pos", "ln_pre_w", "ln_pre_b", "ln_post_w", "ln_post_b", "proj", "conv1_s",
                                  "proj_t_s")] + \
               [("conv1_scale", _f), ("proj_t_scale", _f), ("blocks", ctypes.POINTER(VitLayer))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"excel_b200: {LIB_PATH} is missing -- build it with `python -m excel_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
        _lib = L
    return _lib


def call(name, *args):
    """Call an int-returning entry point; non-zero -> RuntimeError with the library's message."""
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {L.excel_last_error().decode()}")


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("excel_b200: expected a CUDA tensor (the hot path has no CPU implementation)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def _first_cuda_tensor(values):
    for a in values:
        if torch.is_tensor(a):
            if a.is_cuda:
                return a
        elif isinstance(a, (list, tuple)) and a and torch.is_tensor(a[0]) and a[0].is_cuda:
            return a[0]
    return None


def on_tensor_device(fn):
    """Decorator of the public shims: run `fn` with the device of its first CUDA-tensor argument current, so that
    `stream()` / the kernel launches / the function attributes land on the tensors' device, not on whatever device the
    caller left current (tensors on cuda:1 with cuda:0 current)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(list(args) + list(kwargs.values()))
        if t is not None and t.device.index != torch.cuda.current_device():
            with torch.cuda.device(t.device):
                return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def f32c(t):
    """fp32 + contiguous, like the reference's `.float()` calls (no copy when already so)."""
    return t.detach().to(torch.float32).contiguous()


class _PinnedRing:
    """Small ring of pinned host staging buffers per device: index arrays (a few hundred bytes per call) go to the device
    as ONE asynchronous copy from pinned memory instead of several blocking copies from pageable memory.  A slot is reused
    only after the copy that last read it has completed (event)."""

    def __init__(self, slots=8):
        self.slots = [None] * slots
        self.i = 0

    def upload(self, arr, dev):
        """arr: 1-D numpy int32 / int64 array -> device tensor (same dtype); asynchronous on the current stream."""
        import numpy as np
        t_dtype = torch.int32 if arr.dtype == np.int32 else torch.int64
        k = self.i % len(self.slots)
        self.i += 1
        slot = self.slots[k]
        nbytes = max(int(arr.nbytes), 8)
        if slot is None or slot[0].numel() < nbytes:
            slot = [torch.empty(max(nbytes, 4096), dtype=torch.uint8).pin_memory(), None]
            self.slots[k] = slot
        elif slot[1] is not None:
            slot[1].synchronize()
        host = slot[0][:arr.nbytes].view(t_dtype)
        host.copy_(torch.from_numpy(arr))
        out = torch.empty(arr.shape[0], dtype=t_dtype, device=dev)
        out.copy_(host, non_blocking=True)
        slot[1] = torch.cuda.Event()
        slot[1].record(torch.cuda.current_stream(dev))
        return out


_RINGS = {}


def upload_ints(arr, dev):
    dev = torch.device(dev)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ring = _RINGS.get(key)
    if ring is None:
        ring = _RINGS[key] = _PinnedRing()
    return ring.upload(arr, dev)


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])
