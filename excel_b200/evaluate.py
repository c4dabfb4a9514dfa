"""Device-side confusion histogram and mIoU (reference: utils/evaluate.py:9-50) + the cross-rank reduction."""
import numpy as np
import torch

from . import _lib


@_lib.on_tensor_device
def confusion_hist(label_true, label_pred, num_classes, hist=None):
    """Accumulate utils/evaluate.py:_fast_hist on the device.  int64 CUDA tensors of equal numel;
    returns / updates hist [num_classes, num_classes] int64."""
    lt = label_true.to(torch.int64).contiguous()
    lp = label_pred.to(torch.int64).contiguous()
    if lt.numel() != lp.numel():
        raise RuntimeError("confusion_hist: label_true and label_pred differ in size")
    if hist is None:
        hist = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=lp.device)
    _lib.call("excel_confusion_hist", _lib.ptr(lt), _lib.ptr(lp), lt.numel(), num_classes, _lib.ptr(hist), _lib.stream())
    return hist


def all_reduce_hist(hist):
    """The path's only collective: sum the per-rank histograms (NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def scores_from_hist(hist):
    """utils/evaluate.py:21-29: pixel accuracy and mean IoU over classes present in the ground truth."""
    h = hist.detach().cpu().numpy().astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = np.diag(h) / (h.sum(1) + h.sum(0) - np.diag(h))
        acc = np.diag(h).sum() / h.sum()
    valid = h.sum(1) > 0
    return {"pAcc": float(acc), "miou": float(np.nanmean(iu[valid])) if valid.any() else float("nan")}


def shard_slice(global_batch, rank, world):
    """Contiguous slice of a global batch owned by `rank` (cfg3's 64 images over 8 ranks -> 8 x 8); the first
    `global_batch % world` ranks take one extra image."""
    per, extra = divmod(global_batch, world)
    lo = rank * per + min(rank, extra)
    return slice(lo, lo + per + (1 if rank < extra else 0))


def shard_indices(n, rank, world):
    """tools/infer_lam.py:166: rank r takes images r, r+W, r+2W, ..."""
    return list(range(rank, n, world))
