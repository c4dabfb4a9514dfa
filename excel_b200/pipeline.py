"""The whole hot path for a batch: CLIP-surgery ViT -> patch x text CAM -> SVC -> PAR -> pseudo labels
(tools/infer_lam.py:74-94, training-free branch; model/model_excel.py:48-58)."""
import torch

from . import affutils
from .clip import clip_feature_surgery, token_normalize
from .encoder import SurgeryViT
from .par import PAR

PAR_DILATIONS = (1, 2, 4, 8, 12, 24)   # scripts/train_voc.py:112, tools/infer_lam.py:168
PAR_ITERS = 20


class ExCELHotPath:
    """encoder: SurgeryViT; text_attr_t [T,E] (= ExCEL_model.text_attr.permute(1,0), model/model_excel.py:58);
    num_fg = num_classes - 1."""

    def __init__(self, encoder, text_attr_t, num_fg, caa_thre=0.79, par=None):
        self.encoder = encoder
        self.text = text_attr_t.to(encoder.device, torch.float32).contiguous()
        self.num_fg = num_fg
        self.caa_thre = caa_thre
        self.par = par if par is not None else PAR(PAR_DILATIONS, PAR_ITERS)

    @torch.no_grad()
    def cams(self, imgs):
        """ExCEL_model.forward up to attr_maps_raw (model/model_excel.py:55-58):
        (attr_maps_raw [B,n_p,num_fg], attn_weights [L,B,N,N], all_feats [L,B,N,D])."""
        tokens, attn, feats = self.encoder(imgs)
        attr = clip_feature_surgery(token_normalize(tokens), self.text)[:, 1:, :self.num_fg]
        return attr, attn, feats

    @torch.no_grad()
    def __call__(self, imgs, cls_labels, par_imgs=None, out_size=None):
        """imgs [B,3,S,S] (normalised), cls_labels [B,num_fg] one-hot (CPU or CUDA) -> labels [B,H,W] int64.
        The image-level labels steer host logic (which planes exist), so they are read on the host BEFORE the
        encoder is enqueued: with CPU labels the device stream never waits on the host inside a step."""
        cls_lists = affutils._class_lists(cls_labels)
        attr, attn, _ = self.cams(imgs)
        par_imgs = imgs if par_imgs is None else par_imgs
        return affutils.refine_batch(attr, attn, cls_labels, par_imgs, self.par, out_size, self.caa_thre, cls_lists=cls_lists)


class HostPipeline:
    """End-to-end driver for batches that live in HOST memory: `submit(imgs_host, cls_host)` enqueues the H2D copy of
    the batch on a copy stream, the hot path on the compute stream and the D2H copy of the int64 labels on a second copy
    stream, and returns the PREVIOUS batch's labels (pinned host tensor, complete) -- so the copies of batch i+1 / i-1
    overlap the kernels of batch i (a returned tensor stays valid for the next depth-2 submits).  `flush()` returns the
    last batch's labels.  (The reference moves each image
    with a blocking `.cuda()` / `.cpu()`, tools/infer_lam.py:76-77,113-114.)"""

    def __init__(self, hot_path, depth=3):
        self.hp = hot_path
        self.dev = hot_path.encoder.device
        self.depth = depth
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)
        self.slots = [None] * depth      # per slot: device image buffer, pinned label buffer, events
        self.i = 0
        self.pending = None              # (pinned labels, done event) of the previous batch
        self.staged = None               # (slot, ready event, cls_host) of a batch whose H2D copy is already enqueued

    def _slot(self, k, imgs_host, B, H, W):
        s = self.slots[k]
        if s is None or s["img"].shape != imgs_host.shape:
            s = {"img": torch.empty(imgs_host.shape, dtype=torch.float32, device=self.dev),
                 "lab": torch.empty((B, H, W), dtype=torch.int64).pin_memory(),
                 "free": None, "lab_free": None}
            self.slots[k] = s
        return s

    def stage(self, imgs_host, cls_host):
        """Enqueue the H2D copy of the NEXT batch (call before `submit` of the current one to overlap them)."""
        k = self.i % self.depth
        B, _, H, W = imgs_host.shape
        s = self._slot(k, imgs_host, B, H, W)
        if s["free"] is not None:
            self.s_in.wait_event(s["free"])          # the compute that last read this buffer has finished
        with torch.cuda.stream(self.s_in):
            s["img"].copy_(imgs_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.s_in)
        self.staged = (k, ready, cls_host)
        self.i += 1

    def submit(self, imgs_host=None, cls_host=None, stage_next=None):
        """Run the batch staged earlier (or `imgs_host, cls_host` now).  `stage_next=(imgs, cls)` enqueues the next
        batch's H2D copy before this batch's kernels.  Returns the previous batch's labels (or None)."""
        if self.staged is None:
            self.stage(imgs_host, cls_host)
        k, ready, cls = self.staged
        self.staged = None
        s = self.slots[k]
        if stage_next is not None:
            self.stage(*stage_next)
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ready)
        labels = self.hp(s["img"], cls)
        s["free"] = torch.cuda.Event()
        s["free"].record(cur)
        done = torch.cuda.Event()
        self.s_out.wait_event(s["free"])
        labels.record_stream(self.s_out)
        with torch.cuda.stream(self.s_out):
            s["lab"].copy_(labels, non_blocking=True)
            done.record(self.s_out)
        prev = self.pending
        self.pending = (s["lab"], done)
        if prev is not None:
            prev[1].synchronize()
            return prev[0]
        return None

    def flush(self):
        prev, self.pending = self.pending, None
        if prev is None:
            return None
        prev[1].synchronize()
        return prev[0]
