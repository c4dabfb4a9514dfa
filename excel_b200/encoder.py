"""CLIP-surgery vision encoder on sm_100a -- the engine behind ``clip.generate_clip_fts``
(clip/clip.py:348-358; VisionTransformer.forward, clip/clip_surgery_model.py:418-448).

``SurgeryViT`` owns a device copy of the frozen encoder weights (a plain dict of fp32 tensors) and calls
``excel_vit_forward``; ``from_visual`` builds it from a reference / OpenAI-CLIP style ``VisionTransformer``
module, so the reference's ``ExCEL_model.encoder`` can be used as is.
"""
import ctypes
import math

import torch

from . import _lib
from .clip import token_normalize

_BLOCK_FIELDS = (("ln1_w", "ln_1.weight"), ("ln1_b", "ln_1.bias"), ("in_w", "in_proj_weight"), ("in_b", "in_proj_bias"),
                 ("out_w", "out_proj.weight"), ("out_b", "out_proj.bias"), ("ln2_w", "ln_2.weight"), ("ln2_b", "ln_2.bias"),
                 ("fc_w", "c_fc.weight"), ("fc_b", "c_fc.bias"), ("proj_w", "c_proj.weight"), ("proj_b", "c_proj.bias"))


def pack_from_visual(visual):
    """state_dict of a VisionTransformer (clip/clip_surgery_model.py:374) -> flat weight pack.  Works before
    or after ``reload_self_attn`` (the surgery ``Attention`` keeps clones of in_proj / out_proj, :396-405)."""
    sd = {k: v.detach() for k, v in visual.state_dict().items()}
    W = {k: sd[k] for k in ("conv1.weight", "class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias",
                            "ln_post.weight", "ln_post.bias", "proj")}
    L = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    for i in range(L):
        p, o = "transformer.resblocks.%d." % i, "blocks.%d." % i
        if p + "attn.in_proj_weight" in sd:
            names = ("attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias")
        else:
            names = ("attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias")
        for dst, src in zip(("in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias"), names):
            W[o + dst] = sd[p + src]
        for n in ("ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias"):
            W[o + n] = sd[p + n]
        for n in ("c_fc.weight", "c_fc.bias", "c_proj.weight", "c_proj.bias"):
            W[o + n] = sd[p + "mlp." + n]
    heads = getattr(visual, "num_heads", None) or sd["conv1.weight"].shape[0] // 64
    W["meta"] = torch.tensor([L, heads, sd["conv1.weight"].shape[-1]], dtype=torch.int64)
    return W


class SurgeryViT:
    """Frozen CLIP-surgery ViT.  ``weights``: pack as produced by ``pack_from_visual`` (any device)."""

    def __init__(self, weights, n_surgery=5, device="cuda", graph=False):
        self.device = torch.device(device)
        # graph=True: the ~190 launches of a forward are captured once per input shape into a CUDA graph and replayed
        # (no host launch gaps).  The three output tensors are then REUSED by every call with that shape -- meant for
        # pipelines that consume them before the next forward (pipeline.ExCELHotPath), not for the drop-in shims.
        self.graph = graph
        self._graphs = {}
        self.replayed_launches = 0
        L, H, P = (int(v) for v in weights["meta"])
        self.W = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in weights.items() if k != "meta"}
        self.layers, self.heads, self.patch, self.n_surgery = L, H, P, n_surgery
        self.width = self.W["conv1.weight"].shape[0]
        self.embed = self.W["proj"].shape[1]
        self.grid0 = int(round((self.W["positional_embedding"].shape[0] - 1) ** 0.5))
        self._blocks = (_lib.VitLayer * L)()
        self.Ws = {}   # split-fp16 weight copies: the tcgen05 GEMM engine's operand format
        for i in range(L):
            for f, name in _BLOCK_FIELDS:
                setattr(self._blocks[i], f, self.W["blocks.%d.%s" % (i, name)].data_ptr())
            for f, name in (("in", "in_proj_weight"), ("out", "out_proj.weight"), ("fc", "c_fc.weight"), ("proj", "c_proj.weight")):
                ws, scale = self._split("blocks.%d.%s" % (i, name))
                setattr(self._blocks[i], f + "_ws", ws.data_ptr())
                setattr(self._blocks[i], f + "_scale", scale)
        w = _lib.VitWeights()
        w.layers, w.width, w.heads, w.patch, w.embed, w.grid0, w.n_surgery = L, self.width, H, P, self.embed, self.grid0, n_surgery
        for f, name in (("conv1", "conv1.weight"), ("cls", "class_embedding"), ("pos", "positional_embedding"),
                        ("ln_pre_w", "ln_pre.weight"), ("ln_pre_b", "ln_pre.bias"), ("ln_post_w", "ln_post.weight"),
                        ("ln_post_b", "ln_post.bias"), ("proj", "proj")):
            setattr(w, f, self.W[name].data_ptr())
        self.W["conv1.flat"] = self.W["conv1.weight"].reshape(self.width, -1).contiguous()
        self.W["proj.t"] = self.W["proj"].t().contiguous()
        (c1, w.conv1_scale), (pt, w.proj_t_scale) = self._split("conv1.flat"), self._split("proj.t")
        w.conv1_s, w.proj_t_s = c1.data_ptr(), pt.data_ptr()
        w.blocks = ctypes.cast(self._blocks, ctypes.POINTER(_lib.VitLayer))
        self._w = w
        self._ws = None

    def _split(self, name):
        """fp32 weight [out, in] -> (split fp16 [out, 2*round_up(in, 64)] (hi | lo) of scale * W on the device, scale).
        scale = the power of two that puts max|W| at 2^13..2^14: trained weights of magnitude ~0.02 would otherwise have
        their lo halves in fp16's subnormal range (absolute resolution 6e-8 instead of 22 significant bits)."""
        x = self.W[name]
        rows, cols = x.shape
        kp = (cols + 63) // 64 * 64
        amax = float(x.abs().max())
        if not math.isfinite(amax):
            raise RuntimeError(f"SurgeryViT: weight {name} contains inf / nan")
        scale = 2.0 ** (13 - math.frexp(amax)[1] + 1) if amax > 0 else 1.0      # amax * scale in [2^13, 2^14)
        out = torch.empty((rows, 2 * kp), dtype=torch.float16, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("excel_split_f16", _lib.ptr(x), x.stride(0), rows, cols, kp, scale, _lib.ptr(out), _lib.stream())
        self.Ws[name] = out
        return out, scale

    @classmethod
    def from_visual(cls, visual, n_surgery=5, device="cuda"):
        return cls(pack_from_visual(visual), n_surgery, device)

    @torch.no_grad()
    def forward(self, img, ex_feats=None):
        """img [B,3,S,S] -> (tokens [B,N,E] un-normalised, attn [L,B,N,N], feats [L,B,N,D]).
        ex_feats [B,C,h,w] (decoder features, h*w = N-1): LVC branch of the surgery attention
        (clip/clip_surgery_model.py:127-141)."""
        img = img.to(self.device, torch.float32)
        if img.stride(-1) != 1:
            img = img.contiguous()
        B, C, S, S2 = img.shape
        if C != 3 or S != S2:
            raise RuntimeError(f"SurgeryViT: expected [B,3,S,S] square images, got {tuple(img.shape)}")
        if S % self.patch:
            raise RuntimeError(f"SurgeryViT: image size {S} is not a multiple of the patch size {self.patch}")
        with torch.cuda.device(self.device):   # launches, streams and function attributes on the engine's device
            if ex_feats is not None:
                return self._forward(img, lvc_attention(ex_feats, (S // self.patch) ** 2))
            if self.graph:
                return self._forward_graph(img)
            return self._forward(img)

    def _forward_graph(self, img):
        key = tuple(img.shape)
        g = self._graphs.get(key)
        if g is None:
            static_in = torch.empty_like(img)
            static_in.copy_(img)
            # The captured kernels have their workspace pointers baked in: the graph owns a DEDICATED workspace that lives as
            # long as the graph does (the shared self._ws may be re-allocated by a later, larger eager call).
            ws = self._alloc_ws(img.shape[0], img.shape[2])
            self._forward(static_in, ws=ws)               # warm-up outside the capture (function attributes)
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.lib().excel_launch_count()
            with torch.cuda.graph(graph):
                outs = self._forward(static_in, ws=ws)
            g = self._graphs[key] = (graph, static_in, outs, _lib.lib().excel_launch_count() - n0, ws)
        graph, static_in, outs, nlaunch, _ws = g
        static_in.copy_(img)
        graph.replay()
        self.replayed_launches += nlaunch      # kernels launched by graph replays (the library's counter sees host launches only)
        return outs

    def _alloc_ws(self, B, S):
        nbytes = _lib.lib().excel_vit_workspace_bytes(B, S, self.patch, self.width, self.heads)
        return torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)   # 512 B-aligned by the allocator

    def _forward(self, img, lvc_attn=None, ws=None):
        B, C, S, S2 = img.shape
        N = (S // self.patch) ** 2 + 1
        if ws is None:   # eager calls share one grow-only workspace (never referenced by a captured graph)
            nbytes = _lib.lib().excel_vit_workspace_bytes(B, S, self.patch, self.width, self.heads)
            if self._ws is None or self._ws.numel() * 4 < nbytes:
                self._ws = None
                self._ws = self._alloc_ws(B, S)
            ws = self._ws
        tokens = torch.empty((B, N, self.embed), dtype=torch.float32, device=self.device)
        # the attention maps are written by TMA stores: rows padded to a multiple of 16 B; callers get the [.., :N] view
        # (the reference's consumers slice / index this tensor anyway: attn_weights[:, i], [-6:, 1:, 1:], ...)
        npad = (N + 3) // 4 * 4
        attn = torch.empty((self.layers, B, N, npad), dtype=torch.float32, device=self.device)
        feats = torch.empty((self.layers, B, N, self.width), dtype=torch.float32, device=self.device)
        _lib.call("excel_vit_forward", ctypes.byref(self._w), _lib.ptr(img), img.stride(0), img.stride(1), img.stride(2), B, S,
                  _lib.ptr(ws), ws.numel() * 4, _lib.ptr(tokens), _lib.ptr(attn), npad, _lib.ptr(feats), _lib.ptr(lvc_attn),
                  _lib.stream())
        return tokens, attn[..., :N], feats

    __call__ = forward


@_lib.on_tensor_device
def lvc_attention(ex_feats, n_p=None, beta=1.0, gamma=3.0):
    """clip/clip_surgery_model.py:127-137: decoder features [B,C,h,w] -> ex_attn [B,n_p,n_p] (fp32, CUDA)."""
    f = _lib.f32c(ex_feats)
    B, C = f.shape[:2]
    f = f.reshape(B, C, -1)
    np_ = f.shape[2]
    if n_p is not None and np_ != n_p:
        raise RuntimeError(f"ex_feats has {np_} positions, the encoder has {n_p} patches")
    dev = f.device
    qt = torch.empty((B, np_, C), dtype=torch.float32, device=dev)
    rowsum = torch.empty((B * np_,), dtype=torch.float64, device=dev)
    mean = torch.empty((1,), dtype=torch.float32, device=dev)
    out = torch.empty((B, np_, np_), dtype=torch.float32, device=dev)
    _lib.call("excel_lvc_attention", _lib.ptr(f), B, C, np_, float(beta), float(gamma), _lib.ptr(qt), _lib.ptr(rowsum),
              _lib.ptr(mean), _lib.ptr(out), _lib.stream())
    return out


_ENGINES = {}
ENGINE_OPTS = {"graph": False}   # how engine_for builds its engines (install(graph=True): CUDA-graph replay, outputs reused per shape)


def engine_for(model, n_surgery=5):
    """SurgeryViT for a reference ``ExCEL_CLIP`` (or its ``.visual``); cached per module and weight version."""
    if isinstance(model, SurgeryViT):
        return model
    visual = getattr(model, "visual", model)
    key = id(visual)
    ver = tuple((p.data_ptr(), p._version) for p in visual.parameters())
    hit = _ENGINES.get(key)
    if hit is None or hit[0] != ver:
        dev = next(visual.parameters()).device
        _ENGINES[key] = (ver, SurgeryViT(pack_from_visual(visual), n_surgery, device=dev if dev.type == "cuda" else "cuda",
                                         graph=ENGINE_OPTS["graph"]))
    return _ENGINES[key][1]


def generate_clip_fts(inputs, model, return_weights=True, ex_feats=None):
    """clip/clip.py:348-358: (image_features [B,N,E] normalised over the TOKEN axis, attn_weights [L,B,N,N],
    all_feats [L,B,N,D]).  ``model``: a SurgeryViT, or the reference's ExCEL_CLIP module."""
    tokens, attn, feats = engine_for(model)(inputs, ex_feats)
    return token_normalize(tokens), attn, feats
