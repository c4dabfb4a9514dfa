"""Module path + attribute surface of the reference's model/model_excel.py: `ExCEL_model` with `.encoder.visual`
(OpenAI-CLIP VisionTransformer parameter names), `.text_attr [E,T]`, `.num_classes`, `.decoder_fts_fuse`
(SegFormerHead) and `.decoder` (a trained PyTorch module in the reference: here a 1x1-conv stand-in with the same
`(logits, attention list)` return).  forward is a stub: the patched forward is the one that runs at inference."""
import torch
from torch import nn

from .segformer_head import SegFormerHead


class _Block(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(width, heads)
        self.ln_1 = nn.LayerNorm(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, 4 * width))
        self.mlp.add_module("c_proj", nn.Linear(4 * width, width))
        self.ln_2 = nn.LayerNorm(width)


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[_Block(width, heads) for _ in range(layers)])


class VisionTransformer(nn.Module):
    """Parameter container with OpenAI-CLIP's names, filled from an excel_b200 weight pack."""

    def __init__(self, pack):
        super().__init__()
        layers, heads, patch = (int(v) for v in pack["meta"])
        width = pack["conv1.weight"].shape[0]
        self.num_heads = heads
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        self.class_embedding = nn.Parameter(pack["class_embedding"].clone())
        self.positional_embedding = nn.Parameter(pack["positional_embedding"].clone())
        self.ln_pre, self.ln_post = nn.LayerNorm(width), nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads)
        self.proj = nn.Parameter(pack["proj"].clone())
        sd = {"conv1.weight": pack["conv1.weight"]}
        for n in ("ln_pre", "ln_post"):
            sd[n + ".weight"], sd[n + ".bias"] = pack[n + ".weight"], pack[n + ".bias"]
        for i in range(layers):
            o, p = "blocks.%d." % i, "transformer.resblocks.%d." % i
            sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"] = pack[o + "in_proj_weight"], pack[o + "in_proj_bias"]
            sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"] = pack[o + "out_proj.weight"], pack[o + "out_proj.bias"]
            for n in ("ln_1", "ln_2"):
                sd[p + n + ".weight"], sd[p + n + ".bias"] = pack[o + n + ".weight"], pack[o + n + ".bias"]
            for n in ("c_fc", "c_proj"):
                sd[p + "mlp." + n + ".weight"], sd[p + "mlp." + n + ".bias"] = pack[o + n + ".weight"], pack[o + n + ".bias"]
        self.load_state_dict(sd, strict=False)


class _Clip(nn.Module):
    def __init__(self, pack):
        super().__init__()
        self.visual = VisionTransformer(pack)


class _Decoder(nn.Module):
    def __init__(self, width, num_classes):
        super().__init__()
        self.linear_pred = nn.Conv2d(width, num_classes, kernel_size=1)

    def forward(self, x):
        return self.linear_pred(x), []


class ExCEL_model(nn.Module):
    def __init__(self, pack, text_attr, num_classes=21, embedding_dim=256):
        super().__init__()
        self.num_classes = num_classes
        self.encoder = _Clip(pack)
        self.decoder_fts_fuse = SegFormerHead(pack["conv1.weight"].shape[0], embedding_dim, num_classes, int(pack["meta"][0]))
        self.decoder = _Decoder(embedding_dim, num_classes)
        self.register_buffer("text_attr", text_attr.clone())          # [E, T] like model_excel.py:35

    def forward(self, img, ex_feats=None):
        raise NotImplementedError("dropin_tree: ExCEL_model.forward is a stub -- excel_b200.install() did not patch it")
