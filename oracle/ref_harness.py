"""Reference harness: runs the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Nothing under ``excel_b200/`` may import this file.
It exists only in the build container (the GPU box has no /root/reference); it
is used to (1) validate ``oracle/port.py`` against the real reference and
(2) generate the committed fixtures under ``tests/golden/``
(see ``oracle/make_golden.py``).

Shims (none of them edits the reference tree):
  * stub modules for packages absent from this image (ftfy, mmcv, matplotlib,
    pydensecrf, imageio, texttable) -- pulled in by clip/simple_tokenizer.py:6,
    model/segformer_head.py:10, utils/camutils.py:3-6, utils/dcrf.py:1-3;
  * ``Tensor.cuda`` / ``Module.cuda`` -> identity: the path hard-codes .cuda()
    (clip/clip.py:350, model/load_attr.py:94, utils/affutils.py:164-169,209,217);
  * ``clip.load`` -> seeded random-init ExCEL_CLIP (no checkpoint, no network).
"""
import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("EXCEL_REFERENCE_ROOT", "/root/reference")

_STUBS = ["ftfy", "mmcv", "mmcv.cnn", "matplotlib", "matplotlib.pyplot", "pydensecrf",
          "pydensecrf.densecrf", "pydensecrf.utils", "imageio", "imageio.v2", "texttable"]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "utils"))


def _install_stubs():
    for name in _STUBS:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["ftfy"].fix_text = lambda s: s
    sys.modules["mmcv.cnn"].ConvModule = object
    sys.modules["texttable"].Texttable = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pydensecrf.utils"].unary_from_softmax = None
    sys.modules["pydensecrf.utils"].unary_from_labels = None
    sys.modules["pydensecrf"].densecrf = sys.modules["pydensecrf.densecrf"]
    sys.modules["pydensecrf"].utils = sys.modules["pydensecrf.utils"]


_loaded = {}


def load():
    """Import the reference modules; returns a namespace of them."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    _install_stubs()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with _cwd(REF_ROOT):
        import clip  # noqa
        import clip.clip_surgery_model as csm
        from utils import PAR as par_mod
        from utils import affutils, camutils, evaluate, attrutils
        from model import model_excel, load_attr
    _loaded.update(clip=clip, csm=csm, PAR=par_mod.PAR, affutils=affutils, camutils=camutils, attrutils=attrutils,
                   evaluate=evaluate, model_excel=model_excel, load_attr=load_attr)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


VIT_B16 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
               vision_patch_size=16, context_length=77, vocab_size=49408,
               transformer_width=512, transformer_heads=8, transformer_layers=12)
# a 2-layer-text, small-width model for cheap fixtures (vision tower keeps the surgery structure)
VIT_TINY = dict(embed_dim=64, image_resolution=64, vision_layers=8, vision_width=128,
                vision_patch_size=16, context_length=77, vocab_size=49408,
                transformer_width=64, transformer_heads=2, transformer_layers=1)


def build_clip(cfg=VIT_B16, seed=0, sharpen=1.0):
    """Seeded random-init ExCEL_CLIP (clip/clip_surgery_model.py:452)."""
    ref = load()
    torch.manual_seed(seed)
    enc = ref.csm.ExCEL_CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"],
                             cfg["vision_width"], cfg["vision_patch_size"], cfg["context_length"],
                             cfg["vocab_size"], cfg["transformer_width"], cfg["transformer_heads"],
                             cfg["transformer_layers"]).float().eval()
    # the stock initialiser leaves the vision tower at torch defaults; give the biases and
    # LayerNorm affine parameters non-trivial values so parity tests exercise them.
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in enc.visual.named_parameters():
            if name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif "ln_" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
            elif "in_proj_weight" in name and sharpen != 1.0:
                p.mul_(sharpen)
    return enc


def build_model(enc, dataset="pascal_voc", img_size=224, mode="val"):
    """ExCEL_model (model/model_excel.py:15) around a given encoder; real attribute bank."""
    ref = load()
    ref.clip.load = lambda name, device=None, **k: (enc, None)
    ref.clip.clip.load = ref.clip.load
    voc = dataset == "pascal_voc"
    with _cwd(REF_ROOT):
        m = ref.model_excel.ExCEL_model(
            clip_model="ExCEL_ViT-B/16", embedding_dim=256, in_channels=enc.visual.embed_dim,
            dataset_name=dataset, num_classes=21 if voc else 81,
            num_atrr_clusters=112 if voc else 224,
            json_file="./attributes_text/descriptors_%s_gpt4.0_cluster_a_photo_of4.json" % dataset,
            img_size=img_size, mode=mode, device="cpu").eval()
    return m
