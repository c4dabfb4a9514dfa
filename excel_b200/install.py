"""Drop-in installation: make the reference's scripts (scripts/train_*.py, tools/infer_lam.py,
engine/validatation_engine.py) run on the sm_100a hot path WITHOUT editing them.

The reference has no plugin API; its boundary is a set of module-level Python functions imported by name
(SURVEY.md §8b).  ``install()`` imports those reference modules and rebinds the hot-path symbols to this
package's implementations, so that a later ``from utils.affutils import refine_cams_with_aff`` in a script binds
ours -- both branches of every function (training-free and LVC: ``ex_feats`` in the encoder, ``seg_attn`` in SVC).

    python -m excel_b200.run tools/infer_lam.py --infer_set train --training_free true ...
"""
import importlib
import os
import sys


def install(reference_root=None, graph=False):
    """Patch the reference modules in place.  Returns the dict {qualified name: original object}.

    graph=True: the encoder engines built for the reference's CLIP modules replay their ~190 launches as one CUDA graph
    per input shape (no host launch gaps -- what the batch-1 loops of tools/infer_lam.py / engine/validatation_engine.py
    need).  The three tensors a forward returns are then REUSED by the next forward of the same shape: fine for those
    loops (each image's outputs are consumed before the next `model(inputs)`), not for callers that keep them.

    Behavioural differences to the reference after install() (also in INTEGRATION.md): `get_mask_by_radius` returns a CUDA
    fp32 tensor (the reference: numpy float64; `cams_to_affinity_label` accepts both), `attr2cls_embedings` adds the
    foreground text rows (the reference adds all rows, which only broadcasts when there are no background rows), and no
    patched function has a CPU path: CPU tensors raise."""
    from . import encoder as _enc
    _enc.ENGINE_OPTS["graph"] = bool(graph)
    if reference_root:
        reference_root = os.path.abspath(reference_root)
        if reference_root not in sys.path:
            sys.path.insert(0, reference_root)
    from . import affutils as my_aff, attrutils as my_attr, camutils as my_cam, clip as my_clip, encoder as my_enc, par as my_par

    originals = {}

    def patch(module_name, attr, new):
        mod = importlib.import_module(module_name)
        originals[module_name + "." + attr] = getattr(mod, attr)
        setattr(mod, attr, new)

    patch("utils.affutils", "refine_cams_with_aff", my_aff.refine_cams_with_aff)
    patch("utils.affutils", "refine_cams_with_bkg_weclip", my_aff.refine_cams_with_bkg_weclip)
    patch("utils.affutils", "compute_trans_mat", my_aff.compute_trans_mat)
    patch("utils.PAR", "PAR", my_par.PAR)
    patch("utils.camutils", "cure_attr_map", my_cam.cure_attr_map)
    patch("utils.camutils", "cure_attr_map_flip", my_cam.cure_attr_map_flip)
    patch("utils.camutils", "lam_to_label", my_cam.lam_to_label)
    patch("utils.camutils", "cams_to_affinity_label", my_cam.cams_to_affinity_label)
    patch("utils.camutils", "get_mask_by_radius", my_cam.get_mask_by_radius)

    patch("utils.attrutils", "attrmap2clsmap", my_attr.attrmap2clsmap)
    patch("utils.attrutils", "attr2cls_embedings", my_attr.attr2cls_embedings)
    # decoder-side inference: only without autograd and in eval mode (Dropout2d, trainable decoder)
    import torch
    from . import decoder as my_dec
    ref_model = importlib.import_module("model.model_excel")
    ref_head = importlib.import_module("model.segformer_head")
    orig_forward, orig_head_forward = ref_model.ExCEL_model.forward, ref_head.SegFormerHead.forward

    def model_forward(self, img, ex_feats=None):
        if torch.is_grad_enabled() or self.training or not img.is_cuda:
            return orig_forward(self, img, ex_feats)
        return my_dec.excel_model_forward(self, img, ex_feats)

    def head_forward(self, x_all):
        if torch.is_grad_enabled() or self.training or not x_all.is_cuda:
            return orig_head_forward(self, x_all)
        return my_dec.segformer_head(self, x_all)

    originals["model.model_excel.ExCEL_model.forward"] = None   # class attributes: restored explicitly by uninstall()
    _CLASS_PATCHES[:] = [(ref_model.ExCEL_model, "forward", orig_forward), (ref_head.SegFormerHead, "forward", orig_head_forward)]
    ref_model.ExCEL_model.forward = model_forward
    ref_head.SegFormerHead.forward = head_forward

    for mod in ("clip", "clip.clip"):
        patch(mod, "generate_clip_fts", my_enc.generate_clip_fts)
        patch(mod, "clip_feature_surgery", my_clip.clip_feature_surgery)
    return originals


_CLASS_PATCHES = []


def uninstall(originals):
    for cls, attr, obj in _CLASS_PATCHES:
        setattr(cls, attr, obj)
    _CLASS_PATCHES.clear()
    for name, obj in originals.items():
        if obj is None:
            continue
        module_name, attr = name.rsplit(".", 1)
        setattr(importlib.import_module(module_name), attr, obj)
