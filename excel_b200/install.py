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


def install(reference_root=None):
    """Patch the reference modules in place.  Returns the dict {qualified name: original object}."""
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
    for mod in ("clip", "clip.clip"):
        patch(mod, "generate_clip_fts", my_enc.generate_clip_fts)
        patch(mod, "clip_feature_surgery", my_clip.clip_feature_surgery)
    return originals


def uninstall(originals):
    for name, obj in originals.items():
        module_name, attr = name.rsplit(".", 1)
        setattr(importlib.import_module(module_name), attr, obj)
