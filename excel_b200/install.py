"""Drop-in installation: make the reference's scripts (scripts/train_*.py, tools/infer_lam.py,
engine/validatation_engine.py) run on the sm_100a hot path WITHOUT editing them.

The reference has no plugin API; its boundary is a set of module-level Python functions imported by name
(SURVEY.md §8b).  ``install()`` imports those reference modules and rebinds the hot-path symbols to this
package's implementations, so that a later ``from utils.affutils import refine_cams_with_aff`` in a script binds
ours.  The one branch this round has not built (the LVC ``ex_feats`` path of the encoder) keeps
dispatching to the reference's own PyTorch code.

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
    from . import affutils as my_aff, camutils as my_cam, clip as my_clip, encoder as my_enc, par as my_par

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

    ref_clip = importlib.import_module("clip")
    ref_clip_inner = importlib.import_module("clip.clip")
    ref_gen = ref_clip_inner.generate_clip_fts

    def generate_clip_fts(inputs, model, return_weights=True, ex_feats=None):
        if ex_feats is not None:      # LVC branch: reference implementation
            return ref_gen(inputs, model, return_weights, ex_feats)
        return my_enc.generate_clip_fts(inputs, model, return_weights)

    for mod in ("clip", "clip.clip"):
        patch(mod, "generate_clip_fts", generate_clip_fts)
        patch(mod, "clip_feature_surgery", my_clip.clip_feature_surgery)
    return originals


def uninstall(originals):
    for name, obj in originals.items():
        module_name, attr = name.rsplit(".", 1)
        setattr(importlib.import_module(module_name), attr, obj)
