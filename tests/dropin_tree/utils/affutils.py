"""Module path of the reference's utils/affutils.py; bodies are stubs (see ../README.md)."""
from _stub import stub

refine_cams_with_aff = stub("utils.affutils.refine_cams_with_aff")
refine_cams_with_bkg_weclip = stub("utils.affutils.refine_cams_with_bkg_weclip")
compute_trans_mat = stub("utils.affutils.compute_trans_mat")
