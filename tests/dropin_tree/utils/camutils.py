"""Module path of the reference's utils/camutils.py; bodies are stubs (see ../README.md)."""
from _stub import stub

cure_attr_map = stub("utils.camutils.cure_attr_map")
cure_attr_map_flip = stub("utils.camutils.cure_attr_map_flip")
lam_to_label = stub("utils.camutils.lam_to_label")
cams_to_affinity_label = stub("utils.camutils.cams_to_affinity_label")
get_mask_by_radius = stub("utils.camutils.get_mask_by_radius")
