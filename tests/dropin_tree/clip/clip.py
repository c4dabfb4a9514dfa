"""Module path of the reference's clip/clip.py; the two hot-path functions are stubs (see ../README.md)."""
from _stub import stub

__all__ = ["generate_clip_fts", "clip_feature_surgery"]
generate_clip_fts = stub("clip.generate_clip_fts")
clip_feature_surgery = stub("clip.clip_feature_surgery")
