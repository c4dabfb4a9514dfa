"""Module path of the reference's utils/attrutils.py; bodies are stubs (see ../README.md)."""
from _stub import stub

attrmap2clsmap = stub("utils.attrutils.attrmap2clsmap")
attr2cls_embedings = stub("utils.attrutils.attr2cls_embedings")
