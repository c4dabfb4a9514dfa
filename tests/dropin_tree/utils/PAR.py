"""Module path of the reference's utils/PAR.py; the class is a stub (see ../README.md)."""
from torch import nn


class PAR(nn.Module):
    def __init__(self, dilations, num_iter):
        super().__init__()
        raise NotImplementedError("dropin_tree: utils.PAR.PAR is a stub -- excel_b200.install() did not patch it")
