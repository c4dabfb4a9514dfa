"""Module path + state_dict naming of the reference's model/segformer_head.py (12 two-layer MLPs + a 1x1 fuse conv).
forward is a stub: the patched forward (excel_b200.install) is the one that runs at inference."""
from torch import nn


class MLP(nn.Module):
    def __init__(self, input_dim, embed_dim):
        super().__init__()
        self.proj = nn.Linear(input_dim, embed_dim)
        self.proj_2 = nn.Linear(embed_dim, embed_dim)


class SegFormerHead(nn.Module):
    def __init__(self, in_channels=768, embedding_dim=256, num_classes=21, index=12):
        super().__init__()
        self.linears_modulelist = nn.ModuleList([MLP(in_channels, embedding_dim) for _ in range(index)])
        self.linear_fuse = nn.Conv2d(embedding_dim * index, embedding_dim, kernel_size=1)

    def forward(self, x_all):
        raise NotImplementedError("dropin_tree: SegFormerHead.forward is a stub -- excel_b200.install() did not patch it")
