"""Seeded synthetic inputs for the CAM -> SVC -> PAR path.

No dataset or checkpoint is reachable (no network), so every test / bench input is generated
here, on the CPU, from a seed; the oracle and the CUDA path consume the same tensors.
"""
import numpy as np
import torch
import torch.nn.functional as F

# ImageNet statistics the reference datasets normalise with (datasets/transforms.py:7-14)
MEAN = (123.675, 116.28, 103.53)
STD = (58.395, 57.12, 57.375)

# empirical number of present classes per image, from the reference's label files
# (datasets/voc/cls_labels_onehot.npy: 12 031 images, mean 1.55, max 6;
#  datasets/coco/cls_labels_onehot.npy: 122 218 images, mean 2.84, max 18)
N_CLASSES_HIST = {
    "pascal_voc": [0, 7176, 3479, 1090, 233, 49, 4],
    "ms_coco": [0, 26860, 37772, 24758, 14295, 8410, 4796, 2623, 1367, 735, 330, 163, 67, 28, 7, 4, 2, 0, 1],
}


def images(batch, size, seed=0, normalized=True):
    """[B,3,S,S] fp32: uniform RGB noise smoothed by a 9x9 box filter (so PAR affinities are
    non-degenerate), scaled to 0..255 and normalised like the reference datasets."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, size, size, generator=g)
    x = F.avg_pool2d(F.pad(x, [4] * 4, mode="replicate"), 9, stride=1)
    lo = x.amin(dim=(2, 3), keepdim=True)
    hi = x.amax(dim=(2, 3), keepdim=True)
    x = (x - lo) / (hi - lo) * 255.0
    if not normalized:
        return (x / 255.0).contiguous()
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def class_labels(batch, num_fg, seed=0, n_fixed=3, dataset="pascal_voc"):
    """[B,num_fg] float one-hot image-level labels; n per image fixed, or drawn from the
    empirical distribution of the dataset when n_fixed is None."""
    rng = np.random.default_rng(seed)
    hist = np.asarray(N_CLASSES_HIST[dataset], dtype=np.float64)
    out = torch.zeros(batch, num_fg)
    for b in range(batch):
        n = n_fixed if n_fixed is not None else int(rng.choice(len(hist), p=hist / hist.sum()))
        n = max(1, min(n, num_fg))
        out[b, torch.from_numpy(rng.choice(num_fg, size=n, replace=False))] = 1
    return out


def text_bank(T, E, seed=0):
    """[T,E] unit-norm rows: stand-in for ``model.text_attr.permute(1,0)`` where the real
    attribute bank (attributes_text/*.pth in the reference tree) is not reachable."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(T, E, generator=g)
    return (t / t.norm(dim=1, keepdim=True)).contiguous()


def random_visual_weights(layers=12, width=768, patch=16, grid0=14, embed=512, seed=0, sharpen=1.0):
    """Seeded random weight pack with CLIP-like initial scales (no checkpoint is available);
    used where the reference is not importable (GPU box) so bench/tests need no fixture file."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: (torch.randn(*s, generator=g) * std)
    W = {"conv1.weight": rn(width, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
         "class_embedding": rn(width, std=width ** -0.5),
         "positional_embedding": rn(grid0 * grid0 + 1, width, std=width ** -0.5),
         "proj": rn(width, embed, std=width ** -0.5)}
    for n in ("ln_pre", "ln_post"):
        W[n + ".weight"], W[n + ".bias"] = 1 + rn(width, std=0.05), rn(width, std=0.02)
    for i in range(layers):
        o = "blocks.%d." % i
        W[o + "in_proj_weight"] = rn(3 * width, width, std=width ** -0.5) * sharpen
        W[o + "in_proj_bias"] = rn(3 * width, std=0.02)
        W[o + "out_proj.weight"] = rn(width, width, std=width ** -0.5 * (2 * layers) ** -0.5)
        W[o + "out_proj.bias"] = rn(width, std=0.02)
        for n in ("ln_1", "ln_2"):
            W[o + n + ".weight"], W[o + n + ".bias"] = 1 + rn(width, std=0.05), rn(width, std=0.02)
        W[o + "c_fc.weight"], W[o + "c_fc.bias"] = rn(4 * width, width, std=(2 * width) ** -0.5), rn(4 * width, std=0.02)
        W[o + "c_proj.weight"] = rn(width, 4 * width, std=width ** -0.5 * (2 * layers) ** -0.5)
        W[o + "c_proj.bias"] = rn(width, std=0.02)
    W["meta"] = torch.tensor([layers, width // 64, patch], dtype=torch.int64)
    return W
