"""GPU parity: PAR kernels (excel_par_forward / excel_par_labels) vs the CPU oracle and the
fixtures frozen from the reference (utils/PAR.py)."""
import pytest
import torch

from _parity import label_parity
from excel_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu
t = torch.from_numpy


_near_tie_mismatches = label_parity   # (hard, total): the one label gate, margin 1e-5 (tests/_parity.py)


def test_par_golden(golden):
    from excel_b200.par import PAR
    G = golden("par")
    im_a = synth.images(2, 48, seed=11).cuda()
    out = PAR(port.PAR_DILATIONS, 20)(im_a, t(G["mk_a"]).cuda()).cpu()
    assert (out - t(G["out_a"])).abs().max() < 2e-5
    hard, total = 0, 0
    for b in range(2):
        h, n = _near_tie_mismatches(t(G["out_a"])[b], t(G["out_a"])[b].argmax(0), out[b].argmax(0))
        hard, total = hard + h, total + n
    assert hard == 0 and total <= 2
    # resize branch (align_corners=True), non-square, C=5, 3 iterations, all images in one launch group
    out = PAR(port.PAR_DILATIONS, 3, group=0)(t(G["im_b"]).cuda(), t(G["mk_b"]).cuda()).cpu()
    assert (out - t(G["out_b"])).abs().max() < 2e-5


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (3, 2, 37, 91), (2, 7, 130, 70), (1, 1, 5, 3)])
def test_par_vs_oracle(shape):
    from excel_b200.par import PAR, par_affinity
    b, c, h, w = shape
    g = torch.Generator().manual_seed(h * w)
    imgs = synth.images(b, max(h, w), seed=h)[:, :, :h, :w].contiguous()
    masks = torch.softmax(2 * torch.randn(b, c, h, w, generator=g), 1)
    aff = par_affinity(imgs.cuda(), (h, w), port.PAR_DILATIONS).cpu()
    aff_ref = port.par_affinity(imgs, (h, w))
    assert (aff - aff_ref).abs().max() < 5e-6
    for iters, group in ((1, 1), (2, 0), (20, 2)):
        out = PAR(port.PAR_DILATIONS, iters, group=group)(imgs.cuda(), masks.cuda()).cpu()
        ref = port.par_forward(imgs, masks, num_iter=iters)
        assert (out - ref).abs().max() < 2e-5 * 1.01 ** iters, (iters, group)


def test_par_affinity_degenerate_neighbourhoods():
    """utils/PAR.py:67-86 on inputs where the neighbour std vanishes or explodes: constant regions (std = 0 -> the 1e-8 epsilon
    decides), an isolated spike in a flat region (all 48 differences equal and huge after the division), a step edge and a
    large-magnitude image.  The kernel folds 1 / ((std + eps) w1) into one squared weight per channel and exponentiates with the
    bare MUFU: everything must stay finite and equal to the oracle."""
    from excel_b200.par import par_affinity
    h, w = 96, 80
    img = torch.zeros(3, 3, h, w)
    img[0, :, 40, 30] = 1.0                       # spike in a constant image
    img[0, 1, 10, 10] = -3.0
    img[1, :, :, w // 2:] = 2.5                   # step edge, flat on both sides
    img[1, 0, h // 2:, :] += 1e-4                 # and a barely visible one
    img[2] = 1e3 * synth.images(1, 96, seed=5)[0, :, :h, :w]
    aff = par_affinity(img.cuda(), (h, w), port.PAR_DILATIONS).cpu()
    ref = port.par_affinity(img, (h, w))
    assert torch.isfinite(aff).all()
    assert (aff - ref).abs().max() < 5e-6
    assert (aff.sum(1) - 1.01).abs().max() < 1e-5   # softmax over the 48 neighbours + w2 * positional softmax


def test_par_other_dilations_and_strided_image():
    from excel_b200.par import PAR
    imgs = synth.images(2, 80, seed=3)
    view = imgs.cuda()[:, :, 8:72, 4:68]          # non-contiguous rows, unit x stride
    masks = torch.softmax(torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(1)), 1)
    out = PAR([1, 3, 5], 4)(view, masks.cuda()).cpu()
    ref = port.par_forward(imgs[:, :, 8:72, 4:68], masks, dilations=(1, 3, 5), num_iter=4)
    assert (out - ref).abs().max() < 2e-5


def test_par_ragged_planes_and_labels():
    from excel_b200.par import par_refine_planes, par_labels
    imgs = synth.images(3, 96, seed=5)
    counts = [2, 5, 3]
    g = torch.Generator().manual_seed(9)
    planes = [torch.softmax(2 * torch.randn(c, 96, 96, generator=g), 0) for c in counts]
    keys = [torch.tensor([0] + sorted(torch.randperm(20, generator=g)[:c - 1].add(1).tolist())) for c in counts]
    off = torch.tensor([0, 2, 7, 10], dtype=torch.int32).cuda()
    out = par_refine_planes(imgs.cuda(), torch.cat(planes).cuda(), off, 5, port.PAR_DILATIONS, 20, group=1)
    labels = par_labels(out, off, torch.cat(keys).cuda(), 3).cpu()
    out = out.cpu()
    o = 0
    for b, c in enumerate(counts):
        ref = port.par_forward(imgs[b:b + 1], planes[b][None], num_iter=20)[0]
        assert (out[o:o + c] - ref).abs().max() < 5e-5
        lab_ref = keys[b][ref.argmax(0)]
        hard, total = _near_tie_mismatches(ref, lab_ref, labels[b])
        assert hard == 0 and total <= 4, (b, hard, total)
        o += c


def test_par_errors():
    from excel_b200.par import PAR
    with pytest.raises(RuntimeError):
        PAR([1, 2, 4, 8, 12, 24], 2)(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))      # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        PAR([1, 200], 2)(torch.zeros(1, 3, 8, 8).cuda(), torch.zeros(1, 2, 8, 8).cuda())    # halo exceeds smem
    with pytest.raises(RuntimeError):
        PAR(list(range(1, 10)), 2)(torch.zeros(1, 3, 8, 8).cuda(), torch.zeros(1, 2, 8, 8).cuda())
