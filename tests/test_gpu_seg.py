"""GPU parity: multi-scale + flip segmentation inference (tools/infer_seg_voc.py:56-88) -- the merge
kernels against the oracle restatement, and the whole loop (patched model per scale + merge) against the oracle model."""
import pytest
import torch

from excel_b200 import synth
from oracle import port
from oracle.make_golden_cfg import TINY

pytestmark = pytest.mark.gpu


def _label_gate(segs_ref, lab_ref, lab_gpu, tol=1e-5):
    """mismatching labels must be near-ties of the oracle's two best logits (relative margin <= tol)."""
    top2 = segs_ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]) / top2[:, 0].abs().clamp_min(1e-30)
    bad = lab_ref != lab_gpu
    return int((bad & (margin > tol)).sum()), int(bad.sum())


@pytest.mark.parametrize("size,label_size", [((281, 500), None), ((96, 130), (200, 173)), ((320, 320), (320, 320))])
def test_merge_scales_vs_oracle(size, label_size):
    from excel_b200 import segutils
    g = torch.Generator().manual_seed(size[0])
    seg_list = [torch.randn(2, 21, s, s, generator=g) * 3 for s in (20, 14, 24, 30)]      # 320, 224, 384, 480 px at patch 16
    segs_r, lab_r = port.merge_scales(seg_list, size, label_size)
    segs = segutils.merge_scales([s.cuda() for s in seg_list], size)
    lab = segutils.seg_argmax(segs, size if label_size is None else label_size)
    assert segs.shape == segs_r.shape and lab.shape == lab_r.shape and lab.dtype == torch.int64
    assert (segs.cpu() - segs_r).abs().max() < 1e-5
    resized_r = torch.nn.functional.interpolate(segs_r, size=size if label_size is None else label_size, mode="bilinear",
                                                align_corners=False)
    hard, total = _label_gate(resized_r, lab_r, lab.cpu())
    assert hard == 0 and total <= 4, (hard, total)
    # single scale: the base scale alone drops the flipped half (:69-72)
    one = segutils.merge_scales([seg_list[0].cuda()], size).cpu()
    assert (one - torch.nn.functional.interpolate(seg_list[0][:1], size=size, mode="bilinear", align_corners=False)).abs().max() < 1e-5
    with pytest.raises(RuntimeError):
        segutils.merge_scales([seg_list[0]], size)                  # CPU tensors: no fallback


def test_multi_scale_flip_seg_end_to_end_vs_oracle():
    """Whole loop: encoder + SegFormerHead + stand-in decoder per scale on [x, flip(x)], merged -- vs the oracle's encoder
    (port.generate_clip_fts), head (port.segformer_head) and merge (port.merge_scales)."""
    import sys
    from test_gpu_dropin import TREE, _foreign_modules_loaded, _TOP
    if _foreign_modules_loaded():
        pytest.skip("another `model` package is already imported in this process")
    from excel_b200 import decoder, segutils
    sys.path.insert(0, TREE)
    try:
        from model.model_excel import ExCEL_model
        W = synth.random_visual_weights(seed=3, **TINY)
        text = synth.text_bank(45, TINY["embed"], seed=6)
        torch.manual_seed(5)
        model = ExCEL_model(W, text.t().contiguous(), 21, embedding_dim=64).eval()
        img = synth.images(1, 80, seed=31)
        m_gpu = model.cuda()
        segs, lab = segutils.multi_scale_flip_seg(lambda x: decoder.excel_model_forward(m_gpu, x), img.cuda(), scales=(0.6, 1.0, 1.4),
                                                  resize_size=80, label_size=(75, 91))
        # oracle
        sd = {k: v.detach().cpu() for k, v in model.decoder_fts_fuse.state_dict().items()}
        conv = model.decoder.linear_pred.cpu()
        seg_list = []
        with torch.no_grad():
            for s in (80, 48, 112):
                x = torch.nn.functional.interpolate(img, size=[s, s], mode="bilinear", align_corners=False)
                xc = torch.cat([x, x.flip(-1)], 0)
                _, _, feats = port.generate_clip_fts(W, xc)
                L, B, N, D = feats.shape
                x_all = feats[:, :, 1:].permute(0, 1, 3, 2).reshape(L, B, D, s // 16, s // 16)
                seg_list.append(conv(port.segformer_head(sd, x_all)))
            segs_r, lab_r = port.merge_scales(seg_list, (80, 80), (75, 91))
        assert (segs.cpu() - segs_r).abs().max() < 2e-4 * float(segs_r.abs().max())
        resized_r = torch.nn.functional.interpolate(segs_r, size=(75, 91), mode="bilinear", align_corners=False)
        hard, total = _label_gate(resized_r, lab_r, lab.cpu(), tol=1e-3)
        assert hard == 0 and total <= 0.01 * 75 * 91, (hard, total)
    finally:
        sys.path.remove(TREE)
        for m in list(sys.modules):
            if m in _TOP or m.split(".")[0] in _TOP:
                del sys.modules[m]
