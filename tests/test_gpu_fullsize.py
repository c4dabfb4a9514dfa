"""GPU parity at the FULL sizes of the BASELINE.json configs (the other GPU tests run reduced geometries):
cfg2 ViT-B/16 @512^2 (N = 1025), cfg3 ViT-B/16 @448^2 (N = 785, T = 103), cfg5 PAR @1024^2, and the confusion
histogram of row f3 -- each against oracle/port.py on the same seeded inputs."""
import numpy as np
import pytest
import torch

from _parity import label_parity
from excel_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu


def _encoder_cam_labels(S, B, T, K, dataset, seed):
    """Encoder + CAM + SVC + PAR of the GPU path vs the port at image size S, batch B (ViT-B/16, random-init weights)."""
    from excel_b200 import affutils
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath
    W = synth.random_visual_weights(seed=seed)
    text = synth.text_bank(T, 512, seed=seed + 1)
    imgs = synth.images(B, S, seed=seed + 2)
    cls = synth.class_labels(B, K, seed=seed + 3, n_fixed=None, dataset=dataset)
    hp = ExCELHotPath(SurgeryViT(W), text, K)
    g = S // 16
    N = g * g + 1
    attr, attn, feats = hp.cams(imgs.cuda())
    assert attn.shape == (12, B, N, N) and attr.shape == (B, N - 1, K)
    with torch.no_grad():
        tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs)
        attr_r = port.clip_feature_surgery(tok_r, text)[:, 1:, :K]
    assert (attn.cpu() - attn_r).abs().max() < 5e-5
    assert (attn.cpu()[:7].sum(-1) - 1).abs().max() < 1e-4 and (attn.cpu()[7:].sum(-1) - 12).abs().max() < 1e-3
    assert ((feats.cpu() - feats_r).abs().amax(dim=(1, 2, 3)) / feats_r.abs().amax(dim=(1, 2, 3))).max() < 1e-4
    d_cam = (attr.cpu() - attr_r).abs().max().item()
    assert d_cam < 1e-3, d_cam                                         # north_star: fp32 CAM values within 1e-3
    # labels, stage-isolated: the GPU tail (SVC + PAR + argmax) on the ORACLE's CAMs / attention vs the oracle's tail
    par = hp.par
    lab_iso, planes_iso, off, _ = affutils.refine_batch(attr_r.cuda(), attn_r.cuda(), cls, imgs.cuda(), par, return_cams=True)
    # labels, end to end: everything on the GPU
    lab_e2e, planes_e2e, off_e, _ = affutils.refine_batch(attr, attn, cls, imgs.cuda(), par, return_cams=True)
    off = off.cpu().tolist()
    stats = []
    for b in range(B):
        with torch.no_grad():
            lst, cl = port.refine_cams_with_aff(attr_r[b], attn_r[:, b], cls[b], (S, S), caa_thre=0.79)
            lab, cams, ref_planes = port.refine_cams_with_bkg_weclip(lst, imgs[b], cl, (S, S))
        e_iso = (planes_iso[off[b]:off[b + 1]].cpu() - cams).abs().max().item()
        e_e2e = (planes_e2e[off[b]:off[b + 1]].cpu() - cams).abs().max().item()
        assert e_iso < 1e-3, (b, e_iso)
        hard, total = label_parity(ref_planes, lab[0], lab_iso[b].cpu(), plane_err=e_iso)
        # PAR + argmax alone on the oracle's OWN input planes: the strict gate (no upstream difference)
        out_p = par(imgs[b:b + 1].cuda(), cams[None].cuda())[0].cpu()
        assert (out_p - ref_planes).abs().max() < 5e-5
        hard_p, total_p = label_parity(ref_planes, ref_planes.argmax(0), out_p.argmax(0))
        assert hard_p == 0 and total_p <= 1e-4 * S * S, (b, hard_p, total_p)
        hard_e, total_e = label_parity(ref_planes, lab[0], lab_e2e[b].cpu(), plane_err=e_e2e)
        stats.append((b, e_iso, hard, total, e_e2e, hard_e, total_e))
    print("fullsize", S, "cam", d_cam, stats)
    for b, e_iso, hard, total, e_e2e, hard_e, total_e in stats:
        assert hard == 0 and total <= 1e-3 * S * S, stats
        assert hard_e == 0, stats


def test_cfg2_vitb16_512_full_vs_oracle():
    """configs[1] geometry: 512^2, N = 1025 (8 full 128-token blocks + the CLS edge), T = 45, VOC class mix."""
    _encoder_cam_labels(512, 2, 45, 20, "pascal_voc", seed=40)


def test_cfg3_vitb16_448_full_vs_oracle():
    """configs[2] geometry: 448^2, N = 785, T = 103 (80 fg + 23 bg prompts), COCO class mix."""
    _encoder_cam_labels(448, 1, 103, 80, "ms_coco", seed=50)


@pytest.mark.parametrize("iters", [1, 2])
def test_cfg5_par_1024_vs_oracle(iters):
    """configs[4] geometry: PAR at 1024^2, 4 planes, vs utils/PAR.py:64-92 restated (oracle/port.par_forward)."""
    from excel_b200.par import PAR
    imgs = synth.images(1, 1024, seed=61)
    g = torch.Generator().manual_seed(62)
    masks = torch.softmax(2 * torch.randn(1, 4, 1024, 1024, generator=g), 1)
    out = PAR(port.PAR_DILATIONS, iters)(imgs.cuda(), masks.cuda()).cpu()
    with torch.no_grad():
        ref = port.par_forward(imgs, masks, num_iter=iters)
    assert (out - ref).abs().max() < 2e-5
    hard, total = label_parity(ref[0], ref[0].argmax(0), out[0].argmax(0))
    assert hard == 0 and total <= 16, (hard, total)


@pytest.mark.parametrize("nc", [21, 81])
def test_confusion_hist_vs_oracle(nc):
    """Row f3: excel_confusion_hist vs utils/evaluate.py:9-15 (_fast_hist): ground truth 255 (ignore) and negative
    values are dropped, the histogram accumulates across calls."""
    from excel_b200 import evaluate
    rng = np.random.default_rng(nc)
    n = 3 * 333 * 517 + 5
    lt = rng.integers(0, nc, n)
    lt[rng.random(n) < 0.07] = 255                                  # ignore label of the datasets
    lt[rng.random(n) < 0.01] = -1
    lp = rng.integers(0, nc, n)
    ref = port.fast_hist(lt, lp, nc)
    hist = evaluate.confusion_hist(torch.from_numpy(lt).cuda(), torch.from_numpy(lp).cuda(), nc)
    assert hist.dtype == torch.int64 and np.array_equal(hist.cpu().numpy(), ref)
    assert int(hist.sum()) == int(((lt >= 0) & (lt < nc)).sum())
    # second batch accumulates into the same histogram (a whole evaluation run: one histogram, one all-reduce)
    lt2, lp2 = rng.integers(0, nc, 1000), rng.integers(0, nc, 1000)
    evaluate.confusion_hist(torch.from_numpy(lt2).cuda().view(10, 100), torch.from_numpy(lp2).cuda().view(10, 100), nc, hist)
    assert np.array_equal(hist.cpu().numpy(), ref + port.fast_hist(lt2, lp2, nc))
    s_gpu = evaluate.scores_from_hist(hist)
    assert abs(s_gpu["miou"] - port.miou_from_hist(hist.cpu().numpy())) < 1e-12
