"""GPU parity: SVC (utils/affutils.py) and CAM (clip/clip.py:288-310,353) kernels vs the oracle + reference fixtures."""
import numpy as np
import pytest
import torch

from _parity import label_parity
from excel_b200 import synth
from oracle import port

pytestmark = pytest.mark.gpu
t = torch.from_numpy


def test_box_masks_golden_and_random(golden):
    from excel_b200.affutils import box_masks
    G = golden("svc")
    for thr in (0.75, 0.79, 0.88):
        idx = [i for i, x in enumerate(G["thrs"]) if abs(x - thr) < 1e-6]
        maps = t(G["maps"][idx])                                   # [n,8,8] -> attr layout [B=n, n_p, K=1]
        attr = maps.reshape(len(idx), 64, 1).cuda()
        m = box_masks(attr, [torch.tensor([0])] * len(idx), 8, 8, thr).cpu().numpy()
        assert np.array_equal(m, G["masks"][idx])
    rng = np.random.default_rng(3)
    for g in (14, 20, 28, 32, 64):
        cams = rng.random((24, g, g)).astype(np.float32)
        cams[::2] *= (rng.random((12, g, g)) > 0.55)
        cams[1::3] = np.round(cams[1::3] * 255) / 255               # exact uint8 boundaries
        attr = t(cams).reshape(24, g * g, 1).cuda()
        m = box_masks(attr, [torch.tensor([0])] * 24, g, g, 0.79).cpu().numpy()
        ref = np.stack([port.box_mask_cv2(c, 0.79) for c in cams])
        assert np.array_equal(m, ref), g


def test_trans_mat_and_refine_golden(golden):
    from excel_b200 import affutils
    G = golden("svc")
    A = t(G["A"])
    T = affutils.compute_trans_mat(A[:, 1:, 1:].mean(0).cuda()).cpu()
    assert (T - t(G["T"])).abs().max() < 1e-7                       # entries ~1/64^2... compare relative too
    assert ((T - t(G["T"])).abs() / t(G["T"]).abs().clamp_min(1e-12)).max() < 1e-4
    lst, cl = affutils.refine_cams_with_aff(t(G["attr"]).cuda(), A.cuda(), t(G["cls"]).cuda(), (128, 128), caa_thre=0.79)
    assert np.array_equal(cl.numpy(), G["cls_lst"]) and not cl.is_cuda
    out = torch.stack(lst).cpu()
    assert ((out - t(G["refined"])).abs() / t(G["refined"]).abs().clamp_min(1e-6)).max() < 1e-4
    # seg_attn (LVC) branch, utils/affutils.py:182-195, caa_thre 0.75 as in engine/validatation_engine.py:33
    lst, _ = affutils.refine_cams_with_aff(t(G["attr"]).cuda(), A.cuda(), t(G["cls"]).cuda(), (128, 128), caa_thre=0.75,
                                           seg_attn=t(G["seg_attn"]).cuda())
    out = torch.stack(lst).cpu()
    assert ((out - t(G["refined_seg"])).abs() / t(G["refined_seg"]).abs().clamp_min(1e-6)).max() < 1e-4


def test_bkg_weclip_golden(golden):
    from excel_b200 import affutils
    from excel_b200.par import PAR
    G = golden("svc")
    img = synth.images(1, 128, seed=12)[0].cuda()
    par = PAR(port.PAR_DILATIONS, 20)
    lab, cams = affutils.refine_cams_with_bkg_weclip(list(t(G["refined"]).cuda()), img, t(G["cls_lst"]), par, (96, 112))
    assert lab.shape == (1, 96, 112) and lab.dtype == torch.int64
    assert (cams.cpu() - t(G["cams"])).abs().max() < 5e-6
    _, _, ref_planes = port.refine_cams_with_bkg_weclip(list(t(G["refined"])), img.cpu(), t(G["cls_lst"]), (96, 112))
    hard, total = label_parity(ref_planes, t(G["labels"])[0], lab.cpu()[0], plane_err=5e-6)
    assert hard == 0 and total <= 4, (hard, total)
    with pytest.raises(RuntimeError):
        affutils.refine_cams_with_bkg_weclip([], img, torch.zeros(0, dtype=torch.int64), par, (96, 112))


def test_cam_golden_and_oracle(golden):
    from excel_b200 import clip
    G = golden("cam")
    out = clip.clip_feature_surgery(t(G["F"]).cuda(), t(G["T"]).cuda()).cpu()
    assert (out - t(G["cam"])).abs().max() < 1e-5
    g = torch.Generator().manual_seed(0)
    for (B, N, E, T) in ((3, 197, 512, 45), (2, 785, 512, 103), (1, 50, 768, 7), (2, 401, 512, 300)):
        tok = torch.randn(B, N, E, generator=g)
        Fn = clip.token_normalize(tok.cuda())
        assert (Fn.cpu() - tok / tok.norm(dim=1, keepdim=True)).abs().max() < 1e-6
        text = synth.text_bank(T, E, seed=B)
        out = clip.clip_feature_surgery(Fn, text.cuda()).cpu()
        ref = port.clip_feature_surgery(tok / tok.norm(dim=1, keepdim=True), text)
        assert (out - ref).abs().max() < 2e-5, (B, N, E, T)
        assert out.min() == 0 and out.max() == 1


def test_refine_batch_vs_oracle():
    """Batched SVC+PAR on oracle-provided CAMs / attention: labels equal the oracle's except near-ties."""
    from excel_b200 import affutils
    from excel_b200.par import PAR
    B, S, K = 3, 160, 20
    g = S // 16
    N = g * g + 1
    gen = torch.Generator().manual_seed(4)
    attn = torch.rand(8, B, N, N, generator=gen) + 0.01
    attn[3:] *= 12                                                   # head-sum layers (rows sum to 12)
    # smooth blobs so that the box masks are stable and non-trivial
    attr = torch.nn.functional.interpolate(torch.rand(B, K, 4, 4, generator=gen), size=(g, g), mode="bicubic")
    attr = ((attr - attr.amin((2, 3), keepdim=True)) / (attr.amax((2, 3), keepdim=True) - attr.amin((2, 3), keepdim=True)))
    attr = attr.reshape(B, K, g * g).permute(0, 2, 1).contiguous()
    cls = synth.class_labels(B, K, seed=8, n_fixed=None)
    imgs = synth.images(B, S, seed=9)
    labels, planes, plane_off, refined = affutils.refine_batch(attr.cuda(), attn.cuda(), cls.cuda(), imgs.cuda(),
                                                               PAR(port.PAR_DILATIONS, 20), return_cams=True)
    # the production path sorts the images by plane count and launches PAR per count class: identical labels
    labels_sorted = affutils.refine_batch(attr.cuda(), attn.cuda(), cls.cuda(), imgs.cuda(), PAR(port.PAR_DILATIONS, 20))
    assert torch.equal(labels_sorted, labels)
    labels, planes, off = labels.cpu(), planes.cpu(), plane_off.cpu().tolist()
    for b in range(B):
        lst, cl = port.refine_cams_with_aff(attr[b], attn[:, b], cls[b], (S, S), caa_thre=0.79)
        lab, cams, ref_planes = port.refine_cams_with_bkg_weclip(lst, imgs[b], cl, (S, S))
        err = (planes[off[b]:off[b + 1]] - cams).abs().max().item()
        assert err < 2e-4, (b, err)                                      # north_star tolerance: 1e-3 max-abs
        hard, total = label_parity(ref_planes, lab[0], labels[b], plane_err=err)
        assert hard == 0 and total <= 8, (b, hard, total, err)
        # PAR + argmax alone on the oracle's own planes: strict margin
        out_p = PAR(port.PAR_DILATIONS, 20)(imgs[b:b + 1].cuda(), cams[None].cuda())[0].cpu()
        hard, total = label_parity(ref_planes, ref_planes.argmax(0), out_p.argmax(0))
        assert hard == 0 and total <= 4, (b, hard, total)


@pytest.mark.parametrize("out_hw", [(75, 91), (120, 166), (53, 40)])
def test_dropin_refine_at_label_resolution_vs_oracle(out_hw):
    """tools/infer_lam.py:93-94: the per-image calls with PAR at the LABEL resolution (the original image size: any H x W,
    rows that are not 16 B multiples, larger or smaller than the network input) -- the image goes through the
    align_corners=True resize of utils/PAR.py:67, the CAMs through cv2-style bilinear up-sampling (utils/affutils.py:75)."""
    from excel_b200 import affutils
    from excel_b200.par import PAR
    S, K = 96, 20
    g = S // 16
    N = g * g + 1
    gen = torch.Generator().manual_seed(out_hw[0])
    attn = torch.rand(8, N, N, generator=gen) + 0.01
    attr = torch.nn.functional.interpolate(torch.rand(1, K, 3, 3, generator=gen), size=(g, g), mode="bicubic")[0]
    attr = (attr - attr.amin((1, 2), keepdim=True)) / (attr.amax((1, 2), keepdim=True) - attr.amin((1, 2), keepdim=True))
    attr = attr.reshape(K, g * g).t().contiguous()
    cls = torch.zeros(K)
    cls[[2, 7, 11]] = 1
    img = synth.images(1, S, seed=out_hw[1])[0]
    par = PAR(port.PAR_DILATIONS, 20)
    refined, cls_lst = affutils.refine_cams_with_aff(attr.cuda(), attn.cuda(), cls.cuda(), (S, S), caa_thre=0.79)
    labels, planes = affutils.refine_cams_with_bkg_weclip(refined, img.cuda(), cls_lst, par, out_hw)
    lst, cl = port.refine_cams_with_aff(attr, attn, cls, (S, S), caa_thre=0.79)
    lab, cams, ref_planes = port.refine_cams_with_bkg_weclip(lst, img, cl, out_hw)
    assert labels.shape == (1,) + tuple(out_hw) and torch.equal(cls_lst, cl)
    err = (planes.cpu() - cams).abs().max().item()
    assert err < 2e-4, err
    hard, total = label_parity(ref_planes, lab[0], labels[0].cpu(), plane_err=err)
    assert hard == 0 and total <= 8, (hard, total, err)


def test_label_utils_golden(golden):
    """utils/camutils.py:123-143,438-476 on the device vs the reference fixtures (integer outputs: bit-exact)."""
    from excel_b200 import camutils
    G = golden("labels")
    mask = camutils.get_mask_by_radius(6, 7, 2)
    assert np.array_equal(mask.cpu().numpy().astype(np.float64), G["mask"])
    aff = camutils.cams_to_affinity_label(t(G["lab"]).cuda(), mask=mask, ignore_index=255)
    assert np.array_equal(aff.cpu().numpy(), G["aff"])
    assert np.array_equal(camutils.cams_to_affinity_label(t(G["lab"]).cuda(), mask=G["mask"], ignore_index=255).cpu().numpy(), G["aff"])
    v, l = camutils.lam_to_label(t(G["cam"]).cuda(), t(G["cls"]).cuda(), bkg_thre=0.45, high_thre=0.6, low_thre=0.3, ignore_mid=True,
                                 ignore_index=255)
    assert np.array_equal(l.cpu().numpy(), G["l_mid"]) and np.array_equal(v.cpu().numpy(), G["valid"])
    _, l2 = camutils.lam_to_label(t(G["cam"]).cuda(), t(G["cls"]).cuda(), bkg_thre=0.45)
    assert np.array_equal(l2.cpu().numpy(), G["l_bkg"])
    big = camutils.get_mask_by_radius(32, 32, 8).cpu().numpy()
    assert np.array_equal(big.astype(np.float64), port.get_mask_by_radius(32, 32, 8))
