"""GPU parity: LVC branch of the encoder (Attention.forward with ex_feats,
clip/clip_surgery_model.py:127-141) and utils/attrutils.py (a10), vs the reference's golden outputs and the oracle."""
import pytest
import torch

from excel_b200 import synth
from oracle import port
from oracle.make_golden_cfg import TINY

pytestmark = pytest.mark.gpu
t = torch.from_numpy


def test_lvc_attention_vs_oracle():
    """ex_attn: rows are probability vectors over the kept (>= batch mean) neighbours; values match the oracle except where a
    similarity sits within rounding of the batch mean (the reference's `< 0 -> -inf` cut is discontinuous there)."""
    from excel_b200.encoder import lvc_attention
    g = torch.Generator().manual_seed(5)
    ex = torch.randn(3, 32, 14, 14, generator=g)
    ex = ex + torch.nn.functional.avg_pool2d(torch.nn.functional.pad(ex, [2] * 4, mode="replicate"), 5, stride=1)
    got = lvc_attention(ex.cuda()).cpu()
    ref = port.lvc_attention(ex)
    assert (got.sum(-1) - 1).abs().max() < 1e-5
    q = torch.nn.functional.normalize(ex.flatten(2), dim=1)
    sim = torch.einsum("bcm,bcn->bmn", q, q)
    near_cut = ((sim - sim.mean()) * 3).abs() < 1e-5
    rows_ok = ~near_cut.any(-1)
    assert rows_ok.float().mean() > 0.99
    assert (got - ref)[rows_ok].abs().max() < 1e-6


def test_lvc_encoder_tiny_golden(golden):
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    G = golden("lvc")
    enc = SurgeryViT(port.random_visual_weights(seed=3, **TINY))
    imgs = synth.images(2, 96, seed=13).cuda()
    tok, attn, feats = generate_clip_fts(imgs, enc, ex_feats=t(G["ex"]).cuda())
    assert (tok.cpu() - t(G["tok"])).abs().max() < 2e-5
    assert (attn.cpu() - t(G["attn"])).abs().max() < 2e-5
    assert (feats.cpu() - t(G["feats"])).abs().amax(dim=(1, 2, 3)).max() < 2e-4
    tok0, _, _ = generate_clip_fts(imgs, enc)
    assert (tok0 - tok).abs().max() > 1e-3                         # the LVC bias is not a no-op


def test_lvc_model_forward_b16_vs_oracle():
    """ExCEL_model.forward(img, ex_feats) == clip_feature_surgery(generate_clip_fts(img, ex_feats)) (model_excel.py:50-53)."""
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    from excel_b200.clip import clip_feature_surgery
    W = port.random_visual_weights(seed=1)
    imgs = synth.images(2, 224, seed=21)
    ex = torch.randn(2, 256, 14, 14, generator=torch.Generator().manual_seed(8))
    text = synth.text_bank(45, 512, seed=3)
    tok, attn, feats = generate_clip_fts(imgs.cuda(), SurgeryViT(W), ex_feats=ex.cuda())
    attr = clip_feature_surgery(tok, text.cuda())[:, 1:, :20].cpu()
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs, ex_feats=ex)
    attr_r = port.clip_feature_surgery(tok_r, text)[:, 1:, :20]
    assert (attn.cpu() - attn_r).abs().max() < 5e-5 and (tok.cpu() - tok_r).abs().max() < 1e-4
    assert (attr - attr_r).abs().max() < 1e-3                      # north_star: fp32 CAM values within 1e-3


def test_attrutils_vs_reference_golden(golden):
    from excel_b200 import attrutils
    G = golden("lvc")
    clsmap = attrutils.attrmap2clsmap(t(G["flag"]).cuda(), t(G["amap"]).cuda())
    assert clsmap.shape == (2, 36, 20) and (clsmap.cpu() - t(G["clsmap"])).abs().max() < 1e-5
    agg = attrutils.attr2cls_embedings(t(G["tf"]).cuda(), t(G["bank"]).cuda(), 20)
    assert agg.shape == (64, 20) and (agg.cpu() - t(G["agg"])).abs().max() < 1e-6
    with pytest.raises(RuntimeError):
        attrutils.attrmap2clsmap(t(G["flag"]), t(G["amap"]))          # CPU tensors: no fallback


def test_flip_merge_vs_reference_golden_and_oracle(golden):
    """utils/camutils.py:19-26 (cure_attr_map_flip post-processing) as one kernel."""
    from excel_b200.camutils import merge_flipped_maps
    G = golden("cam")
    lam2b = t(G["lam2b"])
    b = lam2b.shape[0] // 2
    g = int(round(lam2b.shape[1] ** 0.5))
    got = merge_flipped_maps(lam2b.cuda(), b, g, g).cpu()
    assert (got - t(G["merged"])).abs().max() < 1e-6
    x = torch.rand(6, 7 * 9, 20, generator=torch.Generator().manual_seed(0))          # non-square grid
    lam = x.permute(0, 2, 1).reshape(6, 20, 7, 9)
    lam = torch.max(lam[:3], lam[3:].flip(-1))
    lam = lam - lam.amin(dim=(2, 3), keepdim=True)
    lam = (lam / (lam.amax(dim=(2, 3), keepdim=True) + 1e-5)).reshape(3, 20, 63).permute(0, 2, 1)
    assert (merge_flipped_maps(x.cuda(), 3, 7, 9).cpu() - lam).abs().max() < 1e-6


class _MLP(torch.nn.Module):           # same attribute layout as model/segformer_head.py:12-26
    def __init__(self, cin, e):
        super().__init__()
        self.proj, self.proj_2 = torch.nn.Linear(cin, e), torch.nn.Linear(e, e)


class _Head(torch.nn.Module):          # same attribute layout as model/segformer_head.py:46-64
    def __init__(self, cin, e, n):
        super().__init__()
        self.linears_modulelist = torch.nn.ModuleList([_MLP(cin, e) for _ in range(n)])
        self.linear_fuse = torch.nn.Conv2d(e * n, e, kernel_size=1)


def test_decoder_inference_vs_reference_golden(golden):
    """SegFormerHead.forward + attn_pred (model/segformer_head.py:66-77, model/model_excel.py:71-76)."""
    from excel_b200 import decoder
    G, GL = golden("decoder"), golden("lvc")
    head = _Head(TINY["width"], 32, TINY["layers"])
    head.load_state_dict({k[5:]: t(v) for k, v in G.items() if k.startswith("head.")})
    head = head.cuda().eval()
    feats = t(GL["feats"])                                          # [L,B,N,D] as generate_clip_fts returns them
    L, B, N, D = feats.shape
    x_all = feats[:, :, 1:].permute(0, 1, 3, 2).reshape(L, B, D, 6, 6).cuda()
    fts = decoder.segformer_head(head, x_all)
    assert fts.shape == (B, 32, 6, 6)
    assert (fts.cpu() - t(G["fts"])).abs().max() < 2e-5 * float(t(G["fts"]).abs().max())
    # token-major entry point (no transposes: the form excel_model_forward uses), CLS rows included and dropped
    fused = decoder.segformer_head_tokens(head, feats.reshape(L, B * N, D).cuda())
    fts2 = fused.reshape(B, N, -1)[:, 1:].permute(0, 2, 1).reshape(B, -1, 6, 6)
    assert (fts2.cpu() - t(G["fts"])).abs().max() < 2e-5 * float(t(G["fts"]).abs().max())
    ap = decoder.attn_pred(t(G["fts"]).cuda())
    assert (ap.cpu() - t(G["attn_pred"])).abs().max() < 1e-6
    assert (port.attn_pred(t(G["fts"])) - t(G["attn_pred"])).abs().max() < 1e-6


def test_excel_model_forward_vs_oracle_composition():
    """decoder.excel_model_forward (ExCEL_model.forward at inference, model/model_excel.py:48-77) == the oracle's pieces
    composed the same way; the trained DecoderTransformer is stood in for by a fixed torch function."""
    import types
    from excel_b200 import decoder
    from excel_b200.encoder import SurgeryViT
    W = port.random_visual_weights(seed=3, **TINY)
    torch.manual_seed(7)
    head = _Head(TINY["width"], 32, TINY["layers"]).eval()
    text_attr = synth.text_bank(45, TINY["embed"], seed=6).t().contiguous()            # model.text_attr: [E, T]
    dec = lambda fts: (fts.mean(dim=1, keepdim=True), None)
    model = types.SimpleNamespace(encoder=SurgeryViT(W), text_attr=text_attr.cuda(), num_classes=21,
                                  decoder_fts_fuse=head.cuda(), decoder=dec)
    imgs = synth.images(2, 96, seed=13)
    seg, attn_fts, attr, attn_w, apred = decoder.excel_model_forward(model, imgs.cuda())
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs)
    attr_r = port.clip_feature_surgery(tok_r, text_attr.t())[:, 1:, :20]
    x_all = feats_r[:, :, 1:].permute(0, 1, 3, 2).reshape(TINY["layers"], 2, TINY["width"], 6, 6)
    fts_r = port.segformer_head({k: v.detach().cpu() for k, v in head.state_dict().items()}, x_all)
    assert (attr.cpu() - attr_r).abs().max() < 1e-3 and (attn_w.cpu() - attn_r).abs().max() < 5e-5
    assert (attn_fts.cpu() - fts_r).abs().max() < 1e-4 * float(fts_r.abs().max())
    assert (seg.cpu() - fts_r.mean(dim=1, keepdim=True)).abs().max() < 1e-4 * float(fts_r.abs().max())
    assert (apred.cpu() - port.attn_pred(fts_r)).abs().max() < 1e-4
    # LVC call form: only attr_maps_raw comes back (model_excel.py:50-53)
    out = decoder.excel_model_forward(model, imgs.cuda(), ex_feats=attn_fts)
    tok_l, _, _ = port.generate_clip_fts(W, imgs, ex_feats=fts_r)
    assert out.shape == attr.shape and (out.cpu() - port.clip_feature_surgery(tok_l, text_attr.t())[:, 1:, :20]).abs().max() < 1e-3
