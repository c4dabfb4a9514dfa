"""GPU parity of the fused attention kernels at the key-block edge cases (attn_tc.cu / attn_pv.cu) and of the host-side
pipelining (pipeline.HostPipeline, par side streams): results must not depend on how the work is scheduled."""
import pytest
import torch

from excel_b200 import synth
from oracle import port
from oracle.make_golden_cfg import TINY

pytestmark = pytest.mark.gpu


# token counts N = g^2 + 1 chosen for the last 128-key block: 37 (one block, <= 64 valid keys -> one 64-key sub-step),
# 101 (one block, two sub-steps), 145 (second block holds 17 keys), 197 (second block holds 69 keys: two sub-steps),
# 257 (third block holds ONE key, like N = 1025 at 512^2)
@pytest.mark.parametrize("size", [96, 160, 192, 224, 256])
def test_attention_key_block_edges_vs_oracle(size):
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    W = port.random_visual_weights(seed=5, **TINY)
    imgs = synth.images(3, size, seed=40 + size)
    tok, attn, feats = generate_clip_fts(imgs.cuda(), SurgeryViT(W))
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs)
    L, H = TINY["layers"], TINY["width"] // 64
    first = L - 5
    assert (attn.cpu() - attn_r).abs().max() < 5e-5
    # rows of the returned maps: head MEAN before the surgery blocks, head SUM inside them (clip_surgery_model.py:146,154)
    assert (attn.cpu()[:first].sum(-1) - 1).abs().max() < 1e-4 and (attn.cpu()[first:].sum(-1) - H).abs().max() < 1e-3
    assert ((feats.cpu() - feats_r).abs().amax(dim=(1, 2, 3)) / feats_r.abs().amax(dim=(1, 2, 3))).max() < 1e-4
    assert (tok.cpu() - tok_r).abs().max() < 1e-4


def test_sharpened_attention_vs_oracle():
    """Peaky softmaxes (in_proj x 2): probabilities close to 1 exercise the split-fp16 P operand."""
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    W = synth.random_visual_weights(seed=6, sharpen=2.0, **TINY)
    imgs = synth.images(2, 224, seed=46)
    tok, attn, feats = generate_clip_fts(imgs.cuda(), SurgeryViT(W))
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs)
    assert attn_r.max() > 0.5                                     # the case really is peaky
    assert (attn.cpu() - attn_r).abs().max() < 1e-4
    assert ((feats.cpu() - feats_r).abs().amax(dim=(1, 2, 3)) / feats_r.abs().amax(dim=(1, 2, 3))).max() < 2e-4
    assert (tok.cpu() - tok_r).abs().max() < 2e-4


def _tiny_hot_path():
    from excel_b200.encoder import SurgeryViT
    from excel_b200.pipeline import ExCELHotPath
    W = synth.random_visual_weights(seed=3, **TINY)
    return ExCELHotPath(SurgeryViT(W), synth.text_bank(45, TINY["embed"], seed=6), 20)


def test_hot_path_is_deterministic_and_schedule_independent():
    """Same labels bit for bit (a) run to run (the head-group reduce-adds of the attention map have a fixed order),
    (b) whether the ragged PAR runs share one stream or fork onto side streams, (c) for CPU or CUDA class labels."""
    from excel_b200 import affutils
    hp = _tiny_hot_path()
    imgs = synth.images(6, 96, seed=50).cuda()
    cls = torch.zeros(6, 20)
    for b, n in enumerate([1, 3, 1, 2, 5, 2]):                      # 2..6 planes: every kernel variant, four runs
        cls[b, torch.randperm(20, generator=torch.Generator().manual_seed(b))[:n]] = 1
    a = hp(imgs, cls)
    b_ = hp(imgs, cls.cuda())
    assert torch.equal(a, b_)
    attr, attn, _ = hp.cams(imgs)
    attr2, attn2, _ = hp.cams(imgs)
    assert torch.equal(attn, attn2) and torch.equal(attr, attr2)
    # per-image calls (one run each, no side streams) must give the batch's labels
    for i in range(6):
        one = affutils.refine_batch(attr[i:i + 1], attn[:, i:i + 1], cls[i:i + 1], imgs[i:i + 1], hp.par)
        assert torch.equal(one[0], a[i]), i


def test_host_pipeline_matches_direct_calls():
    from excel_b200.pipeline import HostPipeline
    hp = _tiny_hot_path()
    batches = [(synth.images(4, 96, seed=60 + i).pin_memory(), synth.class_labels(4, 20, seed=70 + i, n_fixed=None))
               for i in range(4)]
    want = [hp(i.cuda(), c).cpu() for i, c in batches]
    pipe = HostPipeline(hp)
    got = []
    for k, (i, c) in enumerate(batches):
        if pipe.staged is None:
            pipe.stage(i, c)
        nxt = batches[k + 1] if k + 1 < len(batches) else None
        out = pipe.submit(stage_next=nxt)
        if out is not None:
            got.append(out.clone())
    got.append(pipe.flush().clone())
    assert pipe.flush() is None
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert not g.is_cuda and g.dtype == torch.int64 and torch.equal(g, w)


def test_persistent_item_loop_more_items_than_sms():
    """B x row-blocks > 148: every CTA of the fused attention kernel walks several work items (ring / TMEM-buffer phases
    carry over); checked on a 160-image batch against the oracle on a few of its images."""
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    W = port.random_visual_weights(seed=5, **TINY)
    imgs = synth.images(160, 96, seed=77)
    tok, attn, feats = generate_clip_fts(imgs.cuda(), SurgeryViT(W))
    pick = [0, 73, 147, 148, 159]
    tok_r, attn_r, feats_r = port.generate_clip_fts(W, imgs[pick])
    # the token-axis normalisation is per image, so a sub-batch of the oracle is comparable
    assert (attn.cpu()[:, pick] - attn_r).abs().max() < 5e-5
    assert (tok.cpu()[pick] - tok_r).abs().max() < 1e-4
    assert ((feats.cpu()[:, pick] - feats_r).abs().amax(dim=(1, 2, 3)) / feats_r.abs().amax(dim=(1, 2, 3))).max() < 1e-4


@pytest.mark.parametrize("B,S", [(25, 224), (8, 320)])
def test_split_attention_launch_plans_agree(B, S):
    """attn_pv_plan picks the work-item shape by batch: B = 25 at 224^2 runs one item per (image, query block, group of 4
    heads), B = 8 at 320^2 groups of 2 heads, B = 1 single heads -- each with its partial maps summed in a fixed order.
    The plans must agree with each other (and each is checked against the oracle elsewhere) to fp32 summation order."""
    from excel_b200.encoder import SurgeryViT, generate_clip_fts
    enc = SurgeryViT(synth.random_visual_weights(seed=7))
    imgs = synth.images(B, S, seed=80).cuda()
    tok, attn, feats = generate_clip_fts(imgs, enc)
    for i in (0, B // 2, B - 1):
        tok1, attn1, feats1 = generate_clip_fts(imgs[i:i + 1], enc)
        assert (attn[:, i] - attn1[:, 0]).abs().max() < 1e-5
        assert (tok[i] - tok1[0]).abs().max() < 1e-5
        assert ((feats[:, i] - feats1[:, 0]).abs().amax(dim=(1, 2)) / feats1[:, 0].abs().amax(dim=(1, 2))).max() < 2e-5
    torch.cuda.synchronize()
