"""CPU oracle: a torch-CPU / numpy restatement of ExCEL's CAM -> SVC -> PAR hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / the timed CPU baseline.  Nothing under
``excel_b200/`` imports it; the product path is CUDA only.

Parity status: the reference ships NO tests or golden vectors, so
this port is pinned against the reference ITSELF: ``oracle/check_port.py`` runs
the unmodified reference (via ``oracle/ref_harness.py``) and this port on the
same seeded inputs in the build container, and ``oracle/make_golden.py`` freezes
reference outputs into ``tests/golden/`` which ``tests/test_oracle_golden.py``
replays on every run (also on the GPU box, where /root/reference is absent).

Every function cites the reference file:line (relative to /root/reference) it
restates.  Everything is fp32, as in the reference (clip/build_model.py:72).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

try:  # the reference's own third-party arithmetic for contours / resize (utils/affutils.py:3)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

# --------------------------------------------------------------------------- weights


def export_visual_weights(visual):
    """Flatten a (reference or stock-CLIP style) VisionTransformer into a plain dict of
    fp32 tensors -- the weight-pack format shared by the oracle and the CUDA path.

    Block attention parameters are exported under in_proj/out_proj names for both
    nn.MultiheadAttention (blocks before the surgery) and the surgery ``Attention``
    module, whose qkv/proj are clones of them (clip/clip_surgery_model.py:396-405).
    """
    W = {}
    sd = {k: v.detach().float().contiguous() for k, v in visual.state_dict().items()}
    W["conv1.weight"] = sd["conv1.weight"]
    W["class_embedding"] = sd["class_embedding"]
    W["positional_embedding"] = sd["positional_embedding"]
    for n in ("ln_pre", "ln_post"):
        W[n + ".weight"], W[n + ".bias"] = sd[n + ".weight"], sd[n + ".bias"]
    W["proj"] = sd["proj"]
    L = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    for i in range(L):
        p = "transformer.resblocks.%d." % i
        o = "blocks.%d." % i
        if p + "attn.in_proj_weight" in sd:
            W[o + "in_proj_weight"], W[o + "in_proj_bias"] = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            W[o + "out_proj.weight"], W[o + "out_proj.bias"] = sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"]
        else:
            W[o + "in_proj_weight"], W[o + "in_proj_bias"] = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
            W[o + "out_proj.weight"], W[o + "out_proj.bias"] = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        for n in ("ln_1", "ln_2"):
            W[o + n + ".weight"], W[o + n + ".bias"] = sd[p + n + ".weight"], sd[p + n + ".bias"]
        W[o + "c_fc.weight"], W[o + "c_fc.bias"] = sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]
        W[o + "c_proj.weight"], W[o + "c_proj.bias"] = sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"]
    W["meta"] = torch.tensor([L, visual.num_heads if hasattr(visual, "num_heads") else sd["conv1.weight"].shape[0] // 64,
                              sd["conv1.weight"].shape[-1]], dtype=torch.int64)
    return W


from excel_b200.synth import random_visual_weights  # noqa: E402,F401  (one seeded definition for oracle and CUDA path)


# --------------------------------------------------------------------------- ViT (a3-a7)


def _layer_norm(x, w, b):
    # clip/clip_surgery_model.py:271-277 (fp32 LayerNorm, eps 1e-5)
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _mlp(x, W, o):
    # clip/clip_surgery_model.py:291-295,280-282: c_fc -> QuickGELU x*sigmoid(1.702x) -> c_proj
    h = F.linear(x, W[o + "c_fc.weight"], W[o + "c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)
    return F.linear(h, W[o + "c_proj.weight"], W[o + "c_proj.bias"])


def _split_heads(t, B, N, H):
    return t.reshape(B, N, H, -1).permute(0, 2, 1, 3)  # [B,H,N,dh]


def _std_attention(x, W, o, H):
    """Blocks before the surgery: nn.MultiheadAttention(need_weights=True), head-MEAN
    probabilities (clip/clip_surgery_model.py:297-307,332-337).  x: [B,N,D]."""
    B, N, D = x.shape
    qkv = F.linear(x, W[o + "in_proj_weight"], W[o + "in_proj_bias"])
    q, k, v = (_split_heads(t, B, N, H) for t in qkv.split(D, dim=-1))
    q = q * (1.0 / math.sqrt(D // H))  # torch MHA scales q before q.k^T
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    out = (p @ v).permute(0, 2, 1, 3).reshape(B, N, D)
    out = F.linear(out, W[o + "out_proj.weight"], W[o + "out_proj.bias"])
    return out, p.mean(dim=1)


def lvc_attention(ex_feats, beta=1.0, gamma=3.0):
    """LVC bias of the surgery attention (clip/clip_surgery_model.py:127-136): ex_feats [B,C,h,w] decoder features ->
    ex_attn [B,n_p,n_p] = softmax_j of the batch-mean-centred, x3, negatives -> -inf cosine similarity."""
    q_k = F.normalize(ex_feats.flatten(2, 3).float(), dim=1)                   # :130
    sim = torch.einsum("bcm,bcn->bmn", q_k, q_k)                               # :131
    sim = (sim - torch.mean(sim) * beta) * gamma                               # :132 (mean over the WHOLE batch)
    sim = sim.masked_fill(sim < 0.0, float("-inf"))                            # :133
    return torch.softmax(sim, dim=-1)                                          # :137 (same for every head)


def _surgery_attention(x, W, o, H, ex_attn=None):
    """Surgery ``Attention.forward`` (clip/clip_surgery_model.py:95-159); ex_attn = lvc_attention(ex_feats) or None.
    Returns (x_new, x_ori, attn_ori head-SUM)."""
    B, N, D = x.shape
    scale = (D // H) ** -0.5
    qkv = F.linear(x, W[o + "in_proj_weight"], W[o + "in_proj_bias"])
    q, k, v = (_split_heads(t, B, N, H) for t in qkv.split(D, dim=-1))
    p_ori = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)           # :101-102
    p_new = (torch.softmax((q @ q.transpose(-1, -2)) * scale, dim=-1)
             + torch.softmax((k @ k.transpose(-1, -2)) * scale, dim=-1)
             + torch.softmax((v @ v.transpose(-1, -2)) * scale, dim=-1)) / 3   # :119-125
    if ex_attn is not None:                                                    # :139-141 added to every head's patch block
        p_new = p_new.clone()
        p_new[:, :, 1:, 1:] = p_new[:, :, 1:, 1:] + ex_attn.unsqueeze(1)
    p_new = p_new.sum(dim=1, keepdim=True)                                     # :146 (sum over heads)
    x_ori = (p_ori @ v).permute(0, 2, 1, 3).reshape(B, N, D)                   # :148
    x_new = (p_new @ v).permute(0, 2, 1, 3).reshape(B, N, D)                   # :149
    x_new = F.linear(x_new, W[o + "out_proj.weight"], W[o + "out_proj.bias"])  # :151
    x_ori = F.linear(x_ori, W[o + "out_proj.weight"], W[o + "out_proj.bias"])  # :152
    return x_new, x_ori, p_ori.sum(dim=1)                                      # :154


def resize_pos_embed(pos, new_side):
    """clip/clip_surgery_model.py:426-435: bilinear (align_corners=False) resize of the grid part."""
    side = int((pos.shape[0] - 1) ** 0.5)
    if side == new_side:
        return pos
    D = pos.shape[1]
    grid = pos[1:].reshape(1, side, side, D).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, (new_side, new_side), mode="bilinear")
    grid = grid.reshape(D, new_side * new_side).t()
    return torch.cat([pos[:1], grid], 0)


def vit_forward(W, img, n_surgery=5, ex_feats=None):
    """VisionTransformer.forward + Transformer.forward (clip/clip_surgery_model.py:418-448,346-371)
    for a [B,3,S,S] image batch.  Returns (tokens [B,N,E] BEFORE the token-axis normalisation,
    attn list of L x [B,N,N], feats list of L x [B,N,D] with the aliasing of row a5 applied)."""
    L, H, P = (int(v) for v in W["meta"])
    B = img.shape[0]
    x = F.conv2d(img, W["conv1.weight"], stride=P)                  # :421 patch embed, no bias
    D = x.shape[1]
    x = x.reshape(B, D, -1).permute(0, 2, 1)
    x = torch.cat([W["class_embedding"].expand(B, 1, D), x], dim=1)  # :424
    x = x + resize_pos_embed(W["positional_embedding"], int((x.shape[1] - 1) ** 0.5))
    x = _layer_norm(x, W["ln_pre.weight"], W["ln_pre.bias"])         # :438
    attns, feats = [], []
    first = L - n_surgery                                            # :399 range(1, layers) -> last 5 blocks
    ex_attn = None if ex_feats is None else lvc_attention(ex_feats)
    x_ori = None
    for i in range(L):
        o = "blocks.%d." % i
        if i < first:                                                # :332-337
            a, p = _std_attention(_layer_norm(x, W[o + "ln_1.weight"], W[o + "ln_1.bias"]), W, o, H)
            x = x + a
            x = x + _mlp(_layer_norm(x, W[o + "ln_2.weight"], W[o + "ln_2.bias"]), W, o)
            feats.append(x)
        else:                                                        # :309-330
            src = x if x_ori is None else x_ori
            x_res, x_ori_res, p = _surgery_attention(
                _layer_norm(src, W[o + "ln_1.weight"], W[o + "ln_1.bias"]), W, o, H, ex_attn)
            mid = src + x_ori_res
            if i > first:
                # aliasing quirk: all_feats[i-1] is a view of the previous x_ori, which the in-place
                # `x_ori += x_ori_res` (:317) mutates before torch.stack (clip/clip.py:356)
                feats[i - 1] = mid
            x_ori = mid + _mlp(_layer_norm(mid, W[o + "ln_2.weight"], W[o + "ln_2.bias"]), W, o)
            x = x + x_res                                            # new path: no FFN (:319,329)
            feats.append(x_ori)
        attns.append(p)
    if x_ori is not None:
        x = x.clone()
        x[:, 0] = x_ori[:, 0]                                        # :442
        feats[first - 1] = x                                         # in-place `x += x_res` / x[0]=.. alias all_feats[6]
    out = _layer_norm(x, W["ln_post.weight"], W["ln_post.bias"]) @ W["proj"]  # :445-446
    return out, attns, feats


def generate_clip_fts(W, img, n_surgery=5, ex_feats=None):
    """clip/clip.py:348-358: normalise over the TOKEN axis (dim=1), stack the lists."""
    tok, attns, feats = vit_forward(W, img, n_surgery, ex_feats)
    tok = tok / tok.norm(dim=1, keepdim=True)
    return tok, torch.stack(attns, 0), torch.stack(feats, 0)


# --------------------------------------------------------------------------- decoder-side inference (f4)


def segformer_head(Wd, x_all):
    """SegFormerHead.forward at inference (model/segformer_head.py:66-77).  Wd: state_dict-style weights
    (linears_modulelist.{l}.proj / proj_2, linear_fuse); x_all [L,B,C,h,w] -> [B,E,h,w]."""
    L, B, C, h, w = x_all.shape
    outs = []
    for l in range(L):
        x = x_all[l].float().flatten(2).transpose(1, 2)                                                # :22
        x = F.relu(F.linear(x, Wd["linears_modulelist.%d.proj.weight" % l], Wd["linears_modulelist.%d.proj.bias" % l]))
        x = F.linear(x, Wd["linears_modulelist.%d.proj_2.weight" % l], Wd["linears_modulelist.%d.proj_2.bias" % l])
        outs.append(x.permute(0, 2, 1).reshape(B, -1, h, w))                                           # :72
    return F.conv2d(torch.cat(outs, dim=1), Wd["linear_fuse.weight"], Wd["linear_fuse.bias"])         # :74-75 (dropout: eval)


def attn_pred(attn_fts, beta=1.0, gamma=3.0):
    """model/model_excel.py:71-76."""
    f = F.normalize(attn_fts.flatten(2).float(), dim=1)
    a = f.transpose(2, 1).bmm(f)
    return torch.sigmoid((a - torch.mean(a) * beta) * gamma)


def merge_scales(seg_list, size, label_size=None):
    """tools/infer_seg_voc.py:66-86 given the per-scale model outputs: seg_list[0] = base scale ([2,C,g,g], only the
    un-flipped half is used, :69-72), the others are flip-merged (:78-80); mean over scales (:83), resize to the label
    size and argmax (:85-86).  Returns (segs [1,C,h,w], labels [1,H,W])."""
    out = []
    for i, segs in enumerate(seg_list):
        up = F.interpolate(segs.float(), size=size, mode="bilinear", align_corners=False)
        out.append(up[0].unsqueeze(0) if i == 0 else (up[:1] + up[1:].flip(-1)) / 2)
    segs = torch.mean(torch.stack(out, dim=0), dim=0)
    resized = F.interpolate(segs, size=size if label_size is None else label_size, mode="bilinear", align_corners=False)
    return segs, torch.argmax(resized, dim=1)


# --------------------------------------------------------------------------- attribute bank (a10)


def attrmap2clsmap(attri_flag, attr_maps):
    """utils/attrutils.py:11-17: attr_maps [B,n_p,A] @ attri_flag[cls,A]^T."""
    return attr_maps @ attri_flag.unsqueeze(0).permute(0, 2, 1)


def attr2cls_embedings(text_features, text_attri, num_classes):
    """utils/attrutils.py:19-29 with the foreground rows added at :25 (the reference adds ALL text rows there, which only
    broadcasts when there are no background rows -- the two agree in that case; model/load_attr.py:112 is the live form)."""
    fg, bg = text_features[:num_classes], text_features[num_classes:]
    corr = (fg @ text_attri).softmax(dim=-1)
    agg = torch.cat([corr @ text_attri.t() + fg, bg], dim=0)
    return (agg / agg.norm(dim=1, keepdim=True)).permute(1, 0)


# --------------------------------------------------------------------------- CAM (a9)


def clip_feature_surgery(image_features, text_features):
    """clip/clip.py:288-310 restated as GEMM + epilogue.  image_features [B,N,E], text_features [T,E]
    -> [B,N,T] min-max normalised over all N tokens (CLS included), no epsilon."""
    S = image_features @ text_features.t()                    # sum_c f*t  (:301,:306)
    w = torch.softmax(S[:, :1, :] * 2, dim=-1)                # :295-296
    w = w / w.mean(dim=-1, keepdim=True)                      # :297
    Sw = S * w                                                # :302
    sim = Sw - Sw.mean(dim=-1, keepdim=True)                  # :303-304
    lo = sim.min(dim=1, keepdim=True)[0]
    hi = sim.max(dim=1, keepdim=True)[0]
    return (sim - lo) / (hi - lo)                             # :308


def cure_attr_map_flip_post(lam2b, g):
    """utils/camutils.py:19-26: merge the maps of [x, flip(x)]; lam2b [2B,n_p,K] -> [B,n_p,K]."""
    b2, n_p, K = lam2b.shape
    b = b2 // 2
    lam = lam2b.permute(0, 2, 1).reshape(b2, K, g, g)
    lam = torch.max(lam[:b], lam[b:].flip(-1))
    lam = lam - lam.amin(dim=(2, 3), keepdim=True)
    lam = lam / (lam.amax(dim=(2, 3), keepdim=True) + 1e-5)
    return lam.reshape(b, K, n_p).permute(0, 2, 1)


# --------------------------------------------------------------------------- SVC (a13-a15)


def compute_trans_mat(A):
    """utils/affutils.py:8-24: 3 rounds of (column-normalise, row-normalise), symmetrise, square."""
    T = A
    for _ in range(3):
        T = T / T.sum(dim=0, keepdim=True)
        T = T / T.sum(dim=1, keepdim=True)
    T = (T + T.t()) / 2
    return T @ T


def box_mask_cv2(cam, thr):
    """utils/affutils.py:26-53 + :209-212 with the reference's own OpenCV calls."""
    h, w = cam.shape
    img = (cam * 255).astype(np.uint8)[..., None]
    _, binary = cv2.threshold(img, int(thr * np.max(img)), 255, cv2.THRESH_BINARY)
    contours = cv2.findContours(binary, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0]
    mask = np.zeros((h, w), np.float32)
    for c in contours:
        x, y, bw, bh = cv2.boundingRect(c)
        mask[y:min(y + bh, h - 1), x:min(x + bw, w - 1)] = 1
    return mask


def box_mask_cc(cam, thr):
    """The same mask WITHOUT OpenCV: union of the clipped bounding boxes of the 8-connected
    components of ``uint8(cam*255) > int(thr*max)`` (0/3000 mismatches vs cv2).
    This is the algorithm the CUDA kernel implements."""
    h, w = cam.shape
    img = (cam * 255).astype(np.uint8)
    fg = img > int(thr * int(img.max()))
    label = -np.ones((h, w), np.int64)
    mask = np.zeros((h, w), np.float32)
    for sy, sx in zip(*np.nonzero(fg)):
        if label[sy, sx] >= 0:
            continue
        label[sy, sx] = 1
        stack, y0, y1, x0, x1 = [(sy, sx)], sy, sy, sx, sx
        while stack:
            y, x = stack.pop()
            y0, y1, x0, x1 = min(y0, y), max(y1, y), min(x0, x), max(x1, x)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = y + dy, x + dx
                    if 0 <= yy < h and 0 <= xx < w and fg[yy, xx] and label[yy, xx] < 0:
                        label[yy, xx] = 1
                        stack.append((yy, xx))
        mask[y0:min(y1 + 1, h - 1), x0:min(x1 + 1, w - 1)] = 1
    return mask


def svc_attention(attn_weights, seg_attn=None, attn_layers=6):
    """utils/affutils.py:180-198: aggregate the last ``attn_layers`` patch-patch attention maps.
    attn_weights [L,N,N] -> [n_p,n_p]."""
    A = attn_weights[:, 1:, 1:][-attn_layers:]
    if seg_attn is None:
        return A.mean(dim=0)
    diff = (seg_attn - A).flatten(1).sum(dim=1)               # :183-184
    keep = (diff <= diff.mean()).float().reshape(-1, 1, 1)    # :185-189
    A = (keep * A).sum(dim=0) / (keep.expand_as(A).sum(dim=0) + 1e-5)  # :193
    return A * seg_attn.squeeze(0)                            # :195


def refine_cams_with_aff(attr_map, attn_weights, cls_label, size, caa_thre=0.79, attn_layers=6,
                         seg_attn=None, use_cv2=True):
    """utils/affutils.py:177-223.  attr_map [n_p,K], attn_weights [L,N,N], cls_label [K].
    Returns (list of [g_h,g_w] tensors, int64 class indices)."""
    h, w = size
    gh, gw = h // 16, w // 16
    T = compute_trans_mat(svc_attention(attn_weights, seg_attn, attn_layers)).float()
    cls_lst = torch.where(cls_label)[0]
    out = []
    for c in cls_lst:
        cam = attr_map[:, c].numpy().reshape(gh, gw)
        mask = (box_mask_cv2 if (use_cv2 and cv2 is not None) else box_mask_cc)(cam, caa_thre)
        m = torch.from_numpy(mask).reshape(1, -1)
        out.append(((T * m) @ torch.from_numpy(cam).reshape(-1, 1)).reshape(gh, gw))  # :215-221
    return out, cls_lst


# --------------------------------------------------------------------------- PAR (a16-a21)

PAR_DILATIONS = (1, 2, 4, 8, 12, 24)
_TAPS = ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1))  # utils/PAR.py:10-24


def par_neighbors(x, dilations=PAR_DILATIONS):
    """utils/PAR.py:39-49 as a clamped-coordinate gather: out[b,c,k,y,x] = x[b,c,clamp(y+dy*d),clamp(x+dx*d)],
    k = dilation-major, tap order of get_kernel."""
    b, c, h, w = x.shape
    ys, xs = torch.arange(h), torch.arange(w)
    out = []
    for d in dilations:
        for dy, dx in _TAPS:
            yi = (ys + dy * d).clamp(0, h - 1)
            xi = (xs + dx * d).clamp(0, w - 1)
            out.append(x[:, :, yi][:, :, :, xi])
    return torch.stack(out, dim=2)


def par_pos_term(dilations=PAR_DILATIONS, w1=0.3):
    """utils/PAR.py:51-62,74,81,84: the constant 48-vector softmax(-(pos/(std(pos)+1e-8)/w1)^2)."""
    ker = torch.ones(8)
    ker[[0, 2, 5, 7]] = float(np.sqrt(2))
    pos = torch.cat([ker * d for d in dilations])
    pos_aff = -(pos / (pos.std() + 1e-8) / w1) ** 2
    return torch.softmax(pos_aff, dim=0)


def par_affinity(imgs, size, dilations=PAR_DILATIONS, w1=0.3, w2=0.01):
    """utils/PAR.py:67-86: imgs [b,3,hi,wi] -> aff [b,K,H,W] (rows sum to 1+w2)."""
    imgs = F.interpolate(imgs, size=size, mode="bilinear", align_corners=True)   # :67
    nb = par_neighbors(imgs, dilations)                                          # [b,3,K,H,W]
    absd = (nb - imgs.unsqueeze(2)).abs()                                        # :76
    std = nb.std(dim=2, keepdim=True)                                            # :77 (unbiased)
    aff = -(absd / (std + 1e-8) / w1) ** 2                                       # :80
    aff = aff.mean(dim=1)                                                        # :81
    return torch.softmax(aff, dim=1) + w2 * par_pos_term(dilations, w1).view(1, -1, 1, 1)  # :86


def par_forward(imgs, masks, dilations=PAR_DILATIONS, num_iter=20, w1=0.3, w2=0.01):
    """utils/PAR.py:64-92.  imgs [b,3,hi,wi], masks [b,C,H,W] -> [b,C,H,W]."""
    aff = par_affinity(imgs, masks.shape[-2:], dilations, w1, w2).unsqueeze(1)
    for _ in range(num_iter):
        masks = (par_neighbors(masks, dilations) * aff).sum(2)                   # :88-90
    return masks


def scale_cam(cam, out_hw):
    """utils/affutils.py:69-78: min-max with (1e-7 + max), bilinear resize (cv2.resize ==
    F.interpolate(align_corners=False) to <=1.5e-6)."""
    cam = cam - cam.min()
    cam = cam / (1e-7 + cam.max())
    return F.interpolate(cam[None, None], size=out_hw, mode="bilinear", align_corners=False)[0, 0]


def refine_cams_with_bkg_weclip(cam_list, img, cls_lst, out_hw, dilations=PAR_DILATIONS, num_iter=20,
                                use_cv2=True):
    """utils/affutils.py:161-174 (+ :55-67, :80-89).  cam_list: n x [g,g]; img [3,hi,wi];
    returns (labels [1,H,W] int64, cams [C,H,W], refined [C,H,W] PAR output)."""
    H, W_ = out_hw
    if use_cv2 and cv2 is not None:
        ups = []
        for c in cam_list:
            a = c.numpy().astype(np.float32)
            a = a - a.min()
            a = a / (1e-7 + a.max())
            ups.append(torch.from_numpy(cv2.resize(a, (W_, H))))
        cams = torch.stack(ups, 0)
    else:
        cams = torch.stack([scale_cam(c, (H, W_)) for c in cam_list], 0)
    bg = 1 - cams.max(dim=0, keepdim=True)[0]                                    # :165
    cams = torch.cat([bg, cams], 0)                                              # :166
    valid_key = torch.cat([torch.zeros(1, dtype=torch.int64), cls_lst.to(torch.int64) + 1])  # :168
    refined = par_forward(img[None].float(), cams[None].float(), dilations, num_iter)[0]
    labels = valid_key[refined.argmax(dim=0)]                                    # :86-87
    return labels[None], cams, refined


# --------------------------------------------------------------------------- whole path + metric


def hot_path(W, text_attr_t, imgs, cls_labels, num_fg, caa_thre=0.79, num_iter=20, par_imgs=None,
             out_hw=None, use_cv2=True):
    """tools/infer_lam.py:74-94 (training-free branch) for a batch: encoder -> CAM -> per image SVC -> PAR.
    text_attr_t [T,E] (= model.text_attr.permute(1,0)); cls_labels [B,num_fg] one-hot.
    Returns dict(attr_maps_raw, attn_weights, all_feats, labels list, cams list (PAR input planes), refined list (PAR
    output planes))."""
    tok, attn, feats = generate_clip_fts(W, imgs)
    attr = clip_feature_surgery(tok, text_attr_t)[:, 1:, :num_fg]               # model/model_excel.py:58
    par_imgs = imgs if par_imgs is None else par_imgs
    out_hw = tuple(imgs.shape[-2:]) if out_hw is None else out_hw
    labels, cams, refined = [], [], []
    for i in range(imgs.shape[0]):
        cam_list, cls_lst = refine_cams_with_aff(attr[i], attn[:, i], cls_labels[i], imgs.shape[-2:],
                                                 caa_thre=caa_thre, use_cv2=use_cv2)
        lab, cam, ref = refine_cams_with_bkg_weclip(cam_list, par_imgs[i], cls_lst, out_hw,
                                                    num_iter=num_iter, use_cv2=use_cv2)
        labels.append(lab)
        cams.append(cam)
        refined.append(ref)
    return dict(attr_maps_raw=attr, attn_weights=attn, all_feats=feats, labels=labels, cams=cams, refined=refined)


def fast_hist(label_true, label_pred, num_classes):
    """utils/evaluate.py:9-15."""
    lt, lp = np.asarray(label_true).ravel(), np.asarray(label_pred).ravel()
    m = (lt >= 0) & (lt < num_classes)
    return np.bincount(num_classes * lt[m].astype(int) + lp[m], minlength=num_classes ** 2
                       ).reshape(num_classes, num_classes)


def miou_from_hist(hist):
    """utils/evaluate.py:21-24."""
    hist = hist.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
    return float(np.nanmean(iu[hist.sum(1) > 0]))


# --------------------------------------------------------------------------- label utilities (f2)


def get_mask_by_radius(h, w, radius):
    """utils/camutils.py:459-476 without the Python double loop."""
    ys, xs = np.divmod(np.arange(h * w), w)
    return ((np.abs(ys[:, None] - ys[None]) <= radius) & (np.abs(xs[:, None] - xs[None]) <= radius)).astype(np.float64)


def cams_to_affinity_label(cam_label, mask=None, ignore_index=255):
    """utils/camutils.py:438-457."""
    b, h, w = cam_label.shape
    small = F.interpolate(cam_label.unsqueeze(1).float(), size=[h // 16, w // 16], mode="nearest").reshape(b, 1, -1)
    rep = small.repeat([1, small.shape[-1], 1])
    aff = (rep == rep.permute(0, 2, 1)).long()
    for i in range(b):
        if mask is not None:
            aff[i, torch.as_tensor(mask) == 0] = ignore_index
        aff[i, :, rep[i, 0, :] == ignore_index] = ignore_index
        aff[i, rep[i, 0, :] == ignore_index, :] = ignore_index
    return aff


def lam_to_label(cam, cls_label, bkg_thre=0.5, high_thre=None, low_thre=None, ignore_mid=False, ignore_index=None):
    """utils/camutils.py:123-143 (img_box=None)."""
    valid = cls_label[:, :, None, None] * cam
    val, lab = valid.max(dim=1)
    lab = lab + 1
    if ignore_mid:
        lab[val <= high_thre] = ignore_index
        lab[val <= low_thre] = 0
    else:
        lab[val <= bkg_thre] = 0
    return valid, lab
