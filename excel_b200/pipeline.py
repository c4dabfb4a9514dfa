"""The whole hot path for a batch: CLIP-surgery ViT -> patch x text CAM -> SVC -> PAR -> pseudo labels
(tools/infer_lam.py:74-94, training-free branch; model/model_excel.py:48-58)."""
import torch

from . import affutils
from .clip import clip_feature_surgery, token_normalize
from .encoder import SurgeryViT
from .par import PAR

PAR_DILATIONS = (1, 2, 4, 8, 12, 24)   # scripts/train_voc.py:112, tools/infer_lam.py:168
PAR_ITERS = 20


class ExCELHotPath:
    """encoder: SurgeryViT; text_attr_t [T,E] (= ExCEL_model.text_attr.permute(1,0), model/model_excel.py:58);
    num_fg = num_classes - 1."""

    def __init__(self, encoder, text_attr_t, num_fg, caa_thre=0.79, par=None):
        self.encoder = encoder
        self.text = text_attr_t.to(encoder.device, torch.float32).contiguous()
        self.num_fg = num_fg
        self.caa_thre = caa_thre
        self.par = par if par is not None else PAR(PAR_DILATIONS, PAR_ITERS)

    @torch.no_grad()
    def cams(self, imgs):
        """ExCEL_model.forward up to attr_maps_raw (model/model_excel.py:55-58):
        (attr_maps_raw [B,n_p,num_fg], attn_weights [L,B,N,N], all_feats [L,B,N,D])."""
        tokens, attn, feats = self.encoder(imgs)
        attr = clip_feature_surgery(token_normalize(tokens), self.text)[:, 1:, :self.num_fg]
        return attr, attn, feats

    @torch.no_grad()
    def __call__(self, imgs, cls_labels, par_imgs=None, out_size=None):
        """imgs [B,3,S,S] (normalised), cls_labels [B,num_fg] one-hot -> labels [B,H,W] int64."""
        attr, attn, _ = self.cams(imgs)
        par_imgs = imgs if par_imgs is None else par_imgs
        return affutils.refine_batch(attr, attn, cls_labels, par_imgs, self.par, out_size, self.caa_thre)
