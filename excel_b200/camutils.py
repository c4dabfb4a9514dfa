"""CAM wrappers -- drop-ins for the live functions of the reference's ``utils/camutils.py``."""
import torch


def cure_attr_map(model, inputs, ex_feats):
    """utils/camutils.py:93-97."""
    with torch.no_grad():
        return model(inputs, ex_feats=ex_feats)


def merge_flipped_maps(attr_2b, b, gh, gw):
    """utils/camutils.py:19-26: element-max of the maps of x and flip(x) (un-flipped), per-(b,c) min subtracted,
    divided by (max + 1e-5).  attr_2b [2b, n_p, K] -> [b, n_p, K]."""
    lam = attr_2b.permute(0, 2, 1).reshape(2 * b, -1, gh, gw)
    lam = torch.max(lam[:b], lam[b:].flip(-1))
    lam = lam - lam.amin(dim=(2, 3), keepdim=True)
    lam = lam / (lam.amax(dim=(2, 3), keepdim=True) + 1e-5)
    return lam.reshape(b, -1, gh * gw).permute(0, 2, 1)


def cure_attr_map_flip(model, inputs, ex_fts=True, flip=True, raw_fts=None):
    """utils/camutils.py:8-30."""
    b, c, h, w = inputs.shape
    with torch.no_grad():
        if not flip:
            return model(inputs, ex_feats=raw_fts)
        inputs_cat = torch.cat([inputs, inputs.flip(-1)], dim=0)
        if ex_fts:
            ex_feats = model(inputs_cat)[1]
            attr = model(inputs_cat, ex_feats=ex_feats)
        else:
            attr = model(inputs_cat)[2]
        return merge_flipped_maps(attr, b, h // 16, w // 16)
