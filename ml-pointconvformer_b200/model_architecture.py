"""Model wiring with the reference's class / function names and state-dict layout
(/root/reference/model_architecture.py): get_default_configs (13-77), PCF_Backbone (80-245), the
PCF_Tiny..PCF_Large presets (248-342) and PointConvFormer_Segmentation (345-502).  This file is the
*caller* of the hot path; it only routes tensors (edges, inverse maps, VI features) between the layer
modules of layers.py.

Deliberate differences from the reference, all of them bug fixes needed to make the shipped entry
points usable:
  * get_default_configs also defaults PCONV_OPT (False), guided_level (0) and resblocks_back -- the
    reference's presets raise AttributeError without them;
  * inverse maps are optional even when cfg.PCONV_OPT is set: if the caller passes none (the
    reference's eval drivers never build them) the layers build them lazily when a backward needs them.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused_mlp
from .layers import PCFLayer, PointConv, PointConvStridePE, PointConvTransposePE
from .layer_utils import Linear_BN, linear


class EasyDict(dict):
    """Attribute dict (stand-in for easydict.EasyDict, which the reference imports)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


_DEFAULTS = dict(USE_VI=True, USE_PE=False, transformer_type='PCF', attention_type='subtraction',
                 layer_norm_guidance=False, drop_path_rate=0., BATCH_NORM=True, dropout_rate=0., TIME=False,
                 USE_XYZ=True, point_dim=3, mid_dim_back=1, use_level_1=True, USE_CUDA_KERNEL=False,
                 PCONV_OPT=False, guided_level=0, dropout_fc=0.)


def get_default_configs(cfg, num_level=5, base_dim=64):
    """Fill `cfg` with the model defaults (model_architecture.py:13-77)."""
    cfg.num_level = num_level
    cfg.base_dim = base_dim
    if 'feat_dim' not in cfg.keys():
        cfg.feat_dim = [base_dim * (i + 1) for i in range(cfg.num_level + 1)]
    for k, v in _DEFAULTS.items():
        if k not in cfg.keys():
            cfg[k] = v
    if 'resblocks_back' not in cfg.keys():
        cfg.resblocks_back = [0] * (cfg.num_level + 1)
    return cfg


def _inv_kwargs(enabled, inv_n, inv_k, inv_idx, i):
    if not enabled or inv_n is None:
        return {}
    return {"inv_neighbors": inv_n[i], "inv_k": inv_k[i], "inv_idx": inv_idx[i]}


class PCF_Backbone(nn.Module):
    """Encoder (model_architecture.py:80-245)."""

    def __init__(self, cfg, input_feat_dim=3):
        super().__init__()
        self.cfg = cfg
        self.total_level = cfg.num_level
        self.guided_level = cfg.guided_level
        self.input_feat_dim = input_feat_dim + 3 if cfg.USE_XYZ else input_feat_dim
        self.relu = nn.ReLU(inplace=True)
        wn_in = cfg.point_dim + 9 if cfg.USE_VI is True else cfg.point_dim
        if cfg.transformer_type != 'PCF':
            raise NotImplementedError("only transformer_type='PCF' is on the B200 hot path (PointTransformerLayer is an "
                                      "ablation in the reference, layers.py:419-539)")
        if cfg.use_level_1:
            wn0 = [wn_in, cfg.mid_dim[0]]
            self.selfpointconv = PointConv(self.input_feat_dim, cfg.base_dim, cfg, wn0)
            self.selfpointconv_res1 = PointConvStridePE(cfg.base_dim, cfg.base_dim, cfg, wn0)
            self.selfpointconv_res2 = PointConvStridePE(cfg.base_dim, cfg.base_dim, cfg, wn0)
        else:
            self.selfmlp = Linear_BN(self.input_feat_dim, cfg.base_dim, bn_ver='1d')
        self.pointconv = nn.ModuleList()
        self.pointconv_res = nn.ModuleList()
        for i in range(1, self.total_level):
            in_ch, out_ch = cfg.feat_dim[i - 1], cfg.feat_dim[i]
            wn = [wn_in, cfg.mid_dim[i]]
            make = (lambda a, b: PointConvStridePE(a, b, cfg, wn)) if i <= self.guided_level else \
                   (lambda a, b: PCFLayer(a, b, cfg, wn, cfg.num_heads))
            self.pointconv.append(make(in_ch, out_ch))
            self.pointconv_res.append(nn.ModuleList([make(out_ch, out_ch) for _ in range(cfg.resblocks[i])]))

    def forward(self, features, pointclouds, edges_self, edges_forward, norms,
                inv_neighbors_self=None, inv_k_self=None, inv_idx_self=None,
                inv_neighbors_forward=None, inv_k_forward=None, inv_idx_forward=None):
        opt = bool(self.cfg.PCONV_OPT)
        x = torch.cat([features, pointclouds[0]], -1) if self.cfg.USE_XYZ else features
        if self.cfg.use_level_1:
            kw = _inv_kwargs(opt, inv_neighbors_self, inv_k_self, inv_idx_self, 0)
            x, vi = self.selfpointconv(pointclouds[0], x, edges_self[0], norms[0], **kw)
            x, _ = self.selfpointconv_res1(pointclouds[0], x, edges_self[0], norms[0], vi_features=vi, **kw)
            x, _ = self.selfpointconv_res2(pointclouds[0], x, edges_self[0], norms[0], vi_features=vi, **kw)
        else:
            x = self.selfmlp(x, act=fused_mlp.ACT_RELU) if isinstance(self.selfmlp, Linear_BN) else \
                F.relu(linear(x, self.selfmlp.weight, self.selfmlp.bias))          # after replace_batchnorm
        feat_list = [x]
        for i, conv in enumerate(self.pointconv):
            kw = _inv_kwargs(opt, inv_neighbors_forward, inv_k_forward, inv_idx_forward, i)
            x, _ = conv(pointclouds[i], feat_list[-1], edges_forward[i], norms[i], pointclouds[i + 1], norms[i + 1], **kw)
            vi = None                        # VI features of level i+1 are computed by its first res block
            kw = _inv_kwargs(opt, inv_neighbors_self, inv_k_self, inv_idx_self, i + 1)
            for block in self.pointconv_res[i]:
                x, vi_new = block(pointclouds[i + 1], x, edges_self[i + 1], norms[i + 1], vi_features=vi, **kw)
                if vi is None:
                    vi = vi_new
            feat_list.append(x)
        return feat_list


def _preset(num_level, heads, resblocks, mid, grid, base_dim):
    cfg = get_default_configs(EasyDict(), num_level=num_level, base_dim=base_dim)
    cfg.guided_level = 0
    cfg.num_heads = heads
    cfg.resblocks = resblocks
    cfg.mid_dim = [mid] * num_level
    cfg.grid_size = grid
    return PCF_Backbone(cfg), cfg


def PCF_Tiny(input_grid_size, base_dim=64):
    """model_architecture.py:248-268."""
    g = input_grid_size
    return _preset(5, 1, [0, 1, 1, 1, 1], 4, [g, g * 2, g * 4, g * 8, g * 16], base_dim)


def PCF_Small(input_grid_size, base_dim=64):
    """model_architecture.py:273-293."""
    g = input_grid_size
    return _preset(5, 8, [0, 2, 2, 2, 2], 4, [g, g * 2, g * 4, g * 8, g * 16], base_dim)


def PCF_Normal(input_grid_size, base_dim=64):
    """model_architecture.py:298-318."""
    g = input_grid_size
    return _preset(5, 8, [0, 2, 4, 6, 6], 16, [g, g * 2, g * 4, g * 8, g * 16], base_dim)


def PCF_Large(input_grid_size, base_dim=64):
    """model_architecture.py:321-342."""
    g = input_grid_size
    return _preset(6, 8, [0, 2, 4, 6, 6, 2], 16, [g, g * 2.5, g * 5, g * 10, g * 20, g * 40], base_dim)


class PointConvFormer_Segmentation(nn.Module):
    """Backbone + PointConvTranspose decoder + per-point classifier (model_architecture.py:345-502)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.total_level = cfg.num_level
        self.pcf_backbone = PCF_Backbone(cfg)
        wn = [cfg.point_dim + 9 if cfg.USE_VI is True else cfg.point_dim, cfg.mid_dim_back]
        self.pointdeconv = nn.ModuleList()
        self.pointdeconv_res = nn.ModuleList()
        for i in range(self.total_level - 2, -1, -1):
            in_ch = cfg.feat_dim[i + 1]
            out_ch = cfg.base_dim if i == 0 else cfg.feat_dim[i]
            self.pointdeconv.append(PointConvTransposePE(in_ch, out_ch, cfg, wn, [out_ch, out_ch]))
            n_res = cfg.resblocks_back[i] if cfg.resblocks[i] != 0 else 0
            self.pointdeconv_res.append(nn.ModuleList([PointConvStridePE(out_ch, out_ch, cfg, wn) for _ in range(n_res)]))
        self.fc1 = Linear_BN(cfg.base_dim, cfg.base_dim, bn_ver='1d')
        self.dropout_fc = nn.Dropout(p=cfg.dropout_fc) if cfg.dropout_fc > 0. else nn.Identity()
        self.fc2 = nn.Linear(cfg.base_dim, cfg.num_classes)

    def forward(self, features, pointclouds, edges_self, edges_forward, edges_propagate, norms,
                inv_self=None, inv_forward=None, inv_propagate=None):
        # (SyncBatchNorm row counts travel inside each statistics message, fused_mlp.bn_finalize: nothing to set up here)
        opt = bool(self.cfg.PCONV_OPT)
        ins, iks, iis = inv_self if (opt and inv_self is not None) else (None, None, None)
        inf, ikf, iif = inv_forward if (opt and inv_forward is not None) else (None, None, None)
        inp, ikp, iip = inv_propagate if (opt and inv_propagate is not None) else (None, None, None)
        feat_list = self.pcf_backbone(features, pointclouds, edges_self, edges_forward, norms,
                                      ins, iks, iis, inf, ikf, iif)
        x = feat_list[-1]
        for i, deconv in enumerate(self.pointdeconv):
            lvl = self.total_level - 2 - i
            x, _ = deconv(pointclouds[lvl + 1], x, edges_propagate[lvl], norms[lvl + 1], pointclouds[lvl], norms[lvl],
                          feat_list[lvl], **_inv_kwargs(opt, inp, ikp, iip, lvl))
            vi = None
            kw = _inv_kwargs(opt, ins, iks, iis, lvl)
            for block in self.pointdeconv_res[i]:
                x, vi_new = block(pointclouds[lvl], x, edges_self[lvl], norms[lvl], vi_features=vi, **kw)
                if vi is None:
                    vi = vi_new
            feat_list[lvl] = x
        h = self.fc1(x, act=fused_mlp.ACT_RELU) if isinstance(self.fc1, Linear_BN) else F.relu(linear(x, self.fc1.weight, self.fc1.bias))
        return linear(self.dropout_fc(h), self.fc2.weight, self.fc2.bias)
