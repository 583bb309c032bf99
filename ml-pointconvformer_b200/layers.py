"""PointConv / PointConvFormer layer modules with the reference's constructor and forward signatures
(/root/reference/layers.py): PointConv (744-906), PointConvStridePE (542-741), PCFLayer (194-416),
PointConvTransposePE (909-1105), WeightNet (127-191), MultiHeadGuidance (23-68).  Parameter / sub-module
names are the reference's (both PCONV_OPT spellings), so its checkpoints load.

What differs is where the work happens: the gathers of xyz / normals, the localisation and the 12-d VI
transform are one kernel (layer_utils.edge_geometry); gather -> (guidance) -> K-contraction -> Linear is
one kernel (layer_utils.FusedPConvFunction, tcgen05 on B200) whose backward scatters through the kNN
inverse map without atomics; the strided shortcut is a fused max-gather.  The switches keep their
meaning: cfg.USE_CUDA_KERNEL selects the fused contraction, cfg.PCONV_OPT additionally fuses the
Linear and changes the parameter spelling.  With USE_CUDA_KERNEL False the contraction is the
reference's unfused formulation (gather kernel + batched matmul).  CUDA tensors only.
"""
import os

import torch
from torch import nn
import torch.nn.functional as F

from . import fused_mlp
from . import streams as S
from .layer_utils import (FusedPConvFunction, PConvLinearOpt, Linear_BN, UnaryBlock, edge_geometry, gather_max,
                          guidance_input, index_points, linear, resolve_inverse)


def _chain_spec(mods, acts):
    """[(linear, bn_or_None, act)] for a list of Linear_BN / nn.Linear modules, or None if a size is unsupported."""
    out = []
    for m, a in zip(mods, acts):
        lin, bn = (m.c, m.bn) if isinstance(m, Linear_BN) else (m, None)
        out.append((lin, bn, a))
    if not fused_mlp.supported([(l.in_features, l.out_features) for l, _, _ in out]):
        return None
    return out


def _drop_path(cfg):
    rate = getattr(cfg, "drop_path_rate", 0.)
    if rate > 0.:
        try:
            from timm.models.layers import DropPath
            return DropPath(rate)
        except ImportError:
            return _DropPath(rate)
    return nn.Identity()


class _DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath semantics; timm is optional here)."""

    def __init__(self, p):
        super().__init__()
        self.p = p

    def forward(self, x):
        if not self.training or self.p == 0.:
            return x
        keep = 1 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _inv_tuple(nei_inds, n_in, inv_neighbors, inv_k, inv_idx, needed):
    if not needed:
        return None
    return resolve_inverse(nei_inds, n_in, inv_neighbors, inv_k, inv_idx)


PAD_C_IN = os.environ.get("PCFB_PAD_CIN", "1") != "0"     # pad C_in to a multiple of 4 in front of the fused contraction


def _contract(cfg, feats, nei_inds, inv, weights, additional, guidance, lin_w, lin_b):
    """gather -> (x guidance) -> concat additional -> sum_k (.) w -> (Linear).  Fused kernel when
    cfg.USE_CUDA_KERNEL, else the reference's unfused torch formulation on our gather."""
    if cfg.USE_CUDA_KERNEL:
        C_in, pad = feats.shape[-1], (-feats.shape[-1]) % 4
        if PAD_C_IN and pad and lin_w is not None and guidance is None and feats.is_cuda:
            # The first layer of the encoder sees colour + xyz = 6 channels; the warp-specialised forward and the vector path
            # of the contraction backward need C_in % 4 == 0.  Two zero channels (and the matching zero columns of the Linear)
            # change nothing in the result and move the 100 k-point level-0 layer from the 64-point-tile fallback kernel
            # (276 us) to the kernels every other layer uses; autograd slices the gradients back (pad / cat backward).
            M = weights.shape[-1]
            feats = F.pad(feats, (0, pad))
            lin_w = torch.cat([lin_w[:, :C_in * M], lin_w.new_zeros(lin_w.shape[0], pad * M), lin_w[:, C_in * M:]], dim=1)
        return FusedPConvFunction.apply(feats.contiguous(), nei_inds, inv, weights.contiguous(),
                                        None if additional is None else additional.contiguous(),
                                        None if guidance is None else guidance.contiguous(), lin_w, lin_b)
    g = index_points(feats, nei_inds, inv)
    if guidance is not None:
        g = g * guidance.repeat(1, 1, 1, g.shape[-1] // guidance.shape[-1])
    if additional is not None:
        g = torch.cat([g, additional], dim=-1)
    p = torch.matmul(g.permute(0, 1, 3, 2), weights).reshape(g.shape[0], g.shape[1], -1)
    return p if lin_w is None else linear(p, lin_w, lin_b)


class MultiHeadGuidance(nn.Module):
    """sigmoid(MLP_{dim -> 8 -> heads}(query - key)) (layers.py:23-68; sigmoid, not softmax: SURVEY D1)."""

    def __init__(self, cfg, num_heads, num_hiddens):
        super().__init__()
        self.cfg, self.dim, self.num_heads = cfg, num_hiddens, num_heads
        self.layer_norm_q = nn.LayerNorm(num_hiddens) if cfg.layer_norm_guidance else nn.Identity()
        self.layer_norm_k = nn.LayerNorm(num_hiddens) if cfg.layer_norm_guidance else nn.Identity()
        self.mlp = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        dims = [num_hiddens, 8, num_heads]
        for cin, cout in zip(dims[:-1], dims[1:]):
            self.mlp.append(Linear_BN(cin, cout) if cfg.BATCH_NORM else nn.Linear(cin, cout))

    def forward(self, guidance_query, guidance_key):
        """guidance_key None: guidance_query already is query - key (layer_utils.guidance_input)."""
        s = guidance_query if guidance_key is None else self.layer_norm_q(guidance_query) - self.layer_norm_k(guidance_key)
        last = len(self.mlp) - 1
        spec = _chain_spec(list(self.mlp), [fused_mlp.ACT_RELU] * last + [fused_mlp.ACT_SIGMOID])
        if spec is not None:                       # one fused pass per layer (csrc/mlp.cu)
            return fused_mlp.mlp_chain(s, spec, self.training)
        for i, layer in enumerate(self.mlp):
            s = layer(s) if isinstance(layer, Linear_BN) else linear(s, layer.weight, layer.bias)
            s = torch.sigmoid(s) if i == last else F.relu(s)
        return s


class MultiHeadGuidanceQK(nn.Module):
    """QK-style guidance with sigmoid activation (layers.py:77-114)."""

    def __init__(self, cfg, num_heads, num_hiddens, key_dim):
        super().__init__()
        assert num_hiddens % num_heads == 0
        self.cfg, self.dim, self.num_heads, self.key_dim = cfg, num_hiddens, num_heads, key_dim
        self.scale = key_dim ** -0.5
        self.qk_linear = Linear_BN(num_hiddens, key_dim * num_heads)

    def forward(self, q, k):
        B, N, K, _ = q.shape
        qh = self.qk_linear(q).view(B, N, K, self.num_heads, -1)
        kh = self.qk_linear(k).view(B, N, K, self.num_heads, -1)[:, :, :1]
        return torch.sigmoid((qh * kh).sum(-1) * self.scale)


class WeightNet(nn.Module):
    """Per-edge MLP in -> hidden... -> out with ReLU after every Linear_BN, the last included
    (layers.py:127-171).  `efficient` (gradient checkpointing, 173-191) is accepted and ignored: it only
    trades memory for recompute and the maths is identical."""

    def __init__(self, in_channel, out_channel, hidden_unit=[8, 8], efficient=False):
        super().__init__()
        self.efficient = efficient
        self.mlp_convs = nn.ModuleList()
        dims = [in_channel] + list(hidden_unit or []) + [out_channel]
        for cin, cout in zip(dims[:-1], dims[1:]):
            self.mlp_convs.append(Linear_BN(cin, cout))

    def forward(self, localized_xyz):
        spec = _chain_spec(list(self.mlp_convs), [fused_mlp.ACT_RELU] * len(self.mlp_convs))
        if spec is not None:                       # fused Linear+BN(batch stats)+ReLU chain (csrc/mlp.cu)
            return fused_mlp.mlp_chain(localized_xyz, spec, self.training)
        w = localized_xyz
        for conv in self.mlp_convs:
            w = F.relu(conv(w))
        return w


class _PointLayerBase(nn.Module):
    def _geometry(self, xyz_in, nrm_in, nei_inds, xyz_out, nrm_out, vi_features, use_vi):
        """-> (localized_xyz or None, weightNetInput)."""
        if use_vi and vi_features is not None:
            need_r = getattr(self, "_needs_r", False)
            if not need_r:
                return None, vi_features
            return vi_features[..., 9:12], vi_features        # the last 3 VI channels ARE localized_xyz
        r, vi = edge_geometry(xyz_in, nrm_in, nei_inds, xyz_out, nrm_out, use_vi)
        return r, (vi if use_vi else r)


class PCFLayer(_PointLayerBase):
    """PointConvFormer layer (layers.py:194-416)."""

    def __init__(self, in_channel, out_channel, cfg, weightnet=[9, 16], num_heads=4, guidance_feat_len=32):
        super().__init__()
        self.cfg, self.in_channel, self.out_channel, self.num_heads = cfg, in_channel, out_channel, num_heads
        self.drop_path = _drop_path(cfg)
        self.mlp_conv = Linear_BN(12, guidance_feat_len) if cfg.BATCH_NORM else nn.Linear(12, guidance_feat_len)
        self.unary1 = UnaryBlock(in_channel, out_channel // 4, use_bn=True, bn_momentum=0.1) \
            if in_channel != out_channel // 4 else nn.Identity()
        self.guidance_unary = UnaryBlock(out_channel // 4, guidance_feat_len, use_bn=True, bn_momentum=0.1, no_relu=True)
        assert (out_channel // 2) % num_heads == 0
        if cfg.attention_type == 'subtraction':
            self.guidance_weight = MultiHeadGuidance(cfg, num_heads, 2 * guidance_feat_len)
        else:
            self.guidance_weight = MultiHeadGuidanceQK(cfg, num_heads, 2 * guidance_feat_len, key_dim=16)
        self.weightnet = WeightNet(weightnet[0], weightnet[1], efficient=True)
        lin_in = out_channel // 4 * weightnet[-1]
        self.linear = Linear_BN(lin_in, out_channel // 2, bn_ver='1d') if cfg.BATCH_NORM else nn.Linear(lin_in, out_channel // 2)
        self.dropout = nn.Dropout(p=cfg.dropout_rate) if cfg.dropout_rate > 0. else nn.Identity()
        self.unary2 = UnaryBlock(out_channel // 2, out_channel, use_bn=True, bn_momentum=0.1, no_relu=True)
        self.unary_shortcut = UnaryBlock(in_channel, out_channel, use_bn=True, bn_momentum=0.1, no_relu=True) \
            if in_channel != out_channel else nn.Identity()
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, dense_xyz, dense_feats, nei_inds, dense_xyz_norm, sparse_xyz=None, sparse_xyz_norm=None,
                vi_features=None, inv_neighbors=None, inv_k=None, inv_idx=None):
        N = dense_xyz.shape[1]
        strided = sparse_xyz is not None
        c_xyz, c_nrm = (sparse_xyz, sparse_xyz_norm) if strided else (dense_xyz, dense_xyz_norm)
        M, K = c_xyz.shape[1], nei_inds.shape[2]
        nei_inds = nei_inds.contiguous()
        inv = _inv_tuple(nei_inds, N, inv_neighbors, inv_k, inv_idx, torch.is_grad_enabled() and dense_feats.requires_grad)

        _, weightNetInput = self._geometry(dense_xyz, dense_xyz_norm, nei_inds, c_xyz, c_nrm, vi_features, self.cfg.USE_VI is True)
        # three branches that only meet at the guidance / the contraction / the residual add run on side streams
        par = _streams_ok(self)
        br_w = S.fork(lambda: self.weightnet(weightNetInput), 0, par)
        spec = _chain_spec([self.mlp_conv], [fused_mlp.ACT_RELU])
        br_pe = S.fork(lambda: fused_mlp.mlp_chain(weightNetInput, spec, self.training) if spec is not None
                       else F.relu(self.mlp_conv(weightNetInput)), 1, par)
        has_sc = strided or not isinstance(self.unary_shortcut, nn.Identity)
        br_sc = S.fork(lambda: self.unary_shortcut(gather_max(dense_feats, nei_inds, inv) if strided else dense_feats), 2,
                       par and has_sc)
        feats_x = self.unary1(dense_feats)
        guidance_x = self.guidance_unary(feats_x)
        feat_pe = S.join(br_pe)
        fused_qk = (self.cfg.attention_type == 'subtraction' and not self.cfg.layer_norm_guidance
                    and guidance_x.shape[-1] % 4 == 0 and feat_pe.shape[-1] % 4 == 0)
        if fused_qk:
            # gather + cat + key (column 0 = the centre itself when M == N, T6; else max over the neighbours) + subtraction:
            # one kernel (csrc/glue.cu) instead of five passes over a [M, K, 64] tensor
            guidance_score = self.guidance_weight(guidance_input(guidance_x, feat_pe, nei_inds, inv, M != N), None)
        else:
            guidance_feature = torch.cat([index_points(guidance_x, nei_inds, inv), feat_pe], dim=-1)
            if M == N:
                guidance_key = guidance_feature[:, :, :1, :]
            else:
                guidance_key = guidance_feature.max(dim=2, keepdim=True)[0]
            guidance_score = self.guidance_weight(guidance_feature, guidance_key.expand_as(guidance_feature)
                                                  if self.cfg.attention_type != 'subtraction' else guidance_key)
        weights = S.join(br_w)

        if isinstance(self.linear, Linear_BN):
            lin, post_bn = self.linear.c, self.linear
        else:
            lin, post_bn = self.linear, None
        new_feat = _contract(self.cfg, feats_x, nei_inds, inv, weights, None, guidance_score, lin.weight, lin.bias)
        new_feat = _bn_relu(post_bn.bn, new_feat, lin.bias) if post_bn is not None else F.relu(new_feat)
        new_feat = self.dropout(new_feat)
        shortcut = S.join(br_sc)
        return _block_tail(self, new_feat, shortcut), weightNetInput


def _block_tail(block, new_feat, shortcut):
    """leaky_relu(drop_path(unary2(new_feat)) + shortcut) (layers.py:413-415, 737-739): with drop_path off and a wide unary2
    the residual add and the LeakyReLU ride in unary2's BatchNorm apply pass (fused_mlp.bn_act)."""
    if isinstance(block.drop_path, nn.Identity) and block.unary2.can_fuse_tail() and shortcut.shape == new_feat.shape[:-1] + (block.unary2.out_dim,):
        return block.unary2(new_feat, residual=shortcut, act=fused_mlp.ACT_LEAKY)
    return block.leaky_relu(block.drop_path(block.unary2(new_feat)) + shortcut)


def _streams_ok(module):
    """Side streams are used unless the module's BatchNorms exchange statistics across ranks without per-stream exchange
    channels (fused_mlp.SYNC_CHANNELS)."""
    if not S.ENABLED:
        return False
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and not fused_mlp.SYNC_CHANNELS:
        return not any(isinstance(m, nn.SyncBatchNorm) for m in module.modules())
    return True


def _bn_relu(bn, x, pivot=None, skip=None):
    """ReLU(BatchNorm(x)) [+ skip] over the last dim of [B,N,C] with a BatchNorm1d-like module `bn` (batch statistics over all
    points of the packed batch when training, layer_utils.py:276-277; layers.py:708-709,721): two passes of pcfb_bn_*;
    skip (the decoder's `new_feat + dense_feats`, layers.py:1096-1097) is added after the ReLU in the same pass."""
    if fused_mlp.bn_supported(x.shape[-1]):
        return fused_mlp.bn_act(x, bn, fused_mlp.ACT_RELU, pivot=pivot, residual=skip, residual_after_act=True)
    if skip is not None:
        return _bn_relu(bn, x, pivot) + skip
    shape = x.shape
    if isinstance(bn, nn.SyncBatchNorm):
        return F.relu(bn(x.reshape(-1, shape[-1])).reshape(shape))
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return F.relu(F.batch_norm(x.reshape(-1, shape[-1]), bn.running_mean, bn.running_var, bn.weight, bn.bias,
                               bn.training or not bn.track_running_stats, 0.0 if bn.momentum is None else bn.momentum,
                               bn.eps).reshape(shape))


class _PConvLinearMixin:
    """The Linear(+BN) after the contraction in both parameter spellings:
    PCONV_OPT True  -> self.pconv_linear_opt.linear + self.bn ; False -> self.linear (Linear_BN or Linear)."""

    def _build_linear(self, cfg, lin_in, lin_out):
        if cfg.PCONV_OPT:
            self.pconv_linear_opt = PConvLinearOpt(lin_in, lin_out)
            if cfg.BATCH_NORM:
                self.bn = nn.BatchNorm1d(lin_out, momentum=0.1)
        else:
            self.linear = Linear_BN(lin_in, lin_out, bn_ver='1d') if cfg.BATCH_NORM else nn.Linear(lin_in, lin_out)

    def _contract_linear(self, feats, nei_inds, inv, weights, additional, skip=None):
        """-> ReLU(BN(Linear(contraction))) [+ skip]: every caller in the reference applies ReLU right after the BatchNorm
        (layers.py:709,721, 898, 1092), so the activation rides in the BatchNorm's apply pass."""
        cfg = self.cfg
        if cfg.PCONV_OPT:
            lin, bn = self.pconv_linear_opt.linear, (self.bn if cfg.BATCH_NORM else None)
        elif isinstance(self.linear, Linear_BN):
            lin, bn = self.linear.c, self.linear.bn
        else:
            lin, bn = self.linear, None
        y = _contract(cfg, feats, nei_inds, inv, weights, additional, None, lin.weight, lin.bias)
        if bn is None:
            return F.relu(y) if skip is None else F.relu(y) + skip
        return _bn_relu(bn, y, lin.bias, skip)


class PointConvStridePE(_PointLayerBase, _PConvLinearMixin):
    """PointConv bottleneck block with positional-encoding features (layers.py:542-741)."""
    _needs_r = True

    def __init__(self, in_channel, out_channel, cfg, weightnet=[9, 16]):
        super().__init__()
        self.cfg, self.in_channel, self.out_channel = cfg, in_channel, out_channel
        self.drop_path = _drop_path(cfg)
        last_ch = min(out_channel // 4, 32)
        self.pe_convs = WeightNet(3, last_ch, hidden_unit=[out_channel // 4], efficient=True)
        self.unary1 = UnaryBlock(in_channel, out_channel // 4, use_bn=True, bn_momentum=0.1) \
            if in_channel != out_channel // 4 else nn.Identity()
        self.weightnet = WeightNet(weightnet[0], weightnet[1], efficient=True)
        self._build_linear(cfg, (out_channel // 4 + last_ch) * weightnet[-1], out_channel // 2)
        self.dropout = nn.Dropout(p=cfg.dropout_rate) if cfg.dropout_rate > 0. else nn.Identity()
        self.unary2 = UnaryBlock(out_channel // 2, out_channel, use_bn=True, bn_momentum=0.1, no_relu=True)
        self.unary_shortcut = UnaryBlock(in_channel, out_channel, use_bn=True, bn_momentum=0.1, no_relu=True) \
            if in_channel != out_channel else nn.Identity()
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, dense_xyz, dense_feats, nei_inds, dense_xyz_norm, sparse_xyz=None, sparse_xyz_norm=None,
                vi_features=None, inv_neighbors=None, inv_k=None, inv_idx=None):
        N = dense_xyz.shape[1]
        strided = sparse_xyz is not None
        c_xyz, c_nrm = (sparse_xyz, sparse_xyz_norm) if strided else (dense_xyz, dense_xyz_norm)
        nei_inds = nei_inds.contiguous()
        inv = _inv_tuple(nei_inds, N, inv_neighbors, inv_k, inv_idx, torch.is_grad_enabled() and dense_feats.requires_grad)

        localized_xyz, weightNetInput = self._geometry(dense_xyz, dense_xyz_norm, nei_inds, c_xyz, c_nrm, vi_features,
                                                       self.cfg.USE_VI is True)
        par = _streams_ok(self)
        br_w = S.fork(lambda: self.weightnet(weightNetInput), 0, par)
        br_pe = S.fork(lambda: self.pe_convs(localized_xyz), 1, par)
        has_sc = strided or not isinstance(self.unary_shortcut, nn.Identity)
        br_sc = S.fork(lambda: self.unary_shortcut(gather_max(dense_feats, nei_inds, inv) if strided else dense_feats), 2,
                       par and has_sc)
        feats_x = self.unary1(dense_feats)
        feat_pe = S.join(br_pe)
        weights = S.join(br_w)
        new_feat = self.dropout(self._contract_linear(feats_x, nei_inds, inv, weights, feat_pe))
        shortcut = S.join(br_sc)
        return _block_tail(self, new_feat, shortcut), weightNetInput


class PointConv(_PointLayerBase, _PConvLinearMixin):
    """VI_PointConv / PointConv without bottleneck, used for the first layer (layers.py:744-906)."""

    def __init__(self, in_channel, out_channel, cfg, weightnet=[9, 16], USE_VI=None):
        super().__init__()
        self.cfg, self.in_channel, self.out_channel = cfg, in_channel, out_channel
        self.USE_VI = cfg.USE_VI if USE_VI is None else USE_VI
        last_ch = in_channel + ((12 if self.USE_VI else 3) if cfg.USE_PE else 0)
        self.weightnet = WeightNet(weightnet[0], weightnet[1], efficient=True)
        self._build_linear(cfg, last_ch * weightnet[-1], out_channel)
        self.dropout = nn.Dropout(p=cfg.dropout_rate) if cfg.dropout_rate > 0. else nn.Identity()

    def forward(self, dense_xyz, dense_feats, nei_inds, dense_xyz_norm=None, sparse_xyz=None, sparse_xyz_norm=None,
                inv_neighbors=None, inv_k=None, inv_idx=None):
        N = dense_xyz.shape[1]
        c_xyz, c_nrm = (sparse_xyz, sparse_xyz_norm) if sparse_xyz is not None else (dense_xyz, dense_xyz_norm)
        nei_inds = nei_inds.contiguous()
        inv = _inv_tuple(nei_inds, N, inv_neighbors, inv_k, inv_idx, torch.is_grad_enabled() and dense_feats.requires_grad)
        _, weightNetInput = self._geometry(dense_xyz, dense_xyz_norm, nei_inds, c_xyz, c_nrm, None, self.USE_VI is True)
        additional = weightNetInput if self.cfg.USE_PE else None
        weights = self.weightnet(weightNetInput)
        new_feat = self._contract_linear(dense_feats, nei_inds, inv, weights, additional)
        return self.dropout(new_feat), weightNetInput


class PointConvTransposePE(_PointLayerBase, _PConvLinearMixin):
    """Upsampling PointConv: features of the sparse cloud are pulled to the dense points
    (layers.py:909-1105)."""
    _needs_r = True

    def __init__(self, in_channel, out_channel, cfg, weightnet=[9, 16], mlp2=None):
        super().__init__()
        self.cfg, self.in_channel, self.out_channel = cfg, in_channel, out_channel
        self.drop_path = _drop_path(cfg)
        if cfg.USE_PE:
            last_ch = min(out_channel // 4, 32)
            self.pe_convs = WeightNet(3, last_ch, hidden_unit=[out_channel // 4], efficient=True)
        else:
            last_ch = 0
            self.pe_convs = nn.ModuleList()
        self.weightnet = WeightNet(weightnet[0], weightnet[1], efficient=True)
        self._build_linear(cfg, (last_ch + in_channel) * weightnet[-1], out_channel)
        self.dropout = nn.Dropout(p=cfg.dropout_rate) if cfg.dropout_rate > 0. else nn.Identity()
        self.mlp2_convs = nn.ModuleList()
        self.mlp2_bns = nn.ModuleList()
        if mlp2 is not None:
            for i in range(1, len(mlp2)):
                self.mlp2_convs.append(Linear_BN(mlp2[i - 1], mlp2[i], bn_ver='1d') if cfg.BATCH_NORM
                                       else nn.Linear(mlp2[i - 1], mlp2[i]))

    def forward(self, sparse_xyz, sparse_feats, nei_inds, sparse_xyz_norm, dense_xyz, dense_xyz_norm,
                dense_feats=None, vi_features=None, inv_neighbors=None, inv_k=None, inv_idx=None):
        n_in = sparse_xyz.shape[1]
        nei_inds = nei_inds.contiguous()
        inv = _inv_tuple(nei_inds, n_in, inv_neighbors, inv_k, inv_idx, torch.is_grad_enabled() and sparse_feats.requires_grad)
        if inv is not None and inv[2].shape[1] != n_in + 1:
            # compute_knn_inverse pads propagate maps to the dense level's size (common_util.py:303-306)
            inv = (inv[0], inv[1], inv[2][:, :n_in + 1].contiguous())
        localized_xyz, weightNetInput = self._geometry(sparse_xyz, sparse_xyz_norm, nei_inds, dense_xyz, dense_xyz_norm,
                                                       vi_features, self.cfg.USE_VI is True)
        par = _streams_ok(self)
        br_pe = S.fork(lambda: self.pe_convs(localized_xyz) if self.cfg.USE_PE else None, 1, par and self.cfg.USE_PE)
        weights = self.weightnet(weightNetInput)
        feat_pe = S.join(br_pe)
        new_feat = self._contract_linear(sparse_feats, nei_inds, inv, weights, feat_pe, skip=dense_feats)
        new_feat = self.dropout(new_feat)
        for conv in self.mlp2_convs:
            new_feat = conv(new_feat, act=fused_mlp.ACT_RELU) if isinstance(conv, Linear_BN) else \
                F.relu(linear(new_feat, conv.weight, conv.bias))
        return new_feat, weightNetInput
