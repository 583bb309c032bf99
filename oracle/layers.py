"""CPU oracle: PointConv / PointConvFormer layers and the segmentation model as a *functional*
torch-fp32 restatement.  TEST INFRASTRUCTURE (see oracle/__init__.py); PINNED against golden vectors
produced by the unmodified reference (tests/golden/make_golden.py).

All functions take a flat ``params`` dict that uses the reference's state-dict key names (both
spellings of trap T5: ``linear.c.* / linear.bn.*`` for PCONV_OPT=False and
``pconv_linear_opt.linear.* / bn.*`` for PCONV_OPT=True) under a ``prefix``; gradients come from
autograd (never from the reference's CUDA backward, SURVEY.md trap T1).

Tensor conventions are the reference's: leading batch dim 1, features [1,N,C] fp32, nei_inds
[1,M,K] int64 indexing the *input* cloud.

Restated reference code:
  index_points              layer_utils.py:13-30
  VI_coordinate_transform   layer_utils.py:176-231
  Linear_BN                 layer_utils.py:241-277    UnaryBlock layer_utils.py:281-319
  WeightNet                 layers.py:127-191
  MultiHeadGuidance         layers.py:23-68
  PCFLayer                  layers.py:194-416
  PointConvStridePE         layers.py:542-741
  PointConv                 layers.py:744-906
  PointConvTransposePE      layers.py:909-1105
  PCF_Backbone / PointConvFormer_Segmentation   model_architecture.py:80-245, 345-502
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def gather(points, idx):
    """index_points (layer_utils.py:13-30) for B == 1: [1,N,C],[1,M,K] -> [1,M,K,C]."""
    return points[0][idx[0]].unsqueeze(0)


def _has(params, key):
    return key in params


def linear_bn(x, params, prefix, training):
    """Linear_BN.forward (layer_utils.py:272-277): Linear then BatchNorm over the channel (last) dim,
    batch statistics over every other element when training; also accepts a BN-folded / plain
    ``nn.Linear`` (keys prefix.weight / prefix.bias) as produced by Linear_BN.fuse (260-270)."""
    if _has(params, prefix + ".c.weight"):
        y = F.linear(x, params[prefix + ".c.weight"], params[prefix + ".c.bias"])
        return batch_norm_last(y, params, prefix + ".bn", training)
    return F.linear(x, params[prefix + ".weight"], params[prefix + ".bias"])


def batch_norm_last(y, params, prefix, training):
    """BatchNorm over the last (channel) dim, evaluated on the channel-second permuted *view* exactly as
    Linear_BN.forward does (layer_utils.py:274-277).  The layout matters numerically on CPU: torch's
    strided BN kernel accumulates the batch sums in double, the contiguous [rows, C] kernel in float,
    and behind chains of BatchNorms the latter costs ~1e-2 relative accuracy in weight gradients."""
    rm = params[prefix + ".running_mean"].clone()
    rv = params[prefix + ".running_var"].clone()
    perm = (0, 3, 2, 1) if y.dim() == 4 else (0, 2, 1)
    out = F.batch_norm(y.permute(*perm), rm, rv, params[prefix + ".weight"], params[prefix + ".bias"],
                       training, 0.1, BN_EPS)
    return out.permute(*perm)


def unary_block(x, params, prefix, training, relu=True):
    """UnaryBlock (layer_utils.py:281-315): Linear_BN('1d') [+ LeakyReLU(0.1)]; absent -> Identity."""
    if not (_has(params, prefix + ".mlp.c.weight") or _has(params, prefix + ".mlp.weight")):
        return x
    y = linear_bn(x, params, prefix + ".mlp", training)
    return F.leaky_relu(y, 0.1) if relu else y


def weightnet(x, params, prefix, training):
    """WeightNet.real_forward (layers.py:163-171): ReLU after *every* Linear_BN, including the last."""
    i = 0
    while _has(params, "%s.mlp_convs.%d.c.weight" % (prefix, i)) or _has(params, "%s.mlp_convs.%d.weight" % (prefix, i)):
        x = F.relu(linear_bn(x, params, "%s.mlp_convs.%d" % (prefix, i), training))
        i += 1
    assert i > 0, "no weightnet under " + prefix
    return x


def vi_transform(r, n_j, n_i):
    """VI_coordinate_transform (layer_utils.py:176-231).  r [1,M,K,3] localized xyz, n_j [1,M,K,3]
    gathered normals, n_i [1,M,3] centre normals -> [1,M,K,12]."""
    n_i = n_i.unsqueeze(2)
    r_hat = F.normalize(r, dim=3)
    v = n_i - (n_i * r_hat).sum(3, keepdim=True) * r_hat
    v = F.normalize(v, dim=3)
    w = F.normalize(torch.cross(r_hat, v, dim=3), dim=3)
    dot = lambda a, b: (a * b).sum(3, keepdim=True)
    t1 = dot(n_j, n_i)
    t2 = dot(r_hat, n_i)
    t3 = dot(r_hat, n_j)
    t4 = dot(r, n_i)
    t5 = t3
    t6 = dot(n_j, v)
    t7 = dot(n_j, w)
    t8 = dot(r, torch.cross(n_j, n_i.expand_as(n_j), dim=3))
    t9 = r.norm(dim=3, keepdim=True)
    return torch.cat([t1, t2, t3, t4, t5, t6, t7, t8, t9, r], dim=3)


def _edge_geometry(xyz_in, nrm_in, nei, xyz_out, nrm_out, use_vi, vi_features):
    r = gather(xyz_in, nei) - xyz_out.unsqueeze(2)
    if not use_vi:
        return r, r
    if vi_features is not None:
        return r, vi_features
    return r, vi_transform(r, gather(nrm_in, nei), nrm_out)


def pconv(feats, nei, weights, additional=None, guidance=None):
    """P[m, c*C_mid + j] = sum_k cat(x[nei[m,k]] (* guidance[m,k,c % H]), add[m,k])[c] * w[m,k,j]
    (layers.py:386-390 PCF, 713-716 StridePE, 894-897 PointConv, 1086-1089 Transpose)."""
    g = gather(feats, nei)
    if guidance is not None:
        H = guidance.shape[-1]
        C = g.shape[-1]
        g = g * guidance.repeat(1, 1, 1, C // H)       # channel c uses head c % H
    if additional is not None:
        g = torch.cat([g, additional], dim=-1)
    p = torch.einsum("bmkc,bmkj->bmcj", g, weights)
    return p.reshape(p.shape[0], p.shape[1], -1)


def _pconv_linear(P, params, prefix, training, batch_norm=True):
    """The Linear(+BN) after the PConv, either spelling (T5)."""
    if _has(params, prefix + ".pconv_linear_opt.linear.weight"):
        y = F.linear(P, params[prefix + ".pconv_linear_opt.linear.weight"],
                     params[prefix + ".pconv_linear_opt.linear.bias"])
        if _has(params, prefix + ".bn.weight"):
            y = batch_norm_last(y, params, prefix + ".bn", training)
        return y
    return linear_bn(P, params, prefix + ".linear", training)


def point_conv(params, prefix, cfg, xyz, feats, nei, nrm=None, sparse_xyz=None, sparse_nrm=None,
               training=True):
    """PointConv.forward (layers.py:813-906) -> (new_feat, weightNetInput)."""
    c_xyz = xyz if sparse_xyz is None else sparse_xyz
    c_nrm = nrm if sparse_xyz is None else sparse_nrm
    _, wni = _edge_geometry(xyz, nrm, nei, c_xyz, c_nrm, cfg["USE_VI"], None)
    add = wni if cfg["USE_PE"] else None
    w = weightnet(wni, params, prefix + ".weightnet", training)
    P = pconv(feats, nei, w, add)
    return F.relu(_pconv_linear(P, params, prefix, training)), wni


def point_conv_stride_pe(params, prefix, cfg, xyz, feats, nei, nrm, sparse_xyz=None, sparse_nrm=None,
                         vi_features=None, training=True):
    """PointConvStridePE.forward (layers.py:631-741)."""
    c_xyz = xyz if sparse_xyz is None else sparse_xyz
    c_nrm = nrm if sparse_xyz is None else sparse_nrm
    x = unary_block(feats, params, prefix + ".unary1", training)
    r, wni = _edge_geometry(xyz, nrm, nei, c_xyz, c_nrm, cfg["USE_VI"], vi_features)
    pe = weightnet(r, params, prefix + ".pe_convs", training)
    w = weightnet(wni, params, prefix + ".weightnet", training)
    h = F.relu(_pconv_linear(pconv(x, nei, w, pe), params, prefix, training))
    h = unary_block(h, params, prefix + ".unary2", training, relu=False)
    s = gather(feats, nei).max(dim=2)[0] if sparse_xyz is not None else feats
    s = unary_block(s, params, prefix + ".unary_shortcut", training, relu=False)
    return F.leaky_relu(h + s, 0.1), wni


def guidance_scores(q, key, params, prefix, training):
    """MultiHeadGuidance.forward (layers.py:47-68), subtraction attention, sigmoid (D1: not softmax),
    layer_norm_guidance=False."""
    s = q - key
    s = F.relu(linear_bn(s, params, prefix + ".mlp.0", training))
    return torch.sigmoid(linear_bn(s, params, prefix + ".mlp.1", training))


def pcf_layer(params, prefix, cfg, xyz, feats, nei, nrm, sparse_xyz=None, sparse_nrm=None,
              vi_features=None, training=True):
    """PCFLayer.forward (layers.py:306-416), attention_type='subtraction'."""
    strided = sparse_xyz is not None
    c_xyz = sparse_xyz if strided else xyz
    c_nrm = sparse_nrm if strided else nrm
    x = unary_block(feats, params, prefix + ".unary1", training)
    _, wni = _edge_geometry(xyz, nrm, nei, c_xyz, c_nrm, cfg["USE_VI"], vi_features)
    pe = F.relu(linear_bn(wni, params, prefix + ".mlp_conv", training))
    gx = unary_block(x, params, prefix + ".guidance_unary", training, relu=False)
    q = torch.cat([gather(gx, nei), pe], dim=-1)
    M, N = c_xyz.shape[1], xyz.shape[1]
    key = q[:, :, :1] if M == N else q.max(dim=2, keepdim=True)[0]     # layers.py:377-381 (T6)
    g = guidance_scores(q, key, params, prefix + ".guidance_weight", training)
    w = weightnet(wni, params, prefix + ".weightnet", training)
    P = pconv(x, nei, w, None, guidance=g)
    h = F.relu(linear_bn(P, params, prefix + ".linear", training))
    h = unary_block(h, params, prefix + ".unary2", training, relu=False)
    s = gather(feats, nei).max(dim=2)[0] if strided else feats
    s = unary_block(s, params, prefix + ".unary_shortcut", training, relu=False)
    return F.leaky_relu(h + s, 0.1), wni


def point_conv_transpose_pe(params, prefix, cfg, sparse_xyz, sparse_feats, nei, sparse_nrm, dense_xyz,
                            dense_nrm, dense_feats=None, vi_features=None, training=True):
    """PointConvTransposePE.forward (layers.py:1000-1105)."""
    r, wni = _edge_geometry(sparse_xyz, sparse_nrm, nei, dense_xyz, dense_nrm, cfg["USE_VI"], vi_features)
    pe = weightnet(r, params, prefix + ".pe_convs", training) if cfg["USE_PE"] else None
    w = weightnet(wni, params, prefix + ".weightnet", training)
    h = F.relu(_pconv_linear(pconv(sparse_feats, nei, w, pe), params, prefix, training))
    if dense_feats is not None:
        h = h + dense_feats
    i = 0
    while _has(params, "%s.mlp2_convs.%d.c.weight" % (prefix, i)) or _has(params, "%s.mlp2_convs.%d.weight" % (prefix, i)):
        h = F.relu(linear_bn(h, params, "%s.mlp2_convs.%d" % (prefix, i), training))
        i += 1
    return h, wni


def backbone(params, prefix, cfg, features, pcs, e_self, e_fwd, norms, training=True):
    """PCF_Backbone.forward (model_architecture.py:175-245), transformer_type='PCF'."""
    x = torch.cat([features, pcs[0]], -1) if cfg["USE_XYZ"] else features
    if cfg["use_level_1"]:
        x, vi = point_conv(params, prefix + "selfpointconv", cfg, pcs[0], x, e_self[0], norms[0], training=training)
        x, _ = point_conv_stride_pe(params, prefix + "selfpointconv_res1", cfg, pcs[0], x, e_self[0], norms[0],
                                    vi_features=vi, training=training)
        x, _ = point_conv_stride_pe(params, prefix + "selfpointconv_res2", cfg, pcs[0], x, e_self[0], norms[0],
                                    vi_features=vi, training=training)
    else:
        x = F.relu(linear_bn(x, params, prefix + "selfmlp", training))
    feats = [x]
    for i in range(cfg["num_level"] - 1):
        lvl = i + 1
        layer = point_conv_stride_pe if lvl <= cfg["guided_level"] else pcf_layer
        x, _ = layer(params, "%spointconv.%d" % (prefix, i), cfg, pcs[i], feats[-1], e_fwd[i], norms[i],
                     pcs[i + 1], norms[i + 1], training=training)
        vi = None
        for b in range(cfg["resblocks"][lvl]):
            x, vi_new = layer(params, "%spointconv_res.%d.%d" % (prefix, i, b), cfg, pcs[lvl], x, e_self[lvl],
                              norms[lvl], vi_features=vi, training=training)
            vi = vi_new if vi is None else vi
        feats.append(x)
    return feats


def segmentation_model(params, cfg, features, pcs, e_self, e_fwd, e_prop, norms, training=True):
    """PointConvFormer_Segmentation.forward (model_architecture.py:406-502), decoder res-blocks included
    (resblocks_back is all 0 in every shipped config; tests/golden/model_routing.npz exercises them)."""
    feats = backbone(params, "pcf_backbone.", cfg, features, pcs, e_self, e_fwd, norms, training)
    x = feats[-1]
    L = cfg["num_level"]
    for i in range(L - 1):
        lvl = L - 2 - i
        x, _ = point_conv_transpose_pe(params, "pointdeconv.%d" % i, cfg, pcs[lvl + 1], x, e_prop[lvl],
                                       norms[lvl + 1], pcs[lvl], norms[lvl], feats[lvl], training=training)
        n_res = cfg.get("resblocks_back", [0] * L)[lvl] if cfg["resblocks"][lvl] != 0 else 0
        vi = None                                  # decoder res-blocks (model_architecture.py:391-398, 466-494)
        for b in range(n_res):
            x, vi_new = point_conv_stride_pe(params, "pointdeconv_res.%d.%d" % (i, b), cfg, pcs[lvl], x, e_self[lvl],
                                             norms[lvl], vi_features=vi, training=training)
            vi = vi_new if vi is None else vi
        feats[lvl] = x
    h = F.relu(linear_bn(x, params, "fc1", training))
    return F.linear(h, params["fc2.weight"], params["fc2.bias"])
