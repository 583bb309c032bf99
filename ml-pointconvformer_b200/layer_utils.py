"""Host-side mirror of the reference's layer_utils.py (/root/reference/layer_utils.py) for the B200 path:
same public names (index_points, PConvLinearOptFunction, PConvLinearOpt, PCFFunction, PCF, PConvFunction,
PConv, VI_coordinate_transform, Linear_BN, UnaryBlock) and argument meaning, every gather / contraction
routed to libpcf_b200.so through pcf_cuda.py.  Sub-module names (`c`, `bn`, `mlp`, `linear`) are the
reference's, so its checkpoints load unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pcf_cuda
from . import streams as S


# ------------------------------------------------------------------------------------------------
# inverse-map plumbing
# ------------------------------------------------------------------------------------------------
def resolve_inverse(nei_inds, total_points, inv_neighbors=None, inv_k=None, inv_idx=None):
    """Returns (inv_neighbors, inv_k, inv_idx) for `nei_inds`.  Uses the caller's maps when given (casting
    as the reference's tests do with .int()/.byte()/.int(), tests_pointconv/encoder.py:110-112), otherwise
    builds them once and caches them on the index tensor (the reference's eval drivers never build them,
    SURVEY.md T8)."""
    if inv_neighbors is not None:
        return (inv_neighbors.to(torch.int32).contiguous(), inv_k.to(torch.uint8).contiguous(),
                inv_idx.to(torch.int32).contiguous())
    # the cache key carries the tensor's version counter and address: an index tensor that is overwritten in place (static
    # input buffers, reused loader tensors) gets a fresh map instead of a stale one
    key = (total_points, nei_inds._version, nei_inds.data_ptr())
    cached = getattr(nei_inds, "_pcfb_inverse", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    inv = pcf_cuda.compute_knn_inverse(nei_inds.contiguous(), total_points)
    try:
        nei_inds._pcfb_inverse = (key, inv)
    except Exception:
        pass
    return inv


# ------------------------------------------------------------------------------------------------
# index_points (layer_utils.py:13-30) with an atomics-free backward
# ------------------------------------------------------------------------------------------------
class _GatherFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, inv):
        ctx.inv = inv
        ctx.idx = idx
        ctx.n_in = points.shape[1]
        return torch.stack([pcf_cuda.gather(points[b], idx[b]) for b in range(points.shape[0])])

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        inv = ctx.inv if ctx.inv is not None else resolve_inverse(ctx.idx, ctx.n_in)
        g = torch.stack([pcf_cuda.gather_backward(grad[b], (inv[0][b], inv[1][b], inv[2][b]), ctx.n_in)
                         for b in range(grad.shape[0])])
        return g, None, None


def index_points(points, idx, inv=None):
    """points [B,N,C], idx [B,S,K] (or [B,S]) -> [B,S,K,C] (or [B,S,C])."""
    if idx.dim() == 2:
        return index_points(points, idx.unsqueeze(-1), None).squeeze(2)
    points = points.contiguous().float()
    idx = idx.contiguous()
    if points.requires_grad and torch.is_grad_enabled():
        return _GatherFunction.apply(points, idx, inv)
    return torch.stack([pcf_cuda.gather(points[b], idx[b]) for b in range(points.shape[0])])


class _GatherMaxFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx, inv):
        outs, args = zip(*[pcf_cuda.gather_max(points[b], idx[b]) for b in range(points.shape[0])])
        ctx.inv, ctx.idx, ctx.n_in = inv, idx, points.shape[1]
        ctx.save_for_backward(torch.stack(args))
        return torch.stack(outs)

    @staticmethod
    def backward(ctx, grad):
        (arg,) = ctx.saved_tensors
        grad = grad.contiguous()
        inv = ctx.inv if ctx.inv is not None else resolve_inverse(ctx.idx, ctx.n_in)
        K = ctx.idx.shape[2]
        g = torch.stack([pcf_cuda.gather_max_backward(grad[b], arg[b], (inv[0][b], inv[1][b], inv[2][b]), ctx.n_in, K)
                         for b in range(grad.shape[0])])
        return g, None, None


def gather_max(points, idx, inv=None):
    """max_k index_points(points, idx) (the strided shortcut, layers.py:403-408,728-733) without
    materialising the [B,S,K,C] gather."""
    return _GatherMaxFunction.apply(points.contiguous(), idx.contiguous(), inv)


class _GuidanceInputFunction(torch.autograd.Function):
    """q - key with q = cat(index_points(guidance_x, nei), feat_pe), key = q[:, :, :1] (M == N) or max_k q
    (layers.py:372-382) in one kernel; the backward emits d feat_pe and the per-edge gradient of the gathered half, which
    the kNN inverse map sums per input point (no atomics)."""

    @staticmethod
    def forward(ctx, guidance_x, feat_pe, nei, inv, use_max):
        outs, args = zip(*[pcf_cuda.guidance_input(guidance_x[b], feat_pe[b], nei[b], use_max) for b in range(nei.shape[0])])
        ctx.inv, ctx.nei, ctx.use_max = inv, nei, use_max
        ctx.n_in, ctx.G, ctx.P = guidance_x.shape[1], guidance_x.shape[2], feat_pe.shape[3]
        ctx.save_for_backward(*[a for a in args if a is not None])
        return torch.stack(outs) if len(outs) > 1 else outs[0].unsqueeze(0)

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        args = ctx.saved_tensors
        want_gx, want_pe = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_gx, g_pe = [], []
        inv = None
        if want_gx:
            inv = ctx.inv if ctx.inv is not None else resolve_inverse(ctx.nei, ctx.n_in)
        for b in range(grad.shape[0]):
            d_gq, d_pe = pcf_cuda.guidance_input_backward(grad[b], args[b] if ctx.use_max else None, ctx.G, ctx.P, ctx.use_max,
                                                          want_gx, want_pe)
            if want_gx:
                g_gx.append(pcf_cuda.gather_backward(d_gq, (inv[0][b], inv[1][b], inv[2][b]), ctx.n_in))
            g_pe.append(d_pe)
        stack = lambda lst: torch.stack(lst) if len(lst) > 1 else lst[0].unsqueeze(0)
        return (stack(g_gx) if want_gx else None), (stack(g_pe) if want_pe else None), None, None, None


def guidance_input(guidance_x, feat_pe, nei_inds, inv, use_max):
    """[B,N,G], [B,M,K,P], [B,M,K] -> [B,M,K,G+P] = cat(gather, pe) - key."""
    return _GuidanceInputFunction.apply(guidance_x.contiguous(), feat_pe.contiguous(), nei_inds, inv, bool(use_max))


# ------------------------------------------------------------------------------------------------
# fused contraction (+guidance) + Linear
# ------------------------------------------------------------------------------------------------
class FusedPConvFunction(torch.autograd.Function):
    """P = PConv(input[*guidance], additional; weightnet);  Y = P W^T + b  in one kernel (forward) and the
    inverse-map driven backward.  linear_weights may be None (then P is the output)."""

    @staticmethod
    def forward(ctx, input_feat, neighbor_inds, inv, weightnet, additional_features, guidance, linear_weights, linear_bias):
        has_lin = linear_weights is not None
        need_grad = any(t is not None and t.requires_grad for t in
                        (input_feat, weightnet, additional_features, guidance, linear_weights, linear_bias))
        y, p = pcf_cuda.pconv_fused_forward(input_feat, neighbor_inds, weightnet, additional_features, guidance,
                                            linear_weights, linear_bias, want_p=(need_grad or not has_lin))
        ctx.inv = inv
        ctx.has_lin = has_lin
        ctx.has_bias = linear_bias is not None
        ctx.save_for_backward(input_feat, neighbor_inds, weightnet, additional_features, guidance, linear_weights,
                              p if has_lin else None)
        return y if has_lin else p

    @staticmethod
    def backward(ctx, grad_output):
        input_feat, neighbor_inds, weightnet, additional_features, guidance, linear_weights, pconv_output = ctx.saved_tensors
        grad_output = grad_output.contiguous()
        ng = ctx.needs_input_grad
        need = (ng[0], ng[3], ng[4], ng[5], ng[6], ng[7] and ctx.has_bias)
        inv = None
        if need[0]:
            inv = ctx.inv if ctx.inv is not None else resolve_inverse(neighbor_inds, input_feat.shape[1])
        g_in, g_w, g_add, g_gd, g_lw, g_lb = pcf_cuda.pconv_fused_backward(
            grad_output if ctx.has_lin else None, None if ctx.has_lin else grad_output, input_feat, inv, neighbor_inds,
            weightnet, additional_features, guidance, linear_weights, pconv_output, need)
        return g_in, None, None, g_w, g_add, g_gd, g_lw, g_lb


class PConvLinearOptFunction(torch.autograd.Function):
    """Same argument list as the reference's (layer_utils.py:42-70)."""

    @staticmethod
    def forward(ctx, input_feat, neighbor_inds, inverse_neighbors, inverse_k, inverse_idx,
                weightnet, additional_features, linear_weights, linear_bias):
        output, pconv_output = pcf_cuda.pconv_linear_cutlass_forward(
            input_feat, neighbor_inds, weightnet, additional_features, linear_weights, linear_bias)
        ctx.save_for_backward(input_feat, inverse_neighbors, inverse_k, inverse_idx, neighbor_inds, weightnet,
                              additional_features, linear_weights, pconv_output)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        (input_feat, inverse_neighbors, inverse_k, inverse_idx, neighbor_inds, weightnet, additional_features,
         linear_weights, pconv_output) = ctx.saved_tensors
        grads = pcf_cuda.pconv_linear_opt_backward(grad_output.contiguous(), input_feat, inverse_neighbors, inverse_k,
                                                   inverse_idx, neighbor_inds, weightnet, additional_features,
                                                   linear_weights, pconv_output)
        return grads[0], None, None, None, None, grads[1], grads[2], grads[3], grads[4]


class PConvLinearOpt(nn.Module):
    """Fused PConv + Linear layer (layer_utils.py:73-86); parameter name `linear` as in the reference."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features)

    def forward(self, input_features, neighbor_inds, inverse_neighbors, inverse_k, inverse_idx, weightnet,
                additional_features=None, guidance=None):
        inv = resolve_inverse(neighbor_inds, input_features.shape[1], inverse_neighbors, inverse_k, inverse_idx) \
            if (torch.is_grad_enabled() and input_features.requires_grad) else None
        return FusedPConvFunction.apply(input_features.contiguous(), neighbor_inds.contiguous(), inv,
                                        weightnet.contiguous(),
                                        None if additional_features is None else additional_features.contiguous(),
                                        None if guidance is None else guidance.contiguous(),
                                        self.linear.weight, self.linear.bias)


class PCFFunction(torch.autograd.Function):
    """layer_utils.py:89-106."""

    @staticmethod
    def forward(ctx, input_feat, neighbor_inds, guidance, weightnet):
        output = pcf_cuda.pcf_forward(input_feat, neighbor_inds, guidance, weightnet)
        ctx.save_for_backward(input_feat, neighbor_inds, guidance, weightnet)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        grad_input, grad_guidance, grad_weight = pcf_cuda.pcf_backward(grad_output.contiguous(), *ctx.saved_tensors)
        return grad_input, None, grad_guidance, grad_weight


class PCF(nn.Module):
    """Fused gather -> guided contraction (layer_utils.py:109-124)."""

    @staticmethod
    def forward(input_features, neighbor_inds, guidance, weightnet):
        return PCFFunction.apply(input_features, neighbor_inds, guidance, weightnet)


class PConvFunction(torch.autograd.Function):
    """layer_utils.py:127-153."""

    @staticmethod
    def forward(ctx, input_feat, neighbor_inds, weightnet, additional_features):
        output = pcf_cuda.pconv_forward(input_feat, neighbor_inds, weightnet, additional_features)
        ctx.save_for_backward(input_feat, neighbor_inds, weightnet, additional_features)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        grad_input, grad_weight, grad_additional = pcf_cuda.pconv_backward(grad_output.contiguous(), *ctx.saved_tensors)
        return grad_input, None, grad_weight, grad_additional


class PConv(nn.Module):
    """layer_utils.py:156-173."""

    @staticmethod
    def forward(input_features, neighbor_inds, weightnet, additional_features=None):
        if additional_features is None:
            additional_features = torch.zeros(input_features.shape[0], neighbor_inds.shape[1], neighbor_inds.shape[2], 0,
                                              device=input_features.device)
        return PConvFunction.apply(input_features, neighbor_inds, weightnet, additional_features)


# ------------------------------------------------------------------------------------------------
# edge geometry
# ------------------------------------------------------------------------------------------------
def edge_geometry(xyz_in, nrm_in, nei_inds, xyz_out, nrm_out, use_vi):
    """localized_xyz [B,M,K,3] and (if use_vi) the 12-d VI features [B,M,K,12] straight from the clouds and
    the neighbour table -- fuses index_points x2, the subtraction and VI_coordinate_transform
    (layers.py:337-353 and layer_utils.py:176-231).  xyz / normals carry no gradient in this model."""
    rs, vis = [], []
    for b in range(xyz_in.shape[0]):
        r, vi = pcf_cuda.edge_geometry(xyz_in[b].contiguous(), nrm_in[b].contiguous() if use_vi else None,
                                       xyz_out[b].contiguous(), nrm_out[b].contiguous() if use_vi else None,
                                       nei_inds[b].contiguous(), want_r=True, want_vi=use_vi)
        rs.append(r)
        vis.append(vi)
    return torch.stack(rs), (torch.stack(vis) if use_vi else None)


def VI_coordinate_transform(localized_xyz, gathered_norm, sparse_xyz_norm, K):
    """Reference signature (layer_utils.py:176-231) on already-gathered tensors -- compatibility entry; the
    layers call edge_geometry() instead, which never materialises the gathered inputs."""
    n_i = sparse_xyz_norm.unsqueeze(2)
    r_hat = F.normalize(localized_xyz, dim=3)
    v = F.normalize(n_i - (n_i * r_hat).sum(3, keepdim=True) * r_hat, dim=3)
    w = F.normalize(torch.cross(r_hat, v, dim=3), dim=3)
    dot = lambda a, b: (a * b).sum(3, keepdim=True)
    t3 = dot(r_hat, gathered_norm)
    return torch.cat([dot(gathered_norm, n_i), dot(r_hat, n_i), t3, dot(localized_xyz, n_i), t3,
                      dot(gathered_norm, v), dot(gathered_norm, w),
                      dot(localized_xyz, torch.cross(gathered_norm, n_i.expand_as(gathered_norm), dim=3)),
                      localized_xyz.norm(dim=3, keepdim=True), localized_xyz], dim=3).contiguous()


# ------------------------------------------------------------------------------------------------
# dense Linear on the tensor cores (3xTF32: fp32-accurate; torch's fp32 path is a SIMT sgemm)
# ------------------------------------------------------------------------------------------------
class _LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        y = pcf_cuda.gemm_nt(x2, weight, bias)
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, grad):
        x2, weight = ctx.saved_tensors
        g2 = grad.reshape(-1, weight.shape[0])
        if g2.stride(-1) != 1 or g2.stride(0) < g2.shape[1]:
            g2 = g2.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = pcf_cuda.gemm_nt(g2, weight, None, w_is_kn=True).reshape(*grad.shape[:-1], weight.shape[1])
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            # a leaf of the backward graph: off the critical path when streams.LEAF_ASYNC (joined by streams.join_leaves())
            gw, gb = S.fork_leaf(lambda: pcf_cuda.gemm_tn(g2, x2, want_rowsum=ctx.has_bias), inputs=(g2, x2))
        return gx, gw, gb


def linear(x, weight, bias=None):
    """F.linear on the B200 tensor cores, fp32-accurate (pcfb_gemm_nt / pcfb_gemm_tn).  CUDA float32 only."""
    if not x.is_cuda:
        raise RuntimeError("pcf_b200.linear needs CUDA tensors (no CPU path)")
    weight = weight.contiguous()
    return _LinearFunction.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------
# Linear + BatchNorm blocks (layer_utils.py:241-319)
# ------------------------------------------------------------------------------------------------
class Linear_BN(nn.Module):
    """Linear followed by BatchNorm over the channel (last) dim of a [B,N,K,C] ('2d') or [B,N,C] ('1d')
    tensor; fuse() folds the BN for inference (layer_utils.py:260-270)."""

    def __init__(self, in_dim, out_dim, bn_ver='2d', bn_weight_init=1, bn_momentum=0.1):
        super().__init__()
        self.c = nn.Linear(in_dim, out_dim)
        self.bn_ver = bn_ver
        self.bn = (nn.BatchNorm2d if bn_ver == '2d' else nn.BatchNorm1d)(out_dim, momentum=bn_momentum)
        nn.init.constant_(self.bn.weight, bn_weight_init)

    @torch.no_grad()
    def fuse(self):
        scale = self.bn.weight / (self.bn.running_var + self.bn.eps) ** 0.5
        fused = nn.Linear(self.c.in_features, self.c.out_features).to(self.c.weight.device)
        fused.weight.copy_(self.c.weight * scale[:, None])
        fused.bias.copy_(self.bn.bias + (self.c.bias - self.bn.running_mean) * scale)
        return fused

    def forward(self, x, act=0, residual=None):
        """act (fused_mlp.ACT_*): activation applied after the BatchNorm in the same kernel pass (0 = none, the
        reference's module; callers that follow the block with ReLU / LeakyReLU pass it here).  residual: added to the
        BatchNorm output before the activation, in the same pass (the `+ shortcut` of a block's tail)."""
        from . import fused_mlp
        x = linear(x, self.c.weight, self.c.bias)
        if fused_mlp.bn_supported(x.shape[-1]):
            # BatchNorm over the last dim == the reference's permute(0,3,2,1) -> BN2d -> permute back / BN1d over a
            # [rows, C] view (layer_utils.py:272-277): statistics + apply(+activation) as two passes of pcfb_bn_*
            return fused_mlp.bn_act(x, self.bn, act, pivot=self.c.bias, residual=residual)
        if residual is not None:
            raise RuntimeError("Linear_BN: a fused residual needs the pcfb_bn_* path (C % 4 == 0, C <= 1024)")
        shape = x.shape
        if isinstance(self.bn, nn.SyncBatchNorm):          # after convert_sync_batchnorm (DDP, sync_bn: True)
            y = self.bn(x.reshape(-1, shape[-1])).reshape(shape)
        else:
            y = F.batch_norm(x.reshape(-1, shape[-1]), self.bn.running_mean, self.bn.running_var, self.bn.weight,
                             self.bn.bias, self.bn.training or not self.bn.track_running_stats,
                             self._momentum(), self.bn.eps).reshape(shape)
        return _activate(y, act)

    def _momentum(self):
        if self.bn.training and self.bn.track_running_stats and self.bn.num_batches_tracked is not None:
            self.bn.num_batches_tracked.add_(1)
        return 0.0 if self.bn.momentum is None else self.bn.momentum


def _activate(y, act):
    """torch form of the fused_mlp.ACT_* codes (channel counts the pcfb_bn_* kernels do not take)."""
    if act == 1:
        return F.relu(y)
    if act == 2:
        return F.leaky_relu(y, 0.1)
    if act == 3:
        return torch.sigmoid(y)
    return y


class UnaryBlock(nn.Module):
    """Linear_BN ('1d') + LeakyReLU(0.1) (layer_utils.py:281-319)."""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.bn_momentum, self.use_bn, self.no_relu = bn_momentum, use_bn, no_relu
        self.in_dim, self.out_dim = in_dim, out_dim
        self.mlp = Linear_BN(in_dim, out_dim, bn_momentum=bn_momentum, bn_ver='1d') if use_bn else nn.Linear(in_dim, out_dim)
        self.leaky_relu = nn.Identity() if no_relu else nn.LeakyReLU(0.1)

    def can_fuse_tail(self):
        """True if forward(x, residual=..., act=...) can add a residual and apply an activation inside the BatchNorm pass."""
        from . import fused_mlp
        return (isinstance(self.mlp, Linear_BN) and self.no_relu and fused_mlp.bn_supported(self.out_dim)
                and not fused_mlp.supported([(self.in_dim, self.out_dim)]))

    def forward(self, x, residual=None, act=None):
        """residual / act (only when can_fuse_tail()): returns act(block(x) + residual) in the block's own BatchNorm pass --
        the `leaky_relu(unary2(h) + shortcut)` tail of the PointConvFormer / PointConvStridePE blocks (layers.py:413-415)."""
        from . import fused_mlp
        lin, bn = (self.mlp.c, self.mlp.bn) if isinstance(self.mlp, Linear_BN) else (self.mlp, None)
        if residual is not None:
            return self.mlp(x, act=act, residual=residual)
        if fused_mlp.supported([(lin.in_features, lin.out_features)]):
            act = fused_mlp.ACT_NONE if self.no_relu else fused_mlp.ACT_LEAKY
            return fused_mlp.mlp_chain(x, [(lin, bn, act)], self.training)
        if isinstance(self.mlp, Linear_BN):                # wide block: tcgen05 GEMM + BatchNorm/LeakyReLU in one apply pass
            return self.mlp(x, act=fused_mlp.ACT_NONE if self.no_relu else fused_mlp.ACT_LEAKY)
        return self.leaky_relu(linear(x, self.mlp.weight, self.mlp.bias))

    def __repr__(self):
        return 'UnaryBlock(in_feat: {:d}, out_feat: {:d}, BN: {:s}, ReLU: {:s})'.format(
            self.in_dim, self.out_dim, str(self.use_bn), str(not self.no_relu))
