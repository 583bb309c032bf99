"""Op-level API with the names and argument order of the reference's `pcf_cuda` extension
(/root/reference/cpp_wrappers/cpp_pcf_kernel/pcf_cuda.cpp:9-19; signatures include/pcf.h:38-250), every
function a thin torch-tensor wrapper over the C ABI (include/pcf_b200.h).  Outputs are freshly
allocated torch tensors on the current stream like the reference's (pconv_ops.cu:666,707,891-895);
inputs must be contiguous CUDA tensors or RuntimeError is raised (pcf.h:14-24).

Extra ops that have no `pcf_cuda` counterpart (the reference gets them from pykeops / torch / its
CPU extension) are exposed with plain names: knn_packed, gather, gather_max, edge_geometry,
grid_subsample.
"""
import ctypes
import os

import torch

from . import _lib
from . import streams as S
from ._lib import PconvShape, check, lib, ptr, require, stream_ptr, workspace

F32, I64, I32, U8 = torch.float32, torch.int64, torch.int32, torch.uint8

# 0 = auto (tcgen05 when the shape allows it), 1 = exact-fp32 SIMT, 2 = force tcgen05
FORWARD_VARIANT = 0


def _shape(n_in, n_out, K, C_in, C_add, C_mid, C_out, H):
    return PconvShape(int(n_in), int(n_out), int(K), int(C_in), int(C_add), int(C_mid), int(C_out), int(H))


def forward_variant_supported(n_in, n_out, K, C_in, C_add, C_mid, C_out, H, variant):
    sh = _shape(n_in, n_out, K, C_in, C_add, C_mid, C_out, H)
    return bool(lib().pcfb_pconv_forward_supported(ctypes.byref(sh), int(variant)))


def _batched(t):
    """All reference ops carry a leading batch dim (always 1 in the packed representation)."""
    return t.shape[0]


# ------------------------------------------------------------------------------------------------
# fused PConv(+guidance)(+Linear) core
# ------------------------------------------------------------------------------------------------
_TRACE_SHAPES = os.environ.get("PCFB_TRACE_SHAPES")      # debug aid: append every fused-forward shape to this file


def _pconv_fwd(inp, nei, weights, additional, guidance, lin_w, lin_b, want_p, variant=None):
    require(inp, F32, "input"); require(nei, I64, "neighbor_inds"); require(weights, F32, "weights")
    B, N_in, C_in = inp.shape
    _, M, K = nei.shape
    C_mid = weights.shape[3]
    C_add = 0 if additional is None else additional.shape[3]
    H = 0 if guidance is None else guidance.shape[3]
    if C_add:
        require(additional, F32, "additional_features")
    if H:
        require(guidance, F32, "guidance")
    C_out = 0
    if lin_w is not None:
        require(lin_w, F32, "linear_weights")
        C_out = lin_w.shape[0]
        if lin_w.shape[1] != (C_in + C_add) * C_mid:
            raise RuntimeError("linear_weights must be [C_out, C_mid*(C_in+C_add)] = [*, %d], got %s"
                               % ((C_in + C_add) * C_mid, tuple(lin_w.shape)))
        if lin_b is not None:
            require(lin_b, F32, "linear_bias")
    if weights.shape[:3] != (B, M, K) or (C_add and additional.shape[:3] != (B, M, K)):
        raise RuntimeError("weights / additional_features must be [B, N_out, K, *]")
    dev = inp.device
    KK = (C_in + C_add) * C_mid
    out_y = torch.empty(B, M, C_out, device=dev, dtype=F32) if lin_w is not None else None
    out_p = torch.empty(B, M, KK, device=dev, dtype=F32) if (want_p or lin_w is None) else None
    v = FORWARD_VARIANT if variant is None else variant
    sh = _shape(N_in, M, K, C_in, C_add, C_mid, C_out, H)
    if _TRACE_SHAPES:
        with open(_TRACE_SHAPES, "a") as f:
            f.write("pconv_forward n_in=%d n_out=%d K=%d C_in=%d C_add=%d C_mid=%d C_out=%d H=%d want_p=%d\n"
                    % (N_in, M, K, C_in, C_add, C_mid, C_out, H, int(out_p is not None)))
    ws_bytes = lib().pcfb_pconv_forward_workspace(ctypes.byref(sh), v)
    ws = workspace(ws_bytes, dev) if ws_bytes else None
    for b in range(B):
        check(lib().pcfb_pconv_forward(
            ctypes.byref(sh), ptr(inp[b]), ptr(nei[b]), ptr(weights[b]),
            ptr(additional[b]) if C_add else 0, ptr(guidance[b]) if H else 0,
            ptr(lin_w), ptr(lin_b), ptr(out_y[b]) if out_y is not None else 0,
            ptr(out_p[b]) if out_p is not None else 0, ptr(ws), ws_bytes, v, stream_ptr()), "pconv_forward")
    _lib.account(B * (M * (8.0 * K + 4.0 * K * (C_mid + C_add + H) + 4.0 * C_out + (4.0 * KK if out_p is not None else 0.0)) + 4.0 * N_in * C_in),
                 B * M * (2.0 * K * (C_in + C_add) * C_mid + 2.0 * KK * C_out))
    return out_y, out_p


def _pconv_bwd(grad_y, grad_p, inp, inv, nei, weights, additional, guidance, lin_w, pconv_out,
               need=(True, True, True, True, True, True)):
    """-> (g_input, g_weights, g_additional, g_guidance, g_lin_w, g_lin_b); entries not needed are None."""
    B, N_in, C_in = inp.shape
    _, M, K = nei.shape
    C_mid = weights.shape[3]
    C_add = 0 if additional is None else additional.shape[3]
    H = 0 if guidance is None else guidance.shape[3]
    C_out = 0 if lin_w is None else lin_w.shape[0]
    dev = inp.device
    n_in, n_w, n_add, n_gd, n_lw, n_lb = need
    n_add = n_add and C_add > 0
    n_gd = n_gd and H > 0
    n_lw = n_lw and lin_w is not None
    n_lb = n_lb and lin_w is not None
    if grad_y is not None:
        require(grad_y, F32, "grad_output")
    if grad_p is not None:
        require(grad_p, F32, "grad_output")
    g_in = torch.empty_like(inp) if n_in else None
    g_w = torch.empty_like(weights) if n_w else None
    g_add = torch.empty_like(additional) if n_add else None
    g_gd = torch.empty_like(guidance) if n_gd else None
    sh = _shape(N_in, M, K, C_in, C_add, C_mid, C_out, H)
    ws_bytes = lib().pcfb_pconv_backward_workspace(ctypes.byref(sh), 0)
    ws = workspace(ws_bytes, dev)
    g_lw_acc, g_lb_acc = None, None
    # dW = dY^T P / db are leaves of the backward graph: with streams.LEAF_ASYNC they run as a separate product on the leaf
    # stream (joined by streams.join_leaves()) instead of inside the composed backward call
    leaf = (S.ENABLED and S.LEAF_ASYNC and (n_lw or n_lb) and grad_y is not None and pconv_out is not None and 0 < C_out <= 256)
    for b in range(B):
        g_lw = torch.empty_like(lin_w) if (n_lw and not leaf) else None
        g_lb = torch.empty(C_out, device=dev, dtype=F32) if (n_lb and not leaf) else None
        inv_n, inv_k, inv_idx = (None, None, None) if inv is None else (inv[0][b], inv[1][b], inv[2][b])
        check(lib().pcfb_pconv_backward(
            ctypes.byref(sh), ptr(grad_y[b]) if grad_y is not None else 0, ptr(grad_p[b]) if grad_p is not None else 0,
            ptr(inp[b]), ptr(nei[b]), ptr(inv_n), ptr(inv_k), ptr(inv_idx), ptr(weights[b]),
            ptr(additional[b]) if C_add else 0, ptr(guidance[b]) if H else 0, ptr(lin_w),
            ptr(pconv_out[b]) if pconv_out is not None else 0,
            ptr(g_in[b]) if n_in else 0, ptr(g_w[b]) if n_w else 0, ptr(g_add[b]) if n_add else 0,
            ptr(g_gd[b]) if n_gd else 0, ptr(g_lw), ptr(g_lb), ptr(ws), ws_bytes, 0, stream_ptr()), "pconv_backward")
        KK = (C_in + C_add) * C_mid
        _lib.account(M * (8.0 * K + 5.0 * K + 4.0 * C_out + 4.0 * KK + 8.0 * K * (C_mid + C_add + H)) + 8.0 * N_in * C_in + 4.0 * C_out * KK,
                     2.0 * M * (2.0 * K * (C_in + C_add) * C_mid + 2.0 * KK * C_out))
        if leaf:
            g_lw, g_lb = S.fork_leaf(lambda: gemm_tn(grad_y[b], pconv_out[b], want_rowsum=True), inputs=(grad_y, pconv_out))
        if n_lw:
            g_lw_acc = g_lw if g_lw_acc is None else g_lw_acc + g_lw
        if n_lb:
            g_lb_acc = g_lb if g_lb_acc is None else g_lb_acc + g_lb
    return g_in, g_w, g_add, g_gd, g_lw_acc, g_lb_acc


def _empty_add(additional):
    return None if additional is None or additional.shape[-1] == 0 else additional


def _inverse_for(inp, nei, inverse_neighbors, inverse_k, inverse_idx):
    """The reference's opt backward needs the inverse map handed in (layer_utils.py:60-68); when it is
    missing (eval drivers never build it) it is built here on the fly."""
    if inverse_neighbors is None:
        return compute_knn_inverse(nei, inp.shape[1])
    return (require(inverse_neighbors, I32, "inverse_neighbor"), require(inverse_k, U8, "inverse_neighbor_k"),
            require(inverse_idx, I32, "inverse_neighbor_idx"))


# ------------------------------------------------------------------------------------------------
# the nine pcf_cuda entry points
# ------------------------------------------------------------------------------------------------
def pconv_linear_cutlass_forward(input, neighbor_inds, weights, additional_features, linear_weights, linear_bias):
    """pcf.h:243-250 -> (output [B,N,C_out], pconv_output [B,N,C_mid*(C_in+C_add)])."""
    y, p = _pconv_fwd(input, neighbor_inds, weights, _empty_add(additional_features), None,
                      linear_weights, linear_bias, want_p=True)
    return y, p


def pconv_linear_forward(input, neighbor_inds, weights, additional_features, linear_weights, linear_bias):
    """pcf.h:131-138 (the reference's SIMT fused kernel): same contract; served by the exact-fp32 variant."""
    y, p = _pconv_fwd(input, neighbor_inds, weights, _empty_add(additional_features), None,
                      linear_weights, linear_bias, want_p=True, variant=1)
    return y, p


def pconv_linear_opt_backward(grad_output, input, inverse_neighbor, inverse_neighbor_k, inverse_neighbor_idx,
                              neighbor_inds, weights, additional_features, linear_weights, pconv_output):
    """pcf.h:213-224 -> [grad_input, grad_weights, grad_additional, grad_linear_weights, grad_linear_bias]."""
    inv = _inverse_for(input, neighbor_inds, inverse_neighbor, inverse_neighbor_k, inverse_neighbor_idx)
    add = _empty_add(additional_features)
    g_in, g_w, g_add, _, g_lw, g_lb = _pconv_bwd(grad_output, None, input, inv, neighbor_inds, weights, add, None,
                                                 linear_weights, pconv_output)
    if g_add is None:
        g_add = torch.zeros_like(additional_features) if additional_features is not None else None
    return [g_in, g_w, g_add, g_lw, g_lb]


def pconv_linear_backward(grad_output, input, neighbor_inds, weights, additional_features, linear_weights, pconv_output):
    """pcf.h:162-170: as above without a caller-supplied inverse map."""
    return pconv_linear_opt_backward(grad_output, input, None, None, None, neighbor_inds, weights,
                                     additional_features, linear_weights, pconv_output)


def pconv_forward(input, neighbor_inds, weights, additional_features):
    """pcf.h:81-86 -> [B,N,C_mid*(C_in+C_add)]."""
    _, p = _pconv_fwd(input, neighbor_inds, weights, _empty_add(additional_features), None, None, None, want_p=True)
    return p


def pconv_backward(grad_output, input, neighbor_inds, weights, additional_features):
    """pcf.h:106-112 -> [grad_input, grad_weights, grad_additional]."""
    inv = compute_knn_inverse(neighbor_inds, input.shape[1])
    add = _empty_add(additional_features)
    g_in, g_w, g_add, _, _, _ = _pconv_bwd(None, grad_output, input, inv, neighbor_inds, weights, add, None, None, None)
    if g_add is None:
        g_add = torch.zeros_like(additional_features) if additional_features is not None else None
    return [g_in, g_w, g_add]


def pcf_forward(input, neighbor_inds, guidance, weights):
    """pcf.h:38-43 -> [B,N,C_mid*C_in]; head of channel c is c % H."""
    _, p = _pconv_fwd(input, neighbor_inds, weights, None, require(guidance, F32, "guidance"), None, None, want_p=True)
    return p


def pcf_backward(grad_output, input, neighbor_inds, guidance, weights):
    """pcf.h:60-66 -> [grad_input, grad_guidance, grad_weights]."""
    inv = compute_knn_inverse(neighbor_inds, input.shape[1])
    g_in, g_w, _, g_gd, _, _ = _pconv_bwd(None, grad_output, input, inv, neighbor_inds, weights, None, guidance, None, None)
    return [g_in, g_gd, g_w]


def compute_knn_inverse(neighbor_inds, total_points):
    """pcf.h:183-186 -> (inv_neighbors int32 [B,N*K], inv_k uint8 [B,N*K], inv_idx int32 [B,total+1])."""
    require(neighbor_inds, I64, "neighbor_inds")
    B, N, K = neighbor_inds.shape
    total = int(total_points)
    dev = neighbor_inds.device
    inv_n = torch.empty(B, N * K, device=dev, dtype=I32)
    inv_k = torch.empty(B, N * K, device=dev, dtype=U8)
    inv_idx = torch.empty(B, total + 1, device=dev, dtype=I32)
    ws_bytes = lib().pcfb_knn_inverse_workspace(N, K, total)
    ws = workspace(ws_bytes, dev)
    for b in range(B):
        check(lib().pcfb_knn_inverse(ptr(neighbor_inds[b]), N, K, total, ptr(inv_n[b]), ptr(inv_k[b]), ptr(inv_idx[b]),
                                     ptr(ws), ws_bytes, stream_ptr()), "compute_knn_inverse")
        _lib.account(13.0 * N * K + 4.0 * (total + 1))
    return inv_n, inv_k, inv_idx


# ------------------------------------------------------------------------------------------------
# fused-layer entry used by layers.py: contraction (+guidance) + Linear in one call
# ------------------------------------------------------------------------------------------------
def pconv_fused_forward(input, neighbor_inds, weights, additional_features, guidance, linear_weights, linear_bias,
                        want_p=True, variant=None):
    return _pconv_fwd(input, neighbor_inds, weights, _empty_add(additional_features), guidance,
                      linear_weights, linear_bias, want_p, variant)


def pconv_fused_backward(grad_y, grad_p, input, inv, neighbor_inds, weights, additional_features, guidance,
                         linear_weights, pconv_output, need):
    return _pconv_bwd(grad_y, grad_p, input, inv, neighbor_inds, weights, _empty_add(additional_features), guidance,
                      linear_weights, pconv_output, need)


# ------------------------------------------------------------------------------------------------
# ops without a pcf_cuda counterpart
# ------------------------------------------------------------------------------------------------
def knn_packed(ref_xyz, ref_counts, qry_xyz, qry_counts, K, out=None):
    """Exact kNN inside each scene of a packed cloud.  ref_xyz [N_ref,3], qry_xyz [N_qry,3] fp32 CUDA;
    ref_counts / qry_counts: per-scene point counts (python ints).  -> int64 [N_qry,K], values index the
    packed reference cloud (scene offset already added, as prepare() does)."""
    require(ref_xyz, F32, "ref_xyz"); require(qry_xyz, F32, "qry_xyz")
    dev = ref_xyz.device
    n_seg = len(ref_counts)
    assert len(qry_counts) == n_seg
    if sum(map(int, ref_counts)) != ref_xyz.shape[0] or sum(map(int, qry_counts)) != qry_xyz.shape[0]:
        raise RuntimeError("scene counts do not add up to the packed cloud sizes")
    ro_d, qo_d = _offsets(ref_counts, dev), _offsets(qry_counts, dev)
    if out is None:
        out = torch.empty(qry_xyz.shape[0], K, device=dev, dtype=I64)
    check(lib().pcfb_knn_packed(ptr(ref_xyz), ptr(ro_d), ptr(qry_xyz), ptr(qo_d), n_seg, ref_xyz.shape[0],
                                qry_xyz.shape[0], int(K), ptr(out), stream_ptr()), "knn_packed")
    return out


# ------------------------------------------------------------------------------------------------
# dense fp32-accurate products on the tensor cores (3xTF32)
# ------------------------------------------------------------------------------------------------
def gemm_nt(x, w, bias=None, w_is_kn=False, act=0):
    """x [M,K] (row-major, last dim contiguous) times w^T (w [N,K]) or times w (w [K,N], w_is_kn) -> [M,N]."""
    require(w, F32, "w")
    if x.dtype != F32 or not x.is_cuda or x.stride(-1) != 1:
        raise RuntimeError("gemm_nt: x must be a CUDA float32 matrix with contiguous rows")
    M, K = x.shape
    N = w.shape[1] if w_is_kn else w.shape[0]
    if (w.shape[0] if w_is_kn else w.shape[1]) != K:
        raise RuntimeError("gemm_nt: inner dimensions do not match")
    out = torch.empty(M, N, device=x.device, dtype=F32)
    ws_bytes = lib().pcfb_gemm_nt_workspace(N, K)
    prep = S.prep_stream(x.device)
    if prep is None:
        ws = workspace(ws_bytes, x.device)
        check(lib().pcfb_gemm_nt(ptr(x), x.stride(0), ptr(w), w.stride(0), 1 if w_is_kn else 0, ptr(bias), ptr(out), N, M, N, K, int(act),
                                 ptr(ws), ws_bytes, stream_ptr()), "gemm_nt")
    else:
        # the weight preparation only depends on the weights: it runs on the prep stream (ordered after the previous optimizer
        # step, streams.prep_stream), off the critical path; the product waits for its event
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(prep):
            ws = workspace(ws_bytes, x.device)
            check(lib().pcfb_gemm_nt_prepare(ptr(w), w.stride(0), 1 if w_is_kn else 0, M, N, K, ptr(ws), ws_bytes, stream_ptr()), "gemm_nt_prepare")
        cur.wait_stream(prep)
        ws.record_stream(cur)
        check(lib().pcfb_gemm_nt_prepared(ptr(x), x.stride(0), ptr(bias), ptr(out), N, M, N, K, int(act), ptr(ws), ws_bytes, stream_ptr()),
              "gemm_nt_prepared")
    _lib.account(4.0 * (M * K + N * K + M * N), 2.0 * M * N * K)
    return out


def gemm_tn(a, b, want_rowsum=False):
    """a [M,N1], b [M,N2] -> (a^T b [N1,N2], sum_m a[m,:] [N1] or None); deterministic split over M."""
    if a.dtype != F32 or b.dtype != F32 or not a.is_cuda or a.stride(-1) != 1 or b.stride(-1) != 1:
        raise RuntimeError("gemm_tn: operands must be CUDA float32 matrices with contiguous rows")
    M, N1 = a.shape
    N2 = b.shape[1]
    if N1 > 256:
        parts = [gemm_tn(a[:, i:i + 256], b, want_rowsum) for i in range(0, N1, 256)]
        return torch.cat([p[0] for p in parts]), (torch.cat([p[1] for p in parts]) if want_rowsum else None)
    out = torch.empty(N1, N2, device=a.device, dtype=F32)
    rs = torch.empty(N1, device=a.device, dtype=F32) if want_rowsum else None
    ws_bytes = lib().pcfb_gemm_tn_workspace(M, N1, N2, 1 if want_rowsum else 0)
    ws = workspace(ws_bytes, a.device)
    check(lib().pcfb_gemm_tn(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), N2, ptr(rs), M, N1, N2, ptr(ws), ws_bytes,
                             stream_ptr()), "gemm_tn")
    _lib.account(4.0 * (M * N1 + M * N2 + N1 * N2), 2.0 * M * N1 * N2)
    return out, rs


_OFFSETS = {}


def _offsets(counts, dev):
    """Device int32 prefix array of the per-scene counts, cached per (counts, device): the tiny H2D copy happens
    once per distinct packing, never inside a captured CUDA graph."""
    key = (tuple(int(c) for c in counts), str(dev))
    t = _OFFSETS.get(key)
    if t is None:
        if len(_OFFSETS) > 4096:
            _OFFSETS.clear()
        t = torch.tensor([0] + list(key[0]), dtype=torch.int64).cumsum(0).to(torch.int32).to(dev)
        _OFFSETS[key] = t
    return t


class KnnGrid:
    """Uniform grid over a packed reference cloud (pcfb_knn_grid_build); query() returns exactly what
    knn_packed() returns.  cell_hint: cell edge (about 1.75x the point spacing works best; <= 0 = automatic)."""

    def __init__(self, ref_xyz, ref_counts, cell_hint=0.0):
        require(ref_xyz, F32, "ref_xyz")
        self.ref, self.counts, self.n_seg = ref_xyz, list(map(int, ref_counts)), len(ref_counts)
        if sum(self.counts) != ref_xyz.shape[0]:
            raise RuntimeError("scene counts do not add up to the packed cloud size")
        dev = ref_xyz.device
        self.ref_off = _offsets(self.counts, dev)
        self.ws_bytes = lib().pcfb_knn_grid_workspace(self.n_seg, ref_xyz.shape[0])
        self.ws = workspace(self.ws_bytes, dev)
        check(lib().pcfb_knn_grid_build(ptr(ref_xyz), ptr(self.ref_off), self.n_seg, ref_xyz.shape[0], float(cell_hint),
                                        ptr(self.ws), self.ws_bytes, stream_ptr()), "knn_grid_build")

    def order_ptr(self):
        """Device pointer to the references' cell-sorted permutation (lives in this grid's workspace)."""
        return lib().pcfb_knn_grid_order(self.n_seg, self.ref.shape[0], ptr(self.ws)) or 0

    def query(self, qry_xyz, qry_counts, K, order=None):
        """order: the KnnGrid built over qry_xyz itself (its cell order makes the warps spatially coherent), or None."""
        require(qry_xyz, F32, "qry_xyz")
        if order is not None and (order.ref.data_ptr() != qry_xyz.data_ptr() or order.ref.shape[0] != qry_xyz.shape[0]):
            raise RuntimeError("order: not the grid of this query cloud")
        if len(qry_counts) != self.n_seg or sum(map(int, qry_counts)) != qry_xyz.shape[0]:
            raise RuntimeError("query scene counts do not match")
        qo = _offsets(qry_counts, qry_xyz.device)
        out = torch.empty(qry_xyz.shape[0], K, device=qry_xyz.device, dtype=I64)
        check(lib().pcfb_knn_grid_query(ptr(self.ref), self.n_seg, self.ref.shape[0], ptr(qry_xyz), ptr(qo), qry_xyz.shape[0],
                                        int(K), order.order_ptr() if order is not None else 0, ptr(out), ptr(self.ws),
                                        self.ws_bytes, stream_ptr()), "knn_grid_query")
        _lib.account(12.0 * (self.ref.shape[0] + qry_xyz.shape[0]) + 8.0 * K * qry_xyz.shape[0])
        return out


def gather(feats, nei):
    """index_points for B folded: feats [N,C], nei [M,K] -> [M,K,C]."""
    require(feats, F32, "feats"); require(nei, I64, "nei")
    M, K = nei.shape
    out = torch.empty(M, K, feats.shape[1], device=feats.device, dtype=F32)
    check(lib().pcfb_gather(ptr(feats), ptr(nei), feats.shape[0], M, K, feats.shape[1], ptr(out), stream_ptr()), "gather")
    _lib.account(8.0 * M * K + 4.0 * feats.numel() + 4.0 * out.numel())
    return out


def gather_backward(grad_out, inv, n_in):
    require(grad_out, F32, "grad_out")
    M, K, C = grad_out.shape
    g = torch.empty(n_in, C, device=grad_out.device, dtype=F32)
    check(lib().pcfb_gather_backward(ptr(grad_out), ptr(inv[0]), ptr(inv[1]), ptr(inv[2]), n_in, M, K, C, ptr(g),
                                     stream_ptr()), "gather_backward")
    _lib.account(4.0 * grad_out.numel() + 5.0 * M * K + 4.0 * (n_in + 1) + 4.0 * g.numel())
    return g


def gather_max(feats, nei):
    require(feats, F32, "feats"); require(nei, I64, "nei")
    M, K = nei.shape
    C = feats.shape[1]
    out = torch.empty(M, C, device=feats.device, dtype=F32)
    arg = torch.empty(M, C, device=feats.device, dtype=U8)
    check(lib().pcfb_gather_max(ptr(feats), ptr(nei), feats.shape[0], M, K, C, ptr(out), ptr(arg), stream_ptr()), "gather_max")
    _lib.account(8.0 * M * K + 4.0 * feats.numel() + 5.0 * out.numel())
    return out, arg


def gather_max_backward(grad_out, arg, inv, n_in, K):
    require(grad_out, F32, "grad_out")
    M, C = grad_out.shape
    g = torch.empty(n_in, C, device=grad_out.device, dtype=F32)
    check(lib().pcfb_gather_max_backward(ptr(grad_out), ptr(arg), ptr(inv[0]), ptr(inv[1]), ptr(inv[2]), n_in, M, K, C,
                                         ptr(g), stream_ptr()), "gather_max_backward")
    _lib.account(5.0 * grad_out.numel() + 5.0 * M * K + 4.0 * (n_in + 1) + 4.0 * g.numel())
    return g


def guidance_input(gx, pe, nei, use_max):
    """gx [N,G], pe [M,K,P], nei [M,K] -> (cat(gx[nei], pe) - key [M,K,G+P], arg [M,G+P] uint8 or None)
    with key = column 0 (use_max False) or max over the K neighbours (layers.py:372-382)."""
    require(gx, F32, "guidance_x"); require(pe, F32, "feat_pe"); require(nei, I64, "nei")
    M, K = nei.shape
    G, P = gx.shape[1], pe.shape[2]
    out = torch.empty(M, K, G + P, device=gx.device, dtype=F32)
    arg = torch.empty(M, G + P, device=gx.device, dtype=U8) if use_max else None
    check(lib().pcfb_guidance_input(ptr(gx), ptr(pe), ptr(nei), gx.shape[0], M, K, G, P, 1 if use_max else 0, ptr(out), ptr(arg),
                                    stream_ptr()), "guidance_input")
    _lib.account(8.0 * M * K + 4.0 * gx.numel() + 4.0 * pe.numel() + 4.0 * out.numel())
    return out, arg


def guidance_input_backward(ds, arg, G, P, use_max, want_gq=True, want_pe=True):
    """ds [M,K,G+P] -> (d_gq [M,K,G] per-edge gradient of the gathered half, d_pe [M,K,P])."""
    require(ds, F32, "grad")
    M, K, _ = ds.shape
    d_gq = torch.empty(M, K, G, device=ds.device, dtype=F32) if want_gq else None
    d_pe = torch.empty(M, K, P, device=ds.device, dtype=F32) if want_pe else None
    check(lib().pcfb_guidance_input_backward(ptr(ds), ptr(arg), M, K, G, P, 1 if use_max else 0, ptr(d_gq), ptr(d_pe), stream_ptr()),
          "guidance_input_backward")
    _lib.account(8.0 * ds.numel())
    return d_gq, d_pe


def edge_geometry(xyz_in, nrm_in, xyz_out, nrm_out, nei, want_r=True, want_vi=True):
    """-> (localized_xyz [M,K,3] or None, vi_features [M,K,12] or None)."""
    require(xyz_in, F32, "xyz_in"); require(xyz_out, F32, "xyz_out"); require(nei, I64, "nei")
    M, K = nei.shape
    dev = xyz_in.device
    r = torch.empty(M, K, 3, device=dev, dtype=F32) if want_r else None
    vi = None
    if want_vi:
        require(nrm_in, F32, "nrm_in"); require(nrm_out, F32, "nrm_out")
        vi = torch.empty(M, K, 12, device=dev, dtype=F32)
    check(lib().pcfb_edge_geometry(ptr(xyz_in), ptr(nrm_in), ptr(xyz_out), ptr(nrm_out), ptr(nei), xyz_in.shape[0], M, K,
                                   ptr(r), ptr(vi), stream_ptr()), "edge_geometry")
    _lib.account(8.0 * M * K + 24.0 * (xyz_in.shape[0] + M) + (12.0 * M * K if want_r else 0.0) + (48.0 * M * K if want_vi else 0.0))
    return r, vi


def grid_subsample(xyz, feats, counts, dl):
    """Packed grid subsampling: xyz [N,3], feats [N,F] or None, counts = per-scene sizes.
    -> (sub_xyz [M,3], sub_feats [M,F] or None, sub_counts list)."""
    require(xyz, F32, "xyz")
    dev = xyz.device
    n_seg, n_pts = len(counts), xyz.shape[0]
    F = 0 if feats is None else feats.shape[1]
    if F:
        require(feats, F32, "feats")
    off = torch.tensor([0] + list(counts), dtype=torch.int64).cumsum(0).to(torch.int32).to(dev)
    origin = torch.empty(n_seg, 3, device=dev, dtype=F32)
    dims = torch.empty(n_seg, 3, device=dev, dtype=I32)
    ws0 = workspace(n_seg * 24 + 64, dev)
    check(lib().pcfb_gridsub_bounds(ptr(xyz), ptr(off), n_seg, n_pts, float(dl), ptr(origin), ptr(dims), ptr(ws0),
                                    ws0.numel(), stream_ptr()), "gridsub_bounds")
    dims_h = dims.cpu().to(torch.int64)                                  # host sync #1 (grid extent)
    cells = (dims_h[:, 0] * dims_h[:, 1] * dims_h[:, 2])
    cell_off_h = torch.cat([torch.zeros(1, dtype=torch.int64), cells.cumsum(0)])
    total_cells = int(cell_off_h[-1])
    if total_cells >= (1 << 30):
        raise RuntimeError("grid_subsample: %d cells is too many for dense binning" % total_cells)
    cell_off = cell_off_h.to(torch.int32).to(dev)
    ws_bytes = lib().pcfb_gridsub_workspace(n_seg, n_pts, total_cells)
    ws = workspace(ws_bytes, dev)
    out_counts = torch.empty(n_seg, device=dev, dtype=I32)
    check(lib().pcfb_gridsub_count(ptr(xyz), ptr(off), n_seg, n_pts, float(dl), ptr(origin), ptr(dims), ptr(cell_off),
                                   total_cells, ptr(out_counts), ptr(ws), ws_bytes, stream_ptr()), "gridsub_count")
    sub_counts = out_counts.cpu().tolist()                               # host sync #2 (output size)
    m = int(sum(sub_counts))
    out_xyz = torch.empty(m, 3, device=dev, dtype=F32)
    out_f = torch.empty(m, F, device=dev, dtype=F32) if F else None
    check(lib().pcfb_gridsub_emit(ptr(xyz), ptr(feats), n_seg, n_pts, F, total_cells, ptr(out_xyz), ptr(out_f), ptr(ws),
                                  ws_bytes, stream_ptr()), "gridsub_emit")
    return out_xyz, out_f, sub_counts


# ------------------------------------------------------------------------------------------------
# device-resident pyramid construction (voxelisation + grid subsampling without a host read per level)
# ------------------------------------------------------------------------------------------------
def _cells_upper_bound(xyz, counts, cell, margin=2):
    """Host upper bound of the dense voxel table for cell edge `cell`, from the per-scene bounding boxes (ONE device->host
    read of 6 floats per scene).  Barycentres of coarser levels stay inside the box of level 0, so the same boxes bound
    every level of the pyramid."""
    dev = xyz.device
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    boxes = []
    for s in range(len(counts)):
        seg = xyz[off[s]:off[s + 1]]
        boxes.append(torch.cat([seg.min(0).values, seg.max(0).values]) if seg.shape[0] else torch.zeros(6, device=dev))
    return torch.stack(boxes).cpu().double()                            # [n_seg, 6] (min xyz, max xyz)


def cells_for(boxes, cell, margin=2):
    ext = (boxes[:, 3:] - boxes[:, :3]).clamp(min=0) / float(cell)
    dims = ext.floor().to(torch.int64) + 1 + margin
    return int(dims.prod(1).sum().item())


def voxelize(xyz, counts, voxel, boxes=None):
    """voxelize(coord, voxel, hash_type='ravel', mode='deterministic') (util/voxelize.py:44-70) on packed scenes.
    -> (idx int32 [<= N] upper-bound sized, seg_off int32 device [n_seg+1], status device int32, boxes)."""
    require(xyz, F32, "xyz")
    dev = xyz.device
    n_seg, n_pts = len(counts), xyz.shape[0]
    if boxes is None:
        boxes = _cells_upper_bound(xyz, counts, voxel)
    cells_max = max(cells_for(boxes, voxel), 1)
    seg_off = _offsets(counts, dev)
    out_idx = torch.empty(max(n_pts, 1), device=dev, dtype=I32)
    out_off = torch.empty(n_seg + 1, device=dev, dtype=I32)
    status = torch.zeros(1, device=dev, dtype=I32)
    ws_bytes = lib().pcfb_voxelize_workspace(n_seg, n_pts, cells_max)
    ws = workspace(ws_bytes, dev)
    check(lib().pcfb_voxelize(ptr(xyz), ptr(seg_off), n_seg, n_pts, float(voxel), cells_max, ptr(out_idx), ptr(out_off),
                              ptr(status), ptr(ws), ws_bytes, stream_ptr()), "voxelize")
    return out_idx, out_off, status, boxes


def pyramid_level(xyz, feats, seg_off, n_seg, dl, boxes, status):
    """One grid-subsampling level whose input size lives on the device (seg_off[-1]).  Buffers are upper-bound sized.
    -> (sub_xyz, sub_feats, sub_seg_off)."""
    dev = xyz.device
    n_max = xyz.shape[0]
    F = 0 if feats is None else feats.shape[1]
    cells_max = max(cells_for(boxes, dl), 1)
    out_xyz = torch.empty(n_max, 3, device=dev, dtype=F32)
    out_f = torch.empty(n_max, F, device=dev, dtype=F32) if F else None
    out_off = torch.empty(n_seg + 1, device=dev, dtype=I32)
    ws_bytes = lib().pcfb_pyramid_level_workspace(n_seg, n_max, cells_max)
    ws = workspace(ws_bytes, dev)
    check(lib().pcfb_pyramid_level(ptr(xyz), ptr(feats), F, ptr(seg_off), n_seg, n_max, float(dl), cells_max, ptr(out_xyz),
                                   ptr(out_f), ptr(out_off), ptr(status), ptr(ws), ws_bytes, stream_ptr()), "pyramid_level")
    _lib.account(n_max * (12.0 * 3 + 4.0 * F * 2 + 8.0) + 12.0 * cells_max)
    return out_xyz, out_f, out_off
