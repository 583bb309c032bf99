"""Fused small-MLP chains with BatchNorm (WeightNet, pe_convs, mlp_conv, guidance MLP, UnaryBlock) on top of
pcfb_mlp_* (csrc/mlp.cu): one streaming kernel pass per layer in the forward, two in the backward, BatchNorm
statistics accumulated on the fly, every reduction deterministic.  Mirrors what the reference computes with
Linear_BN + activation sequences (/root/reference/layers.py:127-191, 38-68; layer_utils.py:241-319)."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib, ptr, stream_ptr, workspace

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3
F32 = torch.float32
SYNC_CHANNELS = False       # True once the SyncBatchNorm exchange has one channel per stream (streams.side_index)


def supported(dims):
    """dims: [(cin, cout), ...]"""
    return all(lib().pcfb_mlp_supported(int(a), int(b)) for a, b in dims)


def _sync_group(bn):
    if isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return True
    return False


# ---- the SyncBatchNorm statistics exchange --------------------------------------------------------------------------
# ~540 all-reduces of <= 2 x 384 floats per training step: latency only.  When the ranks share an NVLink / NVSwitch domain
# the exchange runs as one single-CTA kernel over symmetric peer memory (csrc/peer_reduce.cu, ~1/3 of the latency of an
# NCCL call inside the captured step); otherwise -- other backend, no peer access, PCFB_PEER_REDUCE=0 -- it is
# dist.all_reduce.  torch.distributed._symmetric_memory only provides the allocation + address exchange (plumbing).
_PEER = {"state": None}       # None = not tried, False = unavailable, dict = ready


def _peer_setup(device):
    import os
    st = {"ok": False}
    try:
        if os.environ.get("PCFB_PEER_REDUCE", "1") == "0" or dist.get_backend() != "nccl":
            raise RuntimeError("disabled")
        import torch.distributed._symmetric_memory as symm_mem
        world, rank = dist.get_world_size(), dist.get_rank()
        nbytes = int(lib().pcfb_peer_buffer_bytes(world))
        if nbytes == 0:
            raise RuntimeError("world size not supported")
        buf = symm_mem.empty(nbytes // 4, dtype=F32, device=device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        bases = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64).to(device)
        torch.cuda.synchronize(device)
        st = {"ok": True, "buf": buf, "hdl": hdl, "bases": bases, "rank": rank, "world": world,
              "maxn": int(lib().pcfb_peer_max_floats())}
    except Exception as e:                                    # every rank must agree: vote below
        st = {"ok": False, "why": "%s: %s" % (type(e).__name__, e)}
    vote = torch.tensor([1.0 if st["ok"] else 0.0], device=device)
    dist.all_reduce(vote, op=dist.ReduceOp.MIN)               # also orders the zero-fill of every buffer before its first use
    if float(vote.item()) < 1.0:
        return False
    return st


def sync_all_reduce(t):
    """In-place sum of a small contiguous float32 CUDA tensor over all ranks (the BatchNorm sums)."""
    if not t.is_cuda:                                         # gloo / CPU tensors (host-logic tests)
        dist.all_reduce(t)
        return t
    st = _PEER["state"]
    if st is None:
        if torch.cuda.is_current_stream_capturing():          # set-up needs host syncs: do it in the eager warm-up steps
            st = False
        else:
            st = _PEER["state"] = _peer_setup(t.device)
    if st and t.dtype == F32 and t.is_contiguous() and t.numel() <= st["maxn"]:
        check(lib().pcfb_peer_allreduce(ptr(t), ptr(t), t.numel(), ptr(st["bases"]), st["rank"], st["world"], stream_ptr()),
              "peer_allreduce")
        return t
    dist.all_reduce(t)
    return t


# ---- global row counts for SyncBatchNorm -------------------------------------------------------------------------
# A BatchNorm over [1, N_l, (K,) C] needs the row count summed over all ranks.  Every tensor of the model has N_l rows
# of some pyramid level l, so ONE all-reduce of the per-level point counts at the start of a forward
# (register_levels, called by PointConvFormer_Segmentation.forward) serves all ~190 BatchNorms of the step; a chain
# whose leading shape is not a registered level falls back to its own all-reduce.
_LEVEL_ROWS = {}          # local N_l -> (1-element double tensor with the global N_l)
_DERIVED_ROWS = {}        # (local N_l, factor) -> global N_l * factor
_LOCAL_COUNTS = {}        # (counts, device) -> device tensor of the local counts


def register_levels(point_counts, device):
    """point_counts: local per-level point counts of this rank's packed batch.  No-op without an initialised
    process group.  The registry is keyed by the LOCAL count, so a rank on which two levels have the same count cannot
    tell them apart; the decision has to be the same on every rank and must not need a host read (the step is captured
    as a CUDA graph), so such a level gets a NaN count on ALL ranks -- the step then fails loudly (NaN loss) instead of
    normalising with a wrong count or hanging in mismatched collectives.  (It takes a scene whose cloud stops shrinking
    between two levels, i.e. <= 16 points, datasetCommon.py:413-414.)"""
    _LEVEL_ROWS.clear()
    _DERIVED_ROWS.clear()
    counts = [int(c) for c in point_counts]
    if not counts or not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    key = (tuple(counts), str(device))
    local = _LOCAL_COUNTS.get(key)
    if local is None:                                 # H2D copy once per distinct packing, never inside a captured graph
        if len(_LOCAL_COUNTS) > 1024:
            _LOCAL_COUNTS.clear()
        ambiguous = [1.0 if counts.count(c) > 1 else 0.0 for c in counts]
        local = _LOCAL_COUNTS[key] = torch.tensor(counts + ambiguous, dtype=torch.float64).to(device)
    g = local.clone()
    dist.all_reduce(g)
    L = len(counts)
    total = torch.where(g[L:] > 0, torch.full_like(g[:L], float("nan")), g[:L])
    for i, c in enumerate(counts):
        _LEVEL_ROWS[c] = total[i:i + 1]


def global_rows(lead_shape, rows, device):
    """1-element device double with the number of rows summed over all ranks (stays on the device: no host sync)."""
    m = int(lead_shape[1]) if len(lead_shape) >= 2 else -1
    if m > 0 and m in _LEVEL_ROWS and rows % m == 0:
        key = (m, rows // m)
        if key not in _DERIVED_ROWS:
            _DERIVED_ROWS[key] = _LEVEL_ROWS[m] * float(rows // m)
        return _DERIVED_ROWS[key]
    d = torch.full((1,), float(rows), device=device, dtype=torch.float64)
    dist.all_reduce(d)
    return d


class _ChainFunction(torch.autograd.Function):
    """args: x2 [E, cin] (rows contiguous), spec (python: list of dict(act, has_bn, eps, momentum, sync)), training,
    then per layer (W, b, gamma, beta) tensors (gamma/beta None without BN); running stats are passed through `buffers`."""

    @staticmethod
    def forward(ctx, x2, spec, training, buffers, lead, *params):
        E = x2.shape[0]
        dev = x2.device
        L = len(spec)
        cur, ld = x2, x2.stride(0)
        in_scale = in_shift = None
        in_act = ACT_NONE
        ys, ctxs = [], []
        world = dist.get_world_size() if any(s["sync"] for s in spec) else 1
        count, d_count = E, None
        if world > 1 and training:                       # global row count stays on the device (no host sync)
            d_count = global_rows(lead, E, dev)
        for l, s in enumerate(spec):
            W, b, gamma, beta = params[4 * l: 4 * l + 4]
            cout, cin = W.shape
            y = torch.empty(E, cout, device=dev, dtype=F32)
            want_stats = s["has_bn"] and training
            nblk = ctypes.c_int(0)
            ws = workspace(lib().pcfb_mlp_workspace(E, cin, cout), dev) if want_stats else None
            check(lib().pcfb_mlp_forward(ptr(cur), ld, E, cin, cout, ptr(W), ptr(b), ptr(in_scale), ptr(in_shift), in_act,
                                         ptr(y), cout, ptr(ws), ctypes.addressof(nblk), stream_ptr()), "mlp_forward")
            scale = shift = mean = invstd = None
            if s["has_bn"]:
                rm, rv, nbt = buffers[l]
                if training:
                    scale = torch.empty(cout, device=dev, dtype=F32); shift = torch.empty_like(scale)
                    mean = torch.empty_like(scale); invstd = torch.empty_like(scale)
                    part, nb = ws, nblk.value
                    if s["sync"] and world > 1:
                        summed = torch.empty(2 * cout, device=dev, dtype=F32)
                        check(lib().pcfb_sum_partials(ptr(ws), nblk.value, 2 * cout, ptr(summed), stream_ptr()), "sum_partials")
                        sync_all_reduce(summed)
                        part, nb = summed, 1
                    check(lib().pcfb_bn_finalize(ptr(part), nb, cout, count, ptr(d_count) if s["sync"] else 0, ptr(b), ptr(gamma), ptr(beta), float(s["eps"]),
                                                 float(s["momentum"]), ptr(rm), ptr(rv), ptr(scale), ptr(shift), ptr(mean),
                                                 ptr(invstd), ptr(nbt), stream_ptr()), "bn_finalize")
                else:
                    invstd = torch.rsqrt(rv + s["eps"])
                    scale = (gamma * invstd).contiguous()
                    shift = (beta - rm * scale).contiguous()
                    mean = rm
            ys.append(y)
            ctxs.append((scale, shift, mean, invstd))
            cur, ld = y, cout
            in_scale, in_shift, in_act = scale, shift, s["act"]
        out = torch.empty_like(ys[-1])
        check(lib().pcfb_bn_act(ptr(ys[-1]), E, ys[-1].shape[1], ptr(in_scale), ptr(in_shift), in_act, ptr(out), stream_ptr()), "bn_act")
        ctx.spec, ctx.training, ctx.d_count, ctx.world = spec, training, d_count, world
        ctx.n_layers = L
        saved = [x2] + ys + list(params)
        for c in ctxs:
            saved += list(c)
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        spec, L = ctx.spec, ctx.n_layers
        saved = ctx.saved_tensors
        x2, ys = saved[0], saved[1:1 + L]
        params = saved[1 + L: 1 + 5 * L]
        cflat = saved[1 + 5 * L:]
        ctxs = [cflat[4 * l: 4 * l + 4] for l in range(L)]
        E = x2.shape[0]
        dev = x2.device
        dA = grad_out.reshape(E, -1)
        if dA.stride(-1) != 1 or dA.stride(0) != dA.shape[1]:
            dA = dA.contiguous()
        grads = [None] * (4 * L)
        train_bn = ctx.training
        zeros_cache = {}

        def bn_ctx(l):
            """(scale, shift, mean, invstd) usable by the kernels; eval-mode BN = fixed affine (mean terms vanish)."""
            scale, shift, mean, invstd = ctxs[l]
            return scale, shift, mean, invstd

        def stats(l, dA_l):
            C = ys[l].shape[1]
            scale, shift, mean, invstd = bn_ctx(l)
            sums = torch.empty(2 * C, device=dev, dtype=F32)
            if not train_bn:
                sums.zero_()
                return sums, sums
            ws = workspace(lib().pcfb_mlp_workspace(E, C, C), dev)
            check(lib().pcfb_mlp_backward_stats(ptr(dA_l), dA_l.stride(0), ptr(ys[l]), C, E, C, ptr(scale), ptr(shift), ptr(mean),
                                                ptr(invstd), spec[l]["act"], ptr(sums), ptr(ws), ws.numel(), stream_ptr()), "mlp_backward_stats")
            local = sums
            if spec[l]["sync"] and ctx.world > 1:
                # dx needs the sums over the GLOBAL batch; dgamma / dbeta stay LOCAL (the DDP gradient all-reduce adds the
                # ranks up, exactly as torch.nn.SyncBatchNorm does)
                local = sums.clone()
                sync_all_reduce(sums)
            return sums, local

        sums, sums_local = stats(L - 1, dA) if spec[L - 1]["has_bn"] else (None, None)
        need_x_grad = ctx.needs_input_grad[0]
        for l in range(L - 1, -1, -1):
            W, b, gamma, beta = params[4 * l: 4 * l + 4]
            cout, cin = W.shape
            scale, shift, mean, invstd = ctxs[l]
            has_bn = spec[l]["has_bn"]
            if l > 0:
                x_prev, ldx = ys[l - 1], ys[l - 1].shape[1]
                p_scale, p_shift, p_mean, p_invstd = ctxs[l - 1]
                in_act = spec[l - 1]["act"]
            else:
                x_prev, ldx = x2, x2.stride(0)
                p_scale = p_shift = p_mean = p_invstd = None
                in_act = ACT_NONE
            want_prev = l > 0 or need_x_grad
            dA_prev = torch.empty(E, cin, device=dev, dtype=F32) if want_prev else None
            prev_has_bn = l > 0 and spec[l - 1]["has_bn"]
            fuse_prev = prev_has_bn and train_bn and cin <= 32
            prev_sums = torch.empty(2 * cin, device=dev, dtype=F32) if fuse_prev else None
            dW = torch.empty_like(W)
            db = torch.empty(cout, device=dev, dtype=F32) if b is not None else None
            ws = workspace(lib().pcfb_mlp_workspace(E, cin, cout), dev)
            inv_scale = 1.0
            check(lib().pcfb_mlp_backward(
                ptr(dA), dA.stride(0), ptr(ys[l]), cout, E, cin, cout, ptr(W),
                ptr(scale) if has_bn else 0, ptr(shift) if has_bn else 0, ptr(mean) if has_bn else 0,
                ptr(invstd) if has_bn else 0, ptr(sums) if has_bn else 0, spec[l]["act"],
                ptr(x_prev), ldx, ptr(p_scale), ptr(p_shift), in_act, ptr(p_mean), ptr(p_invstd),
                ptr(dA_prev), cin, ptr(prev_sums), ptr(dW), ptr(db), ptr(ctx.d_count) if (has_bn and spec[l]["sync"]) else 0,
                ptr(ws), ws.numel(), stream_ptr()), "mlp_backward")
            grads[4 * l] = dW
            grads[4 * l + 1] = db
            if has_bn and gamma is not None:
                # dgamma = sum dz * xhat, dbeta = sum dz  (train mode); eval mode: recompute from dA is not needed
                # because the affine is constant w.r.t. the batch -- dgamma/dbeta then come from the same sums
                C = cout
                if train_bn:
                    grads[4 * l + 2] = sums_local[C:]    # views: AccumulateGrad takes them as they are (no copy kernels)
                    grads[4 * l + 3] = sums_local[:C]
            if l > 0:
                if prev_has_bn:
                    if fuse_prev:
                        sums = sums_local = prev_sums
                        if spec[l - 1]["sync"] and ctx.world > 1:
                            sums_local = prev_sums.clone()
                            sync_all_reduce(prev_sums)             # SyncBatchNorm: dx uses the sums over the global batch
                    else:
                        sums, sums_local = stats(l - 1, dA_prev)
                else:
                    sums = sums_local = None
                dA = dA_prev
        gx = dA_prev if need_x_grad else None
        return (gx, None, None, None, None) + tuple(grads)


def mlp_chain(x, layers, training):
    """x [..., cin]; layers: list of (linear_module, bn_module_or_None, act_code).  Returns act_L(BN_L(...)) with the
    same leading shape.  Raises if a layer size is not supported (callers check `supported`)."""
    if not x.is_cuda:
        raise RuntimeError("pcf_b200 fused MLP needs CUDA tensors (no CPU path)")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    spec, params, buffers = [], [], []
    for lin, bn, act in layers:
        has_bn = bn is not None
        spec.append(dict(act=act, has_bn=has_bn, eps=bn.eps if has_bn else 1e-5,
                         momentum=(bn.momentum if bn.momentum is not None else 0.1) if has_bn else 0.1,
                         sync=_sync_group(bn) if has_bn else False))
        params += [lin.weight, lin.bias, bn.weight if has_bn else None, bn.bias if has_bn else None]
        # num_batches_tracked is incremented by the finalize kernel (one tiny torch add_ per BatchNorm was 273 launches a step)
        buffers.append((bn.running_mean, bn.running_var, bn.num_batches_tracked) if has_bn else (None, None, None))
    out = _ChainFunction.apply(x2, spec, training, buffers, tuple(lead), *params)
    return out.reshape(*lead, out.shape[-1])


# ---- BatchNorm (+ activation) on wide / post-contraction tensors (csrc/mlp.cu, pcfb_bn_*) ---------------------------
def bn_supported(C):
    return bool(lib().pcfb_bn_supported(int(C)))


class _BnActFunction(torch.autograd.Function):
    """out = act(BatchNorm(x2)) for contiguous x2 [rows, C]; cfg = dict(act, training, eps, momentum, sync, lead)."""

    @staticmethod
    def forward(ctx, x2, gamma, beta, pivot, running_mean, running_var, cfg):
        rows, C = x2.shape
        dev = x2.device
        act, training = cfg["act"], cfg["training"]
        world = dist.get_world_size() if cfg["sync"] else 1
        d_count = None
        if training:
            if world > 1:
                d_count = global_rows(cfg["lead"], rows, dev)
            ws = workspace(lib().pcfb_bn_workspace(rows, C), dev)
            nblk = ctypes.c_int(0)
            check(lib().pcfb_bn_stats(ptr(x2), rows, C, ptr(pivot), ptr(ws), ws.numel(), ctypes.addressof(nblk), stream_ptr()), "bn_stats")
            part, nb = ws, nblk.value
            if world > 1:
                summed = torch.empty(2 * C, device=dev, dtype=F32)
                check(lib().pcfb_sum_partials(ptr(ws), nblk.value, 2 * C, ptr(summed), stream_ptr()), "sum_partials")
                sync_all_reduce(summed)
                part, nb = summed, 1
            scale = torch.empty(C, device=dev, dtype=F32); shift = torch.empty_like(scale)
            mean = torch.empty_like(scale); invstd = torch.empty_like(scale)
            check(lib().pcfb_bn_finalize(ptr(part), nb, C, rows, ptr(d_count), ptr(pivot), ptr(gamma), ptr(beta), float(cfg["eps"]),
                                         float(cfg["momentum"]), ptr(running_mean), ptr(running_var), ptr(scale), ptr(shift),
                                         ptr(mean), ptr(invstd), ptr(cfg["nbt"]), stream_ptr()), "bn_finalize")
        else:
            invstd = torch.rsqrt(running_var + cfg["eps"])
            scale = (gamma * invstd).contiguous() if gamma is not None else invstd.contiguous()
            shift = ((beta if beta is not None else 0.) - running_mean * scale).contiguous()
            mean = running_mean
        out = torch.empty_like(x2)
        check(lib().pcfb_bn_act(ptr(x2), rows, C, ptr(scale), ptr(shift), act, ptr(out), stream_ptr()), "bn_act")
        ctx.cfg, ctx.d_count, ctx.world = cfg, d_count, world
        ctx.has_affine = gamma is not None
        ctx.save_for_backward(x2, scale, shift, mean, invstd)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x2, scale, shift, mean, invstd = ctx.saved_tensors
        rows, C = x2.shape
        dev = x2.device
        cfg = ctx.cfg
        dA = grad_out.reshape(rows, C)
        if not dA.is_contiguous():
            dA = dA.contiguous()
        need_affine = ctx.has_affine and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        sums = local = None
        if cfg["training"] or need_affine:
            sums = torch.empty(2 * C, device=dev, dtype=F32)
            ws = workspace(lib().pcfb_bn_workspace(rows, C), dev)
            check(lib().pcfb_bn_backward_stats(ptr(dA), ptr(x2), rows, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), cfg["act"],
                                               ptr(sums), ptr(ws), ws.numel(), stream_ptr()), "bn_backward_stats")
            local = sums
            if cfg["training"] and ctx.world > 1:          # dx: sums over the global batch; dgamma / dbeta stay local
                local = sums.clone()
                sync_all_reduce(sums)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x2)
            check(lib().pcfb_bn_backward(ptr(dA), ptr(x2), rows, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                                         ptr(sums) if cfg["training"] else 0, cfg["act"], ptr(ctx.d_count), ptr(dx), stream_ptr()), "bn_backward")
        dgamma = local[C:] if need_affine else None
        dbeta = local[:C] if need_affine else None
        return dx, dgamma, dbeta, None, None, None, None


def bn_act(x, bn, act, pivot=None):
    """act(bn(x)) over the last dim of x [..., C] with a BatchNorm1d/2d/SyncBatchNorm-like module `bn` (batch statistics
    over all leading dims when bn.training, running statistics otherwise; running stats / num_batches_tracked updated
    like torch).  pivot: optional per-channel offset near the mean (the bias of the Linear that produced x)."""
    if not x.is_cuda:
        raise RuntimeError("pcf_b200 BatchNorm needs CUDA tensors (no CPU path)")
    lead = tuple(x.shape[:-1])
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    training = bn.training or not bn.track_running_stats
    cfg = dict(act=act, training=training, eps=bn.eps, momentum=0.1 if bn.momentum is None else bn.momentum,
               sync=_sync_group(bn), lead=lead,
               nbt=bn.num_batches_tracked if bn.track_running_stats else None)      # incremented by the finalize kernel
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    out = _BnActFunction.apply(x2, bn.weight, bn.bias, pivot.detach() if pivot is not None else None, rm, rv, cfg)
    return out.reshape(*lead, out.shape[-1])
