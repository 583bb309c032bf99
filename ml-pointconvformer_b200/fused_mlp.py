"""Fused small-MLP chains with BatchNorm (WeightNet, pe_convs, mlp_conv, guidance MLP, UnaryBlock) on top of
pcfb_mlp_* (csrc/mlp.cu): one streaming kernel pass per layer in the forward, two in the backward, BatchNorm
statistics accumulated on the fly, every reduction deterministic.  Mirrors what the reference computes with
Linear_BN + activation sequences (/root/reference/layers.py:127-191, 38-68; layer_utils.py:241-319)."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib
from . import streams as S
from ._lib import check, lib, ptr, stream_ptr, workspace

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3
F32 = torch.float32
SYNC_CHANNELS = True        # SyncBatchNorm exchanges are serialised on one stream (_on_exchange_stream): side streams are safe


def supported(dims):
    """dims: [(cin, cout), ...]"""
    return all(lib().pcfb_mlp_supported(int(a), int(b)) for a, b in dims)


def _sync_group(bn):
    if isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return True
    return False


# ---- the SyncBatchNorm statistics exchange --------------------------------------------------------------------------
# ~540 exchanges of <= 3 KB per training step: latency only.  When the ranks share an NVLink / NVSwitch domain and
# PCFB_PEER_REDUCE=1 (bench.py sets it; opt-in elsewhere: the exchange spins on the GPU until every rank arrives, for at
# most PCFB_PEER_TIMEOUT_S seconds, default 600) the partial sums -> exchange -> finalize sequence is ONE kernel over
# symmetric peer memory (csrc/peer_reduce.cu); otherwise -- other backend, no peer access -- the sums go through
# dist.all_reduce.  torch.distributed._symmetric_memory only provides the allocation + address exchange (plumbing).
# The global row count travels inside the message: nothing has to know the pyramid's level sizes.
_PEER = {"state": None}       # None = not tried, False = unavailable, dict = ready


def _peer_setup(device):
    st = {"ok": False}
    try:
        if os.environ.get("PCFB_PEER_REDUCE", "0") != "1" or dist.get_backend() != "nccl":
            raise RuntimeError("disabled")
        import torch.distributed._symmetric_memory as symm_mem
        world, rank = dist.get_world_size(), dist.get_rank()
        nbytes = int(lib().pcfb_syncbn_buffer_bytes(world))
        if nbytes == 0:
            raise RuntimeError("world size not supported")
        buf = symm_mem.empty(nbytes // 4, dtype=F32, device=device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        bases = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64).to(device)
        torch.cuda.synchronize(device)
        st = {"ok": True, "buf": buf, "hdl": hdl, "bases": bases, "rank": rank, "world": world,
              "timeout": float(os.environ.get("PCFB_PEER_TIMEOUT_S", "600"))}
    except Exception as e:                                    # every rank must agree: vote below
        st = {"ok": False, "why": "%s: %s" % (type(e).__name__, e)}
    vote = torch.tensor([1.0 if st["ok"] else 0.0], device=device)
    dist.all_reduce(vote, op=dist.ReduceOp.MIN)               # also orders the zero-fill of every buffer before its first use
    if float(vote.item()) < 1.0:
        return False
    return st


def _peer(device):
    """The peer-memory exchange state, False if unavailable.  Set up in the eager warm-up steps (needs host syncs)."""
    st = _PEER["state"]
    if st is None:
        if device.type != "cuda" or torch.cuda.is_current_stream_capturing():
            return False
        st = _PEER["state"] = _peer_setup(device)
    return st


def peer_error():
    """True if an exchange on this rank gave up waiting for a peer (the kernel then wrote NaN statistics)."""
    st = _PEER["state"]
    return bool(st) and bool(st["buf"][0].view(torch.int32).item() != 0)


_XSTREAM = {}


def _on_exchange_stream(launch, device):
    """Run `launch()` (one exchange kernel) on the rank's single exchange stream.  When the branches of a layer run on side
    streams (streams.py), their SyncBatchNorm exchanges must still execute in ONE order that is the same on every rank: two
    spinning exchange kernels that the hardware happens to serialise in opposite orders on two GPUs would wait for each other
    forever (CUDA does not guarantee that independent streams / graph branches run concurrently).  Python program order is
    the same on every rank (same model, autograd replays nodes by sequence number), so enqueueing every exchange on one
    stream gives that order; the compute branches around them still overlap."""
    if not S.ENABLED:
        return launch()
    cur = torch.cuda.current_stream()
    key = (device.type, device.index)
    xs = _XSTREAM.get(key)
    if xs is None:
        xs = _XSTREAM[key] = torch.cuda.Stream(device=device)
    xs.wait_stream(cur)
    with torch.cuda.stream(xs):
        launch()
    cur.wait_stream(xs)


def sync_all_reduce(t):
    """In-place sum of a small tensor over all ranks through torch.distributed (gloo / CPU tensors, NCCL fallback)."""
    dist.all_reduce(t)
    return t


def bn_finalize(part, nb, C, rows, pivot, gamma, beta, eps, momentum, rm, rv, nbt, sync, dev):
    """Block partials [nb][2][C] -> (scale, shift, mean, invstd, d_count): one launch (csrc/peer_reduce.cu), including the
    cross-rank exchange when `sync`.  d_count: 1-element device double with the global row count (None without sync)."""
    scale = torch.empty(C, device=dev, dtype=F32); shift = torch.empty_like(scale)
    mean = torch.empty_like(scale); invstd = torch.empty_like(scale)
    if momentum is None:                                       # cumulative moving average: 1 / num_batches_tracked
        if nbt is not None:
            nbt.add_(1)
        momentum = -1.0
    world = dist.get_world_size() if sync else 1
    d_count = torch.empty(1, device=dev, dtype=torch.float64) if world > 1 else None

    def launch(partial, nblocks, d_count_in, bases, rank, wld, timeout):
        check(lib().pcfb_bn_finalize(ptr(partial), nblocks, C, rows, ptr(d_count_in), ptr(pivot), ptr(gamma), ptr(beta), float(eps),
                                     float(momentum), ptr(rm), ptr(rv), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), ptr(nbt),
                                     ptr(d_count) if d_count_in is None else 0, bases, rank, wld, 0, timeout, stream_ptr()), "bn_finalize")
    if world == 1:
        if FORCE_XSTREAM:
            _on_exchange_stream(lambda: launch(part, nb, None, 0, 0, 1, 0.0), dev)
        else:
            launch(part, nb, None, 0, 0, 1, 0.0)
        return scale, shift, mean, invstd, None
    st = _peer(dev)
    if st:
        _on_exchange_stream(lambda: launch(part, nb, None, ptr(st["bases"]), st["rank"], st["world"], st["timeout"]), dev)
        return scale, shift, mean, invstd, d_count
    # torch.distributed fallback: local sums -> all_reduce of (sums, count) in double -> finalize from the reduced sums
    local = torch.empty(2 * C, device=dev, dtype=F32)
    check(lib().pcfb_bn_reduce_sums(ptr(part), nb, C, ptr(local), 0, 0, 0, 1, 0, 0.0, stream_ptr()), "bn_reduce_sums")
    msg = torch.cat([local.double(), torch.full((1,), float(rows), device=dev, dtype=torch.float64)])
    sync_all_reduce(msg)
    summed = msg[:2 * C].float()
    d_count = msg[2 * C:].contiguous()
    launch(summed, 1, d_count, 0, 0, 1, 0.0)
    return scale, shift, mean, invstd, d_count


def bn_reduce_sums(part, nb, C, sync, dev):
    """Block partials [nb][2][C] of (sum dz, sum dz*xhat) -> (sums over the global batch, local sums).  dx needs the global
    sums; dgamma / dbeta stay local (the DDP gradient all-reduce adds the ranks up, exactly as torch.nn.SyncBatchNorm does)."""
    local = torch.empty(2 * C, device=dev, dtype=F32)
    world = dist.get_world_size() if sync else 1
    if world == 1:
        run = (lambda f: _on_exchange_stream(f, dev)) if FORCE_XSTREAM else (lambda f: f())
        run(lambda: check(lib().pcfb_bn_reduce_sums(ptr(part), nb, C, ptr(local), 0, 0, 0, 1, 0, 0.0, stream_ptr()), "bn_reduce_sums"))
        return local, local
    st = _peer(dev)
    if st:
        glob = torch.empty(2 * C, device=dev, dtype=F32)
        _on_exchange_stream(lambda: check(lib().pcfb_bn_reduce_sums(ptr(part), nb, C, ptr(local), ptr(glob), ptr(st["bases"]), st["rank"],
                                                                    st["world"], 0, st["timeout"], stream_ptr()), "bn_reduce_sums"), dev)
        return glob, local
    check(lib().pcfb_bn_reduce_sums(ptr(part), nb, C, ptr(local), 0, 0, 0, 1, 0, 0.0, stream_ptr()), "bn_reduce_sums")
    glob = local.clone()
    sync_all_reduce(glob)
    return glob, local


def eval_affine(gamma, beta, rm, rv, eps):
    """(scale, shift, invstd) of a BatchNorm that uses its running statistics: scale = gamma / sqrt(rv + eps), shift = beta -
    rm * scale, in ONE launch (pcfb_bn_eval_affine): torch's rsqrt / mul / mul / sub / contiguous were 1 065 of the 1 764
    launches of an eval-mode forward of configPCF_10cm_lite (scripts/profile_infer.py).  Not cached: the running statistics
    are updated by kernels (and graph replays) that no version counter sees."""
    C = rv.shape[0]
    scale = torch.empty(C, device=rv.device, dtype=F32); shift = torch.empty_like(scale); invstd = torch.empty_like(scale)
    ok = all(t is None or (t.dtype == F32 and t.is_contiguous()) for t in (gamma, beta, rm, rv))
    if not ok:
        with torch.no_grad():
            invstd = torch.rsqrt(rv + eps)
            scale = (gamma * invstd).contiguous() if gamma is not None else invstd.contiguous()
            shift = ((beta if beta is not None else 0.) - rm * scale).contiguous()
        return scale, shift, invstd
    check(lib().pcfb_bn_eval_affine(ptr(gamma), ptr(beta), ptr(rm), ptr(rv), float(eps), C, ptr(scale), ptr(shift), ptr(invstd),
                                    stream_ptr()), "bn_eval_affine")
    return scale, shift, invstd


CHAIN_EVAL = os.environ.get("PCFB_CHAIN_EVAL", "1") != "0"      # one-kernel inference chain (csrc/mlp_eval.cu)


def _chain_eval(x2, spec, buffers, params):
    """The whole three-layer chain in one kernel when no BatchNorm uses batch statistics and no gradient is recorded
    (pcfb_mlp_chain_eval); None if the sizes have no fused kernel."""
    Ws = [params[4 * l] for l in range(3)]
    dims = [Ws[0].shape[1]] + [w.shape[0] for w in Ws]
    if any(Ws[l].shape[1] != dims[l] for l in range(3)) or not lib().pcfb_mlp_chain_eval_supported(*dims):
        return None
    keep, scales, shifts = [], [], []
    for l, s in enumerate(spec):
        gamma, beta = params[4 * l + 2], params[4 * l + 3]
        if s["has_bn"]:                                       # eval-mode BatchNorm: a fixed affine map
            rm, rv, _ = buffers[l]
            scale, shift, _ = eval_affine(gamma, beta, rm, rv, s["eps"])
            keep += [scale, shift]
            scales.append(ptr(scale)); shifts.append(ptr(shift))
        else:
            scales.append(0); shifts.append(0)
    arr = lambda vals: (ctypes.c_void_p * 3)(*[v or None for v in vals])
    Wc = [w if w.is_contiguous() else w.contiguous() for w in Ws]
    w_arr, b_arr = arr([ptr(w) for w in Wc]), arr([ptr(params[4 * l + 1]) for l in range(3)])
    sc_arr, sh_arr = arr(scales), arr(shifts)
    acts = (ctypes.c_int * 3)(*[int(s["act"]) for s in spec])
    E = x2.shape[0]
    out = torch.empty(E, dims[3], device=x2.device, dtype=F32)
    check(lib().pcfb_mlp_chain_eval(ptr(x2), x2.stride(0), E, dims[0], dims[1], dims[2], dims[3], w_arr, b_arr, sc_arr, sh_arr,
                                    acts, ptr(out), dims[3], stream_ptr()), "mlp_chain_eval")
    _lib.account(4.0 * E * (dims[0] + dims[3]), 2.0 * E * (dims[0] * dims[1] + dims[1] * dims[2] + dims[2] * dims[3]))
    return out


class _ChainFunction(torch.autograd.Function):
    """args: x2 [E, cin] (rows contiguous), spec (python: list of dict(act, has_bn, train, eps, momentum, sync)), then per
    layer (W, b, gamma, beta) tensors (gamma/beta None without BN); running stats are passed through `buffers`."""

    @staticmethod
    def forward(ctx, x2, spec, buffers, *params):
        E = x2.shape[0]
        dev = x2.device
        L = len(spec)
        # (grad mode is always off inside Function.forward: whether a backward can follow is decided by mlp_chain())
        if CHAIN_EVAL and L == 3 and spec[0].get("no_grad") and not any(s["train"] for s in spec):
            fused = _chain_eval(x2, spec, buffers, params)
            if fused is not None:
                return fused
        cur, ld = x2, x2.stride(0)
        in_scale = in_shift = None
        in_act = ACT_NONE
        ys, ctxs, counts = [], [], []
        for l, s in enumerate(spec):
            W, b, gamma, beta = params[4 * l: 4 * l + 4]
            cout, cin = W.shape
            y = torch.empty(E, cout, device=dev, dtype=F32)
            want_stats = s["has_bn"] and s["train"]
            nblk = ctypes.c_int(0)
            ws = workspace(lib().pcfb_mlp_workspace(E, cin, cout), dev) if want_stats else None
            check(lib().pcfb_mlp_forward(ptr(cur), ld, E, cin, cout, ptr(W), ptr(b), ptr(in_scale), ptr(in_shift), in_act,
                                         ptr(y), cout, ptr(ws), ctypes.addressof(nblk), stream_ptr()), "mlp_forward")
            _lib.account(4.0 * E * (cin + cout), 2.0 * E * cin * cout)
            scale = shift = mean = invstd = d_count = None
            if s["has_bn"]:
                rm, rv, nbt = buffers[l]
                if want_stats:
                    scale, shift, mean, invstd, d_count = bn_finalize(ws, nblk.value, cout, E, b, gamma, beta, s["eps"], s["momentum"],
                                                                      rm, rv, nbt, s["sync"], dev)
                else:                                        # eval-mode BatchNorm: a fixed affine map
                    scale, shift, invstd = eval_affine(gamma, beta, rm, rv, s["eps"])
                    mean = rm
            ys.append(y)
            ctxs.append((scale, shift, mean, invstd))
            counts.append(d_count)
            cur, ld = y, cout
            in_scale, in_shift, in_act = scale, shift, s["act"]
        out = torch.empty_like(ys[-1])
        check(lib().pcfb_bn_act(ptr(ys[-1]), E, ys[-1].shape[1], ptr(in_scale), ptr(in_shift), in_act, ptr(out), 0, 0, stream_ptr()), "bn_act")
        _lib.account(8.0 * E * ys[-1].shape[1])
        ctx.spec, ctx.counts = spec, counts
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

        def stats(l, dA_l):
            """-> (global sums, local sums) of (dz, dz*xhat) for layer l's BatchNorm (running statistics in eval mode)."""
            C = ys[l].shape[1]
            scale, shift, mean, invstd = ctxs[l]
            ws = workspace(lib().pcfb_mlp_workspace(E, C, C), dev)
            nblk = ctypes.c_int(0)
            check(lib().pcfb_mlp_backward_stats(ptr(dA_l), dA_l.stride(0), ptr(ys[l]), C, E, C, ptr(scale), ptr(shift), ptr(mean),
                                                ptr(invstd), spec[l]["act"], 0, ctypes.addressof(nblk), ptr(ws), ws.numel(),
                                                stream_ptr()), "mlp_backward_stats")
            _lib.account(8.0 * E * C)
            return bn_reduce_sums(ws, nblk.value, C, spec[l]["sync"] and spec[l]["train"], dev)

        def affine_needed(l):
            gamma, beta = params[4 * l + 2], params[4 * l + 3]
            return (gamma is not None and gamma.requires_grad) or (beta is not None and beta.requires_grad)

        top = spec[L - 1]
        sums = sums_local = None
        if top["has_bn"] and (top["train"] or affine_needed(L - 1)):
            sums, sums_local = stats(L - 1, dA)
        need_x_grad = ctx.needs_input_grad[0]
        zeros = {}
        for l in range(L - 1, -1, -1):
            W, b, gamma, beta = params[4 * l: 4 * l + 4]
            cout, cin = W.shape
            scale, shift, mean, invstd = ctxs[l]
            has_bn, train_l = spec[l]["has_bn"], spec[l]["train"]
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
            fuse_prev = prev_has_bn and spec[l - 1]["train"] and cin <= 32
            prev_sums = torch.empty(2 * cin, device=dev, dtype=F32) if fuse_prev else None
            dW = torch.empty_like(W)
            db = torch.empty(cout, device=dev, dtype=F32) if b is not None else None
            ws = workspace(lib().pcfb_mlp_workspace(E, cin, cout), dev)
            sums_dx = None
            if has_bn:
                if train_l:
                    sums_dx = sums
                else:                                        # eval-mode BatchNorm: dy = scale * dz, the batch terms vanish
                    sums_dx = zeros.get(cout)
                    if sums_dx is None:
                        sums_dx = zeros[cout] = torch.zeros(2 * cout, device=dev, dtype=F32)
            d_count = ctx.counts[l] if (has_bn and train_l) else None
            check(lib().pcfb_mlp_backward(
                ptr(dA), dA.stride(0), ptr(ys[l]), cout, E, cin, cout, ptr(W),
                ptr(scale) if has_bn else 0, ptr(shift) if has_bn else 0, ptr(mean) if has_bn else 0,
                ptr(invstd) if has_bn else 0, ptr(sums_dx) if has_bn else 0, spec[l]["act"],
                ptr(x_prev), ldx, ptr(p_scale), ptr(p_shift), in_act, ptr(p_mean), ptr(p_invstd),
                ptr(dA_prev), cin, ptr(prev_sums), ptr(dW), ptr(db), ptr(d_count),
                ptr(ws), ws.numel(), stream_ptr()), "mlp_backward")
            _lib.account(4.0 * E * (2 * cout + cin + (cin if want_prev else 0)), (4.0 if want_prev else 2.0) * E * cin * cout)
            grads[4 * l] = dW
            grads[4 * l + 1] = db
            if has_bn and gamma is not None and sums_local is not None:
                grads[4 * l + 2] = sums_local[cout:]          # dgamma = sum dz * xhat, dbeta = sum dz: views, no copy kernels
                grads[4 * l + 3] = sums_local[:cout]
            if l > 0:
                sums = sums_local = None
                if prev_has_bn:
                    if fuse_prev:
                        sums = sums_local = prev_sums
                        if spec[l - 1]["sync"] and dist.get_world_size() > 1:     # SyncBatchNorm: dx uses the sums over the global batch
                            sums, sums_local = bn_reduce_sums(prev_sums, 1, cin, True, dev)
                    elif spec[l - 1]["train"] or affine_needed(l - 1):
                        sums, sums_local = stats(l - 1, dA_prev)
                dA = dA_prev
        gx = dA_prev if need_x_grad else None
        return (gx, None, None) + tuple(grads)


def mlp_chain(x, layers, training=None):
    """x [..., cin]; layers: list of (linear_module, bn_module_or_None, act_code).  Returns act_L(BN_L(...)) with the
    same leading shape.  Every BatchNorm follows its OWN mode (bn.training or not bn.track_running_stats, as torch does):
    freezing only the BatchNorm modules (bn.eval()) freezes their statistics in the fused chains too; `training` is accepted
    for compatibility and ignored.  Raises if a layer size is not supported (callers check `supported`)."""
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
                         train=has_bn and (bn.training or not bn.track_running_stats),
                         momentum=bn.momentum if has_bn else 0.1,            # None = cumulative moving average (torch semantics)
                         sync=_sync_group(bn) if has_bn else False))
        params += [lin.weight, lin.bias, bn.weight if has_bn else None, bn.bias if has_bn else None]
        # num_batches_tracked is incremented by the finalize kernel (one tiny torch add_ per BatchNorm was 273 launches a step)
        tracked = has_bn and bn.track_running_stats
        buffers.append((bn.running_mean, bn.running_var, bn.num_batches_tracked) if tracked else (None, None, None))
    spec[0]["no_grad"] = not torch.is_grad_enabled()           # inference: the one-kernel chain may be used (no backward follows)
    out = _ChainFunction.apply(x2, spec, buffers, *params)
    return out.reshape(*lead, out.shape[-1])


# ---- BatchNorm (+ activation) on wide / post-contraction tensors (csrc/mlp.cu, pcfb_bn_*) ---------------------------
def bn_supported(C):
    return bool(lib().pcfb_bn_supported(int(C)))


FORCE_XSTREAM = os.environ.get("PCFB_FORCE_XSTREAM", "0") == "1"   # diagnostic: pay the exchange-stream hop at world size 1
SMALL_BN = os.environ.get("PCFB_BN_SMALL", "1") != "0"      # one-kernel BatchNorm for small tensors (csrc/peer_reduce.cu)
_SMALL_ROWS = []


def _small_bn_rows():
    if not _SMALL_ROWS:
        _SMALL_ROWS.append(int(os.environ.get("PCFB_BN_SMALL_ROWS", lib().pcfb_bn_small_max_rows())))
    return _SMALL_ROWS[0]


def _exchange_args(sync, dev):
    """(run, bases, rank, world, timeout) for a kernel that contains a SyncBatchNorm exchange: `run(launch)` executes the
    launch on the exchange stream when there is an exchange.  None if the peer path is unavailable (torch.distributed fallback)."""
    world = dist.get_world_size() if sync else 1
    if world == 1:
        return ((lambda launch: _on_exchange_stream(launch, dev)) if FORCE_XSTREAM else (lambda launch: launch())), 0, 0, 1, 0.0
    st = _peer(dev)
    if not st:
        return None
    return (lambda launch: _on_exchange_stream(launch, dev)), ptr(st["bases"]), st["rank"], st["world"], st["timeout"]


class _BnActFunction(torch.autograd.Function):
    """out = act(BatchNorm(x2)) for contiguous x2 [rows, C]; cfg = dict(act, training, eps, momentum, sync, nbt)."""

    @staticmethod
    def forward(ctx, x2, gamma, beta, pivot, running_mean, running_var, cfg, residual):
        rows, C = x2.shape
        dev = x2.device
        act, training = cfg["act"], cfg["training"]
        d_count = None
        xa = _exchange_args(cfg["sync"], dev) if (training and SMALL_BN and 1 <= rows <= _small_bn_rows()) else None
        ctx.small = xa is not None
        if ctx.small:
            # small tensor: statistics + (exchange) + finalize + apply in ONE launch
            run, bases, rank, world, timeout = xa
            scale = torch.empty(C, device=dev, dtype=F32); shift = torch.empty_like(scale)
            mean = torch.empty_like(scale); invstd = torch.empty_like(scale)
            momentum, nbt = cfg["momentum"], cfg["nbt"]
            if momentum is None:                                   # cumulative moving average: 1 / num_batches_tracked
                if nbt is not None:
                    nbt.add_(1)
                momentum = -1.0
            d_count = torch.empty(1, device=dev, dtype=torch.float64) if world > 1 else None
            out = torch.empty_like(x2)
            after = 1 if cfg.get("res_after") else 0
            run(lambda: check(lib().pcfb_bn_small_forward(
                ptr(x2), rows, C, ptr(pivot), ptr(gamma), ptr(beta), float(cfg["eps"]), float(momentum), ptr(running_mean), ptr(running_var),
                ptr(nbt), act, ptr(residual), after, ptr(out), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), ptr(d_count),
                bases, rank, world, 0, timeout, stream_ptr()), "bn_small_forward"))
            _lib.account((12.0 if residual is None else 16.0) * rows * C)
            ctx.cfg, ctx.d_count = cfg, d_count
            ctx.has_affine = gamma is not None
            ctx.pre_res = residual is not None and not after
            ctx.save_for_backward(x2, scale, shift, mean, invstd, residual if ctx.pre_res else None)
            return out
        if training:
            ws = workspace(lib().pcfb_bn_workspace(rows, C), dev)
            nblk = ctypes.c_int(0)
            check(lib().pcfb_bn_stats(ptr(x2), rows, C, ptr(pivot), ptr(ws), ws.numel(), ctypes.addressof(nblk), stream_ptr()), "bn_stats")
            _lib.account(4.0 * rows * C)
            scale, shift, mean, invstd, d_count = bn_finalize(ws, nblk.value, C, rows, pivot, gamma, beta, cfg["eps"], cfg["momentum"],
                                                              running_mean, running_var, cfg["nbt"], cfg["sync"], dev)
        else:
            scale, shift, invstd = eval_affine(gamma, beta, running_mean, running_var, cfg["eps"])
            mean = running_mean
        out = torch.empty_like(x2)
        after = 1 if cfg.get("res_after") else 0
        check(lib().pcfb_bn_act(ptr(x2), rows, C, ptr(scale), ptr(shift), act, ptr(out), ptr(residual), after, stream_ptr()), "bn_act")
        _lib.account((8.0 if residual is None else 12.0) * rows * C)
        ctx.cfg, ctx.d_count = cfg, d_count
        ctx.has_affine = gamma is not None
        # only a residual added BEFORE the activation enters the backward kernels (act' is evaluated at z + r)
        ctx.pre_res = residual is not None and not after
        ctx.save_for_backward(x2, scale, shift, mean, invstd, residual if ctx.pre_res else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x2, scale, shift, mean, invstd, res = ctx.saved_tensors
        rows, C = x2.shape
        dev = x2.device
        cfg = ctx.cfg
        dA = grad_out.reshape(rows, C)
        if not dA.is_contiguous():
            dA = dA.contiguous()
        need_res = ctx.needs_input_grad[7]
        need_affine = ctx.has_affine and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        sums = local = None
        xa = _exchange_args(cfg["sync"], dev) if ctx.small else None
        if xa is not None:
            # small tensor (training mode): sums + (exchange) + dX / d_residual in ONE launch
            run, bases, rank, world, timeout = xa
            want_dres = need_res and ctx.pre_res
            local = torch.empty(2 * C, device=dev, dtype=F32)
            dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
            d_res = torch.empty_like(x2) if want_dres else None
            run(lambda: check(lib().pcfb_bn_small_backward(
                ptr(dA), ptr(x2), rows, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), cfg["act"], ptr(res), ptr(ctx.d_count),
                ptr(local), ptr(dx), ptr(d_res), bases, rank, world, 0, timeout, stream_ptr()), "bn_small_backward"))
            _lib.account((20.0 + (8.0 if res is not None else 0.0) + (4.0 if want_dres else 0.0)) * rows * C)
            if need_res and not ctx.pre_res:
                d_res = dA
            return (dx, local[C:] if need_affine else None, local[:C] if need_affine else None, None, None, None, None, d_res)
        if cfg["training"] or need_affine:
            ws = workspace(lib().pcfb_bn_workspace(rows, C), dev)
            nblk = ctypes.c_int(0)
            check(lib().pcfb_bn_backward_stats(ptr(dA), ptr(x2), rows, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd), cfg["act"],
                                               ptr(res), 0, ctypes.addressof(nblk), ptr(ws), ws.numel(), stream_ptr()), "bn_backward_stats")
            _lib.account(8.0 * rows * C)
            # dx: sums over the global batch; dgamma / dbeta stay local
            sums, local = bn_reduce_sums(ws, nblk.value, C, cfg["sync"] and cfg["training"], dev)
        dx = d_res = None
        want_dres = need_res and ctx.pre_res
        if ctx.needs_input_grad[0] or want_dres:
            dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
            d_res = torch.empty_like(x2) if want_dres else None
            check(lib().pcfb_bn_backward(ptr(dA), ptr(x2), rows, C, ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                                         ptr(sums) if cfg["training"] else 0, cfg["act"], ptr(ctx.d_count), ptr(res), ptr(dx), ptr(d_res),
                                         stream_ptr()), "bn_backward")
            _lib.account((12.0 + (4.0 if res is not None else 0.0) + (4.0 if want_dres else 0.0)) * rows * C)
        if need_res and not ctx.pre_res:
            d_res = dA                                        # added after the activation: the gradient passes through
        dgamma = local[C:] if need_affine else None
        dbeta = local[:C] if need_affine else None
        return dx, dgamma, dbeta, None, None, None, None, d_res


def bn_act(x, bn, act, pivot=None, residual=None, residual_after_act=False):
    """act(bn(x)) over the last dim of x [..., C] with a BatchNorm1d/2d/SyncBatchNorm-like module `bn` (batch statistics
    over all leading dims when bn.training, running statistics otherwise; running stats / num_batches_tracked updated
    like torch).  pivot: optional per-channel offset near the mean (the bias of the Linear that produced x).
    residual (same shape as x): act(bn(x) + residual), or act(bn(x)) + residual with residual_after_act, in the same pass."""
    if not x.is_cuda:
        raise RuntimeError("pcf_b200 BatchNorm needs CUDA tensors (no CPU path)")
    lead = tuple(x.shape[:-1])
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    training = bn.training or not bn.track_running_stats
    cfg = dict(act=act, training=training, eps=bn.eps, momentum=bn.momentum,      # None = cumulative moving average
               sync=_sync_group(bn), res_after=bool(residual_after_act),
               nbt=bn.num_batches_tracked if bn.track_running_stats else None)      # incremented by the finalize kernel
    rm, rv = (bn.running_mean, bn.running_var) if bn.track_running_stats else (None, None)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, x.shape[-1])
        if not r2.is_contiguous():
            r2 = r2.contiguous()
    out = _BnActFunction.apply(x2, bn.weight, bn.bias, pivot.detach() if pivot is not None else None, rm, rv, cfg, r2)
    return out.reshape(*lead, out.shape[-1])
