"""Multi-GPU plumbing: scenes are the unit of sharding (edges never cross scenes,
knn_post_dataloader_utils.py:194-212), one process per GPU, no data-path collective.  Training adds the
DDP gradient all-reduce only (train_ScanNet_DDP_WarmUP.py:194-195)."""
import torch
import torch.distributed as dist


def shard_scenes(sizes, world_size):
    """Longest-processing-time greedy partition of scene indices over ranks (balanced by point count)."""
    parts = [[] for _ in range(world_size)]
    loads = [0] * world_size
    for i in sorted(range(len(sizes)), key=lambda i: -sizes[i]):
        r = min(range(world_size), key=lambda r: (loads[r], r))
        parts[r].append(i)
        loads[r] += sizes[i]
    for p in parts:
        p.sort()
    return parts


def allreduce_gradients(params, world_size=None, group=None):
    """One flat-bucket gradient all-reduce (mean), the only bandwidth-relevant collective of DDP training
    (PCF_Normal: 5.4 M parameters = 21.7 MB fp32 = one bucket).  Works with nccl and gloo."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    world_size = world_size or dist.get_world_size(group)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world_size)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


class FlatParameters:
    """All trainable parameters of a module as views into ONE contiguous buffer, so that the optimizer, the gradient
    clipping and the DDP all-reduce each touch a single tensor (PCF_Normal has 382 parameter tensors: the per-tensor
    optimizer bookkeeping alone was ~3 ms of a 48 ms training step).  The module keeps using its own parameter
    objects; only their storage moves.  Call after any module surgery (e.g. SyncBatchNorm conversion)."""

    def __init__(self, module, async_weight_grads=False):
        """async_weight_grads: let the weight-gradient GEMMs of the backward run on their own stream (streams.LEAF_ASYNC);
        gather_grads() joins them.  Only for loops that read gradients through gather_grads()."""
        if async_weight_grads:
            from . import streams
            streams.LEAF_ASYNC = True
            streams.PREP_ASYNC = True
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        ref = self.params[0]
        total = sum(p.numel() for p in self.params)
        flat = torch.empty(total, device=ref.device, dtype=ref.dtype)
        off = 0
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                off += n
        self.flat = torch.nn.Parameter(flat)

    def gather_grads(self, world_size=1, group=None, sync=True, accumulate=False):
        """Concatenate the per-parameter gradients into the flat gradient (mean over ranks when world_size > 1), attach
        it to the flat parameter and drop the per-parameter ones.  Gradient accumulation: accumulate=True adds to the flat
        gradient of the previous micro-steps, sync=False skips the all-reduce (DDP's no_sync; the reference all-reduces on
        every micro-step, train_ScanNet_DDP_WarmUP.py:418-424) -- all-reduce once, on the last micro-step."""
        from . import streams
        streams.join_leaves()                           # weight gradients that ran on the leaf stream (streams.fork_leaf)
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
        g = torch.cat(pieces)
        if accumulate and self.flat.grad is not None:
            g += self.flat.grad
        if world_size > 1 and sync:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            g.div_(world_size)
        self.flat.grad = g
        for p in self.params:
            p.grad = None
        return g


class FlatAdamW:
    """clip_grad_norm_(max_norm) + torch.optim.AdamW.step() (train_ScanNet_DDP_WarmUP.py:421-424) on the ONE flat buffer of
    FlatParameters as two kernel launches (csrc/glue.cu: gradient norm partials, then clip + moments + update), replacing
    torch's norm / clamp / mul / fused-Adam / step-increment sequence.  The learning rate and the step counter are device
    scalars, so a step captured in a CUDA graph follows a learning-rate schedule: set_lr() between replays."""

    def __init__(self, flat, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=0.0):
        from ._lib import lib
        self.flat = flat.flat
        p = self.flat.data
        if not p.is_cuda or p.dtype != torch.float32:
            raise RuntimeError("FlatAdamW needs a CUDA float32 parameter buffer (no CPU path)")
        self.exp_avg = torch.zeros_like(p)
        self.exp_avg_sq = torch.zeros_like(p)
        self.step_t = torch.zeros((), device=p.device, dtype=torch.float32)
        self.lr_t = torch.full((), float(lr), device=p.device, dtype=torch.float32)
        self.norm_t = torch.zeros((), device=p.device, dtype=torch.float32)
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, eps, weight_decay, max_norm
        self._ws = torch.empty(int(lib().pcfb_adamw_workspace()), dtype=torch.uint8, device=p.device)

    def set_lr(self, lr):
        self.lr_t.fill_(float(lr))

    def state_tensors(self):
        return [self.exp_avg, self.exp_avg_sq, self.step_t, self.lr_t]

    def zero_grad(self, set_to_none=True):
        self.flat.grad = None

    @torch.no_grad()
    def step(self, grad=None):
        from ._lib import check, lib, ptr, stream_ptr
        g = self.flat.grad if grad is None else grad
        if g is None:
            raise RuntimeError("FlatAdamW.step: no gradient (call FlatParameters.gather_grads first)")
        g = g.contiguous()
        p = self.flat.data
        check(lib().pcfb_adamw_clip_step(ptr(p), ptr(g), ptr(self.exp_avg), ptr(self.exp_avg_sq), p.numel(), ptr(self.lr_t),
                                         ptr(self.step_t), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                         float(self.weight_decay), float(self.max_norm), ptr(self.norm_t), ptr(self._ws),
                                         self._ws.numel(), stream_ptr()), "adamw_clip_step")
        from . import streams
        streams.weights_updated()                       # weight preparations of the next step wait for this update
        return self.norm_t
