"""Fork / join of the independent branches of a layer onto side CUDA streams.

Inside one PointConvFormer layer the WeightNet chain, the positional-encoding chain and the shortcut branch do not depend
on each other until the contraction (/root/reference/layers.py:306-416: `weightnet`, `mlp_conv`, `unary_shortcut` are three
sub-graphs hanging off the same inputs).  On the coarse levels of the pyramid (5 k / 1 k / 184 points: 19 of PCF_Normal's 25
layers) every one of their kernels is a 2..65-CTA launch on a 148-SM GPU, so running the branches one after the other
leaves the machine idle behind a chain of ~6 us launches.  fork() runs a branch on a side stream (ordered after everything
already enqueued on the caller's stream), join() makes the caller's stream wait for it.  Autograd runs each backward node
on the stream of its forward, so the backward passes of the branches overlap the same way; inside a captured CUDA graph
the fork / join pairs become parallel branches of the graph.

Memory safety with the caching allocator: every branch starts with side.wait_stream(main) and is joined into main before
the caller goes on, results are record_stream()'ed on the consumer stream.
"""
import os

import torch

ENABLED = os.environ.get("PCFB_STREAMS", "1") != "0"
N_SIDE = 4
N_WIDE = 16                 # second pool for stages made of many independent jobs (the 13 kNN query sets of a batch)
_POOL = {}
_WIDE = {}
# Weight-gradient products (dW = dY^T X of every Linear, dW of the fused contraction's Linear) are LEAVES of the backward
# graph: nothing downstream waits for them except the optimizer.  With LEAF_ASYNC they run on their own stream and are only
# joined by join_leaves() (called by sharding.FlatParameters.gather_grads before the flat gradient is assembled), taking
# ~120 small GEMM launches off the backward's critical path.  Opt-in: whoever reads .grad must call join_leaves() first.
LEAF_ASYNC = False
N_LEAF = int(os.environ.get("PCFB_N_LEAF", "2"))                  # leaf streams, used round robin: the last (largest, level-0) products of the backward share the tail
_LEAF = {}
_LEAF_NEXT = [0]


def _side_streams(device):
    key = (device.type, device.index)
    pool = _POOL.get(key)
    if pool is None:
        pool = _POOL[key] = [torch.cuda.Stream(device=device) for _ in range(N_SIDE)]
    return pool


def _wide_streams(device):
    key = (device.type, device.index)
    pool = _WIDE.get(key)
    if pool is None:
        # the first half of the pool has high priority: whoever forks many jobs puts the SMALL ones there, so their few
        # CTAs are placed as soon as a slot frees up instead of queueing behind a big kernel's remaining waves
        pool = _WIDE[key] = [torch.cuda.Stream(device=device, priority=-1 if i < N_WIDE // 2 else 0) for i in range(N_WIDE)]
    return pool


def side_index(stream=None):
    """0 for the caller's main stream, 1..N_SIDE for the side streams, N_SIDE+1.. for the wide pool."""
    stream = stream or torch.cuda.current_stream()
    key = (stream.device.type, stream.device.index)
    for base, pool in ((1, _POOL.get(key)), (N_SIDE + 1, _WIDE.get(key))):
        if pool:
            for i, s in enumerate(pool):
                if s == stream:
                    return base + i
    return 0


class Branch:
    __slots__ = ("result", "stream", "main")

    def __init__(self, result, stream=None, main=None):
        self.result, self.stream, self.main = result, stream, main


def fork(fn, slot, enabled=True, wide=False):
    """Run fn() on side stream `slot` (inline when disabled, on the CPU, or when already on a side stream).
    wide: take the stream from the N_WIDE pool -- for stages of many independent jobs without BatchNorm exchanges."""
    if not (ENABLED and enabled) or not torch.cuda.is_available():
        return Branch(fn())
    main = torch.cuda.current_stream()
    if side_index(main) != 0:                       # no nested forks
        return Branch(fn())
    side = _wide_streams(main.device)[slot % N_WIDE] if wide else _side_streams(main.device)[slot % N_SIDE]
    side.wait_stream(main)
    with torch.cuda.stream(side):
        res = fn()
    return Branch(res, side, main)


def _record(x, stream):
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            x.record_stream(stream)
    elif isinstance(x, (list, tuple)):
        for y in x:
            _record(y, stream)


def join(branch):
    """The caller's stream waits for the branch; returns the branch's result."""
    if branch.stream is not None:
        branch.main.wait_stream(branch.stream)
        _record(branch.result, branch.main)
    return branch.result


def fork_leaf(fn, inputs=()):
    """Run fn() (a leaf of the backward graph: a weight gradient) on a leaf stream; NOT joined until join_leaves().
    inputs: the tensors fn reads -- they are recorded on the leaf stream, so the caching allocator does not hand their
    memory to a later allocation of the caller's stream while the leaf kernel may still be reading it (autograd frees the
    incoming gradient as soon as the node returns)."""
    if not (ENABLED and LEAF_ASYNC) or not torch.cuda.is_available():
        return fn()
    cur = torch.cuda.current_stream()
    key = (cur.device.type, cur.device.index)
    pool = _LEAF.get(key)
    if pool is None:
        pool = _LEAF[key] = [torch.cuda.Stream(device=cur.device) for _ in range(N_LEAF)]
    if cur in pool:
        return fn()
    leaf = pool[_LEAF_NEXT[0] % N_LEAF]
    _LEAF_NEXT[0] += 1
    leaf.wait_stream(cur)
    for t in inputs:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t.record_stream(leaf)
    with torch.cuda.stream(leaf):
        return fn()


def join_leaves():
    """The current stream waits for every weight-gradient product enqueued by fork_leaf()."""
    if not torch.cuda.is_available():
        return
    cur = torch.cuda.current_stream()
    for leaf in _LEAF.get((cur.device.type, cur.device.index), []):
        if leaf != cur:
            cur.wait_stream(leaf)
    _LEAF_NEXT[0] = 0


# ---- weight preparation stream ----------------------------------------------------------------------------------------
# The operand preparation of a dense Linear (gemm.cu: weights -> split tf32 in tensor-core order) depends on the weights only.
# With PREP_ASYNC (opt-in, like LEAF_ASYNC: sharding.FlatParameters(async_weight_grads=True) turns both on) it runs on its own
# stream.  That stream must be ordered after the optimizer step that last wrote the weights: whoever updates the weights
# calls weights_updated() (sharding.FlatAdamW.step does), and the first preparation after that waits for the updating stream.
PREP_ASYNC = False
_PREP = {}
_PREP_DIRTY = {}


def prep_stream(device):
    """The preparation stream for `device`, or None when weight preparations run inline."""
    if not (ENABLED and PREP_ASYNC) or device.type != "cuda":
        return None
    cur = torch.cuda.current_stream(device)
    key = (device.type, device.index)
    st = _PREP.get(key)
    if st is None:
        st = _PREP[key] = torch.cuda.Stream(device=device)
        _PREP_DIRTY[key] = True
    if st == cur:
        return None
    if _PREP_DIRTY.get(key, True):                       # first preparation since the weights changed
        st.wait_stream(cur)
        _PREP_DIRTY[key] = False
    elif torch.cuda.is_current_stream_capturing():       # a capture that began after the last update: pull the stream in
        with torch.cuda.stream(st):
            inside = torch.cuda.is_current_stream_capturing()
        if not inside:
            st.wait_stream(cur)
    return st


def weights_updated():
    """Call after the weights were written (optimizer step, load_state_dict): the next preparation waits for that work."""
    for key in list(_PREP_DIRTY):
        _PREP_DIRTY[key] = True
