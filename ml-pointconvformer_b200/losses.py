"""Cross entropy of the training loop (/root/reference/train_ScanNet_DDP_WarmUP.py:417:
torch.nn.CrossEntropyLoss(weight, ignore_index=-100, label_smoothing=0.2)) as two kernels forward and one backward
(csrc/glue.cu) instead of torch's log_softmax / nll / smoothing chain.  Same semantics: mean over the valid rows weighted by
the target-class weight; the label-smoothing term spreads label_smoothing / C over all classes (class-weighted)."""
import torch

from ._lib import check, lib, ptr, require, stream_ptr, workspace

F32 = torch.float32


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, ignore_index, smoothing):
        require(logits, F32, "logits"); require(target, torch.int64, "target")
        N, C = logits.shape
        loss = torch.empty((), device=logits.device, dtype=F32)
        den = torch.empty((), device=logits.device, dtype=F32)
        ws_bytes = lib().pcfb_ce_workspace(N)
        ws = workspace(ws_bytes, logits.device)
        check(lib().pcfb_ce_forward(ptr(logits), ptr(target), ptr(weight), N, C, int(ignore_index), float(smoothing), ptr(loss),
                                    ptr(den), ptr(ws), ws_bytes, stream_ptr()), "ce_forward")
        ctx.save_for_backward(logits, target, weight, den)
        ctx.ignore_index, ctx.smoothing = int(ignore_index), float(smoothing)
        return loss

    @staticmethod
    def backward(ctx, grad):
        logits, target, weight, den = ctx.saved_tensors
        N, C = logits.shape
        grad = grad.to(F32).contiguous()
        d = torch.empty_like(logits)
        check(lib().pcfb_ce_backward(ptr(logits), ptr(target), ptr(weight), N, C, ctx.ignore_index, ctx.smoothing, ptr(den), ptr(grad),
                                     ptr(d), stream_ptr()), "ce_backward")
        return d, None, None, None, None


def cross_entropy(logits, target, weight=None, ignore_index=-100, label_smoothing=0.0):
    """F.cross_entropy(logits [N, C], target [N], weight, ignore_index=..., label_smoothing=..., reduction='mean') on the
    CUDA path (C <= 64).  CUDA float32 logits only."""
    if not logits.is_cuda:
        raise RuntimeError("pcf_b200.losses.cross_entropy needs CUDA tensors (no CPU path)")
    if weight is not None:
        weight = weight.to(device=logits.device, dtype=F32).contiguous()
    return _CrossEntropy.apply(logits.contiguous(), target.contiguous(), weight, ignore_index, label_smoothing)
