"""CUDA-event timings of the two dense products of the fused backward at the model's shapes (L2 flushed per launch)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcf_b200  # noqa
from pcf_b200 import pcf_cuda

WARM = "--warm" in sys.argv          # operands left in L2 between launches (how the products run inside a step)


def t(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if not WARM:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

for (M, C_out, KK) in [(102095, 32, 512), (102095, 64, 288), (23594, 64, 512), (5153, 96, 768), (1026, 128, 1024), (184, 192, 1536)]:
    dy = torch.randn(M, C_out, device="cuda"); W = torch.randn(C_out, KK, device="cuda"); P = torch.randn(M, KK, device="cuda")
    nt = t(lambda: pcf_cuda.gemm_nt(dy, W, None, w_is_kn=True))
    tn = t(lambda: pcf_cuda.gemm_tn(dy, P, want_rowsum=True))
    fw = t(lambda: pcf_cuda.gemm_nt(P, W))
    gb = 1e-6
    print("M=%6d C_out=%3d KK=%4d | dP=dY W %.3f ms (%.0f GB/s) | dW=dY^T P %.3f ms (%.0f GB/s) | Y=P W^T %.3f ms (%.0f GB/s)" % (
        M, C_out, KK, nt, (M * (C_out + KK) * 4) * gb / nt, tn, (M * (C_out + KK) * 4) * gb / tn, fw, (M * (C_out + KK) * 4) * gb / fw))
