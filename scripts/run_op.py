"""Runs one hot op a few times on its BASELINE-size shape (for ncu captures): python scripts/run_op.py fwd|bwd|knn|inv [reps]"""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcf_b200
from pcf_b200 import pcf_cuda, synthetic

def main():
    what = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device("cuda", 0)
    if what in ("gemm_tn", "gemm_nt"):          # the dense products of the fused backward at the level-0 shape
        M, C_out, KK = 102095, 32, 512
        dy = torch.randn(M, C_out, device=dev); Wm = torch.randn(C_out, KK, device=dev); P = torch.randn(M, KK, device=dev)
        for _ in range(reps):
            if what == "gemm_tn":
                pcf_cuda.gemm_tn(dy, P, want_rowsum=True)
            else:
                pcf_cuda.gemm_nt(dy, Wm, None, w_is_kn=True)
        torch.cuda.synchronize()
        print("ok", what)
        return
    if what == "gemm_small":                    # Y = P W^T of a coarse level (latency-bound: few CTAs, long K chain)
        M, C_out, KK = (int(a) for a in (sys.argv[3].split(",") if len(sys.argv) > 3 else (1026, 128, 1024)))
        P = torch.randn(M, KK, device=dev); Wm = torch.randn(C_out, KK, device=dev)
        for _ in range(reps):
            pcf_cuda.gemm_nt(P, Wm)
        torch.cuda.synchronize()
        print("ok", what, M, C_out, KK)
        return
    xyz, _, _ = synthetic.make_scene(1, 100000)
    xyz = torch.from_numpy(xyz).to(dev)
    n = xyz.shape[0]
    K, C_in, C_add, C_mid, C_out = 16, 16, 16, 16, 32
    if len(sys.argv) > 3:
        C_in, C_add, C_mid, C_out = map(int, sys.argv[3].split(","))
    H = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    g = torch.Generator().manual_seed(0)
    grid = pcf_cuda.KnnGrid(xyz, [n], 0.25)
    nei = grid.query(xyz, [n], K)
    feats = torch.randn(1, n, C_in, generator=g).to(dev)
    w = torch.rand(1, n, K, C_mid, generator=g).to(dev)
    add = torch.rand(1, n, K, C_add, generator=g).to(dev) if C_add else None
    gd = torch.rand(1, n, K, H, generator=g).to(dev) if H else None
    W = (torch.randn(C_out, (C_in + C_add) * C_mid, generator=g) * 0.05).to(dev)
    b = torch.randn(C_out, generator=g).to(dev)
    go = torch.randn(1, n, C_out, generator=g).to(dev)
    inv = pcf_cuda.compute_knn_inverse(nei[None], n)
    y, p = pcf_cuda.pconv_fused_forward(feats, nei[None], w, add, gd, W, b, want_p=True)
    torch.cuda.synchronize()
    if what == "fwd_time":                      # CUDA-event timing of the fused forward (L2 flushed before every launch)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pcf_cuda.pconv_fused_forward(feats, nei[None], w, add, gd, W, b, want_p=False)
            b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        ts.sort()
        print("fwd_time ms median %.4f min %.4f" % (ts[len(ts) // 2], ts[0]))
        return
    for _ in range(reps):
        if what == "fwd":
            pcf_cuda.pconv_fused_forward(feats, nei[None], w, add, gd, W, b, want_p=False)
        elif what == "fwdp":                  # training-mode forward: P saved for the backward (the bench's roofline entry)
            pcf_cuda.pconv_fused_forward(feats, nei[None], w, add, gd, W, b, want_p=True)
        elif what == "bwd":
            pcf_cuda.pconv_fused_backward(go, None, feats, inv, nei[None], w, add, gd, W, p, (True,) * 6)
        elif what == "knn":
            pcf_cuda.KnnGrid(xyz, [n], 0.175).query(xyz, [n], K)
        elif what == "knn_brute":
            pcf_cuda.knn_packed(xyz, [n], xyz, [n], K)
        elif what == "inv":
            pcf_cuda.compute_knn_inverse(nei[None], n)
    torch.cuda.synchronize()
    print("ok", what, n)

if __name__ == "__main__":
    main()
