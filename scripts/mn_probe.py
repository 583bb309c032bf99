"""Probe MN-major operand conventions of tcgen05 on hardware (prints relative errors for several hypotheses)."""
import ctypes, sys, os, itertools
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcf_b200
from pcf_b200 import _lib

def idesc_tf32(M, N):
    return (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)

def run(M, N, K, b_mn, fill, desc, kstep_b, unit_rows):
    g = torch.Generator().manual_seed(1)
    A = torch.randn(M, K, generator=g).cuda(); B = torch.randn(N, K, generator=g).cuda()
    raw = torch.zeros(128, N, device="cuda"); status = torch.zeros(1, dtype=torch.int32, device="cuda")
    la, sa, ka = M * 16, 128, 2 * M * 16
    idesc = idesc_tf32(M, N) | ((1 << 16) if b_mn else 0)
    params = (ctypes.c_uint32 * 12)(la, sa, fill[0], fill[1], la, sa, desc[0], desc[1], ka, kstep_b, idesc, 2 if b_mn else 0)
    _lib.check(_lib.lib().pcfb_selftest_umma(A.data_ptr(), B.data_ptr(), raw.data_ptr(), M, N, K, ctypes.cast(params, ctypes.c_void_p), 0, 1, status.data_ptr(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = (A.double() @ B.double().t()).cpu()
    got = raw.cpu().double()
    if M == 64:
        got = got[[(r % 16) + 32 * (r // 16) for r in range(64)]]
    return float((got - want).abs().max() / want.abs().max()), int(status.item())

M, N, K = 64, 32, 32
# fill layout for MN-major B in the kernel: (r%4)*4 + (r/4)*fill_sbo + (k%8)*16 + (k/8)*fill_lbo
# hypothesis space: fill (lbo, sbo) ; desc (lbo, sbo) ; kstep
cands = []
for fl, fs in [(128, K * 16), (N * 16 * 2, 128)]:        # [r/4][k][4] units-of-4-rows major  |  [k/8][r/4][8][4] k-block major
    for dl, ds in [(fl, fs), (fs, fl)]:
        for ks in sorted({fl, fs, 128, 256}):
            cands.append(((fl, fs), (dl, ds), ks))
for fill, desc, ks in cands:
    try:
        e, st = run(M, N, K, True, fill, desc, ks, 4)
    except Exception as ex:
        e, st = -1, str(ex)[:40]
    print("fill(lbo,sbo)=%s desc(lbo,sbo)=%s kstep=%d -> err %.3e status %s" % (fill, desc, ks, e, st))
