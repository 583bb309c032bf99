"""Hardware check of the tcgen05 conventions the fused forward relies on (include/pcf_b200.h:
pcfb_selftest_umma): K-major no-swizzle core-matrix operands (LBO = stride between 16-byte K chunks, SBO =
stride between 8-row groups), K=8 step = +2*LBO, and the accumulator layout in TMEM (M=128: lane = row;
M=64: lane = (row % 16) + 32 * (row / 16))."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def idesc_tf32(M, N):
    return (1 << 4) | (2 << 7) | (2 << 10) | ((N >> 3) << 17) | ((M >> 4) << 24)


def run(M, N, K, split, swap=False, desc_or=0, a_mn=False, b_mn=False):
    from pcf_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    raw = torch.zeros(128, N, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    # K-major: LBO = stride between 16-byte K chunks (rows*16), SBO = 8-row group stride (128)
    # MN-major: 4 rows share a 16-byte unit; 8 k's x 16 B = one 128-byte core (LBO = next 8 k's),
    #           SBO = stride between 4-row units.  Here units are laid out [row/4][K][4] -> SBO = K*16.
    la, sa, ka = (128, K * 16, 128) if a_mn else (M * 16, 128, 2 * M * 16)
    lb, sb, kb = (128, K * 16, 128) if b_mn else (N * 16, 128, 2 * N * 16)
    fill = [la, sa, lb, sb]
    desc = [sa, la, sb, lb] if swap else [la, sa, lb, sb]
    idesc = idesc_tf32(M, N) | ((1 << 15) if a_mn else 0) | ((1 << 16) if b_mn else 0)
    params = (ctypes.c_uint32 * 12)(*(fill + desc + [ka, kb, idesc, (1 if a_mn else 0) | (2 if b_mn else 0)]))
    _lib.check(_lib.lib().pcfb_selftest_umma(A.data_ptr(), B.data_ptr(), raw.data_ptr(), M, N, K,
                                             ctypes.cast(params, ctypes.c_void_p), desc_or, split, status.data_ptr(),
                                             _lib.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    assert int(status.item()) == 0, "tcgen05.mma never completed"
    want = A.double() @ B.double().t()
    return raw.cpu().double(), want.cpu()


def rows_of(raw, M):
    if M == 128:
        return raw
    lanes = [(r % 16) + 32 * (r // 16) for r in range(64)]
    return raw[lanes]


@pytest.mark.parametrize("M,N,K", [(128, 32, 32), (64, 32, 64), (64, 64, 16), (64, 192, 32), (128, 256, 8), (64, 8, 8)])
def test_umma_3xtf32_matches_fp64(M, N, K):
    raw, want = run(M, N, K, split=1)
    got = rows_of(raw, M)
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 2e-6, "3xTF32 error %.3e (M=%d N=%d K=%d)" % (err, M, N, K)


def test_umma_single_pass_tf32_is_tf32_accurate():
    raw, want = run(64, 32, 32, split=0)
    err = float((rows_of(raw, 64) - want).abs().max() / want.abs().max())
    assert 1e-6 < err < 5e-3, err           # tf32-sized error: proves the tensor path (not fp32 FMA) produced it
