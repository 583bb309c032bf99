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


def run(M, N, K, split, swap=False, desc_or=0):
    from pcf_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    B = torch.randn(N, K, generator=g).cuda()
    raw = torch.zeros(128, N, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    lbo_a, lbo_b, sbo = M * 16, N * 16, 128
    fill = [lbo_a, sbo, lbo_b, sbo]
    desc = [sbo, lbo_a, sbo, lbo_b] if swap else [lbo_a, sbo, lbo_b, sbo]
    params = (ctypes.c_uint32 * 11)(*(fill + desc + [2 * lbo_a, 2 * lbo_b, idesc_tf32(M, N)]))
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
