"""GPU parity of the 3xTF32 tcgen05 GEMMs (pcfb_gemm_nt / pcfb_gemm_tn) against float64 matmul: they must be
fp32-accurate (the reference's 1e-4 bar with a wide margin), for the Linear shapes of the model, skinny per-edge
shapes, ragged M and unaligned leading dimensions."""
import pytest
import torch

from gpu_util import max_err_scaled

pytestmark = pytest.mark.gpu


def _pc():
    from pcf_b200 import pcf_cuda
    return pcf_cuda


@pytest.mark.parametrize("M,K,N", [(1000, 64, 128), (100001, 12, 8), (5000, 3, 16), (777, 512, 32), (4096, 1536, 192),
                                    (300, 416, 256), (50000, 64, 64), (1, 8, 8), (129, 33, 20), (2000, 384, 300),
                                    (23594, 64, 512), (1026, 128, 1024), (184, 192, 1536), (3001, 32, 1000),
                                    (1026, 1024, 128), (184, 1536, 192), (5153, 768, 96), (130, 40, 24), (2050, 2100, 48)])
@pytest.mark.parametrize("w_is_kn", [False, True])
def test_gemm_nt(M, K, N, w_is_kn):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(K, N, generator=g) if w_is_kn else torch.randn(N, K, generator=g)).cuda()
    b = torch.randn(N, generator=g).cuda()
    want = x.double() @ (w.double() if w_is_kn else w.double().t()) + b.double()
    got = _pc().gemm_nt(x, w, b, w_is_kn=w_is_kn)
    # 3xTF32 + fp32 accumulation: error grows like sqrt(K) * 2^-23; fp32 cuBLAS shows the same growth
    tol = 3e-6 * max(1.0, (K / 64.0) ** 0.5)
    err = max_err_scaled(got, want)
    ref_err = max_err_scaled(x @ (w if w_is_kn else w.t()) + b, want)
    print("gemm_nt M=%d K=%d N=%d: err %.2e (torch fp32 %.2e) tol %.1e" % (M, K, N, err, ref_err, tol))
    assert err < tol
    got = _pc().gemm_nt(x, w, None, w_is_kn=w_is_kn, act=2)
    want2 = torch.nn.functional.leaky_relu(want - b.double(), 0.1)
    assert max_err_scaled(got, want2) < tol


def test_gemm_nt_strided_rows():
    x = torch.randn(3000, 80, device="cuda")[:, 5:53]          # lda = 80, K = 48, unaligned start
    w = torch.randn(24, 48, device="cuda")
    assert max_err_scaled(_pc().gemm_nt(x, w), x.double() @ w.double().t()) < 3e-6


@pytest.mark.parametrize("M,N1,N2", [(5000, 32, 512), (100000, 64, 64), (777, 128, 96), (3000, 256, 416), (40, 20, 33),
                                      (20000, 8, 12), (1026, 192, 1536), (1, 16, 16)])
def test_gemm_tn(M, N1, N2):
    g = torch.Generator().manual_seed(M + N1 + N2)
    a = torch.randn(M, N1, generator=g).cuda()
    b = torch.randn(M, N2, generator=g).cuda()
    got, rs = _pc().gemm_tn(a, b, want_rowsum=True)
    tol = 3e-6 * max(1.0, (M / 4096.0) ** 0.5)
    err = max_err_scaled(got, a.double().t() @ b.double())
    print("gemm_tn M=%d N1=%d N2=%d: err %.2e tol %.1e" % (M, N1, N2, err, tol))
    assert err < tol
    assert max_err_scaled(rs, a.double().sum(0)) < tol
    got2, none = _pc().gemm_tn(a, b)
    assert none is None and torch.equal(got, got2)               # deterministic
