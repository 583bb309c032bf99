"""Helpers shared by the -m gpu parity tests (CUDA path vs the CPU oracle)."""
import numpy as np
import torch

import pcf_b200  # noqa: F401  (import alias for ml-pointconvformer_b200)


def cuda(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.cuda()


def surface_cloud(n, seed, extent=(4.0, 3.0, 2.5)):
    """Same generator as tests/golden/make_golden.py (kept in sync by test_host_cpu)."""
    rng = np.random.default_rng(seed)
    ex = np.asarray(extent, np.float32)
    face = rng.integers(0, 5, n)
    p = rng.random((n, 3)).astype(np.float32) * ex
    nrm = np.zeros((n, 3), np.float32)
    for f, (ax, val, sgn) in enumerate([(2, 0.0, 1), (0, 0.0, 1), (0, ex[0], -1), (1, 0.0, 1), (1, ex[1], -1)]):
        m = face == f
        p[m, ax] = val
        nrm[m, ax] = sgn
    p += rng.normal(0, 0.005, p.shape).astype(np.float32)
    nrm += rng.normal(0, 0.05, nrm.shape).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    return p.astype(np.float32), nrm.astype(np.float32)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def max_err_scaled(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
