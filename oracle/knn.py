"""CPU oracle: brute-force kNN + packed edge construction.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates
  * knn_keops            /root/reference/knn_post_dataloader_utils.py:22-41
  * compute_knn          /root/reference/knn_post_dataloader_utils.py:43-87
  * compute_knn_packed   /root/reference/knn_post_dataloader_utils.py:171-223
  * listToBatch/prepare  /root/reference/knn_post_dataloader_utils.py:113-167

PARITY UNPINNED at the KeOps boundary (pykeops is a third-party dependency that is not vendored,
not version-pinned and not installed here).  Pinned arithmetic of this restatement:
    d(q, r) = ((qx-rx)*(qx-rx) + (qy-ry)*(qy-ry)) + (qz-rz)*(qz-rz)      fp32, no FMA contraction
    result  = the K references with the smallest (d, index) in lexicographic order, ascending
which is KeOps' published formula ``((x_i - y_j)**2).sum(-1).argKmin(K)`` with ties resolved to
the lowest reference index (what a sequential strict-< scan gives).

What IS pinned: tests/golden/knn_packed.npz holds tables produced by the reference's own, unmodified
compute_knn / compute_knn_packed / listToBatch / prepare with the kNN itself computed by the reference's
sklearn KDTree option (``compute_knn(method='sklearn')``, lines 68-72) on tie-free clouds
(tests/golden/make_golden.py::make_knn, oracle/ref_shim.load_knn_utils); this oracle reproduces them exactly
(tests/test_oracle_golden.py::test_knn_packed_matches_reference) -- the scene x level loop, the offsets and
the neighbour order wherever float32 and float64 distances order the candidates alike.  Tie order and
float32 rounding of the KeOps reduction itself stay unpinned.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_clib = None


def _c_oracle():
    """The plain-C restatement (oracle/pcf_oracle.c), used for sizes numpy cannot hold."""
    global _clib
    if _clib is None:
        path = os.path.join(_HERE, "libpcf_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libpcf_oracle.so missing -- run `make -C oracle` "
                               "(or __graft_entry__.build())")
        _clib = ctypes.CDLL(path)
        _clib.oracle_knn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        _clib.oracle_knn.restype = ctypes.c_int
    return _clib


def sqdist_f32(query, ref):
    """[Nq,3] x [Nr,3] -> [Nq,Nr] fp32 squared distances with the pinned operation order."""
    q = np.ascontiguousarray(query, dtype=np.float32)
    r = np.ascontiguousarray(ref, dtype=np.float32)
    dx = q[:, None, 0] - r[None, :, 0]
    dy = q[:, None, 1] - r[None, :, 1]
    dz = q[:, None, 2] - r[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def knn_numpy(ref, query, K, chunk=512):
    """knn_keops restatement (knn_post_dataloader_utils.py:22-41): [Nq,K] int64, ascending (d, idx)."""
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    nq = query.shape[0]
    out = np.empty((nq, K), dtype=np.int64)
    for s in range(0, nq, chunk):
        d = sqdist_f32(query[s:s + chunk], ref)
        out[s:s + chunk] = np.argsort(d, axis=1, kind="stable")[:, :K]
    return out


def knn_c(ref, query, K, threads=0):
    """Same result as knn_numpy through oracle/pcf_oracle.c (OpenMP over queries)."""
    ref = np.ascontiguousarray(ref, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    out = np.empty((query.shape[0], K), dtype=np.int64)
    rc = _c_oracle().oracle_knn(ref.ctypes.data, ref.shape[0], query.ctypes.data, query.shape[0],
                                int(K), out.ctypes.data, int(threads))
    if rc != 0:
        raise RuntimeError("oracle_knn failed rc=%d" % rc)
    return out


def compute_knn(ref, query, K, use_c=None):
    """compute_knn (knn_post_dataloader_utils.py:43-87), method='keops', dilated_rate=1.

    The reference's `n_ref < K` branch draws *random* indices (np.random.choice, lines 58-66) and
    is not reproducible; the product special-cases it deterministically as "the n_ref neighbours
    in ascending (d, idx) order, repeated cyclically", and so does this oracle.
    """
    ref = np.asarray(ref, dtype=np.float32)
    query = np.asarray(query, dtype=np.float32)
    n_ref = ref.shape[0]
    kk = min(K, n_ref)
    if use_c is None:
        use_c = ref.shape[0] * query.shape[0] > (1 << 24) and os.path.exists(
            os.path.join(_HERE, "libpcf_oracle.so"))
    idx = knn_c(ref, query, kk) if use_c else knn_numpy(ref, query, kk)
    if kk < K:
        idx = idx[:, np.arange(K) % kk]
    return idx


def compute_knn_packed(pointclouds, points_stored, K_self, K_forward, K_propagate, use_c=None):
    """compute_knn_packed + prepare (knn_post_dataloader_utils.py:156-223) in one go.

    pointclouds: list over levels of [1, sum_N_l, 3] (or [sum_N_l, 3]) float arrays (packed scenes)
    points_stored: list over levels of per-scene point counts
    returns (edges_self, edges_forward, edges_propagate): lists of [1, sum_N, K] int64 arrays whose
    values index the packed *reference* level (offsets as in listToBatch, lines 127-152):
      edges_self[l]      : queries level l,   refs level l
      edges_forward[l]   : queries level l+1, refs level l      (l = 0..L-2)
      edges_propagate[l] : queries level l,   refs level l+1    (l = 0..L-2)
    """
    L = len(pointclouds)
    pcs = [np.asarray(p, dtype=np.float32).reshape(-1, 3) for p in pointclouds]
    offs = [np.concatenate([[0], np.cumsum(ps)]).astype(np.int64) for ps in points_stored]
    S = len(points_stored[0])
    e_self = [[] for _ in range(L)]
    e_fwd = [[] for _ in range(L - 1)]
    e_prop = [[] for _ in range(L - 1)]
    for i in range(S):
        pts = [pcs[j][offs[j][i]:offs[j][i + 1]] for j in range(L)]
        for j in range(L):
            e_self[j].append(compute_knn(pts[j], pts[j], K_self[j], use_c) + offs[j][i])
            if j >= 1:
                e_fwd[j - 1].append(compute_knn(pts[j - 1], pts[j], K_forward[j], use_c) + offs[j - 1][i])
                e_prop[j - 1].append(compute_knn(pts[j], pts[j - 1], K_propagate[j], use_c) + offs[j][i])
    cat = lambda lst: np.concatenate(lst, axis=0)[None]
    return [cat(x) for x in e_self], [cat(x) for x in e_fwd], [cat(x) for x in e_prop]
