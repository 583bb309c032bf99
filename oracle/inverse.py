"""CPU oracle: kNN inverse map (CSR transpose of an edge table).  TEST INFRASTRUCTURE.

Restates create_inverse_python (/root/reference/cpp_wrappers/cpp_pcf_kernel/test_kernels.py:177-213)
with the dtypes of knn_inverse_cuda_forward (/root/reference/cpp_wrappers/cpp_pcf_kernel/src/knn.cu:
104-168: inv_neighbors int32 [B, N*K], inv_k uint8 [B, N*K], inv_idx int32 [B, total+1]) and the
call pattern of compute_knn_inverse (/root/reference/util/common_util.py:250-327).

Canonical order inside a segment: ascending output point n, then k -- the order the reference's
Python restatement produces; the reference CUDA kernel fills segments in atomic (nondeterministic)
order and its own test compares after sorting (test_kernels.py:477-512), so the canonical order is a
valid (and the only reproducible) representative.
"""
import numpy as np


def knn_inverse(nei, total_points):
    """nei: [N_out, K] (or [1, N_out, K]) integer array -> (inv_neighbors i32, inv_k u8, inv_idx i32).

    Entries outside [0, total_points) are ignored (knn.cu:38,75); the unused tail of
    inv_neighbors / inv_k (length N_out*K) stays zero as torch::zeros leaves it (knn.cu:113-114).
    """
    nei = np.asarray(nei)
    if nei.ndim == 3:
        assert nei.shape[0] == 1
        nei = nei[0]
    n_out, K = nei.shape
    assert K <= 255
    flat = nei.reshape(-1).astype(np.int64)
    valid = (flat >= 0) & (flat < total_points)
    e = np.nonzero(valid)[0]
    order = e[np.argsort(flat[e], kind="stable")]
    inv_neighbors = np.zeros(n_out * K, dtype=np.int32)
    inv_k = np.zeros(n_out * K, dtype=np.uint8)
    inv_neighbors[:order.size] = (order // K).astype(np.int32)
    inv_k[:order.size] = (order % K).astype(np.uint8)
    inv_idx = np.zeros(total_points + 1, dtype=np.int32)
    np.cumsum(np.bincount(flat[e], minlength=total_points), out=inv_idx[1:])
    return inv_neighbors, inv_k, inv_idx


def compute_knn_inverse(pointclouds, edges_self, edges_forward, edges_propagate):
    """compute_knn_inverse (util/common_util.py:250-327): total_points = pointclouds[j].shape[1]
    for all three edge kinds (for propagate that is the dense level: harmless padding)."""
    def run(edge_list):
        res = ([], [], [])
        for j, e in enumerate(edge_list):
            out = knn_inverse(e, np.asarray(pointclouds[j]).reshape(-1, 3).shape[0])
            for lst, o in zip(res, out):
                lst.append(o[None])
        return [res[0], res[1], res[2]]
    return run(edges_self), run(edges_forward), run(edges_propagate)
