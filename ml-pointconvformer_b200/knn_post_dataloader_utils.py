"""post_knn edge construction on the GPU with the reference's interface
(/root/reference/knn_post_dataloader_utils.py): compute_knn (43-87), compute_knn_packed (171-223),
prepare / listToBatch / tensorize (89-167).

The reference loops scenes x levels x {self, forward, propagate} in Python and hands each slice to
pykeops (13 KeOps reductions per scene), then re-concatenates the per-scene tables with running
offsets.  Here ONE kernel launch per edge set handles every scene of the packed cloud and writes the
offset (packed) indices directly, so `prepare()` only has to add the batch dimension.  To keep the
reference's two-step call pattern (train_ScanNet_DDP_WarmUP.py:382-383) `compute_knn_packed` returns
lists holding a single "scene" entry -- the already-packed table -- which `prepare` passes through.
"""
import numpy as np
import torch

from . import pcf_cuda
from . import streams as S


def _as_cuda_xyz(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("pcf_b200 kNN needs a CUDA device (no CPU fallback)")
        x = x.cuda(non_blocking=True)
    return x.float().contiguous()


def compute_knn(ref_points, query_points, K, dilated_rate=1, method='keops'):
    """compute_knn (43-87): [N_query, K] int64 nearest references, ascending (distance, index).
    `method` is accepted for signature compatibility (every method maps to the exact brute-force kernel).
    If n_ref < K the found neighbours repeat cyclically (the reference draws random indices, 58-66)."""
    if dilated_rate != 1:
        raise NotImplementedError("dilated_rate > 1 is unused by every caller and broken in the reference "
                                  "(knn_post_dataloader_utils.py:81-86)")
    ref = _as_cuda_xyz(ref_points)
    qry = _as_cuda_xyz(query_points)
    if KNN_METHOD == "grid" and K <= 64 and method != "brute":
        return pcf_cuda.KnnGrid(ref, [ref.shape[0]]).query(qry, [qry.shape[0]], K)
    return pcf_cuda.knn_packed(ref, [ref.shape[0]], qry, [qry.shape[0]], K)


# 'grid' (default): exact uniform-grid search; 'brute': the shared-memory tiled brute-force kernel.  Both return
# identical tables (tests/test_gpu_knn.py); brute force is O(N^2) per scene.
KNN_METHOD = "grid"
BRUTE_MAX_REFS = 256           # per-scene reference count below which brute force beats the grid search
CELL_FACTOR = 1.75             # grid cell edge = CELL_FACTOR x the level's voxel size (scripts/time_knn_stage.py)


def compute_knn_packed(pointclouds, points_stored, K_self, K_forward, K_propagate, grid_size=None, method=None):
    """compute_knn_packed (171-223).  pointclouds: list over levels of [1, sum N_l, 3]; points_stored: list
    over levels of per-scene counts.  Returns (nei_self_list, nei_forward_list, nei_propagate_list) in the
    reference's nesting [scene][level]; here a single pseudo-scene carries the packed, offset tables.
    grid_size (optional, cfg.grid_size): per-level voxel size, used as a cell-size hint by the grid search."""
    method = method or KNN_METHOD
    L = len(pointclouds)
    pcs = [_as_cuda_xyz(p.reshape(-1, 3) if isinstance(p, np.ndarray) else p.reshape(-1, 3)) for p in pointclouds]
    counts = [list(map(int, ps)) for ps in points_stored]
    kmax = max(list(K_self) + list(K_forward) + list(K_propagate))
    e_self, e_fwd, e_prop = [], [], []
    if method == "brute" or kmax > 64:
        for j in range(L):
            e_self.append(pcf_cuda.knn_packed(pcs[j], counts[j], pcs[j], counts[j], K_self[j]))
            if j >= 1:
                e_fwd.append(pcf_cuda.knn_packed(pcs[j - 1], counts[j - 1], pcs[j], counts[j], K_forward[j]))
                e_prop.append(pcf_cuda.knn_packed(pcs[j], counts[j], pcs[j - 1], counts[j - 1], K_propagate[j]))
        return [e_self], [e_fwd], [e_prop]
    # Tiny levels (every scene <= BRUTE_MAX_REFS reference points) use the tiled brute-force kernel; above that the grid
    # search wins even when it is latency bound (measured per query set, grid vs brute force: 128 vs 232 us at 1 k
    # references, 185 vs 287 us at 5 k, 105 vs 84 us at 184) -- same table either way.
    small = [max(counts[j]) <= BRUTE_MAX_REFS for j in range(L)]
    hint = lambda j: CELL_FACTOR * float(grid_size[j]) if grid_size is not None else 0.0

    # Two fork / join stages (streams.fork, the wide pool: one stream per job): the five grid builds, then the 13 query
    # sets.  The small sets go first: a 1 k .. 5 k-query set is a handful of CTAs that take ~100 us of pure latency, so they
    # must be resident before the two 100k-query sets (level-0 self, level-0 -> level-1 propagate) fill every CTA slot --
    # queued behind them they would run after the big ones and double the stage (scripts/time_knn_stage.py).
    builds = [S.fork(lambda j=j: None if small[j] else pcf_cuda.KnnGrid(pcs[j], counts[j], hint(j)), j, wide=True) for j in range(L)]
    grids = [S.join(b) for b in builds]
    jobs = []                                                        # (kind, level index in the output list, jr, jq, K)
    for j in range(L):
        jobs.append(("self", j, j, j, K_self[j]))
        if j >= 1:
            jobs.append(("fwd", j - 1, j - 1, j, K_forward[j]))      # level j looks into level j-1
            jobs.append(("prop", j - 1, j, j - 1, K_propagate[j]))   # dense level j-1 looks into level j
    jobs.sort(key=lambda t: pcs[t[3]].shape[0])

    def query(jr, jq, K):
        if grids[jr] is None:
            return pcf_cuda.knn_packed(pcs[jr], counts[jr], pcs[jq], counts[jq], K)
        # the query cloud's own grid (if it has one) supplies a spatially coherent order of the queries
        return grids[jr].query(pcs[jq], counts[jq], K, order=grids[jq] if jq != jr else None)
    running = [(job, S.fork(lambda job=job: query(job[2], job[3], job[4]), i, wide=True)) for i, job in enumerate(jobs)]
    out = {"self": [None] * L, "fwd": [None] * (L - 1), "prop": [None] * (L - 1)}
    for job, br in running:
        out[job[0]][job[1]] = S.join(br)
    return [out["self"]], [out["fwd"]], [out["prop"]]


def tensorizeTensorList(tensor_list):
    return [None if t is None else t.unsqueeze(0) for t in tensor_list]


def tensorize(edges_self, edges_forward, edges_propagate):
    return tensorizeTensorList(edges_self), tensorizeTensorList(edges_forward), tensorizeTensorList(edges_propagate)


def listToBatch(edges_self, edges_forward, edges_propagate):
    """listToBatch (113-154): per-scene tables -> one packed table per level, running offsets added and -1
    kept.  Tables coming from compute_knn_packed above are already packed (one pseudo-scene) and pass
    through; genuinely per-scene lists (e.g. built with compute_knn) are offset + concatenated here."""
    def as_t(x):
        return torch.from_numpy(x) if isinstance(x, np.ndarray) else x
    n = len(edges_self)
    if n == 1:
        return ([as_t(x) for x in edges_self[0]], [as_t(x) for x in edges_forward[0]],
                [as_t(x) for x in edges_propagate[0]])
    L = len(edges_self[0])
    stored = [0] * L
    outs, outf, outp = [[] for _ in range(L)], [[] for _ in range(L - 1)], [[] for _ in range(L - 1)]

    def shifted(t, off):
        t = as_t(t)
        return torch.where(t == -1, t, t + off)
    for i in range(n):
        for j in range(L - 1):
            outf[j].append(shifted(edges_forward[i][j], stored[j]))
            outp[j].append(shifted(edges_propagate[i][j], stored[j + 1]))
        for j in range(L):
            outs[j].append(shifted(edges_self[i][j], stored[j]))
        for j in range(L):
            stored[j] += edges_self[i][j].shape[0]
    cat = lambda lst: torch.cat(lst, dim=0)
    return [cat(x) for x in outs], [cat(x) for x in outf], [cat(x) for x in outp]


def prepare(edges_self, edges_forward, edges_propagate):
    """prepare (156-167) -> three lists of [1, sum N, K] int64 tensors."""
    return tensorize(*listToBatch(edges_self, edges_forward, edges_propagate))
