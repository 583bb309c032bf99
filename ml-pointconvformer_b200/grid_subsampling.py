"""Grid subsampling on the GPU with the reference's interface: grid_subsampling()
(/root/reference/datasetCommon.py:17-67 -> cpp_wrappers/cpp_subsampling) and subsample()
(datasetCommon.py:384-420), plus a packed multi-scene variant."""
import numpy as np
import torch

from . import pcf_cuda


def _cuda(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return x.float().contiguous().cuda()


def grid_subsampling(points, features=None, labels=None, sampleDl=0.1, verbose=0):
    """Barycentre grid subsampling of one cloud.  Returns (sub_points[, sub_features]) as CUDA tensors in
    ascending voxel-key order (the reference's order is hash-map iteration order).  `labels` (majority
    vote in the reference) is not on the hot path and is rejected."""
    if labels is not None:
        raise NotImplementedError("label voting is not part of the B200 hot path")
    p = _cuda(points)
    f = None if features is None else _cuda(features)
    sp, sf, _ = pcf_cuda.grid_subsample(p, f, [p.shape[0]], sampleDl)
    return sp if f is None else (sp, sf)


def subsample(coord, norm, grid_size=[0.1]):
    """subsample() (datasetCommon.py:384-420): level 0 is the input, level j the grid subsampling of level
    j-1 at grid_size[j]; a level with <= 16 points keeps the previous level (413-414)."""
    pts, nrm = [_cuda(coord)], [_cuda(norm)]
    for g in grid_size[1:]:
        sp, sn = grid_subsampling(pts[-1], nrm[-1], sampleDl=g)
        if sp.shape[0] <= 16:
            sp, sn = pts[-1], nrm[-1]
        pts.append(sp)
        nrm.append(sn)
    return pts, nrm


def subsample_packed(coord, norm, counts, grid_size):
    """All scenes of a packed batch at once: -> (points per level, normals per level, counts per level)."""
    pts, nrm, cnt = [_cuda(coord)], [_cuda(norm)], [list(map(int, counts))]
    for g in grid_size[1:]:
        sp, sn, sc = pcf_cuda.grid_subsample(pts[-1], nrm[-1], cnt[-1], g)
        if min(sc) <= 16:
            # the reference keeps the previous level for a scene that collapses (per scene); with packed
            # scenes we apply it to the scenes concerned
            keep = [c <= 16 for c in sc]
            offs_prev = np.concatenate([[0], np.cumsum(cnt[-1])])
            offs_new = np.concatenate([[0], np.cumsum(sc)])
            P, N, C = [], [], []
            for s, k in enumerate(keep):
                if k:
                    P.append(pts[-1][offs_prev[s]:offs_prev[s + 1]]); N.append(nrm[-1][offs_prev[s]:offs_prev[s + 1]]); C.append(cnt[-1][s])
                else:
                    P.append(sp[offs_new[s]:offs_new[s + 1]]); N.append(sn[offs_new[s]:offs_new[s + 1]]); C.append(sc[s])
            sp, sn, sc = torch.cat(P), torch.cat(N), C
        pts.append(sp)
        nrm.append(sn)
        cnt.append(sc)
    return pts, nrm, cnt
