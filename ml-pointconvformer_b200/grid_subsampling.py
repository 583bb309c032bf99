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


def build_pyramid(coord, norm, counts, grid_size, expect=None, boxes=None):
    """Level-0 clouds -> the whole pyramid on the device (datasetCommon.py:384-420 per scene and level on the CPU).
    All levels are enqueued back to back with device-side sizes (pcfb_pyramid_level); the host reads the per-level,
    per-scene counts ONCE at the end (plus one read of the level-0 bounding boxes that bound the voxel tables).
    expect / boxes: the counts per level and the boxes of an earlier call on the same batch -- then nothing is read back
    at all (a training step captured in a CUDA graph; a mismatch is left in the returned status word).
    -> (points per level, normals per level, counts per level, info dict(boxes, status))."""
    p0, n0 = _cuda(coord), _cuda(norm)
    counts = list(map(int, counts))
    n_seg = len(counts)
    if boxes is None:
        boxes = pcf_cuda._cells_upper_bound(p0, counts, grid_size[0])
    dev = p0.device
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    seg_off = pcf_cuda._offsets(counts, dev)
    pts, nrm, offs = [p0], [n0], [seg_off]
    for g in grid_size[1:]:
        sp, sn, so = pcf_cuda.pyramid_level(pts[-1], nrm[-1], offs[-1], n_seg, g, boxes, status)
        pts.append(sp); nrm.append(sn); offs.append(so)
    if expect is None:
        host = torch.stack(offs[1:] + [status.expand(n_seg + 1)]).cpu() if len(offs) > 1 else None      # the one host read
        cnt = [counts]
        if host is not None:
            if int(host[-1, 0]) != 0:
                raise RuntimeError("build_pyramid: the voxel table bound was exceeded (boxes do not contain the cloud)")
            for l in range(len(offs) - 1):
                cnt.append((host[l, 1:] - host[l, :-1]).tolist())
    else:
        cnt = [list(map(int, c)) for c in expect]
    if any(min(c) <= 16 for c in cnt[1:]):
        # a scene that collapses keeps its previous level (datasetCommon.py:413-414): rare, handled by the per-level path
        p, n, c = subsample_packed(p0, n0, counts, grid_size)
        return p, n, c, dict(boxes=boxes, status=status)
    pts = [p[:sum(c)] for p, c in zip(pts, cnt)]
    nrm = [x[:sum(c)] for x, c in zip(nrm, cnt)]
    return pts, nrm, cnt, dict(boxes=boxes, status=status)


def voxelize_packed(coord, counts, voxel_size, boxes=None):
    """voxelize(..., hash_type='ravel', mode='deterministic') (util/voxelize.py:44-70) for every scene of a packed cloud:
    -> (idx_unique int64 [M] packed indices in ascending (scene, key) order, per-scene counts).  One host read (the counts)."""
    p = _cuda(coord)
    idx, off, status, boxes = pcf_cuda.voxelize(p, list(map(int, counts)), voxel_size, boxes)
    host = torch.cat([off, status]).cpu()
    if int(host[-1]) != 0:
        raise RuntimeError("voxelize_packed: the voxel table bound was exceeded")
    off_h = host[:-1]
    return idx[:int(off_h[-1])].to(torch.int64), (off_h[1:] - off_h[:-1]).tolist()
