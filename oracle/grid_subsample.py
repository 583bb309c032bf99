"""CPU oracle: grid (voxel) subsampling with barycentres.  TEST INFRASTRUCTURE.

Restates grid_subsampling() (/root/reference/cpp_wrappers/cpp_subsampling/grid_subsampling/
grid_subsampling.cpp:9-110, SampledData grid_subsampling.h:13-83, min/max_point cloud.cpp:27-66) and
subsample() (/root/reference/datasetCommon.py:384-420).  PINNED against the reference C++ compiled
from its own sources (oracle/Makefile `make ref` -> oracle/_ref/libgridsub_ref.so).

Arithmetic (all fp32 unless noted):  origin = floor(min * (1/dl)) * dl;  i = floor((p - origin)/dl);
key = iX + NX*iY + NX*NY*iZ;  per-voxel sums accumulate sequentially in input order;
barycentre = sum * float(1.0/count) (double reciprocal rounded to float, grid_subsampling.cpp:91);
feature mean = sum / float(count) (grid_subsampling.cpp:94-98).
The reference emits voxels in libstdc++ unordered_map iteration order (not reproducible); the
canonical order here -- and of the CUDA product -- is ascending voxel key.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def voxel_keys(points, dl):
    p = np.ascontiguousarray(points, dtype=f32)
    dl = f32(dl)
    mn, mx = p.min(0), p.max(0)
    origin = np.floor(mn * (f32(1) / dl)) * dl
    nx = np.uint64(np.floor((mx[0] - origin[0]) / dl)) + np.uint64(1)
    ny = np.uint64(np.floor((mx[1] - origin[1]) / dl)) + np.uint64(1)
    ijk = np.floor((p - origin[None]) / dl).astype(np.uint64)
    return ijk[:, 0] + nx * ijk[:, 1] + nx * ny * ijk[:, 2]


def grid_subsample(points, features, dl):
    """-> (sub_points [M,3], sub_features [M,F], keys [M] ascending, counts [M])."""
    p = np.ascontiguousarray(points, dtype=f32)
    f = np.ascontiguousarray(features, dtype=f32) if features is not None else np.zeros((len(p), 0), f32)
    keys = voxel_keys(p, dl)
    order = np.argsort(keys, kind="stable")            # input order preserved inside a voxel
    ks = keys[order]
    start = np.nonzero(np.r_[True, ks[1:] != ks[:-1]])[0]
    counts = np.diff(np.r_[start, len(ks)])
    data = np.concatenate([p, f], axis=1)[order]
    acc = np.zeros((len(start), data.shape[1]), dtype=f32)
    for r in range(int(counts.max())):                   # sequential fp32 accumulation, round r
        sel = counts > r
        acc[sel] = acc[sel] + data[start[sel] + r]
    inv = (1.0 / counts.astype(np.float64)).astype(f32)
    sub_p = acc[:, :3] * inv[:, None]
    sub_f = acc[:, 3:] / counts.astype(f32)[:, None]
    return sub_p, sub_f, ks[start], counts


def subsample(coord, norm, grid_size):
    """subsample() (datasetCommon.py:384-420): level 0 = input; level j = grid_subsampling of level
    j-1 at grid_size[j]; if a level would have <= 16 points the previous level is kept (413-414)."""
    pts, nrm = [np.asarray(coord, f32)], [np.asarray(norm, f32)]
    for g in grid_size[1:]:
        sp, sn, _, _ = grid_subsample(pts[-1], nrm[-1], g)
        if sp.shape[0] <= 16:
            sp, sn = pts[-1], nrm[-1]
        pts.append(sp)
        nrm.append(sn)
    return pts, nrm


def reference_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libgridsub_ref.so"))


def grid_subsample_reference(points, features, dl):
    """Runs the UNMODIFIED reference C++ (oracle/_ref/libgridsub_ref.so); output in the reference's
    own (hash-map) order."""
    lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libgridsub_ref.so"))
    lib.gridsub_ref.restype = ctypes.c_int
    lib.gridsub_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    p = np.ascontiguousarray(points, dtype=f32)
    f = np.ascontiguousarray(features, dtype=f32)
    n, fd = p.shape[0], f.shape[1]
    op = np.empty((n, 3), f32)
    of = np.empty((n, max(fd, 1)), f32)
    m = lib.gridsub_ref(p.ctypes.data, f.ctypes.data, n, fd, float(f32(dl)), op.ctypes.data, of.ctypes.data, n)
    assert m >= 0
    return op[:m].copy(), of[:m, :fd].copy()
