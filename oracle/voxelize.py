"""CPU restatement of the reference's voxelisation (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this).  Follows /root/reference/util/voxelize.py:26-70 -- ravel_hash_vec + voxelize(mode=
'deterministic'): discrete = floor(coord / voxel) (numpy promotes float32 / np.array(voxel) to float64), key = Fortran-style
ravel of discrete - min, one point per key.  The reference takes the FIRST entry of an unstable np.argsort among equal keys
(implementation defined); the canonical choice here and on the GPU is the smallest input index, i.e. a stable sort.
Pinned by tests/golden/voxelize.npz (the unmodified reference run on two clouds: same occupied voxels, same order, and on
every voxel a point of that voxel)."""
import numpy as np


def ravel_keys(coord, voxel):
    d = np.floor(np.asarray(coord, np.float32).astype(np.float64) / float(voxel))
    d -= d.min(0)
    d = d.astype(np.uint64)
    mx = d.max(0) + np.uint64(1)
    keys = np.zeros(len(d), np.uint64)
    for j in range(d.shape[1] - 1):
        keys += d[:, j]
        keys *= mx[j + 1]
    keys += d[:, -1]
    return keys


def voxelize(coord, voxel):
    """-> idx_unique (ascending key order, smallest index per voxel)."""
    keys = ravel_keys(coord, voxel)
    order = np.argsort(keys, kind="stable")
    ks = keys[order]
    first = np.ones(len(ks), bool)
    first[1:] = ks[1:] != ks[:-1]
    return order[first]


def voxelize_packed(coord, counts, voxel):
    out, cnt, off = [], [], 0
    for c in counts:
        idx = voxelize(coord[off:off + c], voxel)
        out.append(idx + off)
        cnt.append(len(idx))
        off += c
    return np.concatenate(out), cnt
