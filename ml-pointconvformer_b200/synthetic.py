"""Synthetic ScanNet-shaped scenes (no dataset is available offline): a box room with axis-aligned
furniture boxes, points uniform on the faces with 5 mm jitter, face normals, random colours -- then
voxelised at the finest grid size (one point per voxel, as the reference's `voxelize(...,
mode='deterministic')`, util/voxelize.py:44-70, leaves it).  Host-side numpy; not on the hot path."""
import numpy as np


def make_room(seed, extent=(8.0, 6.0, 3.0), n_boxes=12, density=400.0):
    """-> (xyz [N,3] f32, normals [N,3] f32, colors [N,3] f32); density = samples per m^2 of surface."""
    rng = np.random.default_rng(seed)
    ex = np.asarray(extent, np.float64)
    faces = []        # (origin, u, v, normal)
    X, Y, Z = ex
    faces.append((np.zeros(3), np.array([X, 0, 0]), np.array([0, Y, 0]), np.array([0, 0, 1.0])))          # floor
    faces.append((np.zeros(3), np.array([X, 0, 0]), np.array([0, 0, Z]), np.array([0, 1.0, 0])))          # walls
    faces.append((np.array([0, Y, 0]), np.array([X, 0, 0]), np.array([0, 0, Z]), np.array([0, -1.0, 0])))
    faces.append((np.zeros(3), np.array([0, Y, 0]), np.array([0, 0, Z]), np.array([1.0, 0, 0])))
    faces.append((np.array([X, 0, 0]), np.array([0, Y, 0]), np.array([0, 0, Z]), np.array([-1.0, 0, 0])))
    for _ in range(n_boxes):
        sz = rng.uniform([0.4, 0.4, 0.3], [2.0, 1.5, 1.6])
        o = np.array([rng.uniform(0, max(X - sz[0], 0.1)), rng.uniform(0, max(Y - sz[1], 0.1)), 0.0])
        a, b, c = np.array([sz[0], 0, 0]), np.array([0, sz[1], 0]), np.array([0, 0, sz[2]])
        faces += [(o + c, a, b, np.array([0, 0, 1.0])), (o, a, c, np.array([0, -1.0, 0])), (o + b, a, c, np.array([0, 1.0, 0])),
                  (o, b, c, np.array([-1.0, 0, 0])), (o + a, b, c, np.array([1.0, 0, 0]))]
    pts, nrm = [], []
    for o, u, v, n in faces:
        area = np.linalg.norm(np.cross(u, v))
        m = max(int(area * density), 4)
        s, t = rng.random(m), rng.random(m)
        pts.append(o[None] + s[:, None] * u[None] + t[:, None] * v[None])
        nrm.append(np.repeat(n[None], m, 0))
    p = np.concatenate(pts) + rng.normal(0, 0.005, (sum(len(x) for x in pts), 3))
    n = np.concatenate(nrm) + rng.normal(0, 0.03, p.shape)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    p -= p.min(0, keepdims=True)
    col = rng.random(p.shape)
    return p.astype(np.float32), n.astype(np.float32), col.astype(np.float32)


def voxelize_first(xyz, voxel):
    """Indices of the first point of every occupied voxel (deterministic mode of util/voxelize.py)."""
    key = np.floor(xyz / np.float32(voxel)).astype(np.int64)
    key -= key.min(0, keepdims=True)
    mx = key.max(0) + 1
    flat = (key[:, 0] * mx[1] + key[:, 1]) * mx[2] + key[:, 2]
    _, idx = np.unique(flat, return_index=True)
    return np.sort(idx)


def make_scene(seed, target_points, voxel=0.1, n_boxes_per_100m2=25):
    """A room whose level-0 (voxelised) cloud has roughly `target_points` points."""
    # ~ 2.1 occupied 10 cm voxels per dm^2... calibrate floor area from the target: surface ~ 3.3 x floor area
    area = max(target_points * (voxel ** 2) / 3.0, 4.0)
    X = float(np.sqrt(area * 4.0 / 3.0))
    Y = area / X
    for _ in range(6):
        xyz, nrm, col = make_room(seed, (X, Y, 3.0), n_boxes=max(int(X * Y / 100.0 * n_boxes_per_100m2), 2),
                                  density=4.0 / (voxel ** 2))
        keep = voxelize_first(xyz, voxel)
        n = len(keep)
        if abs(n - target_points) <= 0.03 * target_points:
            break
        scale = np.sqrt(target_points / n)
        X, Y = X * scale, Y * scale
    return xyz[keep], nrm[keep], col[keep]
