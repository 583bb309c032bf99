"""GPU parity: packed brute-force kNN + inverse map + grid subsampling (integer / bit-exact work)
against the CPU oracle, through the C ABI (pcf_cuda.* -> libpcf_b200.so)."""
import os

import numpy as np
import pytest
import torch

from oracle import knn as OK, inverse as OI, grid_subsample as OG
from gpu_util import cuda, surface_cloud

pytestmark = pytest.mark.gpu


def _pc():
    from pcf_b200 import pcf_cuda
    return pcf_cuda


def _knn_gpu(ref, qry, ref_counts, qry_counts, K):
    return _pc().knn_packed(cuda(ref), ref_counts, cuda(qry), qry_counts, K).cpu().numpy()


def _knn_oracle_packed(ref, qry, ref_counts, qry_counts, K):
    ro = np.concatenate([[0], np.cumsum(ref_counts)])
    qo = np.concatenate([[0], np.cumsum(qry_counts)])
    out = []
    for s in range(len(ref_counts)):
        out.append(OK.compute_knn(ref[ro[s]:ro[s + 1]], qry[qo[s]:qo[s + 1]], K) + ro[s])
    return np.concatenate(out)


@pytest.mark.parametrize("K", [1, 8, 16, 32, 64, 100])
def test_knn_random_cloud(K):
    rng = np.random.default_rng(K)
    ref = rng.standard_normal((5000, 3)).astype(np.float32)
    qry = rng.standard_normal((1777, 3)).astype(np.float32)
    got = _knn_gpu(ref, qry, [5000], [1777], K)
    assert got.dtype == np.int64 and got.shape == (1777, K)
    assert np.array_equal(got, OK.compute_knn(ref, qry, K))


def test_knn_self_first_and_ties():
    """Tie-heavy inputs: a perfect grid plus exact duplicates; ties must resolve to the lowest index and the
    query itself must come first for self-kNN (layers.py:377-378 relies on it)."""
    grid = np.stack(np.meshgrid(*[np.arange(13)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.1
    cloud = np.concatenate([grid, grid[:300]])
    got = _knn_gpu(cloud, cloud, [len(cloud)], [len(cloud)], 16)
    assert np.array_equal(got, OK.compute_knn(cloud, cloud, 16))
    assert np.array_equal(got[:len(grid), 0], np.arange(len(grid)))


def test_knn_packed_ragged_scenes():
    """Several scenes of ragged size in one launch (scene boundaries inside a CTA), forward and propagate
    directions, including a scene with fewer references than K (cyclic fill)."""
    rng = np.random.default_rng(7)
    ref_counts = [700, 5, 1300, 129, 40]
    qry_counts = [200, 9, 310, 33, 64]
    ref = np.concatenate([surface_cloud(n, 10 + i)[0] + i for i, n in enumerate(ref_counts)])
    qry = np.concatenate([surface_cloud(n, 20 + i)[0] + i for i, n in enumerate(qry_counts)])
    for K in (16, 8):
        assert np.array_equal(_knn_gpu(ref, qry, ref_counts, qry_counts, K), _knn_oracle_packed(ref, qry, ref_counts, qry_counts, K))
        assert np.array_equal(_knn_gpu(qry, ref, qry_counts, ref_counts, K), _knn_oracle_packed(qry, ref, qry_counts, ref_counts, K))


def test_knn_drop_in_interface():
    """compute_knn_packed() + prepare() (knn_post_dataloader_utils.py:156-223) against the oracle's packed
    restatement, on a 3-level pyramid of 3 scenes."""
    from pcf_b200 import knn_post_dataloader_utils as KU
    stored = [[900, 400, 1500], [230, 100, 380], [60, 30, 90]]
    pcs = [np.concatenate([surface_cloud(n, 100 * l + i)[0] for i, n in enumerate(stored[l])])[None] for l in range(3)]
    Ks = [16, 16, 16]
    es, ef, ep = KU.prepare(*KU.compute_knn_packed([torch.from_numpy(p) for p in pcs], stored, Ks, Ks, Ks))
    oes, oef, oep = OK.compute_knn_packed(pcs, stored, Ks, Ks, Ks)
    for got, want in zip(es + ef + ep, oes + oef + oep):
        assert got.dtype == torch.int64 and tuple(got.shape) == want.shape
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("method", ["grid", "brute"])
def test_knn_drop_in_matches_reference_golden(golden_dir, method):
    """compute_knn_packed() + prepare() against tables produced by the reference's OWN compute_knn_packed + prepare
    (kNN on its sklearn KDTree option, tie-free clouds; tests/golden/knn_packed.npz)."""
    from pcf_b200 import knn_post_dataloader_utils as KU
    g = np.load(os.path.join(golden_dir, "knn_packed.npz"))
    pcs = [torch.from_numpy(g["pc%d" % l]) for l in range(3)]
    Ks = g["Ks"].tolist()
    es, ef, ep = KU.prepare(*KU.compute_knn_packed(pcs, g["stored"].tolist(), Ks, Ks, Ks, method=method))
    want = [g["es%d" % l] for l in range(3)] + [g["ef%d" % l] for l in range(2)] + [g["ep%d" % l] for l in range(2)]
    for got, w in zip(es + ef + ep, want):
        assert got.dtype == torch.int64 and tuple(got.shape) == w.shape
        assert np.array_equal(got.cpu().numpy(), w)


def test_knn_full_size_properties():
    """BASELINE size (100k-point scene, K=16): size-independent properties + exact agreement with the oracle
    on a random subset of queries."""
    from pcf_b200 import synthetic
    xyz, _, _ = synthetic.make_scene(3, 100000)
    n = len(xyz)
    got = _knn_gpu(xyz, xyz, [n], [n], 16)
    assert np.array_equal(got[:, 0], np.arange(n))                       # self first
    assert got.min() >= 0 and got.max() < n
    assert all(len(np.unique(r)) == 16 for r in got[::997])              # no repeats
    d = ((xyz[:, None, :] - xyz[got]) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    assert np.all(np.diff(d, axis=1) >= 0)                               # ascending distance
    ties = np.diff(d, axis=1) == 0
    assert np.all(np.diff(got, axis=1)[ties] > 0)                        # ties by ascending index
    sub = np.random.default_rng(0).choice(n, 1500, replace=False)
    assert np.array_equal(got[sub], OK.knn_c(xyz, xyz[sub], 16))


def _grid(ref, qry, ref_counts, qry_counts, K, hint=0.0):
    return _pc().KnnGrid(cuda(ref), ref_counts, hint).query(cuda(qry), qry_counts, K).cpu().numpy()


def test_grid_knn_query_order_does_not_change_the_table():
    """pcfb_knn_grid_query's optional query order (the cell order of the query cloud's own grid: coherent warps) is a
    scheduling hint only: same table as the natural order and as the oracle, ragged packed scenes included."""
    pc = _pc()
    ref_counts, qry_counts = [900, 5, 1500, 129], [300, 9, 410, 33]
    ref = np.concatenate([surface_cloud(n, 30 + i)[0] + i for i, n in enumerate(ref_counts)])
    qry = np.concatenate([surface_cloud(n, 40 + i)[0] + i for i, n in enumerate(qry_counts)])
    r, q = cuda(ref), cuda(qry)
    g_ref, g_qry = pc.KnnGrid(r, ref_counts, 0.2), pc.KnnGrid(q, qry_counts, 0.1)
    want = _knn_oracle_packed(ref, qry, ref_counts, qry_counts, 16)
    for K in (3, 16, 33):
        plain = g_ref.query(q, qry_counts, K).cpu().numpy()
        ordered = g_ref.query(q, qry_counts, K, order=g_qry).cpu().numpy()
        assert np.array_equal(plain, ordered)
        if K == 16:
            assert np.array_equal(ordered, want)
    assert g_qry.order_ptr() != 0
    with pytest.raises(RuntimeError):
        g_ref.query(q, qry_counts, 16, order=g_ref)                      # not the grid of the query cloud


@pytest.mark.parametrize("hint", [0.0, 0.05, 0.3, 5.0])
def test_grid_knn_equals_brute_force_and_oracle(hint):
    """The grid search must return exactly the brute-force table whatever the cell size: random cloud,
    surface clouds, tie-heavy grid with duplicates, ragged packed scenes with a scene smaller than K."""
    rng = np.random.default_rng(3)
    ref = rng.standard_normal((4000, 3)).astype(np.float32)
    qry = rng.standard_normal((900, 3)).astype(np.float32) * 1.5            # some queries outside the ref bbox
    for K in (1, 16, 33, 64):
        assert np.array_equal(_grid(ref, qry, [4000], [900], K, hint), OK.compute_knn(ref, qry, K))
    g = np.stack(np.meshgrid(*[np.arange(13)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * 0.1
    cloud = np.concatenate([g, g[:300]])
    assert np.array_equal(_grid(cloud, cloud, [len(cloud)], [len(cloud)], 16, hint), _knn_gpu(cloud, cloud, [len(cloud)], [len(cloud)], 16))
    ref_counts, qry_counts = [700, 5, 1300, 129, 40], [200, 9, 310, 33, 64]
    ref = np.concatenate([surface_cloud(n, 10 + i)[0] + i for i, n in enumerate(ref_counts)])
    qry = np.concatenate([surface_cloud(n, 20 + i)[0] + i for i, n in enumerate(qry_counts)])
    assert np.array_equal(_grid(ref, qry, ref_counts, qry_counts, 16, hint), _knn_oracle_packed(ref, qry, ref_counts, qry_counts, 16))
    assert np.array_equal(_grid(qry, ref, qry_counts, ref_counts, 8, hint), _knn_oracle_packed(qry, ref, qry_counts, ref_counts, 8))


def test_grid_knn_degenerate_clouds():
    line = np.zeros((3000, 3), np.float32); line[:, 0] = np.linspace(0, 300, 3000)     # 1-D cloud: very elongated grid
    assert np.array_equal(_grid(line, line, [3000], [3000], 16), _knn_gpu(line, line, [3000], [3000], 16))
    same = np.ones((500, 3), np.float32)                                               # all points identical
    assert np.array_equal(_grid(same, same, [500], [500], 16), _knn_gpu(same, same, [500], [500], 16))


def test_grid_knn_full_size_equals_brute_force():
    """BASELINE size: the whole 13-edge-set pyramid of a 100k-point scene, grid vs brute force, exact."""
    from pcf_b200 import synthetic, knn_post_dataloader_utils as KU, grid_subsampling as GS
    xyz, nrm, _ = synthetic.make_scene(4, 100000)
    gs = [0.1, 0.2, 0.4, 0.8, 1.6]
    pts, _ = GS.subsample(xyz, nrm, gs)
    pcs = [p[None] for p in pts]
    stored = [[p.shape[0]] for p in pts]
    a = KU.prepare(*KU.compute_knn_packed(pcs, stored, [16] * 5, [16] * 5, [16] * 5, grid_size=gs, method="grid"))
    b = KU.prepare(*KU.compute_knn_packed(pcs, stored, [16] * 5, [16] * 5, [16] * 5, method="brute"))
    c = KU.prepare(*KU.compute_knn_packed(pcs, stored, [16] * 5, [16] * 5, [16] * 5, method="grid"))     # automatic cell size
    for x, y, z in zip(a[0] + a[1] + a[2], b[0] + b[1] + b[2], c[0] + c[1] + c[2]):
        assert torch.equal(x, y) and torch.equal(z, y)


# ---------------------------------------------------------------------------------------------------
def _inv_gpu(nei, total):
    n, k, idx = _pc().compute_knn_inverse(cuda(nei)[None], total)
    return n[0].cpu().numpy(), k[0].cpu().numpy(), idx[0].cpu().numpy()


def test_inverse_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "inverse.npz"))
    for tag in ("self", "fwd", "prop"):
        nei, total = g[tag + "_nei"], int(g[tag + "_total"])
        n, k, idx = _inv_gpu(nei, total)
        E = nei.size
        assert n.dtype == np.int32 and k.dtype == np.uint8 and idx.dtype == np.int32
        assert np.array_equal(idx, g[tag + "_inv_idx"]) and np.array_equal(n[:E], g[tag + "_inv_neighbors"])
        assert np.array_equal(k[:E], g[tag + "_inv_k"])


@pytest.mark.parametrize("n_out,K,total", [(1, 1, 1), (513, 16, 300), (4000, 16, 16000), (300, 64, 5), (10000, 3, 100000), (77, 255, 50)])
def test_inverse_random(n_out, K, total):
    rng = np.random.default_rng(n_out + K)
    nei = rng.integers(-1, total, (n_out, K)).astype(np.int64)       # includes -1 padding
    nei[rng.random(nei.shape) < 0.01] = total + 3                       # and out-of-range entries
    got = _inv_gpu(nei, total)
    want = OI.knn_inverse(nei, total)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_inverse_skewed_degree():
    """One input point referenced by every edge (segment far longer than a warp)."""
    nei = np.zeros((3000, 16), np.int64)
    nei[:, 1:] = np.random.default_rng(1).integers(0, 50, (3000, 15))
    got = _inv_gpu(nei, 50)
    for a, b in zip(got, OI.knn_inverse(nei, 50)):
        assert np.array_equal(a, b)


def test_inverse_drop_in_interface():
    from pcf_b200 import common_util as CU
    rng = np.random.default_rng(5)
    pcs = [torch.zeros(1, n, 3) for n in (1000, 250, 60)]
    es = [torch.from_numpy(rng.integers(0, n, (1, n, 16))) for n in (1000, 250, 60)]
    ef = [torch.from_numpy(rng.integers(0, a, (1, b, 16))) for a, b in ((1000, 250), (250, 60))]
    ep = [torch.from_numpy(rng.integers(0, b, (1, a, 16))) for a, b in ((1000, 250), (250, 60))]
    got = CU.compute_knn_inverse(pcs, es, ef, ep)
    want = OI.compute_knn_inverse([p.numpy() for p in pcs], [e.numpy() for e in es], [e.numpy() for e in ef], [e.numpy() for e in ep])
    for gk, wk in zip(got, want):
        for gl, wl in zip(gk, wk):
            for a, b in zip(gl, wl):
                assert np.array_equal(a.cpu().numpy(), b)


def test_inverse_full_size_properties():
    rng = np.random.default_rng(11)
    N, K = 100000, 16
    nei = rng.integers(0, N, (N, K)).astype(np.int64)
    n, k, idx = _inv_gpu(nei, N)
    assert idx[0] == 0 and idx[-1] == N * K and np.all(np.diff(idx) >= 0)
    assert np.array_equal(np.diff(idx), np.bincount(nei.reshape(-1), minlength=N))
    assert np.array_equal(nei[n, k], np.repeat(np.arange(N), np.diff(idx)))      # every entry points back
    e = n.astype(np.int64) * K + k
    seg_start = np.zeros(N * K, bool); seg_start[idx[:-1][np.diff(idx) > 0]] = True
    assert np.all((np.diff(e) > 0) | seg_start[1:])                              # ascending (n,k) inside segments


# ---------------------------------------------------------------------------------------------------
def test_grid_subsample_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "grid_subsample.npz"))
    for i in range(3):
        p, f, dl = g["in_p%d" % i], g["in_f%d" % i], float(g["dl%d" % i])
        sp, sf, counts = _pc().grid_subsample(cuda(p), cuda(f), [len(p)], dl)
        sp, sf = sp.cpu().numpy(), sf.cpu().numpy()
        assert counts == [len(g["out_p%d" % i])]
        o = np.lexsort(sp.T[::-1])
        assert np.array_equal(sp[o], g["out_p%d" % i]) and np.array_equal(sf[o], g["out_f%d" % i])   # bit-exact


def test_grid_subsample_packed_order():
    """Packed scenes, canonical output order (ascending scene, voxel key) == the oracle's, bit for bit."""
    counts = [3000, 17, 1200]
    clouds = [surface_cloud(n, 50 + i, extent=(5.0, 4.0, 2.5)) for i, n in enumerate(counts)]
    p = np.concatenate([c[0] - 1.3 for c in clouds]); f = np.concatenate([c[1] for c in clouds])
    sp, sf, sc = _pc().grid_subsample(cuda(p), cuda(f), counts, 0.25)
    off = np.concatenate([[0], np.cumsum(counts)])
    want_p, want_f, want_c = [], [], []
    for s in range(3):
        a, b, _, _ = OG.grid_subsample(p[off[s]:off[s + 1]], f[off[s]:off[s + 1]], 0.25)
        want_p.append(a); want_f.append(b); want_c.append(len(a))
    assert sc == want_c
    assert np.array_equal(sp.cpu().numpy(), np.concatenate(want_p)) and np.array_equal(sf.cpu().numpy(), np.concatenate(want_f))


def test_subsample_drop_in():
    from pcf_b200 import grid_subsampling as GS
    p, n = surface_cloud(6000, 77, extent=(6.0, 5.0, 2.5))
    pts, nrm = GS.subsample(p, n, [0.1, 0.2, 0.4, 0.8, 1.6])
    wp, wn = OG.subsample(p, n, [0.1, 0.2, 0.4, 0.8, 1.6])
    for a, b in zip(pts + nrm, wp + wn):
        assert np.array_equal(a.cpu().numpy(), b)


def test_voxelize_matches_oracle_and_reference(golden_dir):
    """GPU voxelisation (pcfb_voxelize) = the oracle exactly (index for index), on the reference's golden clouds -- whose
    occupied voxels / order it therefore reproduces (tests/test_oracle_golden.py::test_voxelize_matches_reference) -- and
    on a packed batch of three ragged scenes."""
    from oracle import voxelize as OV
    from pcf_b200 import grid_subsampling as GS
    g = np.load(os.path.join(golden_dir, "voxelize.npz"))
    for i in range(2):
        p, voxel = g["p%d" % i], float(g["voxel%d" % i])
        idx, cnt = GS.voxelize_packed(p, [len(p)], voxel)
        assert cnt == [len(g["idx%d" % i])]
        assert np.array_equal(idx.cpu().numpy(), OV.voxelize(p, voxel))
        assert np.array_equal(OV.ravel_keys(p, voxel)[idx.cpu().numpy()], g["keys%d" % i])
    rng = np.random.default_rng(4)
    counts = [3000, 1, 4500]
    clouds = [(rng.random((n, 3)) * [4, 3, 2.5] - [1.0, 2.0, 0.5]).astype(np.float32) for n in counts]
    packed = np.concatenate(clouds)
    idx, cnt = GS.voxelize_packed(packed, counts, 0.1)
    ref_idx, ref_cnt = OV.voxelize_packed(packed, counts, 0.1)
    assert cnt == ref_cnt and np.array_equal(idx.cpu().numpy(), ref_idx)


def test_build_pyramid_equals_per_level_subsampling():
    """The device-sized pyramid (one host read for all levels) must give exactly the per-level grid subsampling (two host
    reads per level), which is bit-exact against the reference C++ (test_grid_subsample_golden); a second call with the
    counts / boxes of the first reads nothing back."""
    from pcf_b200 import grid_subsampling as GS
    from gpu_util import surface_cloud
    grid = [0.1, 0.2, 0.4, 0.8, 1.6]
    clouds = [surface_cloud(9000, 31, extent=(7.0, 5.0, 2.6)), surface_cloud(5000, 32, extent=(4.0, 6.0, 2.6))]
    p = np.concatenate([c[0] for c in clouds]); n = np.concatenate([c[1] for c in clouds])
    counts = [len(c[0]) for c in clouds]
    pts, nrm, cnt, info = GS.build_pyramid(p, n, counts, grid)
    rp, rn, rc = GS.subsample_packed(p, n, counts, grid)
    assert cnt == [list(map(int, c)) for c in rc]
    for a, b in zip(pts + nrm, rp + rn):
        assert torch.equal(a, b)
    pts2, nrm2, cnt2, info2 = GS.build_pyramid(p, n, counts, grid, expect=cnt, boxes=info["boxes"])
    assert int(info2["status"].item()) == 0 and cnt2 == cnt
    for a, b in zip(pts2 + nrm2, pts + nrm):
        assert torch.equal(a, b)
