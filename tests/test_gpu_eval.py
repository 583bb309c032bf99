"""GPU tests of the evaluation-side callers (pcf_b200/eval_utils.py): the reference's inference timing protocol
(test_ScanNet_simple.py:139-174), voting assembly (test_ScanNet_voting.py:201-268) and the kNN dataloader benchmark
(knn_post_benchmark.py:94-151)."""
import numpy as np
import pytest
import torch

from gpu_util import surface_cloud

pytestmark = pytest.mark.gpu


def _small_model():
    import model_variants
    from pcf_b200 import model_architecture as MA
    c = dict(model_variants.cfg_of("small"), USE_CUDA_KERNEL=True, PCONV_OPT=False, K_self=[16] * 5, K_forward=[16] * 5,
             K_propagate=[16] * 5, grid_size=[0.1, 0.2, 0.4, 0.8, 1.6])
    cfg = MA.get_default_configs(MA.EasyDict(c), c["num_level"], c["base_dim"])
    torch.manual_seed(3)
    return MA.PointConvFormer_Segmentation(cfg).cuda(), cfg


def _scene(n, seed):
    from oracle import voxelize as OV
    p, nrm = surface_cloud(n, seed, extent=(6.0, 5.0, 2.6))
    keep = OV.voxelize(p, 0.1)
    col = np.random.default_rng(seed).random((n, 3)).astype(np.float32)
    return p[keep], nrm[keep], col[keep]


def test_voxelize_multiple_covers_every_point():
    """mode='multiple' (util/voxelize.py:61-67): part 0 is the deterministic voxelisation, the parts together cover every
    raw point, and no part holds two points of one voxel."""
    from oracle import voxelize as OV
    from pcf_b200 import eval_utils as EU
    p, _ = surface_cloud(5000, 8, extent=(3.0, 2.5, 2.0))
    parts = [x.cpu().numpy() for x in EU.voxelize_multiple(p, 0.1)]
    assert np.array_equal(np.sort(parts[0]), np.sort(OV.voxelize(p, 0.1)))
    assert len(np.unique(np.concatenate(parts))) == len(p)
    keys = OV.ravel_keys(p, 0.1)
    for part in parts:
        assert len(np.unique(keys[part])) == len(part)


def test_timed_inference_and_voting_protocols():
    from pcf_b200 import eval_utils as EU
    model, cfg = _small_model()
    scenes = [_scene(7000, 41), _scene(5000, 42)]
    model.eval()
    with torch.no_grad():
        pcs, nrms, es, ef, ep = EU.prepare_scene(scenes[0][0], scenes[0][1], cfg)
        want = torch.softmax(model(torch.from_numpy(scenes[0][2]).cuda()[None], pcs, es, ef, ep, nrms)[0], -1)
    probs, times, mean_s = EU.timed_inference(model, scenes, cfg, fold_bn=True, warmup=1)
    assert len(probs) == 2 and len(times) == 2 and mean_s > 0
    assert probs[0].shape == (len(scenes[0][0]), 20)
    torch.testing.assert_close(probs[0], want, rtol=2e-3, atol=2e-4)          # BatchNorm folding keeps the eval-mode output
    # voting over the raw (un-voxelised) cloud: every raw point gets a distribution; two rotations add up
    raw_p, raw_n = surface_cloud(6000, 43, extent=(5.0, 4.0, 2.6))
    raw_c = np.random.default_rng(1).random((6000, 3)).astype(np.float32)
    total, t = EU.vote_inference(model, raw_p, raw_n, raw_c, cfg, rotate_degs=(0.0, 0.25), fold_bn=False)
    assert total.shape == (6000, 20) and len(t) >= 2
    torch.testing.assert_close(total.sum(-1), torch.full((6000,), 2.0, device="cuda"), rtol=1e-4, atol=1e-4)
    i, u, tg = EU.intersectionAndUnion(total.argmax(1).cpu().numpy(), np.zeros(6000, np.int64), 20, 255)
    i2, u2, tg2 = EU.intersectionAndUnionGPU(total.argmax(1), torch.zeros(6000, dtype=torch.long, device="cuda"), 20, 255)
    assert np.array_equal(i, i2.cpu().numpy().astype(np.int64)) and np.array_equal(u, u2.cpu().numpy().astype(np.int64))


def test_knn_post_benchmark_runs():
    from pcf_b200 import eval_utils as EU, grid_subsampling as GS
    _, cfg = _small_model()
    p, nrm, col = _scene(6000, 44)
    pts, nrms, stored, _ = GS.build_pyramid(p, nrm, [len(p)], cfg.grid_size)
    batch = ([x.cpu()[None] for x in pts], stored, [torch.from_numpy(col)])
    mean_s, timing = EU.knn_post_benchmark([batch, batch], cfg, iters=3)
    assert len(timing) == 3 and mean_s > 0
