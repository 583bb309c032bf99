"""Evaluation-side callers of the hot path, with the reference's protocols:

* timed_inference     -- test_ScanNet_simple.py:139-174: model.eval(), replace_batchnorm (every Linear_BN folded), batch 1, wall
                         clock between two torch.cuda.synchronize() around model(...) ONLY (edges are built before the clock
                         starts, as the reference's DataLoader does), softmax, mean over scenes -- the protocol behind the
                         published 70.5 / 110.0 / 281.9 ms numbers.
* vote_inference      -- test_ScanNet_voting.py:201-268: several rotations x the 'multiple' voxelisation parts of a raw scene
                         (util/voxelize.py:61-67: every raw point is covered by at least one part), softmax probabilities
                         accumulated per raw point, normalised per rotation, summed over rotations, argmax.
* knn_post_benchmark  -- knn_post_benchmark.py:94-151: per "epoch" compute_knn_packed + prepare + host->device copies of every
                         batch between two CUDA events, 51 epochs, mean without the first.
* intersectionAndUnion(GPU) -- util/common_util.py:57-85 (the mIoU bookkeeping of both drivers).

The scene preparation (voxelise -> pyramid -> 13 edge sets) runs on the GPU (grid_subsampling.build_pyramid /
voxelize_packed, knn_post_dataloader_utils.compute_knn_packed) instead of the reference's DataLoader workers.
"""
import time

import numpy as np
import torch
import torch.nn.functional as F

from . import common_util as CU
from . import grid_subsampling as GS
from . import knn_post_dataloader_utils as KU


def intersectionAndUnion(output, target, K, ignore_index=255):
    """util/common_util.py:57-70 (numpy)."""
    output = np.asarray(output).flatten().copy()
    target = np.asarray(target).flatten()
    output[np.where(target == ignore_index)[0]] = ignore_index
    intersection = output[np.where(output == target)[0]]
    area_intersection, _ = np.histogram(intersection, bins=np.arange(K + 1))
    area_output, _ = np.histogram(output, bins=np.arange(K + 1))
    area_target, _ = np.histogram(target, bins=np.arange(K + 1))
    return area_intersection, area_output + area_target - area_intersection, area_target


def intersectionAndUnionGPU(output, target, K, ignore_index=255):
    """util/common_util.py:73-85 (torch)."""
    output = output.reshape(-1).clone()
    target = target.reshape(-1)
    output[target == ignore_index] = ignore_index
    intersection = output[output == target]
    area_intersection = torch.histc(intersection.float(), bins=K, min=0, max=K - 1)
    area_output = torch.histc(output.float(), bins=K, min=0, max=K - 1)
    area_target = torch.histc(target.float(), bins=K, min=0, max=K - 1)
    return area_intersection, area_output + area_target - area_intersection, area_target


def prepare_scene(coord, norm, cfg, counts=None):
    """Level-0 cloud(s) (already voxelised) -> (pointclouds, norms, edges_self, edges_forward, edges_propagate) in the
    model's input format ([1, N_l, 3] per level, [1, N, K] int64 tables), everything built on the device."""
    coord = torch.as_tensor(coord, dtype=torch.float32).cuda()
    norm = torch.as_tensor(norm, dtype=torch.float32).cuda()
    counts = [coord.shape[0]] if counts is None else list(map(int, counts))
    pts, nrm, stored, _ = GS.build_pyramid(coord, norm, counts, cfg.grid_size)
    pcs = [p.unsqueeze(0) for p in pts]
    es, ef, ep = KU.prepare(*KU.compute_knn_packed(pcs, stored, cfg.K_self, cfg.K_forward, cfg.K_propagate, grid_size=cfg.grid_size))
    return pcs, [n.unsqueeze(0) for n in nrm], es, ef, ep


def fold_batchnorm(model):
    """model.eval() + replace_batchnorm (test_ScanNet_simple.py:139-141)."""
    model.eval()
    CU.replace_batchnorm(model)
    return model


@torch.no_grad()
def timed_inference(model, scenes, cfg, fold_bn=True, warmup=1):
    """scenes: iterable of (coord [N,3], norm [N,3], color [N,3]) level-0 clouds.  -> (list of softmax probabilities [N, classes],
    list of seconds per scene, mean seconds).  `warmup` untimed passes over the first scene absorb one-time costs (the
    reference's first scene pays them inside its mean)."""
    if fold_bn:
        fold_batchnorm(model)
    else:
        model.eval()
    probs, times = [], []
    scenes = list(scenes)
    for i, (coord, norm, color) in enumerate([scenes[0]] * warmup + scenes):
        pcs, nrms, es, ef, ep = prepare_scene(coord, norm, cfg)
        feats = torch.as_tensor(color, dtype=torch.float32).cuda().unsqueeze(0)
        torch.cuda.synchronize()
        st = time.time()
        pred = model(feats, pcs, es, ef, ep, nrms)
        torch.cuda.synchronize()
        et = time.time()
        if i >= warmup:
            times.append(et - st)
            probs.append(F.softmax(pred.contiguous().view(-1, pred.shape[-1]), dim=-1))
    return probs, times, float(np.mean(times))


def rotate_scene(coord, norm, rotate_deg):
    """scannet_data_loader_color_DDP.py:176-182: rotation about z by rotate_deg * 360 - 180 degrees (rotate_deg in [0, 1])."""
    coord, norm = np.array(coord, np.float32, copy=True), np.array(norm, np.float32, copy=True)
    if rotate_deg != 0.:
        rad = np.deg2rad(rotate_deg * 360) - np.pi
        c, s = np.cos(rad), np.sin(rad)
        j = np.array([[c, s], [-s, c]], np.float32)
        coord[:, :2] = coord[:, :2] @ j
        norm[:, :2] = norm[:, :2] @ j
    return coord, norm


def voxelize_multiple(coord, voxel_size):
    """voxelize(coord, voxel_size, mode='multiple') (util/voxelize.py:61-67): index sets, part i takes the (i mod count)-th
    point of every voxel, so count.max() parts cover every raw point at least once.  Keys / per-voxel order on the device:
    ravel keys, points of a voxel in ascending input order."""
    p = torch.as_tensor(coord, dtype=torch.float32).cuda()
    d = torch.floor(p.double() / float(voxel_size))
    d = (d - d.min(0).values).long()
    mx = d.max(0).values + 1
    key = (d[:, 0] * mx[1] + d[:, 1]) * mx[2] + d[:, 2]
    order = torch.argsort(key, stable=True)
    ks = key[order]
    start = torch.ones_like(ks, dtype=torch.bool)
    start[1:] = ks[1:] != ks[:-1]
    first = torch.nonzero(start).flatten()
    count = torch.diff(torch.cat([first, torch.tensor([ks.numel()], device=ks.device)]))
    parts = []
    for i in range(int(count.max())):
        parts.append(order[first + (i % count)])
    return parts


@torch.no_grad()
def vote_inference(model, coord, norm, color, cfg, rotate_degs=(0.0,), num_classes=20, fold_bn=True):
    """Multi-rotation, multi-part voting over ONE raw (un-voxelised) scene (test_ScanNet_voting.py:201-268).
    -> (summed probabilities [N_raw, classes], list of seconds per model call)."""
    if fold_bn:
        fold_batchnorm(model)
    else:
        model.eval()
    n_raw = len(coord)
    total = torch.zeros(n_raw, num_classes, device="cuda")
    times = []
    for deg in rotate_degs:
        c_r, n_r = rotate_scene(coord, norm, deg)
        c_r -= c_r.min(0)                                             # input normalize (scannet_data_loader_color_DDP.py:205-207)
        pred = torch.zeros(n_raw, num_classes, device="cuda")
        for idx in voxelize_multiple(c_r, cfg.grid_size[0]):
            idx_h = idx.cpu().numpy()
            pcs, nrms, es, ef, ep = prepare_scene(c_r[idx_h], n_r[idx_h], cfg)
            feats = torch.as_tensor(np.asarray(color, np.float32)[idx_h]).cuda().unsqueeze(0)
            torch.cuda.synchronize()
            st = time.time()
            part = model(feats, pcs, es, ef, ep, nrms)
            torch.cuda.synchronize()
            times.append(time.time() - st)
            pred[idx] += F.softmax(part.contiguous().view(-1, num_classes), dim=-1)
        total += pred / (pred.sum(-1)[:, None] + 1e-8)
    return total, times


def knn_post_benchmark(batches, cfg, iters=51):
    """knn_post_benchmark.py:94-151 on pre-collated host batches: every epoch runs, for each batch
    (pointclouds: list of [1, N_l, 3] CPU tensors, points_stored, other host tensors to upload), compute_knn_packed + prepare
    + the host->device copies, bracketed by CUDA events.  -> mean seconds per epoch without the first (the reference's
    "Average time for keops")."""
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timing = []
    for _ in range(iters):
        torch.cuda.synchronize()
        start.record()
        for pointclouds, points_stored, extras in batches:
            es, ef, ep = KU.prepare(*KU.compute_knn_packed(pointclouds, points_stored, cfg.K_self, cfg.K_forward, cfg.K_propagate,
                                                           grid_size=getattr(cfg, "grid_size", None)))
            CU.to_device(pointclouds, non_blocking=True)
            CU.to_device(list(extras), non_blocking=True)
        torch.cuda.synchronize()
        end.record()
        torch.cuda.synchronize()
        timing.append(start.elapsed_time(end) / 1000)
    return float(sum(timing[1:]) / max(len(timing) - 1, 1)), timing
