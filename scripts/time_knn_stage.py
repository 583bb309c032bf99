"""Time the edge-construction stage (5 grid builds + 13 query sets, knn_post_dataloader_utils.compute_knn_packed) of one
bench scene in isolation, for several cell-size factors.  usage: python scripts/time_knn_stage.py [factors...]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pcf_b200 import configs, grid_subsampling as GS, knn_post_dataloader_utils as KU

cfg = configs.make_cfg(configs.CONFIG_PCF_OPT_10CM)
h = bench.host_scenes(0, 100000, cfg.grid_size, 1)
coord, norm = torch.as_tensor(h["points0"]).cuda(), torch.as_tensor(h["normals0"]).cuda()
pts, nrm, stored, _ = GS.build_pyramid(coord, norm, h["stored0"], cfg.grid_size)
pcs = [p.unsqueeze(0) for p in pts]
print("levels", [p.shape[0] for p in pts])
for f in [float(a) for a in sys.argv[1:]] or [2.5]:
    KU.CELL_FACTOR = f
    run = lambda: KU.compute_knn_packed(pcs, stored, cfg.K_self, cfg.K_forward, cfg.K_propagate, grid_size=cfg.grid_size)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g.replay()
    b.record(); torch.cuda.synchronize()
    print("cell factor %.2f: stage %.1f us (graph replay)" % (f, a.elapsed_time(b) * 50))

# kernel timeline of one replay at the last factor
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.replay()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
for e in ev:
    print("%8.1f %8.1f  %s" % (e.time_range.start - t0, e.time_range.end - t0, e.name[:60]))
