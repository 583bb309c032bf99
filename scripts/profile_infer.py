"""Per-kernel CUDA time of one inference forward of configPCF_2cm_PTF2 (BatchNorm folded, ~250 k points), CUPTI.
usage: python scripts/profile_infer.py [points]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcf_b200 import configs, synthetic, eval_utils as EU, model_architecture as MA
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
cfgd = configs.CONFIG_PCF_2CM_PTF2
cfg = configs.make_cfg(cfgd)
torch.manual_seed(1)
model = MA.PointConvFormer_Segmentation(cfg).cuda()
EU.fold_batchnorm(model)
xyz, nrm, col = synthetic.make_scene(20, n, voxel=cfgd["grid_size"][0])
pcs, nrms, es, ef, ep = EU.prepare_scene(xyz, nrm, cfg)
print("levels", [p.shape[1] for p in pcs])
feats = torch.from_numpy(col).cuda()[None]
with torch.no_grad():
    for _ in range(3):
        model(feats, pcs, es, ef, ep, nrms)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(feats, pcs, es, ef, ep, nrms)
        torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name.split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print("kernel time sum %.2f ms in %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-72s %5d %9.1f us %5.1f%%" % (k, c, us, 100 * us / tot))
