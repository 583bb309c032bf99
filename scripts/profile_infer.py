"""Per-kernel CUDA time (CUPTI) of one inference forward: configPCF_2cm_PTF2 (BatchNorm folded, one ~250 k-point scene) or
configPCF_10cm_lite (eval mode, 16 packed rooms, edge construction included).
usage: python scripts/profile_infer.py [points] [ptf2|lite]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcf_b200 import configs, synthetic, eval_utils as EU, model_architecture as MA
from torch.profiler import profile, ProfilerActivity

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
which = sys.argv[2] if len(sys.argv) > 2 else "ptf2"
torch.manual_seed(1)
if which == "ptf2":
    cfgd = configs.CONFIG_PCF_2CM_PTF2
    cfg = configs.make_cfg(cfgd)
    model = MA.PointConvFormer_Segmentation(cfg).cuda()
    EU.fold_batchnorm(model)
    xyz, nrm, col = synthetic.make_scene(20, n, voxel=cfgd["grid_size"][0])
    pcs, nrms, es, ef, ep = EU.prepare_scene(xyz, nrm, cfg)
    feats = torch.from_numpy(col).cuda()[None]
    run = lambda: model(feats, pcs, es, ef, ep, nrms)
else:
    import bench
    from pcf_b200 import grid_subsampling as GS, knn_post_dataloader_utils as KU
    cfgd = configs.CONFIG_PCF_10CM_LITE
    cfg = configs.make_cfg(cfgd)
    model = MA.PointConvFormer_Segmentation(cfg).cuda().eval()
    h = bench.host_scenes(3, min(n, 100000), cfgd["grid_size"], 16)
    p0, n0 = torch.from_numpy(h["points0"]).cuda(), torch.from_numpy(h["normals0"]).cuda()
    feats = torch.from_numpy(h["colors"]).cuda()[None]

    def run():
        pts, nrm, stored, _ = GS.build_pyramid(p0, n0, h["stored0"], cfgd["grid_size"])
        pcs = [p[None] for p in pts]
        es, ef, ep = KU.prepare(*KU.compute_knn_packed(pcs, stored, cfgd["K_self"], cfgd["K_forward"], cfgd["K_propagate"],
                                                       grid_size=cfgd["grid_size"]))
        return model(feats, pcs, es, ef, ep, [x[None] for x in nrm])
with torch.no_grad():
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run()
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
