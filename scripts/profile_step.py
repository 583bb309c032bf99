"""Per-kernel time breakdown of one bench step with torch.profiler (CUPTI): writes a table sorted by CUDA time."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity

def main():
    stacks = "stacks" in os.environ.get("PROFILE_STEP_FLAGS", "")
    sys.argv = ["bench.py", "--steps", "1", "--warmup", "2", "--no-cpu-baseline"] + sys.argv[1:]
    args = bench.parse()
    # reuse bench internals: run_ours builds everything; instead replicate the step here
    import pcf_b200
    from pcf_b200 import configs, model_architecture as MA, knn_post_dataloader_utils as KU, common_util as CU, pcf_cuda
    pcf_cuda.FORWARD_VARIANT = args.variant
    dev = torch.device("cuda", 0)
    cfgd = configs.CONFIG_PCF_OPT_10CM
    model = MA.PointConvFormer_Segmentation(configs.make_cfg(cfgd)).to(dev).train()
    from pcf_b200 import sharding
    flat = sharding.FlatParameters(model)
    opt = torch.optim.AdamW([flat.flat], lr=1e-3, weight_decay=0.05, fused=True)
    host = bench.host_pyramid(1, args.points, cfgd["grid_size"], args.scenes)
    pts = [torch.from_numpy(p).to(dev) for p in host["points"]]
    nrm = [torch.from_numpy(p).to(dev) for p in host["normals"]]
    col, lab = torch.from_numpy(host["colors"]).to(dev), torch.from_numpy(host["labels"]).to(dev)
    def step():
        pcs = [p.unsqueeze(0) for p in pts]; nrms = [p.unsqueeze(0) for p in nrm]
        with torch.profiler.record_function("edges_knn"):
            es, ef, ep = KU.prepare(*KU.compute_knn_packed(pcs, host["stored"], cfgd["K_self"], cfgd["K_forward"], cfgd["K_propagate"], grid_size=cfgd["grid_size"]))
        with torch.profiler.record_function("edges_inverse"):
            inv = CU.compute_knn_inverse(pcs, es, ef, ep)
        with torch.profiler.record_function("forward"):
            logits = model(col.unsqueeze(0), pcs, es, ef, ep, nrms, *inv)
            loss = torch.nn.functional.cross_entropy(logits[0], lab, label_smoothing=0.2)
        with torch.profiler.record_function("backward"):
            loss.backward()
        with torch.profiler.record_function("optimizer"):
            flat.gather_grads()
            torch.nn.utils.clip_grad_norm_([flat.flat], 10.0)
            opt.step()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=stacks) as prof:
        step()
        torch.cuda.synchronize()
    out = os.path.join(ROOT, "gpurun_out", "prof_table.txt")
    with open(out, "w") as f:
        f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
    print(open(out).read()[:6000])
    if stacks:                                      # who launches the small elementwise kernels?
        with open(os.path.join(ROOT, "gpurun_out", "prof_stacks.txt"), "w") as f:
            for ev in sorted(prof.key_averages(group_by_stack_n=8), key=lambda e: -e.count):
                if ev.key in ("aten::fill_", "aten::zero_", "aten::copy_", "aten::add_", "aten::clone", "aten::contiguous", "aten::cat", "aten::mul", "aten::add") and ev.count >= 8:
                    f.write("%s count=%d cuda=%.1fus\n" % (ev.key, ev.count, ev.device_time_total))
                    for fr in ev.stack[:8]:
                        f.write("    %s\n" % fr)

if __name__ == "__main__":
    main()
