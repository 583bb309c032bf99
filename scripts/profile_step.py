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
    # per (kernel, grid) durations from the chrome trace: separates the level-0 launches from the latency-bound coarse levels
    import collections, re
    trace = os.path.join(ROOT, "gpurun_out", "prof_trace.json")
    prof.export_chrome_trace(trace)
    agg = collections.defaultdict(list)
    for ev in json.load(open(trace))["traceEvents"]:
        if ev.get("cat") == "kernel":
            name = re.sub(r"\(.*$", "", ev["name"]).replace("void ", "").replace("pcfb::", "")
            agg[(name[:60], str(ev.get("args", {}).get("grid")))].append(ev["dur"])
    os.remove(trace)
    with open(os.path.join(ROOT, "gpurun_out", "prof_kernels_by_grid.txt"), "w") as f:
        f.write("# one training step, CUPTI kernel durations grouped by (kernel, grid): calls, mean us, total us\n")
        for (name, grid), d in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-62s %-16s %4d %9.1f %10.1f\n" % (name, grid, len(d), sum(d) / len(d), sum(d)))
    if stacks:                                      # who launches the small elementwise kernels?
        with open(os.path.join(ROOT, "gpurun_out", "prof_stacks.txt"), "w") as f:
            for ev in sorted(prof.key_averages(group_by_stack_n=8), key=lambda e: -e.count):
                if ev.key in ("aten::fill_", "aten::zero_", "aten::copy_", "aten::add_", "aten::clone", "aten::contiguous", "aten::cat", "aten::mul", "aten::add") and ev.count >= 8:
                    f.write("%s count=%d cuda=%.1fus\n" % (ev.key, ev.count, ev.device_time_total))
                    for fr in ev.stack[:8]:
                        f.write("    %s\n" % fr)

if __name__ == "__main__":
    main()
