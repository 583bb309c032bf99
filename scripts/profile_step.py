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
    args.points = args.points or 100000
    # reuse bench internals: run_ours builds everything; instead replicate the step here
    import pcf_b200
    from pcf_b200 import configs, model_architecture as MA, knn_post_dataloader_utils as KU, common_util as CU, pcf_cuda
    pcf_cuda.FORWARD_VARIANT = args.variant
    dev = torch.device("cuda", 0)
    cfgd = configs.CONFIG_PCF_OPT_10CM
    model = MA.PointConvFormer_Segmentation(configs.make_cfg(cfgd)).to(dev).train()
    from pcf_b200 import sharding
    flat = sharding.FlatParameters(model, async_weight_grads=True)
    opt = sharding.FlatAdamW(flat, lr=1e-3, weight_decay=0.05, max_norm=10.0)
    from pcf_b200 import losses
    host = bench.host_scenes(1, args.points, cfgd["grid_size"], args.scenes)
    from pcf_b200 import grid_subsampling as GS
    p0, n0 = torch.from_numpy(host["points0"]).to(dev), torch.from_numpy(host["normals0"]).to(dev)
    col, lab = torch.from_numpy(host["colors"]).to(dev), torch.from_numpy(host["labels"]).to(dev)
    _, _, stored, pyr = GS.build_pyramid(p0, n0, host["stored0"], cfgd["grid_size"])
    host["stored"] = stored
    def step():
        with torch.profiler.record_function("pyramid"):
            pts, nrm, _, _ = GS.build_pyramid(p0, n0, host["stored0"], cfgd["grid_size"], expect=stored, boxes=pyr["boxes"])
        pcs = [p.unsqueeze(0) for p in pts]; nrms = [p.unsqueeze(0) for p in nrm]
        with torch.profiler.record_function("edges_knn"):
            es, ef, ep = KU.prepare(*KU.compute_knn_packed(pcs, host["stored"], cfgd["K_self"], cfgd["K_forward"], cfgd["K_propagate"], grid_size=cfgd["grid_size"]))
        with torch.profiler.record_function("edges_inverse"):
            inv = CU.compute_knn_inverse(pcs, es, ef, ep)
        with torch.profiler.record_function("forward"):
            logits = model(col.unsqueeze(0), pcs, es, ef, ep, nrms, *inv)
            loss = losses.cross_entropy(logits[0], lab, label_smoothing=0.2)
        with torch.profiler.record_function("backward"):
            loss.backward()
        with torch.profiler.record_function("optimizer"):
            opt.step(flat.gather_grads())
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    use_graph = "eager" not in os.environ.get("PROFILE_STEP_FLAGS", "")
    if use_graph:                                   # profile one REPLAY of the captured step: the timeline the bench times
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        graph.replay()
        torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=stacks) as prof:
        graph.replay() if use_graph else step()
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
    spans = []
    for ev in json.load(open(trace))["traceEvents"]:
        if ev.get("cat") == "kernel":
            name = re.sub(r"\(.*$", "", ev["name"]).replace("void ", "").replace("pcfb::", "")
            grid = ev.get("args", {}).get("grid")
            agg[(name[:60], str(grid))].append(ev["dur"])
            ctas = 1
            for gdim in (grid or [1]):
                ctas *= int(gdim)
            spans.append((ev["ts"], ev["ts"] + ev["dur"], ev["dur"], ctas, ev.get("args", {}).get("stream"), name[:44]))
    os.remove(trace)
    # timeline summary: is the step bound by work or by dependent-launch latency?
    spans.sort()
    wall = max(x[1] for x in spans) - spans[0][0]
    busy, cur_s, cur_e = 0.0, None, None
    for s0, e0, _, _, _, _ in spans:
        if cur_e is None or s0 > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s0, e0
        else:
            cur_e = max(cur_e, e0)
    busy += cur_e - cur_s
    total = sum(x[2] for x in spans)
    sub = sum(x[2] for x in spans if x[3] < 148)
    # approximate critical path: from the last kernel walk back to the kernel that finished latest before it started
    import bisect
    by_end = sorted(spans, key=lambda x: x[1])
    ends = [x[1] for x in by_end]
    cur = by_end[-1]
    crit = collections.defaultdict(lambda: [0, 0.0])
    gap_total, n_crit = 0.0, 0
    while True:
        crit[cur[5]][0] += 1; crit[cur[5]][1] += cur[2]; n_crit += 1
        i = bisect.bisect_right(ends, cur[0] + 0.5) - 1
        while i >= 0 and by_end[i] is cur:
            i -= 1
        if i < 0:
            break
        gap_total += max(0.0, cur[0] - by_end[i][1])
        cur = by_end[i]
    with open(os.path.join(ROOT, "gpurun_out", "prof_timeline.txt"), "w") as f:
        f.write("kernels %d  wall %.1f us  GPU busy (union of kernel spans) %.1f us  idle %.1f us  sum of kernel durations %.1f us "
                "(avg concurrency %.2f)  sub-wave (<148 CTAs) kernels: %d launches, %.1f us\n"
                % (len(spans), wall, busy, wall - busy, total, total / max(busy, 1e-9), sum(1 for x in spans if x[3] < 148), sub))
        f.write("approximate critical path: %d kernels, %.1f us of kernels + %.1f us of gaps; by kernel:\n"
                % (n_crit, sum(v[1] for v in crit.values()), gap_total))
        for k, v in sorted(crit.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write("  %-46s %4d launches %9.1f us\n" % (k, v[0], v[1]))
    print(open(os.path.join(ROOT, "gpurun_out", "prof_timeline.txt")).read())
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
