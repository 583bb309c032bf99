"""Per-kernel CUPTI durations of one fused MLP chain (forward + backward) at a given row count:
python scripts/time_chain.py [rows] [dims, e.g. 12,8,8,16]   (default: the level-0 WeightNet, 1.63 M edge rows)"""
import os, sys, json, re, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcf_b200  # noqa
from pcf_b200 import fused_mlp
from torch.profiler import profile, ProfilerActivity

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 102095 * 16
dims = list(map(int, sys.argv[2].split(","))) if len(sys.argv) > 2 else [12, 8, 8, 16]
need_x = len(sys.argv) > 3
torch.manual_seed(0)
mods = []
for a, b in zip(dims[:-1], dims[1:]):
    mods.append((torch.nn.Linear(a, b).cuda(), torch.nn.BatchNorm1d(b).cuda(), fused_mlp.ACT_RELU))
x = torch.randn(rows, dims[0], device="cuda", requires_grad=need_x)
go = torch.randn(rows, dims[-1], device="cuda")
def run():
    out = fused_mlp.mlp_chain(x, mods, True)
    out.backward(go)
for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
trace = os.path.join(ROOT, "gpurun_out", "chain_trace.json")
prof.export_chrome_trace(trace)
tot = 0.0
for ev in sorted((e for e in json.load(open(trace))["traceEvents"] if e.get("cat") == "kernel"), key=lambda e: e["ts"]):
    name = re.sub(r"\(.*$", "", ev["name"]).replace("void ", "").replace("pcfb::", "")
    print("%-50s grid %-14s %8.1f us" % (name[:50], ev["args"].get("grid"), ev["dur"]))
    tot += ev["dur"]
os.remove(trace)
print("rows %d dims %s total %.1f us" % (rows, dims, tot))
