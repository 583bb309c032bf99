#!/usr/bin/env python
"""Turn the raw profiler outputs a GPU call left under gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/summarize_profiles.py r01

* launches_<tag>.csv  (ncu --metrics gpu__time_duration.sum --csv launch list of one bench.py step sequence)
      -> profiles/launches_<tag>_by_kernel.csv  (per kernel: launches, total / mean / share of the GPU time)
         profiles/launches_<tag>.csv.gz         (the launch list itself)
* prof_*_<tag>.ncu-rep (ncu --set full captures)  -> profiles/ncu_<name>_<tag>.txt (selected raw metrics per launch)
"""
import csv
import gzip
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct",
]


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("void ", "").replace("pcfb::", "")
    name = re.sub(r"at::native::(\(anonymous namespace\)::)?", "at::", name)
    return name[:90]


def launches(tag):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        print("no", path)
        return
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    head = rows[0]
    i_name, i_val, i_unit = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= i_val:
            continue
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        unit = r[i_unit]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        k = short(r[i_name])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    os.makedirs(PROF, exist_ok=True)
    dst = os.path.join(PROF, "launches_%s_by_kernel.csv" % tag)
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none launch list of "
                "`bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline`, aggregated per kernel "
                "(all launches of the process: warm-up + timed steps + the per-kernel roofline section; cold-cache, serialised)\n")
        f.write("kernel,launches,total_us,mean_us,share_pct\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.2f,%.2f\n' % (k, n, us, us / n, 100.0 * us / total))
        f.write('"TOTAL",%d,%.1f,,100.0\n' % (sum(a[0] for a in agg.values()), total))
    with gzip.open(os.path.join(PROF, "launches_%s.csv.gz" % tag), "wt") as f:
        f.write("".join(lines))
    print("wrote", dst, "kernels:", len(agg), "total ms %.2f" % (total / 1000.0))


def ncu_report(path, name, tag):
    cmd = ["ncu", "-i", path, "--page", "raw", "--csv"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        print("ncu failed on", path, res.stderr[:300])
        return
    rows = list(csv.reader(io.StringIO(res.stdout)))
    head, units = rows[0], rows[1]
    dst = os.path.join(PROF, "ncu_%s_%s.txt" % (name, tag))
    with open(dst, "w") as f:
        f.write("# extracted with `ncu -i %s --page raw --csv` from an `ncu --set full --clock-control none "
                "--import-source on` capture\n" % os.path.basename(path))
        for r in rows[2:]:
            if len(r) < len(head):
                continue
            d = dict(zip(head, r))
            f.write("\n== launch %s  %s  grid %s block %s\n" % (d.get("ID"), short(d.get("Kernel Name", "")),
                                                              d.get("Grid Size"), d.get("Block Size")))
            for m in RAW_METRICS:
                if m in d:
                    f.write("%-82s %s %s\n" % (m, d[m], units[head.index(m)]))
    print("wrote", dst)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    for fn in sorted(os.listdir(OUT)):
        m = re.match(r"prof_(.+)_%s\.ncu-rep$" % re.escape(tag), fn)
        if m:
            ncu_report(os.path.join(OUT, fn), m.group(1), tag)
    for fn in ("bench_%s.json" % tag, "bench_reference_%s.json" % tag, "step_kernels_%s.txt" % tag, "step_kernels_by_grid_%s.txt" % tag,
               "knn_sweep_%s.json" % tag, "step_timeline_%s.txt" % tag, "step_timeline_nopdl_%s.txt" % tag, "chain_times_%s.txt" % tag,
               "knn_stage_%s.txt" % tag, "gemm_times_%s.txt" % tag, "infer_kernels_ptf2_%s.txt" % tag, "infer_kernels_lite_%s.txt" % tag,
               "bench_single_%s.json" % tag, "bench_tiny_%s.json" % tag, "bench_5cm_%s.json" % tag, "bench_ptf2_infer_%s.json" % tag):
        src = os.path.join(OUT, fn)
        if os.path.exists(src):
            with open(src) as f, open(os.path.join(PROF, fn), "w") as g:
                g.write(f.read())
    clk = os.path.join(OUT, "clocks_%s.csv" % tag)
    if os.path.exists(clk):
        lines = open(clk).read().splitlines()
        with open(os.path.join(PROF, "clocks_%s.csv" % tag), "w") as g:
            g.write("\n".join(lines[:1] + lines[1::max(1, len(lines) // 40)]) + "\n")


if __name__ == "__main__":
    main()
