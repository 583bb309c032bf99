"""CPU checks of bench.py's reference arm (the leg the driver runs first on the GPU box): the JSON line carries the
contract's keys, and the bounded CPU sample is chosen from the budget."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-points", "2500"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["higher_is_better"] is True
    assert line["metric"] == "points_per_sec_fwd_bwd_PCF_Normal_10cm" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]


def test_cpu_sample_follows_the_budget():
    sys.path.insert(0, ROOT)
    import bench
    import argparse
    args = argparse.Namespace(cpu_points=0, cpu_budget=0.0, points=100000, scenes=1, config="10cm")
    assert bench.choose_cpu_points(args, n_steps=1, budget_s=1e-3) == 6000          # nothing fits: the smallest sample
    assert bench.choose_cpu_points(args, n_steps=1, budget_s=1e9) == 100000         # everything fits: the full scene
