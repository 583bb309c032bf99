#!/bin/bash
# 2-GPU check: fused-forward tests, N=1 bench, then N=2 eager and N=2 graph bench with a watchdog
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pconv.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/quick_test.log 2>&1; echo "tests exit $?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/quick_test.log | tail -n 12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 exit $?"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 330 $T 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-graph --watchdog 240 > gpurun_out/bench_n2_nograph.json 2> gpurun_out/bench_n2_nograph.err; echo "n2 nograph exit $?"
grep -E "bench rank|Error|error|File \"/root|line" gpurun_out/bench_n2_nograph.err | tail -n 40
timeout 330 $T 29512 bench.py --gpus 2 --steps 10 --warmup 3 --watchdog 240 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?"
grep -E "bench rank|Error|error|File \"/root|line" gpurun_out/bench_n2.err | tail -n 40
python - <<'PY'
import json
for n in ("bench_n1", "bench_n2_nograph", "bench_n2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f pts/s  ms/step %.2f  e2e %.0f  launches %d gpus %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["n_gpus"]))
        if n == "bench_n1":
            for k, v in d["kernels"].items():
                print("  %-24s %s" % (k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
    except Exception as e:
        print(n, "unreadable", e)
PY
