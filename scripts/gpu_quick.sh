#!/bin/bash
# usage: gpu_quick.sh "<pytest args>" [bench steps]   -- a subset of tests + optional bench
mkdir -p gpurun_out
timeout 1500 python -m pytest $1 -q -m gpu --no-header -p no:cacheprovider -rP > gpurun_out/quick_test.log 2>&1; echo "tests exit $?"
grep -E "^gemm_|passed|failed|FAILED|Error" gpurun_out/quick_test.log | tail -n 60
if [ -n "$2" ]; then
  timeout 900 python bench.py --steps $2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.0f pts/s  ms/step %.2f  e2e %.0f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
for k, v in d["kernels"].items():
    print("  %-24s %s" % (k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
PY
fi
