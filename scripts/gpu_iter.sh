#!/bin/bash
# one development iteration on one B200: selected tests, bench line, optional ncu --set full capture of one op
# usage: gpu_iter.sh "<pytest files>" <bench steps or 0> [op-for-ncu] [kernel regex] [tag]
mkdir -p gpurun_out
timeout 1200 python -m pytest $1 -q -m gpu --no-header -p no:cacheprovider > gpurun_out/quick_test.log 2>&1; echo "tests exit $?"
grep -E "passed|failed|FAILED|Error" gpurun_out/quick_test.log | tail -n 15
if [ "$2" != "0" ]; then
  timeout 900 python bench.py --steps $2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value %.0f pts/s  ms/step %.2f  e2e %.0f  launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"]))
for k, v in d["kernels"].items():
    print("  %-24s %s" % (k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
PY
fi
if [ -n "$3" ]; then
  python scripts/run_op.py $3 3 > gpurun_out/run_op.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$4" -s 1 -c 1 -o gpurun_out/prof_$3_$5 -f python scripts/run_op.py $3 3 > gpurun_out/ncu_$3.log 2>&1
  echo "ncu $3 exit $?"
fi
