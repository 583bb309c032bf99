#!/bin/bash
# kNN iteration: parity tests, bench line, the N x K sweep
mkdir -p gpurun_out
bash scripts/gpu_quick.sh "tests/test_gpu_knn.py tests/test_gpu_eval.py" ${1:-10} 2>&1 | grep -v "^  [a-z_0-9]* *{.ms.: [0-9.]*, .alg"
timeout 200 python bench.py --knn-sweep 2>/dev/null > gpurun_out/knn_sweep_new.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/knn_sweep_new.json").read().strip().splitlines()[-1])
for r in d["rows"]:
    if r["cloud"] == "room" or r["N"] == 100000:
        print(r["cloud"], r["N"], r["K"], round(r["grid_ms"], 3), round(r["grid_Mpts_per_s"], 1), r.get("grid_equals_brute"))
PY
