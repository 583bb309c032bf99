#!/bin/bash
# N-GPU check of bench.py under torchrun (graph mode), with a watchdog so a hung collective cannot burn the budget
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 400 $T 29512 bench.py --gpus $N --steps 10 --warmup 3 --watchdog 300 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit $?"
grep -E "bench rank 0|Error|error|Timeout" gpurun_out/bench_n$N.err | tail -n 12
timeout 400 $T 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref n$N exit $?"
python - <<PY
import json
for n in ("bench_n$N", "bench_ref_n$N"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f pts/s  ms/step %.2f  e2e %.0f  launches %d gpus %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["n_gpus"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
