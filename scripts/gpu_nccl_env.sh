#!/bin/bash
# N=2: does the small-message NCCL configuration change the SyncBatchNorm cost?  (one bench line per configuration)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
run() {  # name, env...
  name=$1; shift
  BENCH_EXTRA=""; for kv in "$@"; do case $kv in BENCH_EXTRA=*) BENCH_EXTRA=${kv#BENCH_EXTRA=};; esac; done
  env "$@" timeout 300 $T 29520 bench.py --gpus 2 --steps 10 --warmup 3 --watchdog 200 $BENCH_EXTRA > gpurun_out/nccl_$name.json 2> gpurun_out/nccl_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/nccl_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step %.2f  value %.0f" % (d["ms_per_step"], d["value"]))
except Exception as e:
    print("$name", "unreadable", e)
PY
}
run default FOO=1
run nvls_off NCCL_NVLS_ENABLE=0
run ll NCCL_NVLS_ENABLE=0 NCCL_PROTO=LL
run ll_1ch NCCL_NVLS_ENABLE=0 NCCL_PROTO=LL NCCL_MAX_NCHANNELS=2
run nosyncbn FOO=1 BENCH_EXTRA=--no-sync-bn
