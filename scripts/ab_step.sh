#!/bin/bash
# A/B of one environment switch on the replayed training step: step time and the CUPTI total of one kernel family
# usage: ab_step.sh <kernel substring> VAR=a VAR=b ...
K=$1; shift
for kv in "$@"; do
env $kv timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
tk=[k for k in d['step']['top_kernels'] if '$K' in k['kernel']]
print('$kv', 'ms/step', round(d['ms_per_step'],3), [(k['kernel'], k['launches'], round(k['ms'],3)) for k in tk])"
done
