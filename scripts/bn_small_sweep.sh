#!/bin/bash
# step time against the row threshold of the one-kernel BatchNorm (PCFB_BN_SMALL_ROWS; 0 = off)
for r in ${@:-0 1100 5300 8192 30000}; do
PCFB_BN_SMALL_ROWS=$r timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small rows $r ms/step', round(d['ms_per_step'],3), d['gpu_launches'])"
done
