#!/bin/bash
# time the warp-specialised forward with parts switched off (PCFB_WS_DEBUG bits: 1 no gather, 2 no MMA hand-off, 4 no FMA loop)
mkdir -p gpurun_out
for d in ${DBG_LIST:-0 1 2 4 3 6 7}; do
  PCFB_WS_DEBUG=$d timeout 300 python scripts/run_op.py fwd_time 20 > gpurun_out/ablate_$d.log 2>&1
  echo "dbg=$d $(tail -n 1 gpurun_out/ablate_$d.log)"
done
