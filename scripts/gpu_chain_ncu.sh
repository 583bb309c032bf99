#!/bin/bash
# ncu --set full capture of the chain backward kernels at the level-0 WeightNet size + the step timeline
mkdir -p gpurun_out
python scripts/time_chain.py > gpurun_out/chain_times.txt 2>&1; tail -n 14 gpurun_out/chain_times.txt
ncu --set full --clock-control none --import-source on -k regex:"mlp_bwd_fused|mlp_fwd_kernel" -c 8 -o gpurun_out/prof_chain_r02 -f python scripts/time_chain.py > gpurun_out/ncu_chain.log 2>&1; echo "ncu exit $?"
timeout 300 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; cat gpurun_out/prof_timeline.txt
