#!/bin/bash
# Round-end evidence run on one B200: full -m gpu suite, smoke(), bench (both arms + the extra configs), the step timeline /
# per-kernel table of one replayed step, the kNN sweep, ncu --set full captures of the hot kernels, the ncu launch list.
R=${1:-r02}
mkdir -p gpurun_out
rm -f gpurun_out/summary.log
timeout 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$R.csv &
SMI=$!
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$R.json 2> gpurun_out/bench_reference.err; echo "bench reference exit $?" >> gpurun_out/summary.log
kill $SMI
for c in single tiny 5cm ptf2_infer; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${c}_$R.json 2> gpurun_out/bench_$c.err; echo "bench $c exit $?" >> gpurun_out/summary.log
done
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1 && cp gpurun_out/prof_table.txt gpurun_out/step_kernels_$R.txt && cp gpurun_out/prof_kernels_by_grid.txt gpurun_out/step_kernels_by_grid_$R.txt && cp gpurun_out/prof_timeline.txt gpurun_out/step_timeline_$R.txt
PCFB_PDL=0 timeout 600 python scripts/profile_step.py > gpurun_out/profile_step_nopdl.log 2>&1 && cp gpurun_out/prof_timeline.txt gpurun_out/step_timeline_nopdl_$R.txt
timeout 300 python scripts/profile_infer.py 250000 ptf2 > gpurun_out/infer_kernels_ptf2_$R.txt 2>&1; echo "profile infer ptf2 exit $?" >> gpurun_out/summary.log
timeout 300 python scripts/profile_infer.py 19000 lite > gpurun_out/infer_kernels_lite_$R.txt 2>&1; echo "profile infer lite exit $?" >> gpurun_out/summary.log
timeout 300 python bench.py --knn-sweep > gpurun_out/knn_sweep_$R.json 2> gpurun_out/knn_sweep.err; echo "knn sweep exit $?" >> gpurun_out/summary.log
for op in fwdp bwd; do
  python scripts/run_op.py $op 3 > gpurun_out/run_op.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"pconv_fwd_ws|pconv_fwd_umma2|pconv_bwd2" -s 1 -c 1 -o gpurun_out/prof_${op}_$R -f python scripts/run_op.py $op 3 > gpurun_out/ncu_$op.log 2>&1
  echo "ncu $op exit $?" >> gpurun_out/summary.log
done
python scripts/run_op.py knn 3 > gpurun_out/run_op.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn_grid_query -s 1 -c 1 -o gpurun_out/prof_knn_$R -f python scripts/run_op.py knn 3 > gpurun_out/ncu_knn.log 2>&1
echo "ncu knn exit $?" >> gpurun_out/summary.log
python scripts/run_op.py gemm_small 3 > gpurun_out/run_op.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_pipe -s 2 -c 1 -o gpurun_out/prof_gemmpipe_$R -f python scripts/run_op.py gemm_small 3 > gpurun_out/ncu_gemmpipe.log 2>&1
echo "ncu gemm pipe exit $?" >> gpurun_out/summary.log
timeout 300 python scripts/time_knn_stage.py 1.75 2.5 > gpurun_out/knn_stage_$R.txt 2>&1; echo "knn stage exit $?" >> gpurun_out/summary.log
timeout 300 python scripts/time_gemm.py > gpurun_out/gemm_times_$R.txt 2>&1; echo "gemm times exit $?" >> gpurun_out/summary.log
python scripts/time_chain.py > gpurun_out/chain_times_$R.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"mlp_bwd_fused_kernel|mlp_fwd_kernel|mlp_bwd_stats" -c 7 -o gpurun_out/prof_chain_$R -f python scripts/time_chain.py > gpurun_out/ncu_chain.log 2>&1
echo "ncu chain exit $?" >> gpurun_out/summary.log
# the launch list last and capped (~3 eager steps): compare SHARES, not absolute times (cold-cache, serialised)
timeout 900 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 15000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -n 2
