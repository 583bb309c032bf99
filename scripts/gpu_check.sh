#!/bin/bash
# One GPU session: hardware self-test of the tcgen05 conventions, the -m gpu parity tests (one process per
# file so that a faulting kernel cannot poison the other files), smoke(), and a short bench run.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.log 2>&1
for t in test_umma_selftest test_gpu_knn test_gpu_pconv test_gpu_layers; do
  timeout 1200 python -m pytest tests/$t.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log
for f in gpurun_out/*.log; do echo "== $f"; tail -n 5 $f; done
timeout 600 python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1; echo "profile exit $?" >> gpurun_out/summary.log
