#!/bin/bash
# last build of round 2 (kinematic bodies, aggregates, acceleration getters added): GPU tests + one bench line per BASELINE configuration.  The CPU baselines of
# configs 1 / 3 / 4 / 5 are the ones of tools/gpu_final.sh earlier in the round (same scenes, same box type; config 4 alone costs 6 minutes of host time).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r40_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r40_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r40_pytest_gpu.log; tail -3 $O/r40_pytest_gpu.log
timeout 600 python bench.py > $O/r40_c2_default.json 2> $O/r40_c2_default.err; echo "default rc=$?"; cut -c1-200 $O/r40_c2_default.json
for c in 5 1 4 3; do
  timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/r40_c$c.json 2> $O/r40_c$c.err; echo "config $c rc=$?"; cut -c1-200 $O/r40_c$c.json
done
timeout 600 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 10 --no-cpu-baseline > $O/r40_c2_churn.json 2> $O/r40_c2_churn.err; cut -c1-200 $O/r40_c2_churn.json
