#!/bin/bash
# round-2 baseline on one B200: GPU tests, then one bench line per BASELINE config (+ churn, relaxed partitioning for comparison)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
for c in 2 5 1 4 3; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 > $O/bench_c$c.json 2> $O/bench_c$c.err; echo "config $c rc=$?"; cut -c1-400 $O/bench_c$c.json
done
timeout 600 python bench.py --config 2 --churn 0.05 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_c2_churn.json 2> $O/bench_c2_churn.err; cut -c1-300 $O/bench_c2_churn.json
for c in 4 3; do timeout 600 python bench.py --config $c --partitioning relaxed --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c${c}_relaxed.json 2> $O/bench_c${c}_relaxed.err; cut -c1-300 $O/bench_c${c}_relaxed.json; done
