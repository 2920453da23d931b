#!/bin/bash
# final pass of the round on one B200: GPU tests, one bench line per BASELINE configuration (our arm with cpu_baseline + second_baseline, then the reference arm), churn variant
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/final_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/final_pytest_gpu.log; tail -3 $O/final_pytest_gpu.log
timeout 600 python bench.py > $O/final_c2_default.json 2> $O/final_c2_default.err; echo "default rc=$?"; cut -c1-300 $O/final_c2_default.json
for c in 5 1 4 3; do
  timeout 900 python bench.py --config $c --steps 30 --warmup 5 > $O/final_c$c.json 2> $O/final_c$c.err; echo "config $c rc=$?"; cut -c1-300 $O/final_c$c.json
done
timeout 600 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 10 --no-cpu-baseline > $O/final_c2_churn.json 2> $O/final_c2_churn.err; cut -c1-300 $O/final_c2_churn.json
timeout 600 python bench.py --config 2 --solver pgs --steps 100 --warmup 10 --no-cpu-baseline > $O/final_c2_pgs.json 2> $O/final_c2_pgs.err; cut -c1-300 $O/final_c2_pgs.json
for c in 2 1; do
  timeout 600 python bench.py --impl reference --config $c --steps 3 --warmup 1 > $O/final_ref_c$c.json 2> $O/final_ref_c$c.err; echo "ref config $c rc=$?"; cut -c1-300 $O/final_ref_c$c.json
done
ls -la $O | head -40
