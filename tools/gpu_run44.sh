#!/bin/bash
# end-of-round build (candidate-list broadphase included): full GPU tests, smoke, one bench line per BASELINE configuration, reference arm of config 2
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r44_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r44_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r44_pytest_gpu.log; tail -3 $O/r44_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r44_smoke.log 2>&1; tail -1 $O/r44_smoke.log
timeout 600 python bench.py > $O/r44_c2_default.json 2> $O/r44_c2_default.err; echo "default rc=$?"; cut -c1-170 $O/r44_c2_default.json
for c in 5 1 4 3; do
  timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/r44_c$c.json 2> $O/r44_c$c.err; echo "config $c rc=$?"; cut -c1-170 $O/r44_c$c.json
done
timeout 600 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 10 --no-cpu-baseline > $O/r44_c2_churn.json 2> $O/r44_c2_churn.err; cut -c1-170 $O/r44_c2_churn.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r44_ref_c2.json 2> $O/r44_ref_c2.err; cut -c1-250 $O/r44_ref_c2.json
