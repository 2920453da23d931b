#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
for b in 0 40 150; do for c in 4 3; do
  PXB_COLOUR_BACKOFF_NS=$b timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/r2_c${c}_b$b.json 2> $O/r2_c${c}_b$b.err; echo "config $c backoff $b rc=$?"
  python -c "
import json,sys
d=json.loads(open('$O/r2_c${c}_b$b.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), d['stage_ms'], d['details']['partitions'])"
done; done
timeout 300 python bench.py --config 1 --steps 200 --warmup 20 --no-cpu-baseline > $O/r2_c1.json 2> $O/r2_c1.err; cut -c1-200 $O/r2_c1.json
timeout 300 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 20 --no-cpu-baseline > $O/r2_c2_churn.json 2> $O/r2_c2_churn.err; python -c "
import json
d=json.loads(open('$O/r2_c2_churn.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['stage_ms'], d['details'])"
