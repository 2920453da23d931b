#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
for c in 2 5; do
  timeout 300 python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline > $O/r13_c${c}.json 2> $O/r13.err
  python -c "
import json
d=json.loads(open('$O/r13_c${c}.json').read().strip().splitlines()[-1]); print('config $c:', round(d['ms_per_step'],4), d['stage_ms'], 'e2e %.4g'%d['e2e']['value'])"
done
