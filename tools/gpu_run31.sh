#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r31_tests.log 2>&1; tail -25 $O/r31_tests.log
python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r31_c2.json 2> $O/r31.err; python -c "
import json; d=json.loads(open('$O/r31_c2.json').read().strip().splitlines()[-1]); print('config 2', d['ms_per_step'], d['stage_ms'])"
