#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
for e in export copy; do
timeout 600 python bench.py --config 2 --steps 100 --warmup 10 --no-cpu-baseline --e2e $e > $O/r3_c2_$e.json 2> $O/r3_c2_$e.err; echo "rc=$?"; tail -3 $O/r3_c2_$e.err
python -c "
import json
d=json.loads(open('$O/r3_c2_$e.json').read().strip().splitlines()[-1]); print('$e', round(d['ms_per_step'],4), d['stage_ms'], 'e2e %.4g'%d['e2e']['value'])"
done
timeout 600 python bench.py --config 5 --steps 100 --warmup 10 --no-cpu-baseline > $O/r3_c5.json 2> $O/r3_c5.err; python -c "
import json
d=json.loads(open('$O/r3_c5.json').read().strip().splitlines()[-1]); print('c5', round(d['ms_per_step'],4), d['stage_ms'], 'e2e %.4g'%d['e2e']['value'])"
timeout 300 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 20 --no-cpu-baseline > $O/r3_c2_churn.json 2> $O/r3_c2_churn.err; python -c "
import json
d=json.loads(open('$O/r3_c2_churn.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), d['stage_ms'], 'e2e %.4g'%d['e2e']['value'], d['details'])"
