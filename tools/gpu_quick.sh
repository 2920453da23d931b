#!/bin/bash
# quick GPU check: parity suite, env timing breakdown, bench (no CPU baseline)
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
[ -f physx_b200/libphysx_b200_timing.so ] && timeout 120 python tools/env_timing.py 2>&1 | tail -12
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms_per_step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value']); print(d['stage_ms']); print(d['roofline'])"
