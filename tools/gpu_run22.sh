#!/bin/bash
# box-box regeneration over a compacted worklist (device-wide path): parity + config 4 / 3 A/B
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r22_tests.log 2>&1; tail -3 $O/r22_tests.log
for c in 4 3; do for p in 1 0; do
  PXB_BOX_PHASES=$p python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/r22_c${c}_b$p.json 2> $O/r22_c${c}_b$p.err
  python - <<PY
import json
d=json.loads(open("$O/r22_c${c}_b$p.json").read().strip().splitlines()[-1])
print("config $c boxphases=$p", d["ms_per_step"], d["stage_ms"])
PY
done; done
