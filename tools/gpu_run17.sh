#!/bin/bash
# three-phase GJK narrowphase: parity + config 3 A/B
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r17_tests.log 2>&1; tail -5 $O/r17_tests.log
for p in 1 0; do
  PXB_GJK_PHASES=$p python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline > $O/r17_c3_p$p.json 2> $O/r17_c3_p$p.err
  python - <<PY
import json
d=json.loads(open("$O/r17_c3_p$p.json").read().strip().splitlines()[-1])
print("phases=$p", d["ms_per_step"], d["stage_ms"], d["e2e"]["stage_ms"])
PY
done
