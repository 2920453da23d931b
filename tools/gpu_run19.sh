#!/bin/bash
# warp-cooperative first-fit colouring: parity + config 4 / 3 A/B
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r19_tests.log 2>&1; tail -3 $O/r19_tests.log
for c in 4 3; do for p in 1 0; do
  PXB_COLOUR_WARP=$p python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/r19_c${c}_w$p.json 2> $O/r19_c${c}_w$p.err
  python - <<PY
import json
d=json.loads(open("$O/r19_c${c}_w$p.json").read().strip().splitlines()[-1])
print("config $c warp=$p", d["ms_per_step"], d["stage_ms"])
PY
done; done
