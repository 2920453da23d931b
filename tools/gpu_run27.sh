#!/bin/bash
# thread-per-body first-fit colouring: parity + config 4 / 3 A/B + backoff sweep
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devicewide or stat or pile or giant or config4 or hull or primitives or all_geometry" > $O/r27_tests.log 2>&1; tail -3 $O/r27_tests.log
for c in 4 3; do for cfg in "1 100" "1 0" "1 400" "0 0"; do set -- $cfg
  PXB_COLOUR_BY_BODY=$1 PXB_COLOUR_BODY_BACKOFF_NS=$2 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/r27_c${c}_$1_$2.json 2> $O/r27.err
  python - <<PY
import json
d=json.loads(open("$O/r27_c${c}_$1_$2.json").read().strip().splitlines()[-1])
print("config $c bybody=$1 backoff=$2", round(d["ms_per_step"],3), d["stage_ms"]["colouring"], d["details"].get("partitions"))
PY
done; done
