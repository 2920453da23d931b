#!/bin/bash
# distance-adaptive backoff of the first-fit dataflow: config 4 sweep
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devicewide or stat or pile or giant" > $O/r20_tests.log 2>&1; tail -2 $O/r20_tests.log
for cfg in "400 400" "400 1000" "200 1000" "100 1000" "0 1000" "100 2000" "200 2000" "0 400"; do set -- $cfg
  PXB_COLOUR_BACKOFF_NS=$1 PXB_COLOUR_FAR_NS=$2 python bench.py --config 4 --steps 15 --warmup 5 --no-cpu-baseline > $O/r20_c4_$1_$2.json 2> $O/r20.err
  python - <<PY
import json
d=json.loads(open("$O/r20_c4_$1_$2.json").read().strip().splitlines()[-1])
print("config 4 near=$1 far=$2", round(d["ms_per_step"],3), d["stage_ms"]["colouring"])
PY
done
