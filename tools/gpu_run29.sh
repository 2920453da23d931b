#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for rep in 1 2; do for p in 1 2; do for c in 2 5; do
  PXB_BOX_PHASES=$p python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline > $O/r29_c${c}_p${p}_$rep.json 2> $O/r29.err
  python - <<PY
import json
d=json.loads(open("$O/r29_c${c}_p${p}_$rep.json").read().strip().splitlines()[-1])
print("config $c boxphases=$p rep $rep", round(d["ms_per_step"],4), d["stage_ms"])
PY
done; done; done
PXB_BOX_PHASES=2 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 10 --no-cpu-baseline > $O/r29_churn_p2.json 2>> $O/r29.err; PXB_BOX_PHASES=1 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 10 --no-cpu-baseline > $O/r29_churn_p1.json 2>> $O/r29.err
python - <<PY
import json
for p in (1,2):
    d=json.loads(open("$O/r29_churn_p%d.json"%p).read().strip().splitlines()[-1]); print("churn boxphases", p, round(d["ms_per_step"],4), d["stage_ms"])
PY
