#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r33_tests.log 2>&1; tail -25 $O/r33_tests.log
python - > $O/r33_tumble.log 2>&1 <<'PY'
import sys; sys.path.insert(0, "tests")
import numpy as np, util
from physx_b200 import engine, scenes
# teacher-forced GPU vs the reference on the tumbling-box golden: every step from the reference's state
for name in ("tumble_12",):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    worst = 0.0; n_bad = 0
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t]); gpu.setConstraintOrder(util.golden_order(z, t)); gpu.step()
        e = float(np.abs(gpu.getStates()[:, :7] - z["states"][t + 1][:, :7]).max()); worst = max(worst, e); n_bad += e > 1e-5
    print(name, "GPU teacher-forced vs reference: worst pose error", worst, "steps > 1e-5:", n_bad)
PY
cat $O/r33_tumble.log
for c in 2 4; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu-baseline > $O/r33_c$c.json 2> $O/r33.err; python -c "
import json; d=json.loads(open('$O/r33_c$c.json').read().strip().splitlines()[-1]); print('config $c', d['ms_per_step'], d['stage_ms'])"; done
