#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
python - <<'PY' > $O/r8_c3_pairs.txt 2>&1
import sys; sys.path.insert(0, '.')
from physx_b200 import engine, scenes
sc = scenes.falling_primitives(128, 64, 128, kinds=("sphere", "capsule", "convex"))
g = engine.Scene(sc, max_pairs=16 * len(sc.actors))
for t in range(160):
    g.step()
    if t % 10 == 9: print(t + 1, "pairs", len(g.getPairs()), "constraints", g.num_constraints, "partitions", g.num_partitions, flush=True)
PY
cat $O/r8_c3_pairs.txt
PXB_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3700 -c 240 --csv --log-file $O/r8_launches_c3.csv python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > $O/r8_ncu_c3.log 2>&1
python tools/launch_summary.py $O/r8_launches_c3.csv 16
