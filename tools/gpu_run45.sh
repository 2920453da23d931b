#!/bin/bash
# warp-aggregated pair emission of the device-wide broadphase: full GPU tests + configs 4 / 3 (compare with r44 of the build before it)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r45_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r45_pytest_gpu.log; tail -3 $O/r45_pytest_gpu.log
for c in 4 3; do
  timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/r45_c$c.json 2> $O/r45_c$c.err; echo "config $c rc=$?"
done
timeout 300 python bench.py --config 2 --path devicewide --steps 100 --warmup 10 --no-cpu-baseline > $O/r45_c2_dw.json 2> $O/r45_c2_dw.err
python - <<'PY'
import json
for f in ["r45_c4","r44_c4","r45_c3","r44_c3","r45_c2_dw"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],4), d.get("stage_ms"))
    except Exception as ex: print(f, "ERR", ex)
PY
