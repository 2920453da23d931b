#!/bin/bash
# k_env_solve residency A/B: 8 CTAs/SM at 128 registers (default) vs 6 CTAs/SM at 168 registers (PXB_ENV_CTAS64=6 build, loaded through PXB_LIB)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for rep in 1 2; do for v in default v6; do for c in 2 5; do
  if [ $v = v6 ]; then export PXB_LIB=$(pwd)/physx_b200/libphysx_b200_v6.so; else unset PXB_LIB; fi
  python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline > $O/r30_c${c}_${v}_$rep.json 2> $O/r30.err
  python - <<PY
import json
d=json.loads(open("$O/r30_c${c}_${v}_$rep.json").read().strip().splitlines()[-1])
print("config $c $v rep $rep", round(d["ms_per_step"],4), d["stage_ms"]["solve"])
PY
done; done; done
