#!/bin/bash
# k_env_solve<64> with a register cap of 144 (7 CTAs/SM, -DPXB_ENV_MAXNREG=144 build loaded through PXB_LIB) against the default (128 registers, 8 CTAs/SM), config 2
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for rep in 1 2; do for v in default v144; do
  if [ $v = v144 ]; then export PXB_LIB=$(pwd)/physx_b200/libphysx_b200_v144.so; else unset PXB_LIB; fi
  python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r52_c2_${v}_$rep.json 2> $O/r52.err
  python - <<PY
import json
d=json.loads(open("$O/r52_c2_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep", round(d["ms_per_step"],4), "solve", d["stage_ms"]["solve"])
PY
done; done
