#!/bin/bash
# wave quantisation of k_env_solve<64> on config 2 (4096 CTAs): padding the dynamic shared memory request lowers the CTAs per SM (8 -> 7 -> 6); PXB_ENV_SMEM_PAD hook
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for rep in 1 2; do for pad in 0 3000 5000 9000; do
  PXB_ENV_SMEM_PAD=$pad python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r51_c2_pad${pad}_$rep.json 2> $O/r51.err
  python - <<PY
import json
d=json.loads(open("$O/r51_c2_pad${pad}_$rep.json").read().strip().splitlines()[-1])
print("pad $pad rep $rep", round(d["ms_per_step"],4), "solve", d["stage_ms"]["solve"])
PY
done; done
