#!/bin/bash
# k_env_bp candidate lists, rebuild loop with the expanded test first: environment-path tests + A/B (PXB_ENV_BP_CAND=0 = all pairs every step)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "env or broadphase or config2 or config5 or abp or kinematic or aggregates or removed or added" > $O/r43_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r43_pytest_gpu.log; tail -3 $O/r43_pytest_gpu.log
for v in 1 0; do
  export PXB_ENV_BP_CAND=$v
  timeout 300 python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r43_c2_cand$v.json 2> $O/r43_c2_cand$v.err
  timeout 300 python bench.py --config 2 --churn 0.05 --steps 200 --warmup 20 --no-cpu-baseline > $O/r43_c2_churn_cand$v.json 2> $O/r43_c2_churn_cand$v.err
  timeout 300 python bench.py --config 5 --steps 60 --warmup 10 --no-cpu-baseline > $O/r43_c5_cand$v.json 2> $O/r43_c5_cand$v.err
done
python - <<'PY'
import json
for f in ["c2_cand1","c2_cand0","c5_cand1","c5_cand0","c2_churn_cand1","c2_churn_cand0"]:
    try:
        d=json.loads(open(f"gpurun_out/r43_{f}.json").read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],4), "bp", d.get("stage_ms",{}).get("broadphase"), "e2e", round(d["e2e"]["value"]/1e6,1))
    except Exception as ex: print(f, "ERR", ex)
PY
