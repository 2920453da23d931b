#!/bin/bash
# k_env_bp temporal coherence (candidate lists): GPU tests, then A/B on one box (PXB_ENV_BP_CAND=0 = all pairs every step) for configs 2 / 5 and the churn variant
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r42_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/r42_pytest_gpu.log; tail -3 $O/r42_pytest_gpu.log
for v in 1 0; do
  export PXB_ENV_BP_CAND=$v
  timeout 300 python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r42_c2_cand$v.json 2> $O/r42_c2_cand$v.err; cut -c1-170 $O/r42_c2_cand$v.json
  timeout 300 python bench.py --config 5 --steps 60 --warmup 10 --no-cpu-baseline > $O/r42_c5_cand$v.json 2> $O/r42_c5_cand$v.err; cut -c1-170 $O/r42_c5_cand$v.json
  timeout 300 python bench.py --config 2 --churn 0.05 --steps 100 --warmup 10 --no-cpu-baseline > $O/r42_c2_churn_cand$v.json 2> $O/r42_c2_churn_cand$v.err; cut -c1-170 $O/r42_c2_churn_cand$v.json
done
python - <<'PY'
import json
for f in ["c2_cand1","c2_cand0","c5_cand1","c5_cand0","c2_churn_cand1","c2_churn_cand0"]:
    try:
        d=json.loads(open(f"gpurun_out/r42_{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d.get("stage_ms"))
    except Exception as ex: print(f, "ERR", ex)
PY
