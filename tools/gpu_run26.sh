#!/bin/bash
# A/B of the headline configuration on ONE box: HEAD vs the build before the box-box worklist / contact data / local poses (_ab_old), alternating; then the tests the -x stop skipped; config 1 with the CTA broadphase
cd "$(dirname "$0")/.."
R=$(pwd); O=$R/gpurun_out; mkdir -p $O
for rep in 1 2 3; do
  for v in new old; do
    if [ $v = new ]; then cd $R; else cd $R/_ab_old; fi
    python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r26_c2_${v}_$rep.json 2> $O/r26_c2_${v}_$rep.err
    python - <<PY
import json
d=json.loads(open("$O/r26_c2_${v}_$rep.json").read().strip().splitlines()[-1])
print("$v $rep", round(d["ms_per_step"],4), d["stage_ms"], round(d["e2e"]["value"]/1e6,1))
PY
  done
done
cd $R
python bench.py --config 1 --steps 200 --warmup 20 --no-cpu-baseline > $O/r26_c1.json 2> $O/r26_c1.err; python -c "
import json; d=json.loads(open('$O/r26_c1.json').read().strip().splitlines()[-1]); print('config 1', d['ms_per_step'], d['stage_ms'])"
python -m pytest tests -x -q -m gpu > $O/r26_tests.log 2>&1; tail -4 $O/r26_tests.log
