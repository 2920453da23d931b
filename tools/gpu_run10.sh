#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor" 2>&1 | tail -2
for w in 0 2 4 8; do for b in 0 100 400; do
  PXB_COLOUR_PREFIX=0 PXB_COLOUR_WINDOW=$w PXB_COLOUR_BACKOFF_NS=$b timeout 600 python bench.py --config 4 --steps 10 --warmup 3 --no-cpu-baseline > $O/r10.json 2> $O/r10.err
  python -c "
import json
d=json.loads(open('$O/r10.json').read().strip().splitlines()[-1]); print('window $w backoff $b: step', round(d['ms_per_step'],3), 'colouring', d['stage_ms']['colouring'])"
done; done
