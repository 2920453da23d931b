#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
for c in 4 3; do for pf in 1 0; do
  CUDA_VISIBLE_DEVICES=0 PXB_COLOUR_PREFIX=$pf timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/r7_c${c}_p$pf.json 2> $O/r7_c${c}_p$pf.err; echo "config $c prefix $pf rc=$?"
  python -c "
import json
d=json.loads(open('$O/r7_c${c}_p$pf.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), d['stage_ms'], d['details']['partitions'], d['details']['constraints_per_gpu'])"
done; done
for g in graph peer-copy; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --gather $g > $O/r7_n${N}_$g.json 2> $O/r7_n${N}_$g.err; echo "gather $g rc=$?"; grep -v "OMP_NUM\|^\*\*\*" $O/r7_n${N}_$g.err | tail -3 | cut -c1-300
  python -c "
import json
d=json.loads(open('$O/r7_n${N}_$g.json').read().strip().splitlines()[-1]); print('$g', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['details']['multi_gpu'][:160])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 5 --steps 200 --warmup 20 > $O/r7_n${N}_c5.json 2> $O/r7_n${N}_c5.err; python -c "
import json
d=json.loads(open('$O/r7_n${N}_c5.json').read().strip().splitlines()[-1]); print('c5', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'])"
