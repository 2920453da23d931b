#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
N=${1:-2}
for g in fused peer-copy nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --gather $g > $O/r4_n${N}_$g.json 2> $O/r4_n${N}_$g.err; echo "gather $g rc=$?"; tail -2 $O/r4_n${N}_$g.err | cut -c1-300
  python -c "
import json
d=json.loads(open('$O/r4_n${N}_$g.json').read().strip().splitlines()[-1]); print('$g', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], d['details']['multi_gpu'][:120])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 5 --steps 200 --warmup 20 > $O/r4_n${N}_c5.json 2> $O/r4_n${N}_c5.err; python -c "
import json
d=json.loads(open('$O/r4_n${N}_c5.json').read().strip().splitlines()[-1]); print('c5', round(d['ms_per_step'],4), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'])"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --config 2 --steps 200 --warmup 20 --no-cpu-baseline > $O/r4_n1.json 2> $O/r4_n1.err; python -c "
import json
d=json.loads(open('$O/r4_n1.json').read().strip().splitlines()[-1]); print('n1', round(d['ms_per_step'],4), 'e2e', d['e2e'])"
