#!/bin/bash
# lean 8-GPU run: config 2 and config 5 with the default exchange (peer copy + consumer release)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for c in 2 5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --config $c --steps 200 --warmup 20 --no-cpu-baseline > $O/final_n8_c$c.json 2> $O/final_n8_c$c.err
  echo "config $c rc=$?"; tail -c 700 $O/final_n8_c$c.json
done
