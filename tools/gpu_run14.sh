#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
PXB_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_ -s 40 -c 2 -o $O/r14_env_c5 -f python bench.py --config 5 --steps 6 --warmup 3 --no-cpu-baseline > $O/r14_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O/r14_env_c5.ncu-rep
