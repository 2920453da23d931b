#!/bin/bash
# final launch lists of all configurations + source-level capture of the config-1 step kernels
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; T=/tmp/ncu; mkdir -p $T
bash tools/gpu_launch_lists.sh r28 > $O/r28_lists.log 2>&1; tail -3 $O/r28_lists.log
export PXB_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_env_solve|k_env_bp" -s 60 -c 2 -o $T/c1 -f python bench.py --config 1 --steps 30 --warmup 5 --no-cpu-baseline > $O/r28_ncu_c1.log 2>&1; echo "rc=$?"
ncu -i $T/c1.ncu-rep --page raw --csv > $O/r28_c1_raw.csv 2>/dev/null
ncu -i $T/c1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r28_c1_source.csv.gz
ls -la $O | tail -20
