#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; T=/tmp/ncu; mkdir -p $T
export PXB_NO_GRAPH=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gjk_" -s 210 -c 3 -o $T/gjk3 -f python bench.py --config 3 --steps 8 --warmup 5 --no-cpu-baseline > $O/r18_ncu.log 2>&1; echo "rc=$?"
ncu -i $T/gjk3.ncu-rep --page raw --csv > $O/r18_gjk3_raw.csv 2>/dev/null
ncu -i $T/gjk3.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $O/r18_gjk3_source.csv.gz
ls -la $O
