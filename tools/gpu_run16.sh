#!/bin/bash
# full ncu captures of the dominant kernel of configs 3, 4, 2 (dram bytes for roofline.traffic + source pages); CSV pages extracted on the box
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; T=/tmp/ncu; mkdir -p $T
export PXB_NO_GRAPH=1
cap() {  # name, kernel regex, skip, count, config, steps
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o $T/$1 -f python bench.py --config $5 --steps $6 --warmup 5 --no-cpu-baseline > $O/r16_ncu_$1.log 2>&1; echo "$1 rc=$?"
  ncu -i $T/$1.ncu-rep --page raw --csv > $O/r16_$1_raw.csv 2>/dev/null
  ncu -i $T/$1.ncu-rep --page source --csv > $O/r16_$1_source.csv 2>/dev/null
}
cap gjk_c3 k_narrowphase_gjk 70 1 3 8
cap solve_c4 "k_solve_tgs|k_colour_firstfit|k_prep_rows" 120 3 4 8
cap env_c2 k_env_solve 20 1 2 30
cap c1 "k_env_|k_narrowphase|k_solve|k_colour|k_prep" 40 12 1 30
rm -f $O/r16_c1_source.csv
ls -la $O/
