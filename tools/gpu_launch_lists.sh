#!/bin/bash
# ncu launch lists (duration + dram bytes per launch) of one bench step window per configuration -> gpurun_out/rNN_launches_c*.csv
cd "$(dirname "$0")/.."
TAG=${1:-r21}; O=gpurun_out; mkdir -p $O
export PXB_NO_GRAPH=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio
#      config skip count
for spec in "2 100 20" "5 100 20" "1 100 20" "3 4200 130" "4 2400 120"; do set -- $spec
  timeout 900 ncu --metrics $M --clock-control none -s $2 -c $3 --csv --log-file $O/${TAG}_launches_c$1.csv python bench.py --config $1 --steps 10 --warmup 5 --no-cpu-baseline > $O/${TAG}_launches_c$1.log 2>&1; echo "c$1 rc=$?"
done
ls -la $O
