#!/bin/bash
# end-of-round build: ncu launch lists of configs 2 / 5 (the step's kernel shares with the candidate-list broadphase) + one ncu --set full capture of k_env_bp
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; T=/tmp/ncu; mkdir -p $T
export PXB_NO_GRAPH=1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio
for spec in "2 100 20" "5 100 20"; do set -- $spec
  timeout 600 ncu --metrics $M --clock-control none -s $2 -c $3 --csv --log-file $O/r46_launches_c$1.csv python bench.py --config $1 --steps 10 --warmup 5 --no-cpu-baseline > $O/r46_launches_c$1.log 2>&1; echo "c$1 rc=$?"
  python tools/launch_summary.py $O/r46_launches_c$1.csv > $O/r46_launches_c$1.summary.txt 2>&1; cat $O/r46_launches_c$1.summary.txt | head -12
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_env_bp" -s 20 -c 1 -o $T/bp -f python bench.py --config 2 --steps 10 --warmup 5 --no-cpu-baseline > $O/r46_ncu_bp.log 2>&1; echo "rc=$?"
ncu -i $T/bp.ncu-rep --page raw --csv > $O/r46_k_env_bp_c2_raw.csv 2>/dev/null
