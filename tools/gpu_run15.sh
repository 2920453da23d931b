#!/bin/bash
# prefetch experiment: device-wide solver, configs 3/4/1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "devicewide or stat or pile or long" > gpurun_out/r15_tests.log 2>&1; tail -3 gpurun_out/r15_tests.log
for c in 4 3 1; do
  python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r15_bench_c$c.json 2> gpurun_out/r15_bench_c$c.err; tail -c 1500 gpurun_out/r15_bench_c$c.json
done
