#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu -k "local_poses" > $O/r25_tests.log 2>&1; tail -40 $O/r25_tests.log
