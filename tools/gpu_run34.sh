#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r34_tests.log 2>&1; tail -8 $O/r34_tests.log
