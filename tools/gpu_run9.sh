#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -v -k "tensor or material" > $O/pytest_sus.log 2>&1; echo "rc=$?"; grep -n "PASSED\|FAILED\|Abort\|test_gpu_parity.py\", line" $O/pytest_sus.log | head -30; grep -n "free()\|corrupt\|terminate\|what()" $O/pytest_sus.log | head
