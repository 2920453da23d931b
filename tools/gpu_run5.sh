#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_parity.py -m gpu -x -q -k "plugin or broadphase_object or state_export or stream_ordered" > $O/pytest_plugin.log 2>&1; echo "rc=$?" >> $O/pytest_plugin.log; tail -40 $O/pytest_plugin.log
# timing of the unmodified host with and without the GPU broadphase plugin on config 2 at 1024 envs
python - <<'PY' > gpurun_out/r5_plugin_timing.txt 2>&1
import os, subprocess, tempfile, json, sys
sys.path.insert(0, '.')
from physx_b200 import scenes
sc = scenes.env_grid_stacks(n_envs=1024)
with tempfile.TemporaryDirectory() as d:
    p = d + "/s.bin"; sc.save(p)
    for extra in ([], ["--gpu-plugin", "plugin/_build/libPhysXGpu_64.so", "--gpu-bp"]):
        r = subprocess.run(["oracle/_ref_gpu/ref_harness", "run", p, "--steps", "20", "--warmup", "3", "--threads", "8"] + extra, capture_output=True, text=True)
        print(extra, r.returncode, r.stdout.strip()[-400:], r.stderr.strip()[-300:])
PY
cat gpurun_out/r5_plugin_timing.txt
