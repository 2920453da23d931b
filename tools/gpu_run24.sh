#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
python - > $O/r24_local.log 2>&1 <<'PY'
import sys; sys.path.insert(0, "tests")
import numpy as np, util, oracle_lib
from physx_b200 import engine, scenes
for name in ("local_poses_mix", "pgs_local_poses_mix"):
    z, sc = util.load_golden(name)
    for env_path in (True, False):
        gpu = engine.Scene(sc, env_path=env_path); o = oracle_lib.OracleScene(sc)
        print(name, "env" if gpu.uses_env_path else "devicewide")
        for t in range(40):
            gpu.step(); o.step()
            a, b = gpu.getStates(), o.getStates()
            ca, cb = gpu.getContacts(), o.getContacts()
            if t < 12 or t % 5 == 0:
                print("  step", t, "state diff", float(np.abs(a - b).max()), "contacts equal", np.array_equal(ca, cb), "counts equal", np.array_equal(ca[:, 0], cb[:, 0]), "max contact diff", float(np.abs(ca - cb).max()) if ca.shape == cb.shape else None)
        # teacher-forced: same states in, one step
        worst = 0
        for t in range(60):
            st = z["states"][t]
            gpu.setStates(st); o.setStates(st)
            gpu.step(util.golden_order(z, t)) if False else None
PY
cat $O/r24_local.log | tail -70
