import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from physx_b200 import engine, scenes
import oracle_lib
np.set_printoptions(linewidth=220, precision=7, suppress=True)
sc = scenes.tumbling_boxes(n=12, seed=7)
gpu, cpu = engine.Scene(sc), oracle_lib.OracleScene(sc)
for t in range(150):
    gpu.step(); cpu.step()
    cg, cc = gpu.getContacts(), cpu.getContacts()
    d = np.abs(cg - cc).max(1)
    bad = np.argwhere(d > 1e-6).ravel()
    ds = np.abs(gpu.getStates() - cpu.getStates()).max()
    if len(bad) or ds > 0 or t % 25 == 0: print(t, "state diff", ds, "bad pairs", bad.tolist(), "pairs", len(cg))
    for i in bad[:2]:
        print(' pair', gpu.getPairs()[i]); print('  gpu', cg[i]); print('  cpu', cc[i])
    if len(bad): break
