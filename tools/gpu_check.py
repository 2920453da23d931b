"""Diagnostic: GPU engine vs CPU oracle on a small scene, step by step (run on the GPU box)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from physx_b200 import engine, scenes
import oracle_lib

def compare(sc, steps, tag):
    gpu = engine.Scene(sc); cpu = oracle_lib.OracleScene(sc)
    worst = 0
    for t in range(steps):
        gpu.step(); cpu.step()
        a, b = gpu.getStates(), cpu.getStates()
        pg, pc = gpu.getPairs(), cpu.getPairs()
        same = pg.shape == pc.shape and np.array_equal(pg, pc)
        cg_, cc = gpu.getContacts(), cpu.getContacts()
        cnt_same = same and np.array_equal(cg_[:, 0], cc[:, 0])
        d = np.abs(a - b)
        worst = max(worst, d.max())
        if t < 3 or t % 20 == 0 or not same or not cnt_same:
            print(f"[{tag}] step {t}: pairs {len(pg)}/{len(pc)} same={same} counts_same={cnt_same} parts {gpu.num_partitions}/{cpu.num_partitions} cons {gpu.num_constraints}/{cpu.num_constraints} "
                  f"pos {d[:, :3].max():.2e} quat {d[:, 3:7].max():.2e} lin {d[:, 7:10].max():.2e} ang {d[:, 10:].max():.2e} launches {gpu.num_launches}")
    print(f"[{tag}] worst abs diff over {steps} steps: {worst:.3e}")
    return worst

if __name__ == "__main__":
    compare(scenes.box_stacks(n_stacks=1, height=1, half_extent=0.25, spacing=1.0), 5, "1box")
    compare(scenes.box_stacks(n_stacks=1, height=2, half_extent=0.25, spacing=1.0, jitter=0.01), 5, "2box")
    compare(scenes.box_stacks(n_stacks=4, height=8, half_extent=0.25, spacing=1.0, jitter=0.01), 120, "4x8")
    compare(scenes.env_grid_stacks(n_envs=16, jitter=0.01), 60, "16env")
    # timing at config-2 scale
    sc = scenes.env_grid_stacks(n_envs=4096)
    t0 = time.time(); gpu = engine.Scene(sc); print("scene create", time.time() - t0)
    for i in range(5):
        t0 = time.time(); gpu.step(); print("step", i, (time.time() - t0) * 1e3, "ms pairs", len(gpu.getPairs()) if i == 0 else "", "cons", gpu.num_constraints, "parts", gpu.num_partitions)
    t0 = time.time()
    for i in range(20): gpu.step()
    print("20 steps avg ms", (time.time() - t0) / 20 * 1e3)
