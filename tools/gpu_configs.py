"""BASELINE configs 3 and 4 (device-wide path: no environment ids) at full size on the GPU: ms/step and per-stage times.
usage: python tools/gpu_configs.py [pile|fall|fall_hulls|fall_sb|pile_pgs] [scale] [exact]   (scale 1.0 = full size; default = PXB_FLAG_RELAXED_PARTITIONING)"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from physx_b200 import engine, scenes

def run(name, sc, max_pairs, warm, steps):
    t0 = time.time(); g = engine.Scene(sc, max_pairs=max_pairs); t_create = time.time() - t0
    for _ in range(warm): g.step()
    t0 = time.time()
    for _ in range(steps): g.step()
    wall = (time.time() - t0) / steps * 1e3
    g.setProfiling(True)
    acc = {}
    for _ in range(10):
        g.step()
        for k, v in g.getStageTimes().items(): acc[k] = acc.get(k, 0) + v / 10
    st = g.getStates()
    print(json.dumps({"config": name, "bodies": g.num_dynamic, "pairs": len(g.getPairs()), "constraints": g.num_constraints, "partitions": g.num_partitions,
                      "env_path": g.uses_env_path, "relaxed_partitioning": relaxed, "ms_per_step_wall": round(wall, 3), "body_steps_per_s": round(g.num_dynamic / wall * 1e3), "stage_ms": {k: round(v, 3) for k, v in acc.items()},
                      "finite": bool(np.isfinite(st).all()), "max_speed": float(np.abs(st[:, 7:10]).max()), "min_y": float(st[:, 1].min()), "create_s": round(t_create, 1)}))

which = sys.argv[1] if len(sys.argv) > 1 else "pile"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
relaxed = not (len(sys.argv) > 3 and sys.argv[3] == "exact")
if which.startswith("pile"):
    n = max(4, int(round(100 * scale ** (1 / 3))))
    sc = scenes.box_pile(n, max(2, int(round(20 * scale ** (1 / 3)))), n, solver=scenes.SOLVER_PGS if which.endswith("pgs") else scenes.SOLVER_TGS, relaxed_partitioning=relaxed)
    run("config 4: dense box pile " + ("PGS" if which.endswith("pgs") else "TGS"), sc, 16 * len(sc.actors), 30, 50)
else:
    n = max(4, int(round(128 * scale ** (1 / 3))))
    kinds = {"fall": ("sphere", "capsule", "box"), "fall_hulls": ("sphere", "capsule", "convex"), "fall_sb": ("sphere", "box")}[which]   # fall_hulls = BASELINE config 3 proper
    sc = scenes.falling_primitives(n, max(2, n // 2), n, kinds=kinds, relaxed_partitioning=relaxed)
    run("config 3 shape: falling " + "/".join(kinds), sc, 8 * len(sc.actors), 60, 50)
