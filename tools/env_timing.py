"""Per-phase clock64 breakdown of k_env_solve (build with EXTRA=-DPXB_ENV_TIMING into a separate .so)."""
import ctypes, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from physx_b200 import engine, scenes
engine._LIB_NAME = "libphysx_b200_timing.so"
sc = scenes.env_grid_stacks(n_envs=4096)
g = engine.Scene(sc)
for _ in range(5): g.step()
out = np.zeros((4096, 16), np.uint64)
g._lib.pxb_debug_env_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
g._lib.pxb_debug_env_timing(g._h, out.ctypes.data_as(ctypes.c_void_p))
names = ["prelude", "compaction", "colour", "static+order", "prep", "solve", "writeback", "finalize"]
m = out[:, :8].astype(np.float64)
for i, n in enumerate(names): print(f"{n:14s} mean {m[:, i].mean():9.0f} cycles  median {np.median(m[:, i]):9.0f}")
print("total mean", m.sum(1).mean())
