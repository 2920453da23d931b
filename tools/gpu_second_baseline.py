"""Second baseline: the reference's OWN GPU plugin (oracle/_ref_gpu/reference_plugin/libPhysXGpu_64.so, built for sm_100 by
oracle/ref_gpu_build.mk) inside the unmodified host SDK on this B200, on BASELINE configs 1 / 2 / 5-shard, next to the host's CPU path and our
plugin's GPU broadphase.  usage: python tools/gpu_second_baseline.py [out.json]"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from physx_b200 import scenes
H = os.path.join(ROOT, "oracle", "_ref_gpu", "ref_harness")
REFP = os.path.join(ROOT, "oracle", "_ref_gpu", "reference_plugin", "libPhysXGpu_64.so")
OURP = os.path.join(ROOT, "plugin", "_build", "libPhysXGpu_64.so")
threads = min(8, os.cpu_count() or 1)
out = []
for name, sc, steps in (("config 1: 10x10 unit boxes", scenes.box_stacks(), 200), ("config 2: 4096 envs x 64 boxes", scenes.env_grid_stacks(n_envs=4096), 30),
                        ("config 5 shard: 4096 envs x 128 boxes", scenes.env_grid_stacks(n_envs=4096, stacks_per_env=16), 20)):
    with tempfile.TemporaryDirectory() as d:
        p = d + "/s.bin"; sc.save(p)
        for label, extra in (("reference CPU (eABP + CPU TGS)", []), ("reference GPU plugin: eGPU broadphase + eENABLE_GPU_DYNAMICS", ["--gpu-plugin", REFP, "--gpu-bp", "--gpu-dynamics"]),
                             ("reference GPU plugin: eGPU broadphase (shift 0) + eENABLE_GPU_DYNAMICS", ["--gpu-plugin", REFP, "--gpu-bp", "--gpu-dynamics", "--gpu-bp-shift", "0"]),
                             ("reference GPU plugin: eGPU broadphase + eENABLE_GPU_DYNAMICS + eENABLE_DIRECT_GPU_API", ["--gpu-plugin", REFP, "--gpu-bp", "--gpu-dynamics", "--direct-gpu-api"]),
                             ("reference GPU plugin: eGPU broadphase (shift 0) + eENABLE_GPU_DYNAMICS + eENABLE_DIRECT_GPU_API", ["--gpu-plugin", REFP, "--gpu-bp", "--gpu-dynamics", "--direct-gpu-api", "--gpu-bp-shift", "0"]),
                             ("reference GPU plugin: eGPU broadphase, CPU dynamics", ["--gpu-plugin", REFP, "--gpu-bp"]), ("our plugin: eGPU broadphase, CPU dynamics", ["--gpu-plugin", OURP, "--gpu-bp"])):
            direct = "--direct-gpu-api" in extra   # with the direct GPU API the host objects are not updated: no state dump
            r = subprocess.run([H, "run", p, "--steps", str(steps), "--warmup", "5", "--threads", str(threads)] + ([] if direct else ["--states", d + "/st"]) + extra, capture_output=True, text=True)
            rec = {"scene": name, "arm": label, "rc": r.returncode}
            if r.returncode == 0:
                j = json.loads(r.stdout.strip().splitlines()[-1]); rec.update(ms_per_step=j["ms_per_step"], ms_median=j["ms_median"], body_steps_per_s=j["body_steps_per_s"], bodies=j["bodies"], threads=threads)
                if not direct:
                    import numpy as np
                    st = np.fromfile(d + "/st", "<f4").reshape(-1, sc.n_dynamic, 13)
                    rec["finite"] = bool(np.isfinite(st).all()); rec["max_drop_m"] = float((st[0, :, 1] - st[-1, :, 1]).max())
            else:
                rec["stderr"] = r.stderr[-600:]
            print(json.dumps(rec), flush=True); out.append(rec)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
