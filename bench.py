#!/usr/bin/env python
"""bench.py -- rigid-body-steps/s of the simulate()+fetchResults() hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--churn F] [--impl ours|reference]

Workloads (BASELINE.json `configs`; synthetic seeded scenes, TGS 4 position + 1 velocity iteration unless --solver pgs, 60 Hz):
  1  10 stacks x 10 unit boxes on a ground plane (100 bodies; the reference's own CPU-runnable case)
  2  4096 independent envs x 64 boxes = 262 144 bodies per GPU  (DEFAULT: the configuration the metric is quoted on)
  3  1 048 576 mixed spheres / capsules / convex hulls falling into a walled bin (broadphase + narrowphase stress)
  4  200 000-box dense pile in a walled bin, one giant island (solver partitioning stress)
  5  32 768 envs x 128 boxes env-partitioned over 8 GPUs = 4096 envs x 128 boxes per GPU (--config 5 at N=1 is one GPU's shard)
Our arm: one process per GPU (torchrun for N > 1), every rank owns its own scene (weak scaling, no physics coupling); configs 2 / 5
all-gather the Direct-GPU-API state tensor over NVLink every step.  One JSON line is printed by rank 0.

Reference arm (--impl reference): the UNMODIFIED reference CPU SDK (oracle/_ref/ref_harness: eABP broadphase, CPU TGS,
PxDefaultCpuDispatcher) on the SAME scene -- configs 1 / 2 / 4 / 5 at full size, config 3 at 1/8 of the bodies (stated in the line: the
metric is per body) -- with the dispatcher thread count chosen by a short sweep over {1, 8, nproc/2, nproc-1}; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rigid-body-steps/sec"
UNIT = "bodies*steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of each config's dominant kernel, from the committed `ncu --set full` captures under
# profiles/ (profiler numbers: quoted, never timed under ncu).  None = not captured for that configuration.
NCU_DRAM_BYTES_PER_LAUNCH = {}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")) as _f:
        NCU_DRAM_BYTES_PER_LAUNCH = json.load(_f)
except Exception:
    pass


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------------------------
# workloads
SETTLE = {1: 0, 2: 0, 3: 60, 4: 30, 5: 0}   # untimed steps before the warm-up: config 3's lattice is falling into the bin, config 4's pile closes its 1 mm gaps


def build_scene(config, envs, rank=0, solver="tgs", scale=1.0, relaxed=False, world=1):
    """The scene one GPU (our arm) or the host (reference arm, `world` GPUs' worth of environments) simulates."""
    from physx_b200 import scenes
    sv = scenes.SOLVER_PGS if solver == "pgs" else scenes.SOLVER_TGS
    if config == 1:
        return scenes.box_stacks(solver=sv), 0
    if config in (2, 5):
        return scenes.env_grid_stacks(n_envs=envs * world, stacks_per_env=8 if config == 2 else 16, seed=1234 + rank, solver=sv), 0
    if config == 3:
        n = max(4, int(round(128 * scale ** (1 / 3))))
        sc = scenes.falling_primitives(n, max(2, n // 2), n, kinds=("sphere", "capsule", "convex"), solver=sv, relaxed_partitioning=relaxed)
        return sc, 16 * len(sc.actors)   # the settling pile reaches ~9 broadphase pairs per body
    if config == 4:
        n = max(4, int(round(100 * scale ** (1 / 3))))
        sc = scenes.box_pile(n, max(2, int(round(20 * scale ** (1 / 3)))), n, solver=sv, relaxed_partitioning=relaxed)
        return sc, 16 * len(sc.actors)
    raise SystemExit(f"unknown config {config}")


def workload_string(config, envs, solver, churn):
    s = solver.upper()
    w = {1: f"config 1: 10 stacks x 10 unit boxes on a ground plane = 100 bodies, {s} 4 pos/1 vel iterations, 60 Hz",
         2: f"config 2: {envs} envs x 64 boxes = {envs * 64} bodies per GPU, shared ground plane, GPU broadphase + {s} 4 pos/1 vel iterations, 60 Hz",
         3: f"config 3: 1048576 mixed spheres / capsules / convex hulls falling into a walled bin, {s} 4 pos/1 vel iterations, 60 Hz",
         4: f"config 4: dense pile of 200000 boxes (100 x 20 x 100, 1 mm gaps) in a walled bin, one island, {s} 4 pos/1 vel iterations, 60 Hz",
         5: f"config 5: {envs} envs x 128 boxes = {envs * 128} bodies per GPU (32768 envs over 8 GPUs), shared ground plane, GPU broadphase + {s} 4 pos/1 vel iterations, per-step state all-gather, 60 Hz"}[config]
    if churn > 0:
        w += f"; churn: {churn:.0%} of the environments reset every step through pxb_set_rigid_dynamic_data_device"
    return w


# ---------------------------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference CPU SDK on the host cores
def run_harness(scene, steps, threads, warmup):
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return None
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.bin")
        scene.save(p)
        out = subprocess.run([harness, "run", p, "--steps", str(steps), "--warmup", str(warmup), "--threads", str(threads)], capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def thread_candidates():
    n = os.cpu_count() or 1
    return sorted({t for t in (1, 8, n // 2, n - 1, n) if 1 <= t <= n})


def sweep_threads(config, solver):
    """Short sweep of PxDefaultCpuDispatcher thread counts on a small scene of the same shape (T in {1, 8, nproc-1}); returns the
    best count and the table."""
    small, _ = {1: lambda: build_scene(1, 0, solver=solver), 2: lambda: build_scene(2, 512, solver=solver), 5: lambda: build_scene(5, 256, solver=solver),
                3: lambda: build_scene(3, 0, solver=solver, scale=1 / 64), 4: lambda: build_scene(4, 0, solver=solver, scale=1 / 32)}[config]()
    table = {}
    for t in thread_candidates():
        r = run_harness(small, 6 if config != 1 else 200, t, 2 + (SETTLE[config] if config == 3 else 0))
        if r is None:
            return None, {}
        table[t] = r["body_steps_per_s"]
    return max(table, key=table.get), table


def reference_run(config, envs, solver, steps, warmup, world=1):
    """(harness result, description) of the reference CPU SDK on this arm's workload.  Config 3: 1/8 of the bodies, stated."""
    best, table = sweep_threads(config, solver)
    if best is None:
        return None, None, None
    scale = 1 / 8 if config == 3 else 1.0
    sc, _ = build_scene(config, envs, solver=solver, scale=scale, world=world)
    r = run_harness(sc, steps, best, warmup + SETTLE[config])
    sample = (f"the same scene at {'1/8 of the bodies' if scale != 1.0 else 'FULL size'} ({sc.n_dynamic} bodies), {steps} steps after {warmup + SETTLE[config]} untimed, unmodified PhysX 5.6.1 CPU "
              f"(eABP + CPU {solver.upper()} 4+1, PxDefaultCpuDispatcher({best}) = best of the sweep {json.dumps({str(k): round(v) for k, v in table.items()})} body-steps/s on a small scene of the same shape; {os.cpu_count()} host threads available)")
    return r, best, sample


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    t0 = time.time()
    steps = min(args.steps, 30) if args.config != 1 else args.steps
    r, threads, sample = reference_run(args.config, args.envs, args.solver, steps, args.warmup, world=world if args.config in (2, 5) else 1)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness is not built (needs /root/reference at build time)"}))
        return
    v = r["body_steps_per_s"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.config, args.envs, args.solver, 0.0)},
            "details": {"implementation": "unmodified PhysX 5.6.1 CPU path: eABP broadphase, CPU TGS/PGS, PxDefaultCpuDispatcher", "threads": threads, "timed_steps": steps, "bodies": r["bodies"],
                        "ms_median": r.get("ms_median"), "ms_p95": r.get("ms_p95")},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


def second_baseline_block(config, envs, solver):
    """SURVEY 8d "second baseline": the reference's OWN GPU plugin (built for sm_100 from its sources by oracle/ref_gpu_build.mk) inside the unmodified
    host SDK on this GPU -- PxBroadPhaseType::eGPU + PxSceneFlag::eENABLE_GPU_DYNAMICS through the public PxScene API, simulate + fetchResults wall
    time on the same scene.  Reported next to `cpu_baseline`; None when the binaries are not there or the config is not one it is run on."""
    harness = os.path.join(ROOT, "oracle", "_ref_gpu", "ref_harness")
    plugin = os.path.join(ROOT, "oracle", "_ref_gpu", "reference_plugin", "libPhysXGpu_64.so")
    if config not in (1, 2, 5) or not (os.path.exists(harness) and os.path.exists(plugin)):
        return None
    sc, _ = build_scene(config, envs, solver=solver)
    threads = min(8, os.cpu_count() or 1)
    steps = 20 if config != 1 else 200
    try:
        out = {}
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "s.bin")
            sc.save(p)
            # two of the reference's modes: the SDK reads the state back to its host objects every step (what `e2e` is compared with), and
            # PxSceneFlag::eENABLE_DIRECT_GPU_API, where the state stays on the device (what `value` is compared with)
            for key, extra in (("host_readback", []), ("direct_gpu_api", ["--direct-gpu-api"])):
                r = subprocess.run([harness, "run", p, "--steps", str(steps), "--warmup", "5", "--threads", str(threads), "--gpu-plugin", plugin, "--gpu-bp", "--gpu-dynamics"] + extra,
                                   capture_output=True, text=True, timeout=600)
                if r.returncode != 0:
                    return {"value": None, "error": r.stderr[-300:]}
                out[key] = json.loads(r.stdout.strip().splitlines()[-1])
        j, jd = out["host_readback"], out["direct_gpu_api"]
        return {"value": jd["body_steps_per_s"], "unit": UNIT, "ms_per_step": jd["ms_per_step"], "kind": "reference GPU plugin (PhysX 5.6.1 libPhysXGpu_64.so, sm_100 build) in the unmodified host SDK",
                "with_host_readback": {"value": j["body_steps_per_s"], "ms_per_step": j["ms_per_step"]},
                "sample": f"the same scene at full size ({j['bodies']} bodies), {steps} steps after 5 untimed, eGPU broadphase (default PxGpuBroadPhaseDesc) + eENABLE_GPU_DYNAMICS, "
                          f"PxDefaultCpuDispatcher({threads}); value: with eENABLE_DIRECT_GPU_API (state stays on the device), with_host_readback: the SDK updates its host objects every step"}
    except Exception as e:   # pragma: no cover
        return {"value": None, "error": repr(e)[:300]}


def cpu_baseline_block(config, envs, solver):
    """Bounded sample for the `cpu_baseline` key of our own line (N = 1 only): 10 timed steps of the reference arm's run."""
    r, threads, sample = reference_run(config, envs, solver, 10 if config != 1 else 300, 3)
    if r is None:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "reference", "sample": "oracle/_ref/ref_harness not built"}
    return {"value": r["body_steps_per_s"], "unit": UNIT, "cores": threads, "kind": "reference", "ms_per_step": r["ms_per_step"], "sample": sample}


# ---------------------------------------------------------------------------------------------------------------------------------------
def bench_ours(args):
    import numpy as np
    import torch
    from physx_b200 import engine, scenes, multi_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    cfg, n_envs = args.config, args.envs
    relaxed = args.partitioning == "relaxed"
    sc, max_pairs = build_scene(cfg, n_envs, rank=rank, solver=args.solver, relaxed=relaxed)   # every rank owns its own scene (weak scaling)
    scene = engine.Scene(sc, device=local, env_path=(args.path != "devicewide"), max_pairs=max_pairs)
    nb = scene.num_dynamic
    stream = torch.cuda.ExternalStream(scene.stream(), device=dev)
    gather, gather_kind = None, None
    if dist is not None and cfg in (2, 5):
        # one packed [n, 13] tensor (pose + linear + angular velocity) exchanged every step, double-buffered on communication streams
        gather, gather_kind = multi_gpu.make_state_gather(dist, nb, 13, dev, stream, kind=args.gather, scene=scene)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    # churn (RL resets): every step a rotating window of environments is put back to a lifted copy of its initial state -- the boxes of a stack
    # separated by more than the contact offset -- through the Direct-GPU-API device write, so pairs are lost and found, manifolds rebuilt
    # and the environments recoloured all the time.
    churn = None
    if args.churn > 0 and cfg in (2, 5):
        per_env = nb // n_envs
        k = max(1, int(round(args.churn * n_envs)))
        st0 = scene.getStates()
        level = (np.arange(nb) % per_env) % 8
        pose0 = np.concatenate([st0[:, 3:7], st0[:, 0:3]], axis=1).astype(np.float32)
        pose0[:, 5] += 0.045 * (level + 1)
        churn = {"k": k, "per_env": per_env, "cursor": 0,
                 "pose": torch.from_numpy(pose0).to(dev), "zero": torch.zeros((nb, 3), dtype=torch.float32, device=dev),
                 "idx": torch.arange(nb, dtype=torch.int32, device=dev)}

    def apply_churn():
        if churn is None:
            return 0
        k, pe = churn["k"], churn["per_env"]
        e0 = churn["cursor"]; churn["cursor"] = (e0 + k) % n_envs
        e1 = min(e0 + k, n_envs)
        lo, hi = e0 * pe, e1 * pe
        n = hi - lo
        with torch.cuda.stream(stream):
            idx = churn["idx"][lo:hi]
            scene.setRigidDynamicDataDevice(engine.RD_GLOBAL_POSE, churn["pose"][lo:hi].data_ptr(), n, idx.data_ptr())
            scene.setRigidDynamicDataDevice(engine.RD_LINEAR_VELOCITY, churn["zero"][lo:hi].data_ptr(), n, idx.data_ptr())
            scene.setRigidDynamicDataDevice(engine.RD_ANGULAR_VELOCITY, churn["zero"][lo:hi].data_ptr(), n, idx.data_ptr())
        return 3

    def one_step():
        apply_churn()
        if gather is not None:
            gather.pre_step()
        scene.simulate()
        if gather is not None:
            gather.step(lambda view: scene.getStatesDevice(view.data_ptr()))
        scene.fetchResults(True)

    warm = max(args.warmup, 3)
    for _ in range(SETTLE[cfg] + warm + (40 if churn else 0)):
        one_step()
    torch.cuda.synchronize(dev)

    # ---- device-timed region: K steps, CUDA events on the scene stream, L2 flushed between steps ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    step_ms, stage_acc, launches = [], {}, 0
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches += apply_churn()
        if gather is not None:
            gather.pre_step()
        scene.simulate()
        if gather is not None:
            gather.step(lambda view: scene.getStatesDevice(view.data_ptr()))   # fused: only raises this rank's flag; copy-based kinds: exchange on communication streams
        e1.record(stream)
        scene.fetchResults(True)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        launches += scene.num_launches + (1 if gather is not None else 0)
    if gather is not None:
        # the exchanges of the last two steps may still be in flight: their completion is part of the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gather.wait()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    torch.cuda.synchronize(dev)
    n_pairs = len(scene.getPairs())
    churn_stats = {"pairs_created_last_timed_step": len(scene.getCreatedPairs()), "pairs_deleted_last_timed_step": len(scene.getDeletedPairs()), "constraints_last_timed_step": scene.num_constraints}
    if dist is not None:
        dist.barrier()
    if gather is not None:
        # validity of the exchange (outside the timed region): the gathered tensor must equal a plain NCCL all-gather of every rank's packed state
        mine = torch.empty((nb, 13), dtype=torch.float32, device=dev)
        scene.getStatesDevice(mine.data_ptr())
        torch.cuda.current_stream(dev).wait_stream(stream)
        ref = torch.empty((world * nb, 13), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ref, mine)
        torch.cuda.synchronize(dev)
        if not torch.equal(ref, gather.latest()):
            raise RuntimeError(f"rank {rank}: gathered state tensor differs from the NCCL reference all-gather")
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(step_ms))
    per_step = sorted(step_ms[:args.steps])
    # ---- per-stage device times (CUDA events inside the engine, direct launches) for the roofline of the
    #      dominant kernel; separate from the timed region, which replays the captured CUDA graph ----
    scene.setProfiling(True)
    prof_steps = 20 if cfg in (1, 2, 5) else 8
    for _ in range(prof_steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        apply_churn()
        scene.simulate()
        scene.fetchResults(True)
        for k, v in scene.getStageTimes().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    scene.setProfiling(False)
    P = scene.num_constraints            # contact patches = touching pairs (one patch per pair)
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        nbt = torch.tensor([nb], dtype=torch.int64, device=dev)
        dist.all_reduce(nbt)
        total_bodies = int(nbt.item())
    else:
        total_bodies = nb
    value = total_bodies * args.steps / (total_ms / 1e3)

    if gather is not None and hasattr(gather, "close"):
        gather.close()
    # ---- end-to-end through the public API with HOST buffers (pinned): per step H2D of the velocity "action" tensors (stream-ordered
    #      pxb_set_rigid_dynamic_data_async: the copy overlaps bounds / broadphase / narrowphase), simulate + fetchResults, and the step's packed
    #      state block (pose + velocities, 52 bytes per body) stored by the step itself into MAPPED PINNED host memory (pxb_scene_set_state_export):
    #      the device-to-host transfer overlaps the solve instead of following it. ----
    state_h = torch.empty((nb, 13), dtype=torch.float32).pin_memory()
    lin_h = torch.empty((nb, 3), dtype=torch.float32).pin_memory()
    ang_h = torch.empty((nb, 3), dtype=torch.float32).pin_memory()
    lib, h = scene._lib, scene._h
    lib.pxb_get_rigid_dynamic_data(h, lin_h.data_ptr(), None, engine.RD_LINEAR_VELOCITY, nb)
    lib.pxb_get_rigid_dynamic_data(h, ang_h.data_ptr(), None, engine.RD_ANGULAR_VELOCITY, nb)
    e2e_steps = max(5, min(args.steps, 50))
    e2e_api = "pxb_set_rigid_dynamic_data_async(lin,ang) -> pxb_scene_simulate (state export into mapped pinned host memory) -> pxb_scene_fetch_results (one host sync per step), pinned host buffers"
    if args.e2e == "copy":
        pose_h = torch.empty((nb, 7), dtype=torch.float32).pin_memory()
        e2e_api = "pxb_set_rigid_dynamic_data_async(lin,ang) -> pxb_scene_simulate -> pxb_get_rigid_dynamic_data_async(pose,lin,ang) -> pxb_scene_fetch_results (one host sync per step), pinned host buffers"
    else:
        scene.setStateExport([state_h.data_ptr()], 0)
    for _ in range(3):   # untimed: graph capture of the export variant of the step
        scene.simulate(); scene.fetchResults(True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # stream-ordered calls on pinned host buffers, ONE host synchronisation per step (fetchResults), as with PxDirectGPUAPI's events
        apply_churn()
        lib.pxb_set_rigid_dynamic_data_async(h, lin_h.data_ptr(), engine.RD_LINEAR_VELOCITY, nb)   # H2D of the action tensors (export leg: the resting velocities read before the loop; copy leg: the values read back last step)
        lib.pxb_set_rigid_dynamic_data_async(h, ang_h.data_ptr(), engine.RD_ANGULAR_VELOCITY, nb)
        scene.simulate()
        if args.e2e == "copy":
            lib.pxb_get_rigid_dynamic_data_async(h, pose_h.data_ptr(), engine.RD_GLOBAL_POSE, nb)      # D2H
            lib.pxb_get_rigid_dynamic_data_async(h, lin_h.data_ptr(), engine.RD_LINEAR_VELOCITY, nb)
            lib.pxb_get_rigid_dynamic_data_async(h, ang_h.data_ptr(), engine.RD_ANGULAR_VELOCITY, nb)
        scene.fetchResults(True)                                                                   # waits for the step (and the copies)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_stage = {}
    scene.setProfiling(True)
    for _ in range(10):   # per-stage device times of the step as the end-to-end leg runs it (export on): how much the in-kernel transfer costs the solve
        scene.simulate(); scene.fetchResults(True)
        for k, v in scene.getStageTimes().items():
            e2e_stage[k] = e2e_stage.get(k, 0.0) + v / 10
    scene.setProfiling(False)
    if args.e2e != "copy":
        chk = scene.getStates()
        if not np.array_equal(chk, state_h.numpy()):
            raise RuntimeError("exported host state differs from pxb_scene_get_states")
        scene.setStateExport(())
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_bodies * e2e_steps / e2e_s
    assert np.isfinite(state_h.numpy() if args.e2e != "copy" else pose_h.numpy()).all()

    if rank == 0:
        peak, peak_src = measured_peaks()
        iters = int(sc.header["posIters"]) + int(sc.header["velIters"])
        stages = {k: round(v / prof_steps, 4) for k, v in stage_acc.items()}
        # SURVEY.md 8d algorithmic bytes.  Contact points / friction rows per patch: 4 / 4 for the box scenes (face manifolds, 2 anchors x 2
        # tangents); config 3's mix averages about 1.5 points per touching pair (spheres / capsules: 1-2, hulls: up to 4) -> C = F = 2 P is used.
        C = F = (4 * P if cfg != 3 else 2 * P)
        solve_bytes = (116 * P + 60 * C + 48 * F + 72 * nb) * iters
        prep_bytes = (64 * P + 16 * C + 224 * P) + (116 * P + 56 * C + 44 * F)
        integ_bytes = 132 * nb
        np_bytes = n_pairs * (16 + 2 * 48 + 2 * 32 + 2 * 256 + 32) + 64 * P + 16 * C
        key = None
        if scene.uses_env_path:
            # environment path: ONE kernel does pre-integration, colouring, prep, all solver iterations, write-back and integration
            kernel_name, stage = "k_env_solve (a12-a18 fused: prep + all solver iterations + integration, environments on chip, rows in registers)", "solve"
            kernel_bytes = prep_bytes + solve_bytes + integ_bytes
            note = ("algorithmic bytes = SURVEY 8d prep + 5 solver iterations + integration; the kernel keeps every solver row on chip (registers), "
                    "so only contacts, friction patches and body state cross HBM once: see `traffic` (ncu dram bytes per launch)")
            key = f"k_env_solve/config{cfg}" if ((n_envs == 4096 or cfg == 1) and args.solver == "tgs" and not churn) else None
        elif cfg == 3:
            kernel_name, stage = "k_narrowphase + k_gjk_refresh / k_gjk_query / k_gjk_epa / k_gjk_manifold (a8-a11: contact generation over all broadphase pairs)", "narrowphase"
            kernel_bytes = np_bytes
            note = "algorithmic bytes = SURVEY 8d narrowphase: per pair 16 + 2*48 + 2*32 read, 2*256 manifold r/w, 32 + 64 P + 16 C written"
            key = "narrowphase/config3" if args.partitioning == "exact" else None
        else:
            kernel_name, stage = "k_solve_tgs / k_solve_pgs (all solver iterations, one cooperative launch)", "solve"
            kernel_bytes = solve_bytes
            note = "device-wide path: rows re-streamed from HBM every iteration"
            key = f"k_solve/config{cfg}" if args.solver == "tgs" else None
        kernel_ms = stage_acc[stage] / prof_steps
        traffic = NCU_DRAM_BYTES_PER_LAUNCH.get(key) if key else None
        achieved = kernel_bytes / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(cfg, n_envs, args.solver, args.churn if churn else 0.0)},
            "details": {"bodies_total": total_bodies, "pairs_per_gpu": n_pairs, "constraints_per_gpu": P, "partitions": scene.num_partitions, "path": "environment (pxb_env.cuh)" if scene.uses_env_path else "device-wide",
                        "partitioning": ("relaxed (Jones-Plassmann rounds)" if relaxed else "exact first-fit (the reference's order-preserving greedy colouring)") if not scene.uses_env_path else "exact first-fit per environment",
                        **churn_stats,
                        "env_broadphase": ("all pairs every step (PXB_ENV_BP_CAND=0)" if os.environ.get("PXB_ENV_BP_CAND", "1")[:1] == "0" else "temporal coherence: per-environment candidate pair lists, all-pairs rebuild when a bound moved beyond the margin (identical pair sets)") if scene.uses_env_path else None,
                        "settle_steps": SETTLE[cfg], "ms_per_step_median": per_step[len(per_step) // 2], "ms_per_step_p95": per_step[min(len(per_step) - 1, int(len(per_step) * 0.95))],
                        "timing": "CUDA events on the scene stream per step, max over ranks; L2 flushed (256 MiB memset) between timed steps",
                        "multi_gpu": ("env-partitioned, one scene per GPU, per-step all-gather of the packed pose+linear+angular velocity tensor (13 floats/body) by " + str(gather_kind) + ", double-buffered on communication streams (overlaps the next step)") if gather is not None else ("independent replicas" if world > 1 else "single scene")},
            "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kernel_bytes, "kernel_ms": kernel_ms,
                         "dram_frac": (traffic / (kernel_ms / 1e3) / 1e9 / peak) if traffic else None, "note": note},
            "stage_ms": stages,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nb * 24), "d2h_bytes_per_step": int(nb * (28 + 24)), "steps": e2e_steps,
                    "api": e2e_api, "ms_per_step": e2e_s / e2e_steps * 1e3, "stage_ms": {k: round(v, 4) for k, v in e2e_stage.items()}},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(cfg, n_envs, args.solver)
            sb = second_baseline_block(cfg, n_envs, args.solver)
            if sb is not None:
                line["second_baseline"] = sb
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json workload (2 = the configuration the metric is quoted on)")
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU (configs 2 / 5)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stacks", type=int, default=0, help="deprecated: --stacks 16 = --config 5")
    ap.add_argument("--churn", type=float, default=0.0, help="fraction of the environments reset every step (configs 2 / 5)")
    ap.add_argument("--partitioning", default="exact", choices=["exact", "relaxed"], help="configs 3 / 4: the reference's first-fit (default) or PXB_FLAG_RELAXED_PARTITIONING")
    ap.add_argument("--path", default="auto", choices=["auto", "devicewide"], help="devicewide forces the path used by scenes without environment ids (comparison runs)")
    ap.add_argument("--gather", default="auto", choices=["auto", "graph", "fused", "peer", "peer-copy", "nccl"], help="multi-GPU state exchange")
    ap.add_argument("--e2e", default="export", choices=["export", "copy"], help="end-to-end leg: state export into mapped pinned host memory (default) or explicit D2H copies after the step")
    ap.add_argument("--solver", default="tgs", choices=["tgs", "pgs"], help="PxSolverType of the scene (headline metric: tgs)")
    args = ap.parse_args()
    if args.stacks == 16:
        args.config = 5
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
