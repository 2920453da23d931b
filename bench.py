#!/usr/bin/env python
"""bench.py -- rigid-body-steps/s of the simulate()+fetchResults() hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl ours|reference]

Our arm (default): BASELINE config 2, 4096 independent envs x 64 boxes (262144 bodies) per GPU, GPU
broadphase + TGS 4+1 iterations, 60 Hz, synthetic seeded scene.  N > 1 (under torchrun): one process per
GPU, each rank owns its own 4096 envs (weak scaling, no physics coupling) and, per step, the ranks
all-gather the Direct-GPU-API state tensors over NCCL.  One JSON line is printed by rank 0.

Reference arm (--impl reference): the UNMODIFIED reference CPU SDK (oracle/_ref/ref_harness: eABP
broadphase, CPU TGS, PxDefaultCpuDispatcher with all host threads) on a bounded sample of the same
workload; rank 0 only.
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
BOXES_PER_ENV = 64
# dram__bytes_read.sum + dram__bytes_write.sum of one k_env_solve launch at config 2 (4096 envs x 64 boxes), from the committed
# ncu --set full capture profiles/r01_env_kernels_full_raw.csv (a profiler number: quoted, never timed under ncu)
ENV_SOLVE_DRAM_BYTES_PER_LAUNCH = 122_967_552   # 87.23 MB read + 35.73 MB written


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


def run_reference_sample(n_envs, steps, threads, warmup=2):
    from physx_b200 import scenes
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    if not os.path.exists(harness):
        return None
    sc = scenes.env_grid_stacks(n_envs=n_envs)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.bin")
        sc.save(p)
        out = subprocess.run([harness, "run", p, "--steps", str(steps), "--warmup", str(warmup), "--threads", str(threads)], capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def cpu_baseline_block(steps=20, n_envs=1024):
    threads = os.cpu_count() or 1
    r = run_reference_sample(n_envs, steps, threads)
    if r is None:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref/ref_harness not built"}
    return {"value": r["body_steps_per_s"], "unit": UNIT, "cores": threads, "kind": "reference", "ms_per_step": r["ms_per_step"],
            "sample": f"{n_envs} envs x {BOXES_PER_ENV} boxes = {n_envs * BOXES_PER_ENV} bodies, {steps} steps after 2 warm-up, unmodified PhysX 5.6.1 CPU (eABP + CPU TGS 4+1, PxDefaultCpuDispatcher({threads}))"}


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = args.ref_envs
    t0 = time.time()
    vals, ms = [], []
    # each "step" of this arm = one simulate+fetchResults of the bounded sample scene
    r = run_reference_sample(n_envs, args.steps, threads, warmup=args.warmup)
    if r is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness is not built (needs /root/reference at build time)"}))
        return
    v = r["body_steps_per_s"]
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n_envs} envs x {BOXES_PER_ENV} boxes ({n_envs * BOXES_PER_ENV} bodies): bounded sample of config 2 (4096 envs x 64 boxes), TGS 4+1, 60 Hz",
                       "implementation": "unmodified PhysX 5.6.1 CPU path: eABP broadphase, CPU TGS, PxDefaultCpuDispatcher", "threads": threads},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference", "sample": f"{n_envs * BOXES_PER_ENV} bodies x {args.steps} steps"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


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

    n_envs = args.envs
    solver = scenes.SOLVER_PGS if args.solver == "pgs" else scenes.SOLVER_TGS
    sc = scenes.env_grid_stacks(n_envs=n_envs, stacks_per_env=args.stacks, seed=1234 + rank, solver=solver)  # every rank owns its own envs (weak scaling)
    scene = engine.Scene(sc, device=local, env_path=(args.path != "devicewide"))
    nb = scene.num_dynamic
    stream = torch.cuda.ExternalStream(scene.stream(), device=dev)
    gather = None
    if dist is not None:
        # one packed [n, 13] tensor (pose + linear + angular velocity) -> ONE NCCL all-gather per step, double-buffered on a
        # communication stream so that it overlaps the next step's kernels
        gather, gather_kind = multi_gpu.make_state_gather(dist, nb, 13, dev, stream, kind=args.gather, scene=scene)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def one_step():
        scene.simulate()
        if gather is not None:
            gather.step(lambda view: scene.getStatesDevice(view.data_ptr()))
        scene.fetchResults(True)

    for _ in range(max(args.warmup, 3)):
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
        scene.simulate()
        if gather is not None:
            gather.step(lambda view: scene.getStatesDevice(view.data_ptr()))   # pack kernel on the scene stream, NCCL on the comm stream
        e1.record(stream)
        scene.fetchResults(True)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        launches += scene.num_launches + (1 if gather is not None else 0)
    if gather is not None:
        # the collectives of the last two steps may still be in flight: their completion is part of the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gather.wait()
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
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
    # ---- per-stage device times (CUDA events inside the engine, direct launches) for the roofline of the
    #      dominant kernel; separate from the timed region, which replays the captured CUDA graph ----
    scene.setProfiling(True)
    prof_steps = 20
    for _ in range(prof_steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        scene.simulate()
        scene.fetchResults(True)
        for k, v in scene.getStageTimes().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    scene.setProfiling(False)
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

    # ---- end-to-end through the public API with HOST buffers (pinned): per step H2D of the velocity "action"
    #      tensors, simulate+fetchResults, D2H of pose + velocities ----
    pose_h = torch.empty((nb, 7), dtype=torch.float32).pin_memory()
    lin_h = torch.empty((nb, 3), dtype=torch.float32).pin_memory()
    ang_h = torch.empty((nb, 3), dtype=torch.float32).pin_memory()
    lib, h = scene._lib, scene._h
    lib.pxb_get_rigid_dynamic_data(h, lin_h.data_ptr(), None, engine.RD_LINEAR_VELOCITY, nb)
    lib.pxb_get_rigid_dynamic_data(h, ang_h.data_ptr(), None, engine.RD_ANGULAR_VELOCITY, nb)
    e2e_steps = max(5, min(args.steps, 50))
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        # stream-ordered calls on pinned host buffers, ONE host synchronisation per step (fetchResults), as with PxDirectGPUAPI's events
        lib.pxb_set_rigid_dynamic_data_async(h, lin_h.data_ptr(), engine.RD_LINEAR_VELOCITY, nb)   # H2D (identity action: values just read back)
        lib.pxb_set_rigid_dynamic_data_async(h, ang_h.data_ptr(), engine.RD_ANGULAR_VELOCITY, nb)
        scene.simulate()
        lib.pxb_get_rigid_dynamic_data_async(h, pose_h.data_ptr(), engine.RD_GLOBAL_POSE, nb)      # D2H
        lib.pxb_get_rigid_dynamic_data_async(h, lin_h.data_ptr(), engine.RD_LINEAR_VELOCITY, nb)
        lib.pxb_get_rigid_dynamic_data_async(h, ang_h.data_ptr(), engine.RD_ANGULAR_VELOCITY, nb)
        scene.fetchResults(True)                                                                   # waits for the step and the copies
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_bodies * e2e_steps / e2e_s
    assert np.isfinite(pose_h.numpy()).all()

    if rank == 0:
        peak, peak_src = measured_peaks()
        P = scene.num_constraints            # contact patches = touching pairs (one patch per primitive pair)
        C = 4 * P                            # contact points (box-plane / box-box face manifolds: 4)
        F = 4 * P                            # friction rows (2 anchors x 2 tangents)
        iters = int(sc.header["posIters"]) + int(sc.header["velIters"])
        # SURVEY.md §8d algorithmic bytes: solver per iteration 116 P + 60 C + 48 F + 72 N_b; contact prep reads 64 P + 16 C + 2*112 P
        # and writes 116 P + 56 C + 44 F; integration 132 N_b.
        solve_bytes = (116 * P + 60 * C + 48 * F + 72 * nb) * iters
        prep_bytes = (64 * P + 16 * C + 224 * P) + (116 * P + 56 * C + 44 * F)
        integ_bytes = 132 * nb
        solve_ms = stage_acc["solve"] / prof_steps
        if scene.uses_env_path:
            # environment path: ONE kernel does pre-integration, colouring, prep, all TGS iterations, write-back and integration
            kernel_name = "k_env_solve (a12-a18 fused: prep + all TGS iterations + integration, one CTA per environment, rows in registers)"
            kernel_bytes = prep_bytes + solve_bytes + integ_bytes
            note = ("algorithmic bytes = SURVEY 8d prep + 5 solver iterations + integration; the kernel keeps every solver row on chip (registers), "
                    "so only contacts, friction patches and body state cross HBM once: see `traffic` (ncu dram bytes per launch, profiles/r01_env_kernels_full_raw.csv)")
            traffic = ENV_SOLVE_DRAM_BYTES_PER_LAUNCH if (n_envs == 4096 and args.stacks == 8 and args.solver == "tgs") else None
        else:
            kernel_name = "k_solve (all TGS iterations, one cooperative launch)"
            kernel_bytes = solve_bytes
            note = "device-wide path: rows re-streamed from HBM every iteration"
            traffic = None
        achieved = kernel_bytes / (solve_ms / 1e3) / 1e9
        stages = {k: round(v / prof_steps, 4) for k, v in stage_acc.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config {2 if args.stacks == 8 else 5}: {n_envs} envs x {args.stacks * 8} boxes = {nb} bodies per GPU, shared ground plane, GPU broadphase + {args.solver.upper()} 4 pos/1 vel iterations, 60 Hz",
                       "bodies_total": total_bodies, "constraints_per_gpu": P, "partitions": scene.num_partitions, "path": "environment (pxb_env.cuh)" if scene.uses_env_path else "device-wide",
                       "timing": "CUDA events on the scene stream per step, max over ranks; L2 flushed (256 MiB memset) between timed steps",
                       "multi_gpu": ("env-partitioned, one scene per GPU, per-step all-gather of the packed pose+linear+angular velocity tensor (13 floats/body) by " + gather_kind + ", double-buffered on a communication stream (overlaps the next step)") if world > 1 else "single scene"},
            "roofline": {"kernel": kernel_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kernel_bytes, "kernel_ms": solve_ms,
                         "dram_frac": (traffic / (solve_ms / 1e3) / 1e9 / peak) if traffic else None, "note": note},
            "stage_ms": stages,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nb * 24), "d2h_bytes_per_step": int(nb * (28 + 24)), "steps": e2e_steps,
                    "api": "pxb_set_rigid_dynamic_data_async(lin,ang) -> pxb_scene_simulate -> pxb_get_rigid_dynamic_data_async(pose,lin,ang) -> pxb_scene_fetch_results (one host sync per step), pinned host buffers"},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--ref-envs", type=int, default=1024, help="environments in the reference arm's bounded sample")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stacks", type=int, default=8, help="stacks of 8 boxes per environment (8 = config 2, 16 = the per-GPU shard of config 5)")
    ap.add_argument("--path", default="auto", choices=["auto", "devicewide"], help="devicewide forces the path used by scenes without environment ids (comparison runs)")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "peer-copy", "nccl"], help="multi-GPU state exchange: peer-memory copies or NCCL all-gather")
    ap.add_argument("--solver", default="tgs", choices=["tgs", "pgs"], help="PxSolverType of the scene (headline metric: tgs)")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
