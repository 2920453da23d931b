"""The drop-in boundary proper: the UNMODIFIED reference host SDK (oracle/_ref_gpu/ref_harness: PhysX 5.6.1 built with
PX_SUPPORT_GPU_PHYSX, public API only) loads the repo's libPhysXGpu_64.so (plugin/) through its own module loader
(PxSetPhysXGpuLoadHook -> dlopen -> PxCreateCudaContextManager / PxCreatePhysXGpu) and simulates with PxBroadPhaseType::eGPU: the scene's
Bp::BroadPhase and Bp::AABBManagerBase are the plugin's, running the B200 broadphase kernels of libphysx_b200.so.  Compared with the same host
running its own CPU ABP broadphase on the same scene."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

import util
from physx_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref_gpu", "ref_harness")
PLUGIN = os.path.join(ROOT, "plugin", "_build", "libPhysXGpu_64.so")
EXPORTS = ["PxCreatePhysXGpu", "PxCreateCudaContextManager", "PxGetSuggestedCudaDeviceOrdinal", "PxSetPhysXGpuProfilerCallback", "PxSetPhysXGpuFoundationInstance",
           "PxGpuCudaRegisterFunction", "PxGpuCudaRegisterFatBinary", "PxGpuGetCudaFunctionTable", "PxGpuGetCudaFunctionTableSize", "PxGpuGetCudaModuleTable",
           "PxGpuGetCudaModuleTableSize", "PxGpuCreatePhysicsGpu"]


def test_plugin_exports_the_twelve_symbols_of_the_reference_boundary():
    """physxgpu/include/PxPhysXGpu.h:207-237; the loader treats the first three as mandatory (PxPhysXGpuModuleLoader.cpp:233)."""
    if not os.path.exists(PLUGIN):
        pytest.skip("plugin/_build/libPhysXGpu_64.so not built (needs /root/reference at build time: make -f plugin/Makefile)")
    out = subprocess.run(["nm", "-D", "--defined-only", PLUGIN], capture_output=True, text=True, check=True).stdout
    have = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(EXPORTS) <= have, sorted(set(EXPORTS) - have)


def _run(scene, steps, gpu, threads=2, kin_targets=None):
    with tempfile.TemporaryDirectory() as d:
        sp = os.path.join(d, "s.bin")
        scene.save(sp)
        cmd = [HARNESS, "run", sp, "--steps", str(steps), "--threads", str(threads), "--states", d + "/st", "--contacts", d + "/con"]
        if kin_targets is not None:
            np.ascontiguousarray(kin_targets, dtype="<f4").tofile(d + "/kin"); cmd += ["--kin-targets", d + "/kin"]
        if gpu:
            cmd += ["--gpu-plugin", PLUGIN, "--gpu-bp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        info = json.loads(r.stdout.strip().splitlines()[-1])
        states = np.fromfile(d + "/st", "<f4").reshape(steps + 1, scene.n_dynamic, 13)
        import struct
        cb = open(d + "/con", "rb").read(); off = 0; pairs = []
        for _ in range(steps):
            n, = struct.unpack_from("<I", cb, off); off += 4
            cur = {}
            for _p in range(n):
                a0, a1, k = struct.unpack_from("<III", cb, off); off += 12 + 40 * k
                if k:
                    cur[(min(a0, a1), max(a0, a1))] = k
            pairs.append(cur)
        return info, states, pairs, r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["config1", "tumble", "envs", "aggregates", "kinematic"])
def test_unmodified_host_sdk_steps_with_the_b200_broadphase_plugin(name):
    """BASELINE config 1 (SnippetHelloWorld stacks), tumbling boxes (pairs created and lost) and a config-2-shaped environment grid through the
    unmodified host with PxBroadPhaseType::eGPU served by the plugin.  The touching pair set (contact reports) equals the CPU-ABP run's at every
    step while the trajectories agree; poses stay within the config-1 tolerance (the order in which created pairs reach the host differs between
    the two broadphases, which reorders island edges: last-bit effects only on stacks; the tumbling scene is compared over its regular prefix)."""
    if not (os.path.exists(HARNESS) and os.path.exists(PLUGIN)):
        pytest.skip("oracle/_ref_gpu/ref_harness or the plugin is not built (needs /root/reference at build time)")
    sc, steps, tol = {"config1": (scenes.box_stacks(), 120, 1e-3), "tumble": (scenes.tumbling_boxes(n=12, seed=7), 60, 5e-2),
                      "envs": (scenes.env_grid_stacks(n_envs=16, jitter=0.01), 60, 1e-3),
                      # real PxAggregates in the host (one without, one with self collisions): the plugin's AABB manager flattens them and drops the pairs inside the quiet one
                      "aggregates": (scenes.aggregates_mix(), 40, 1e-3),
                      # kinematic bodies moved with setKinematicTarget: the host's filter table (no kinematic-static / kinematic-kinematic pairs) reaches the plugin's broadphase
                      "kinematic": (scenes.kinematic_mix(), 40, 1e-3)}[name]
    kt = scenes.kinematic_targets(sc, steps) if name == "kinematic" else None
    info_g, st_g, pairs_g, err = _run(sc, steps, True, kin_targets=kt)
    assert "GPU plugin" in err and "physx_b200" not in err.replace("GPU plugin", ""), err[-1500:]
    info_c, st_c, pairs_c, _ = _run(sc, steps, False, kin_targets=kt)
    assert info_g["bodies"] == info_c["bodies"] == sc.n_dynamic
    same = sum(pg == pc for pg, pc in zip(pairs_g, pairs_c))
    if name == "aggregates":
        agg = sc.actors["aggregate"]
        inside = [(a, b) for p in pairs_g for (a, b) in p if agg[a] and agg[a] == agg[b]]
        assert inside and all(agg[a] & 0x80000000 for a, b in inside), "pairs inside an aggregate: only where self collisions are enabled"
        assert pairs_g[:15] == pairs_c[:15]     # until the falling boxes hit the columns (chaotic afterwards, like the tumbling scene)
    elif name == "kinematic":
        assert pairs_g[:25] == pairs_c[:25] and any(pairs_g[:25])
    elif name != "tumble":
        assert same == steps, f"touching pair sets differ on {steps - same} of {steps} steps"
    else:
        assert pairs_g[:20] == pairs_c[:20]
    assert util.rel_err(st_g[-1][:, :3], st_c[-1][:, :3]) < tol, "final positions"
    assert np.isfinite(st_g).all()
