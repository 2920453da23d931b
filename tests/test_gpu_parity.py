"""GPU parity tests (run on the B200): the CUDA path through the C ABI against (1) the CPU oracle on the same
seeded inputs -- expected bit-exact, asserted <= 1e-6 -- (2) golden data from the unmodified reference, and
(3) size-independent properties at BASELINE.json's full config-2 size."""
import numpy as np
import pytest

import util
from physx_b200 import engine, scenes

pytestmark = pytest.mark.gpu

TOL_POSE, TOL_LINVEL, TOL_ANGVEL = 1e-4, 1e-2, 5e-2   # vs the reference (see test_oracle_vs_reference.py)
TOL_ORACLE = 1e-6                                      # vs our own oracle: same arithmetic, same order
TOL_STEP = 2e-5                                        # one step from identical inputs (libm vs CUDA sinf/cosf/acosf: 1 ulp)


def _scenes_small():
    return {
        "1box": (scenes.box_stacks(n_stacks=1, height=1, half_extent=0.25, spacing=1.0), 10),
        "stacks_4x8_jitter": (scenes.box_stacks(n_stacks=4, height=8, half_extent=0.25, spacing=1.0, jitter=0.01), 120),
        "unit_stacks_10x10": (scenes.box_stacks(), 60),                     # BASELINE config 1
        "envs_16": (scenes.env_grid_stacks(n_envs=16, jitter=0.01), 60),    # BASELINE config 2 shape
        "tumble_12": (scenes.tumbling_boxes(n=12, seed=7), 150),            # pairs created/lost, rotated contacts
        "free_fall": (_free_fall(), 5),
        "pile_6x4x6": (scenes.box_pile(6, 4, 6), 60),                       # BASELINE config 4 shape: one dense island, walled bin
        "fall_5x4x5": (scenes.falling_primitives(5, 4, 5), 120),            # BASELINE config 3 shape: mixed primitives into a bin
    }


def _free_fall():
    s = scenes.box_stacks(n_stacks=2, height=1, half_extent=0.25)
    s.actors["pos"][1:, 1] = 5.0
    return s


@pytest.mark.parametrize("path", ["auto", "devicewide"])
@pytest.mark.parametrize("name", list(_scenes_small()))
def test_gpu_matches_oracle(oracle, name, path):
    """path auto: scenes of at most 288 actors without environment ids run as ONE environment on the fused environment path, larger ones and
    environment scenes as before; devicewide: the device-wide path forced (PXB_FLAG_NO_ENV_PATH)."""
    sc, steps = _scenes_small()[name]
    gpu, cpu = engine.Scene(sc, max_pairs=16 * len(sc.actors), env_path=(path == "auto")), oracle.OracleScene(sc)
    assert path == "auto" or not gpu.uses_env_path
    exact = True
    for t in range(steps):
        gpu.step()
        cpu.step()
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        assert np.array_equal(gpu.getCreatedPairs(), cpu.getCreatedPairs()), f"created, step {t}"
        assert np.array_equal(gpu.getDeletedPairs(), cpu.getDeletedPairs()), f"deleted, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]), f"contact counts, step {t}"
        assert np.abs(cg - cc).max(initial=0) < 1e-5, f"contacts, step {t}"
        assert gpu.num_constraints == cpu.num_constraints and gpu.num_partitions == cpu.num_partitions
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        exact = exact and np.array_equal(sg, cpu.getStates())
        cpu.setStates(sg)   # re-synchronise: every step is checked from identical inputs (sinf/cosf/acosf differ by an ulp between libm and CUDA)
    if name not in ("tumble_12", "fall_5x4x5"):
        assert exact, "stack / free-fall scenes are expected to be bit-identical to the oracle"


@pytest.mark.parametrize("name", ["stacks_4x8_jitter", "stacks_3x5_exact", "envs_4"])
def test_gpu_matches_reference_golden(name):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    steps = z["states"].shape[0] - 1
    for t in range(steps):
        gpu.setConstraintOrder(util.golden_order(z, t))   # island-manager order, as the plugin shim would pass it
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert util.rel_err(st[:, :3], ref[:, :3]) < TOL_POSE, f"position, step {t}"
        assert util.rel_err(st[:, 3:7], ref[:, 3:7]) < TOL_POSE, f"orientation, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < TOL_LINVEL and np.abs(st[:, 10:] - ref[:, 10:]).max() < TOL_ANGVEL, f"velocity, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"


@pytest.mark.parametrize("name", ["stacks_4x8_jitter", "stacks_3x5_exact", "envs_4", "tumble_12"])
def test_gpu_broadphase_matches_abp_bit_exact(name):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    for t in range(z["bounds"].shape[0]):
        gpu.broadphase(z["bounds"][t])
        assert np.array_equal(gpu.getCreatedPairs(), util.golden_created(z, t)), f"created, step {t}"
        assert np.array_equal(gpu.getDeletedPairs(), util.golden_deleted(z, t)), f"deleted, step {t}"


@pytest.mark.parametrize("name", ["stacks_4x8_jitter", "tumble_12"])
def test_gpu_bounds_match_reference_bit_exact(name):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    for t in range(0, z["bounds"].shape[0], 7):
        gpu.setStates(z["states"][t])
        assert np.array_equal(gpu.computeBounds(), z["bounds"][t]), f"step {t}"


def test_config2_full_size_properties(oracle):
    """4096 envs x 64 boxes = 262144 bodies (BASELINE config 2)."""
    sc = scenes.env_grid_stacks(n_envs=4096)
    gpu = engine.Scene(sc)
    gpu.step()
    pairs = gpu.getPairs()
    # pair set identical to the oracle's broadphase on the same bounds (stage level, full size)
    cpu = oracle.OracleScene(sc)
    cpu.computeBounds()
    cpu.broadphase()
    assert np.array_equal(pairs, cpu.getPairs())
    keys = pairs[:, 0].astype(np.int64) << 32 | pairs[:, 1]
    assert np.all(np.diff(keys) > 0), "sorted and unique"
    env = sc.actors["envId"]
    e0, e1 = env[pairs[:, 0]], env[pairs[:, 1]]
    assert np.all((e0 == e1) | (e0 == scenes.NO_ENV) | (e1 == scenes.NO_ENV)), "no cross-environment pair"
    assert len(pairs) == 262144 and gpu.num_constraints == 262144
    st0 = gpu.getStates()
    for _ in range(30):
        gpu.step()
    st = gpu.getStates()
    assert np.isfinite(st).all()
    assert np.abs(st[:, :3] - st0[:, :3]).max() < 0.15, "stacks stay standing (a collapse moves boxes by > 0.5 m)"
    assert np.abs(np.linalg.norm(st[:, 3:7], axis=1) - 1).max() < 1e-5
    # every environment evolves independently and identically shaped: determinism across two scenes
    gpu2 = engine.Scene(sc)
    for _ in range(31):
        gpu2.step()
    assert np.array_equal(gpu2.getStates(), st), "bitwise deterministic"


def test_direct_gpu_api_roundtrip():
    sc = scenes.env_grid_stacks(n_envs=8)
    gpu = engine.Scene(sc)
    gpu.step()
    pose = gpu.getRigidDynamicData(engine.RD_GLOBAL_POSE)
    st = gpu.getStates()
    assert np.array_equal(pose[:, :4], st[:, 3:7]) and np.array_equal(pose[:, 4:], st[:, :3])   # PxTransform = (q.xyzw, p.xyz)
    assert np.array_equal(gpu.getRigidDynamicData(engine.RD_LINEAR_VELOCITY), st[:, 7:10])
    idx = np.array([5, 3, 100], np.uint32)
    sub = gpu.getRigidDynamicData(engine.RD_ANGULAR_VELOCITY, idx)
    assert np.array_equal(sub, st[idx, 10:])
    new = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], np.float32)
    gpu.setRigidDynamicData(engine.RD_LINEAR_VELOCITY, new, idx)
    assert np.array_equal(gpu.getRigidDynamicData(engine.RD_LINEAR_VELOCITY, idx), new)
    with pytest.raises(engine.PhysxB200Error):
        gpu.getRigidDynamicData(engine.RD_GLOBAL_POSE, np.array([10 ** 6], np.uint32))
    gpu.simulate()
    with pytest.raises(engine.PhysxB200Error):   # illegal while the simulation is running (NpDirectGPUAPI.cpp:63-78)
        gpu.getRigidDynamicData(engine.RD_GLOBAL_POSE)
    gpu.fetchResults(True)


def test_pair_capacity_overflow_is_reported():
    sc = scenes.env_grid_stacks(n_envs=16)
    gpu = engine.Scene(sc, max_pairs=100)
    with pytest.raises(engine.PhysxB200Error) as e:
        gpu.step()
    assert "capacity" in str(e.value)


def test_reset_via_set_states_reproduces_trajectory():
    sc = scenes.box_stacks(n_stacks=2, height=4, half_extent=0.25, spacing=1.0, jitter=0.01)
    a = engine.Scene(sc)
    st0 = a.getStates()
    for _ in range(20):
        a.step()
    ref = a.getStates()
    b = engine.Scene(sc)
    b.setStates(st0)
    for _ in range(20):
        b.step()
    assert np.array_equal(b.getStates(), ref)


@pytest.mark.parametrize("name", ["spheres_capsules", "capsule_row", "spheres_boxes", "capsules_on_boxes", "capsules_on_boxes_4", "capsules_boxes_tumbling", "all_primitives", "capsules_into_boxes"])
def test_gpu_matches_oracle_primitives(oracle, name):
    sc = {"spheres_capsules": scenes.mixed_primitives(n=14, seed=3, kinds=("sphere", "capsule")),
          "capsules_on_boxes": scenes.capsules_on_boxes(seed=3),                                   # a10: capsule-box through GJK (k_narrowphase_gjk)
          "capsules_on_boxes_4": scenes.capsules_on_boxes(n_boxes=8, per_box=4, seed=4),
          "capsules_boxes_tumbling": scenes.mixed_primitives(n=14, seed=5, kinds=("capsule", "box")),
          "all_primitives": scenes.mixed_primitives(n=18, seed=3),
          "capsules_into_boxes": scenes.capsules_into_boxes(seed=2),                              # deep penetration: EPA on the device
          "capsule_row": scenes.mixed_primitives(n=6, seed=5, kinds=("capsule",), spread=0.05),
          "spheres_boxes": scenes.mixed_primitives(n=24, seed=11, kinds=("sphere", "box", "box", "sphere"))}[name]
    gpu, cpu = engine.Scene(sc), oracle.OracleScene(sc)
    for t in range(150):
        gpu.step()
        cpu.step()
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]), f"contact counts, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        cpu.setStates(sg)


@pytest.mark.parametrize("name", ["capsule_row", "spheres_capsules_14", "capsules_on_boxes", "hulls_on_plane", "hulls_and_spheres", "hulls_and_capsules", "tumble_12", "tumble_mixed_14"])
def test_gpu_teacher_forced_steps_match_reference(name):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert np.abs(st[:, :7] - ref[:, :7]).max() < 1e-5, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < 1e-4 and np.abs(st[:, 10:] - ref[:, 10:]).max() < 1e-3, f"velocity, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"


def _env_scenes():
    return {"envs_16": (scenes.env_grid_stacks(n_envs=16, jitter=0.01), 40),
            "envs_37x3x5": (scenes.env_grid_stacks(n_envs=37, stacks_per_env=3, height=5, jitter=0.02), 40),
            "ragged": (scenes.env_ragged(), 150),
            "wide_200_bodies": (scenes.env_grid_stacks(n_envs=3, stacks_per_env=25, height=8, jitter=0.01), 25),   # 200 constraints per env: 256-thread CTAs
            "piles_100": (scenes.env_piles(), 40),   # ~300 constraints / ~650 pairs per env: rows and lists through global scratch
            "env_hulls": (scenes.env_hulls(), 120)}  # library hulls + spheres + capsules + boxes per environment: hull bounds in k_env_bp, every a10 pair type


@pytest.mark.parametrize("name", list(_env_scenes()))
def test_env_path_is_bit_identical_to_device_wide_path(name):
    """Environment path (one warp / CTA per environment, rows in shared memory) vs the device-wide path (grid broadphase,
    global colouring, cooperative solve): same pairs, events, contacts and states, bit for bit, every step."""
    sc, steps = _env_scenes()[name]
    env, glob = engine.Scene(sc, max_pairs=16 * len(sc.actors)), engine.Scene(sc, max_pairs=16 * len(sc.actors), env_path=False)
    for t in range(steps):
        env.step()
        glob.step()
        assert env.uses_env_path and not glob.uses_env_path
        assert np.array_equal(env.getPairs(), glob.getPairs()), f"pairs, step {t}"
        assert np.array_equal(env.getCreatedPairs(), glob.getCreatedPairs()), f"created, step {t}"
        assert np.array_equal(env.getDeletedPairs(), glob.getDeletedPairs()), f"deleted, step {t}"
        assert np.array_equal(env.getContacts(), glob.getContacts()), f"contacts + applied forces, step {t}"
        assert env.num_constraints == glob.num_constraints and env.num_partitions == glob.num_partitions
        assert np.array_equal(env.getStates(), glob.getStates()), f"states, step {t}"


def test_env_path_matches_oracle_ragged(oracle):
    sc = scenes.env_ragged()
    gpu, cpu = engine.Scene(sc), oracle.OracleScene(sc)
    for t in range(120):
        gpu.step()
        cpu.step()
        assert gpu.uses_env_path
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        assert np.array_equal(gpu.getCreatedPairs(), cpu.getCreatedPairs()) and np.array_equal(gpu.getDeletedPairs(), cpu.getDeletedPairs()), f"events, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]), f"contact counts, step {t}"
        assert gpu.num_constraints == cpu.num_constraints and gpu.num_partitions == cpu.num_partitions
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        cpu.setStates(sg)


def test_env_path_falls_back_when_order_is_given_and_oversize_environments():
    """(1) a host constraint order switches a running scene from the environment path to the device-wide path (pair list
    converted, manifolds kept); (2) environments larger than the shared-memory row store use global rows, same results."""
    sc = scenes.env_grid_stacks(n_envs=9, jitter=0.01)
    a, b = engine.Scene(sc), engine.Scene(sc, env_path=False)
    for _ in range(10):
        a.step(); b.step()
    assert a.uses_env_path
    order = a.getPairs()
    a.setConstraintOrder(order); b.setConstraintOrder(order)
    for t in range(10):
        a.step(); b.step()
        assert not a.uses_env_path
        assert np.array_equal(a.getCreatedPairs(), b.getCreatedPairs()) and len(a.getCreatedPairs()) == 0
        assert np.array_equal(a.getStates(), b.getStates()), f"step {t}"
    # 64 constraints per environment: 32 threads per CTA -> rows stream through global scratch instead of registers;
    # 8 list slots -> constraint lists in global scratch instead of shared memory
    c, c2, d = engine.Scene(sc, env_threads=32), engine.Scene(sc, env_row_cap=8), engine.Scene(sc)
    for t in range(20):
        c.step(); c2.step(); d.step()
        assert c.uses_env_path and np.array_equal(c.getStates(), d.getStates()) and np.array_equal(c.getContacts(), d.getContacts()), f"step {t}"
        assert c2.uses_env_path and np.array_equal(c2.getStates(), d.getStates()) and np.array_equal(c2.getContacts(), d.getContacts()), f"step {t}"


# ---- PGS solver ----
def _pgs_scenes():
    P = scenes.SOLVER_PGS
    return {"stacks_4x8": (scenes.box_stacks(n_stacks=4, height=8, half_extent=0.25, spacing=1.0, jitter=0.01, solver=P), 100, False),
            "envs_16": (scenes.env_grid_stacks(n_envs=16, jitter=0.01, solver=P), 60, True),
            "ragged": (scenes.env_ragged(solver=P), 120, True),
            "spheres_boxes": (scenes.mixed_primitives(n=24, seed=11, kinds=("sphere", "box", "box", "sphere"), solver=P), 120, False),
            "vel_iters_0_and_3": (scenes.box_stacks(n_stacks=2, height=5, half_extent=0.25, spacing=1.0, jitter=0.01, solver=P, pos_iters=6, vel_iters=3), 40, False)}


@pytest.mark.parametrize("name", list(_pgs_scenes()))
def test_pgs_gpu_matches_oracle(oracle, name):
    sc, steps, env = _pgs_scenes()[name]
    gpu, cpu = engine.Scene(sc, env_path=env), oracle.OracleScene(sc)
    glob = engine.Scene(sc, env_path=False) if env else None
    for t in range(steps):
        gpu.step()
        cpu.step()
        assert gpu.uses_env_path == env
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]), f"contact counts, step {t}"
        assert gpu.num_constraints == cpu.num_constraints and gpu.num_partitions == cpu.num_partitions
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        assert np.abs(cg - cc).max(initial=0) < 1e-4, f"contacts / applied forces, step {t}"
        cpu.setStates(sg)
        if glob is not None:   # environment path == device-wide path, bit for bit
            glob.step()
            assert np.array_equal(glob.getStates(), sg) and np.array_equal(glob.getContacts(), cg), f"env vs device-wide, step {t}"


@pytest.mark.parametrize("name", ["pgs_stacks_3x6_jitter", "pgs_envs_4"])
def test_pgs_gpu_matches_reference_golden(name):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert util.rel_err(st[:, :3], ref[:, :3]) < TOL_POSE and util.rel_err(st[:, 3:7], ref[:, 3:7]) < TOL_POSE, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < TOL_LINVEL and np.abs(st[:, 10:] - ref[:, 10:]).max() < TOL_ANGVEL, f"velocity, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"


def test_pgs_gpu_teacher_forced_steps_match_reference():
    z, sc = util.load_golden("pgs_spheres_capsules_12")
    gpu = engine.Scene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert np.abs(st[:, :7] - ref[:, :7]).max() < 1e-5, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < 1e-4 and np.abs(st[:, 10:] - ref[:, 10:]).max() < 1e-3, f"velocity, step {t}"


@pytest.mark.parametrize("name", ["pile_tgs", "pile_pgs", "fall_tgs"])
def test_relaxed_partitioning_matches_oracle(oracle, name):
    """PXB_FLAG_RELAXED_PARTITIONING (giant islands, BASELINE configs 3 / 4): Jones-Plassmann rounds instead of the sequential
    first-fit.  The oracle runs the same rounds, so the GPU stays checkable bit for bit; the partitioning is valid (no body twice
    in a partition) and the pile settles like the first-fit one."""
    sc = {"pile_tgs": scenes.box_pile(6, 4, 6, relaxed_partitioning=True),
          "pile_pgs": scenes.box_pile(5, 4, 5, relaxed_partitioning=True, solver=scenes.SOLVER_PGS),
          "fall_tgs": scenes.falling_primitives(5, 4, 5, relaxed_partitioning=True)}[name]
    gpu, cpu = engine.Scene(sc, max_pairs=16 * len(sc.actors)), oracle.OracleScene(sc)
    for t in range(80):
        gpu.step()
        cpu.step()
        assert not gpu.uses_env_path
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        assert gpu.num_constraints == cpu.num_constraints and gpu.num_partitions == cpu.num_partitions, f"partition count, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        assert np.abs(gpu.getContacts() - cpu.getContacts()).max(initial=0) < 1e-4, f"contacts / applied forces, step {t}"
        cpu.setStates(sg)
    if name.startswith("pile"):
        st = gpu.getStates()
        assert np.isfinite(st).all() and np.abs(st[:, 7:10]).max() < 0.5 and st[:, 1].min() > 0.2, "the pile rests in the bin"


# ---- sleeping (a18) ----
def _sleep_scenes():
    drop = scenes.box_stacks(n_stacks=1, height=3, half_extent=0.25, spacing=1.0, jitter=0.0, sleep_threshold=0.005)
    drop.actors["pos"][3, 1] = 4.0
    envs = scenes.env_grid_stacks(n_envs=6, stacks_per_env=2, height=3, jitter=0.01, sleep_threshold=0.005)
    envs.actors["pos"][1 + 6 * 2 + 5, 1] = 3.0        # env 2: one box starts high and lands on its sleeping stack later
    return {"stacks": (scenes.box_stacks(n_stacks=2, height=3, half_extent=0.25, spacing=1.0, jitter=0.01, sleep_threshold=0.005), 60, False),
            "drop": (drop, 110, False),
            "envs": (envs, 110, True),
            "envs_pgs": (scenes.env_grid_stacks(n_envs=5, stacks_per_env=2, height=3, jitter=0.01, sleep_threshold=0.005, solver=scenes.SOLVER_PGS), 70, True),
            "stacks_pgs": (scenes.box_stacks(n_stacks=3, height=4, half_extent=0.25, spacing=1.0, jitter=0.01, sleep_threshold=0.005, solver=scenes.SOLVER_PGS), 70, False)}


@pytest.mark.parametrize("name", list(_sleep_scenes()))
def test_sleeping_gpu_matches_oracle(oracle, name):
    """Wake counters, asleep flags and states against the oracle (itself pinned against the reference's getWakeCounter /
    isSleeping, tests/test_oracle_vs_reference.py): islands fall asleep and wake on the same step; no resynchronisation."""
    sc, steps, env = _sleep_scenes()[name]
    gpu, cpu = engine.Scene(sc, env_path=env), oracle.OracleScene(sc)
    ever_slept = woke = False
    prev = None
    for t in range(steps):
        gpu.step()
        cpu.step()
        assert gpu.uses_env_path == env
        (wg, ag), (wc, ac) = gpu.getSleep(), cpu.getSleep()
        assert np.array_equal(ag, ac), f"asleep flags, step {t}"
        assert np.abs(wg - wc).max() < 1e-6, f"wake counters, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < 1e-4, f"state, step {t}"
        if ag.any():
            ever_slept = True
            assert np.all(sg[ag == 1][:, 7:] == 0), "sleeping bodies have zero velocity"
            if prev is not None:
                still = (ag == 1) & (prev[1] == 1)
                assert np.array_equal(sg[still][:, :7], prev[0][still][:, :7]), "sleeping bodies do not move"
        if prev is not None and np.any((prev[1] == 1) & (ag == 0)):
            woke = True
        prev = (sg, ag)
    assert ever_slept
    if name in ("drop", "envs"):
        assert woke, "the impact wakes the sleeping island"


def test_sleeping_gpu_matches_reference_golden():
    z, sc = util.load_golden("sleep_stacks_2x3")
    gpu = engine.Scene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        w, a = gpu.getSleep()
        assert np.array_equal(a, z["asleep"][t]) and np.abs(w - z["wake"][t]).max() < 1e-6, f"step {t}"


def _lock_scenes():
    env = scenes.env_grid_stacks(n_envs=6, stacks_per_env=3, height=5, jitter=0.01)
    dyn = np.nonzero(env.actors["flags"] & 1)[0]
    for k, lock in ((1, 5), (3, 56), (7, 1), (11, 16), (16, 61), (22, 2)):
        scenes.set_lock_flags(env.actors, dyn[k], lock)
    env_pgs = scenes.Scene(env.header, env.actors.copy()); env_pgs.header["solverType"] = scenes.SOLVER_PGS
    return {
        "lock_stacks": (scenes.locked_stacks(), 80),
        "pgs_lock_stacks": (scenes.locked_stacks(solver=scenes.SOLVER_PGS), 80),
        "lock_primitives": (scenes.locked_primitives(seed=4), 100),
        "pgs_lock_primitives": (scenes.locked_primitives(seed=3, solver=scenes.SOLVER_PGS), 100),
        "lock_envs": (env, 60),              # environment path
        "pgs_lock_envs": (env_pgs, 60),
    }


@pytest.mark.parametrize("name", list(_lock_scenes()))
def test_lock_flags_gpu_matches_oracle(oracle, name):
    """PxRigidDynamicLockFlags on both paths and both solvers: same pairs / contacts, states within one-step rounding of the oracle."""
    sc, steps = _lock_scenes()[name]
    gpu, cpu = engine.Scene(sc, max_pairs=16 * len(sc.actors), env_path=name.endswith("_envs")), oracle.OracleScene(sc)
    for t in range(steps):
        gpu.step()
        cpu.step()
        assert gpu.uses_env_path == name.endswith("_envs")
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        assert np.array_equal(gpu.getContacts()[:, 0], cpu.getContacts()[:, 0]), f"contact counts, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        cpu.setStates(sg)


@pytest.mark.parametrize("name", ["lock_stacks", "pgs_lock_stacks"])
def test_lock_flags_gpu_matches_reference_golden(name):
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert util.rel_err(st[:, :3], ref[:, :3]) < TOL_POSE and util.rel_err(st[:, 3:7], ref[:, 3:7]) < TOL_POSE, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < TOL_LINVEL and np.abs(st[:, 10:] - ref[:, 10:]).max() < TOL_ANGVEL, f"velocity, step {t}"


@pytest.mark.parametrize("name", ["stacks", "pgs_stacks", "envs", "pgs_envs", "primitives"])
def test_external_forces_gpu_matches_oracle(oracle, name):
    """PxDirectGPUAPI eFORCE / eTORQUE writes on both paths and both solvers: bit-level agreement with the oracle every step; forces last one step."""
    sc = {"stacks": scenes.box_stacks(n_stacks=3, height=4, half_extent=0.25, spacing=1.0, jitter=0.01),
          "pgs_stacks": scenes.box_stacks(n_stacks=3, height=4, half_extent=0.25, spacing=1.0, jitter=0.01, solver=scenes.SOLVER_PGS),
          "envs": scenes.env_grid_stacks(n_envs=5, stacks_per_env=3, height=4, jitter=0.01),
          "pgs_envs": scenes.env_grid_stacks(n_envs=5, stacks_per_env=3, height=4, jitter=0.01, solver=scenes.SOLVER_PGS),
          "primitives": scenes.mixed_primitives(n=12, seed=3, kinds=("sphere", "capsule"))}[name]
    F = scenes.test_forces(sc.n_dynamic)
    gpu, cpu = engine.Scene(sc, env_path=name.endswith("envs")), oracle.OracleScene(sc)
    for t in range(60):
        if t % 9 != 8:   # every ninth step nothing is written: the previous step's forces must not persist
            gpu.setForces(F[t % len(F), :, :3], F[t % len(F), :, 3:])
            cpu.setForces(F[t % len(F), :, :3], F[t % len(F), :, 3:])
        gpu.step()
        cpu.step()
        assert gpu.uses_env_path == name.endswith("envs")
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        cpu.setStates(sg)


@pytest.mark.parametrize("name", ["forces_stacks", "pgs_forces_stacks"])
def test_external_forces_gpu_matches_reference_golden(name):
    z, sc = util.load_golden(name)
    F = z["forces"]
    gpu = engine.Scene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setForces(F[t % len(F), :, :3], F[t % len(F), :, 3:])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert util.rel_err(st[:, :3], ref[:, :3]) < TOL_POSE and util.rel_err(st[:, 3:7], ref[:, 3:7]) < TOL_POSE, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < TOL_LINVEL and np.abs(st[:, 10:] - ref[:, 10:]).max() < TOL_ANGVEL, f"velocity, step {t}"


def test_stream_ordered_host_api_matches_blocking_api():
    """pxb_get/set_rigid_dynamic_data_async: enqueued around simulate with ONE host sync (fetchResults) per step, same results."""
    import ctypes
    sc = scenes.env_grid_stacks(n_envs=8)
    a, b = engine.Scene(sc), engine.Scene(sc)
    nb = a.num_dynamic
    lin = np.zeros((nb, 3), np.float32); ang = np.zeros((nb, 3), np.float32); pose = np.zeros((nb, 7), np.float32)
    rng = np.random.RandomState(0)
    for t in range(5):
        act = (rng.uniform(-0.2, 0.2, (nb, 3))).astype(np.float32)
        b.setRigidDynamicData(engine.RD_LINEAR_VELOCITY, act)
        b.step()
        ptr = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        assert a._lib.pxb_set_rigid_dynamic_data_async(a._h, ptr(act), engine.RD_LINEAR_VELOCITY, nb) == 0
        a.simulate()
        for arr, ty in ((pose, engine.RD_GLOBAL_POSE), (lin, engine.RD_LINEAR_VELOCITY), (ang, engine.RD_ANGULAR_VELOCITY)):
            assert a._lib.pxb_get_rigid_dynamic_data_async(a._h, ptr(arr), ty, nb) == 0
        a.fetchResults(True)
        assert np.array_equal(pose, b.getRigidDynamicData(engine.RD_GLOBAL_POSE)), f"step {t}"
        assert np.array_equal(lin, b.getRigidDynamicData(engine.RD_LINEAR_VELOCITY)) and np.array_equal(ang, b.getRigidDynamicData(engine.RD_ANGULAR_VELOCITY))


def test_actors_added_to_a_running_env_scene():
    """pxb_scene_add_actors between steps: the environment lists are rebuilt, existing pairs keep their manifolds (no spurious
    created / deleted events) and the result equals the device-wide path's."""
    sc = scenes.env_grid_stacks(n_envs=6, stacks_per_env=2, height=3, jitter=0.01)
    extra = scenes.env_grid_stacks(n_envs=6, stacks_per_env=1, height=2, jitter=0.0).actors[1:].copy()   # 2 more boxes per env
    extra["pos"][:, 0] += 3.0
    scs = [engine.Scene(sc, max_actors=len(sc.actors) + len(extra)), engine.Scene(sc, max_actors=len(sc.actors) + len(extra), env_path=False)]
    for g in scs:
        for _ in range(10):
            g.step()
        assert g._lib.pxb_scene_add_actors(g._h, np.ascontiguousarray(extra).ctypes.data, len(extra)) == 0
        g.num_dynamic = int(g._lib.pxb_scene_num_dynamic(g._h)); g.num_actors = int(g._lib.pxb_scene_num_actors(g._h))
    for t in range(20):
        for g in scs:
            g.step()
        assert scs[0].uses_env_path and not scs[1].uses_env_path
        assert np.array_equal(scs[0].getPairs(), scs[1].getPairs()) and np.array_equal(scs[0].getCreatedPairs(), scs[1].getCreatedPairs()), f"step {t}"
        assert np.array_equal(scs[0].getDeletedPairs(), scs[1].getDeletedPairs()) and len(scs[0].getDeletedPairs()) == 0
        assert np.array_equal(scs[0].getStates(), scs[1].getStates()), f"states, step {t}"
    assert len(scs[0].getStates()) == 6 * 6 + 12


def test_env_path_pair_capacity_overflow_is_reported():
    sc = scenes.env_grid_stacks(n_envs=16)
    gpu = engine.Scene(sc, max_pairs=100)
    with pytest.raises(engine.PhysxB200Error) as e:
        gpu.step()
    assert "capacity" in str(e.value)


def test_convex_hulls_gpu_matches_oracle_and_reference(oracle):
    """Cooked convex hulls through pxb_scene_set_convex_meshes: tight bounds bit-identical to the reference's, broadphase events equal, plane-hull
    contacts and states equal to the oracle every step (the hull fixture carries the reference's cooking output; nothing is cooked here)."""
    z, sc = util.load_golden("hulls_on_plane")
    gpu, cpu = engine.Scene(sc, env_path=False), oracle.OracleScene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        assert np.array_equal(gpu.computeBounds(), z["bounds"][t]), f"bounds, step {t}"
        cpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        cpu.step(util.golden_order(z, t))
        assert not gpu.uses_env_path
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]) and np.abs(cg[:, :4 + 5 * 4] - cc[:, :4 + 5 * 4])[:, [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12]].max(initial=0) < 1e-6, f"contacts, step {t}"
        assert np.abs(gpu.getStates() - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"


@pytest.mark.parametrize("name", ["hulls_and_spheres", "spheres_into_hulls", "hulls_and_capsules", "capsules_into_hulls", "hull_pile", "box_hull_pile", "big_hull_pile", "config3_small"])
def test_sphere_convex_gpu_matches_oracle(oracle, name):
    """pcmContactSphereConvex / pcmContactCapsuleConvex on the device (hull support mapping, GJK, EPA, face + edge-edge contacts) against the
    oracle, teacher-forced from the golden states (hull-hull pairs of the *_into_* scenes excepted: they stop the step, see below)."""
    z, sc = util.load_golden(name)
    gpu, cpu = engine.Scene(sc), oracle.OracleScene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t]); cpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        cpu.step(util.golden_order(z, t))
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]), f"contact counts, step {t}"
        assert np.abs(cg[:, 1:8] - cc[:, 1:8]).max(initial=0) < 1e-6, f"first contact, step {t}"   # (hull_pile / box_hull_pile: pcmContactConvexConvex / BoxConvex)
        assert np.abs(gpu.getStates() - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"


def test_config3_shape_with_hulls_gpu_matches_oracle(oracle):
    """BASELINE config 3 at small size (spheres / capsules / library hulls falling into a walled bin, relaxed partitioning): every pair type of
    a10 incl. the SAT branch of the hull-hull manifold generation; GPU and oracle agree step by step."""
    sc = scenes.falling_primitives(6, 4, 6, kinds=("sphere", "capsule", "convex"), relaxed_partitioning=True)
    gpu, cpu = engine.Scene(sc, max_pairs=32 * len(sc.actors)), oracle.OracleScene(sc)
    for t in range(100):
        gpu.step(); cpu.step()
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        assert np.array_equal(gpu.getContacts()[:, 0], cpu.getContacts()[:, 0]), f"contact counts, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < 5e-5, f"state, step {t}"
        cpu.setStates(sg)
    assert cpu.unsupported_pairs == 0 and sg[:, 1].min() > 0.0


def test_big_hulls_free_running_gpu_matches_oracle(oracle):
    """Hulls of more than 32 vertices (hill-climbing support over Gu::BigConvexRawData: cube-map start sample + vertex adjacency) mixed with
    boxes, spheres and capsules: GPU and oracle agree step by step, free running, on the device-wide path and (environment ids set) on the
    fused environment path."""
    z, sc = util.load_golden("big_hull_pile")
    assert sum("samples" in h for h in sc.cooked_hulls()) >= 2
    for env in (False, True):
        a = sc.actors.copy()
        if env:
            a["envId"][a["flags"] & scenes.ACTOR_DYNAMIC != 0] = 0
        scn = scenes.Scene(sc.header, a, sc.hulls, sc.cooked)
        gpu, cpu = engine.Scene(scn, env_path=env), oracle.OracleScene(scn)
        for t in range(120):
            gpu.step(); cpu.step()
            assert gpu.uses_env_path == env
            assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
            assert np.array_equal(gpu.getContacts()[:, 0], cpu.getContacts()[:, 0]), f"contact counts, step {t}"
            sg = gpu.getStates()
            assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
            cpu.setStates(sg)
        assert cpu.unsupported_pairs == 0


def test_hull_without_cooked_data_is_rejected():
    """A convex actor whose hull was not uploaded (pxb_scene_set_convex_meshes) is refused when it is added -- nothing is silently skipped."""
    z, sc = util.load_golden("hulls_on_plane")
    bare = scenes.Scene(sc.header, sc.actors.copy(), sc.hulls)      # hull point clouds only, no cooked section
    with pytest.raises(engine.PhysxB200Error):
        engine.Scene(bare)


def test_all_geometry_types_free_running_gpu_matches_oracle(oracle):
    """Spheres, capsules, boxes and hulls in one pile (BASELINE config 3's pair types): GPU and oracle agree step by step, free running."""
    z, sc = util.load_golden("box_hull_pile")
    rng = np.random.RandomState(3)
    a = sc.actors.copy()
    scenes.set_sphere(a, 3, 0.15); scenes.set_capsule(a, 6, 0.1, 0.2)       # turn two of the pile's bodies into a sphere and a capsule
    mixed = scenes.Scene(sc.header, a, sc.hulls, sc.cooked)
    gpu, cpu = engine.Scene(mixed), oracle.OracleScene(mixed)
    for t in range(120):
        gpu.step(); cpu.step()
        assert cpu.unsupported_pairs == 0
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        assert np.array_equal(gpu.getContacts()[:, 0], cpu.getContacts()[:, 0]), f"contact counts, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        cpu.setStates(sg)


def test_cpp_host_mirror_snippet_hello_world():
    """examples/snippet_hello_world.cpp (C++ host code over the C ABI, include/physx_b200.hpp) steps the SnippetHelloWorld scene
    (BASELINE config 1) to the same state as the Python binding, bit for bit."""
    import json, os, subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "snippet_hello_world")
    if not os.path.exists(exe):
        subprocess.run(["make", "-f", "examples/Makefile"], cwd=os.path.dirname(os.path.dirname(exe)), check=True)
    for solver, arg in ((scenes.SOLVER_TGS, "tgs"), (scenes.SOLVER_PGS, "pgs")):
        out = json.loads(subprocess.run([exe, "60", "10", "10", arg], check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1])
        g = engine.Scene(scenes.box_stacks(solver=solver))
        for _ in range(60):
            g.step()
        st = g.getStates()
        assert out["bodies"] == 100 and out["pairs"] == len(g.getPairs()) and out["constraints"] == g.num_constraints and out["partitions"] == g.num_partitions
        assert np.float32(out["top_y"]) == st[-1, 1]   # %.9g round-trips a float32
        assert abs(out["checksum"] - float(st[:, :7].astype(np.float64).sum())) < 1e-4
        assert out["min_y"] > 0.49 and out["max_speed"] < 0.5   # the stacks stand (PGS leaves a larger residual wobble than TGS)


# ---- round 2: horizons and configs pinned to the reference directly (VERDICT r1 "next" 1) ----
@pytest.mark.parametrize("name", list(util.LONG_HORIZON))
def test_gpu_long_horizon_matches_reference_golden(name):
    """SURVEY 8d horizons on the GPU: config 1 proper (300 steps), config 2 / 5 shapes (120 steps), config 4 shape = dense pile with the exact
    first-fit in the reference's own order (TGS and PGS, 60 steps).  Bars: tests/util.py LONG_HORIZON."""
    z, sc = util.load_golden(name)
    tp, tl, ta = util.LONG_HORIZON[name]
    gpu = engine.Scene(sc, max_pairs=32 * len(sc.actors))
    for t in range(z["states"].shape[0] - 1):
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert util.rel_err(st[:, :3], ref[:, :3]) < tp and util.rel_err(st[:, 3:7], ref[:, 3:7]) < tp, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < tl and np.abs(st[:, 10:] - ref[:, 10:]).max() < ta, f"velocity, step {t}"
        if name != "pgs_pile_6x4x6":
            assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"


def test_gpu_pile_teacher_forced_steps_match_reference():
    z, sc = util.load_golden("pile_6x4x6")
    gpu = engine.Scene(sc, max_pairs=32 * len(sc.actors))
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert np.abs(st[:, :7] - ref[:, :7]).max() < 2e-5, f"pose, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"


@pytest.mark.parametrize("name", ["pile_12x6x12_stats", "fall_6x5x6_stats"])
@pytest.mark.parametrize("relaxed", [False, True])
def test_gpu_giant_island_statistics_match_reference(name, relaxed):
    """BASELINE configs 3 / 4 at test size against the REFERENCE's run of the same scene (energy / penetration statistics): the exact
    first-fit (canonical order) and PXB_FLAG_RELAXED_PARTITIONING both stay inside the bars of test_oracle_vs_reference.check_statistics."""
    from test_oracle_vs_reference import check_statistics
    z, sc = util.load_stats_golden(name)
    h = sc.header.copy(); h["reserved"][0] = 1 if relaxed else 0
    scn = scenes.Scene(h, sc.actors, sc.hulls, sc.cooked)
    gpu = engine.Scene(scn, max_pairs=32 * len(scn.actors))

    def run(steps):
        for _ in range(steps):
            gpu.step()
            yield util.run_statistics(scn, gpu.getStates(), gpu.getContacts())
    check_statistics(z, scn, None, run, name)


@pytest.mark.parametrize("name,types,bar", [("hull_pile", [5, 5], 0.0), ("box_hull_pile", [3, 5], 0.0), ("big_hull_pile", None, 0.0), ("config3_small", "hull", 0.0),
                                            ("hulls_and_capsules", [2, 5], 0.0), ("capsules_into_hulls", [2, 5], 0.0), ("hulls_and_spheres", [0, 5], 1e-6), ("hulls_on_plane", [1, 5], 1e-6)])
def test_gpu_hull_contacts_match_reference_golden_all_points(name, types, bar):
    """GPU hull contacts against the REFERENCE's contact reports directly (no oracle in between): same contact count per pair at every step and
    EVERY point / separation of the manifold equal to the reference's (bit-identical for the polygonal and capsule pairs, 1e-6 for sphere / plane
    vs hull), teacher-forced states."""
    z, sc = util.load_golden(name)
    gt = sc.actors["geomType"]
    gpu = engine.Scene(sc)
    checked = 0
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        ours = {(int(a), int(b)): c for (a, b), c in zip(gpu.getPairs(), gpu.getContacts())}
        ref = util.golden_contacts(z, t)
        for k in set(ref) | {k for k, c in ours.items() if c[0] > 0}:
            if (5 not in (int(gt[k[0]]), int(gt[k[1]]))) if types == "hull" else (types and sorted((int(gt[k[0]]), int(gt[k[1]]))) != types):
                continue
            assert k in ref and k in ours and int(ours[k][0]) == len(ref[k][1]), f"contact count, pair {k}, step {t}"
            n = len(ref[k][1]); op = ours[k][4:4 + 5 * n].reshape(n, 5); rp = ref[k][1]
            err = max(min(np.abs(op[i, :3] - rp[j, :3]).max() + abs(op[i, 3] - rp[j, 6]) for j in range(n)) for i in range(n))
            assert err <= bar, f"points, pair {k}, step {t}: {err}"
            nrm = min(np.abs(ours[k][1:4] - rp[0, 3:6]).max(), np.abs(ours[k][1:4] + rp[0, 3:6]).max())
            assert nrm <= max(bar, 1e-6), f"normal, pair {k}, step {t}"
            checked += 1
    assert checked >= 30


# ---- fused state export / stream-ordered velocity writes (a19) ----
@pytest.mark.parametrize("kind", ["envs", "envs_ragged", "devicewide", "pgs_devicewide"])
def test_state_export_matches_get_states(kind):
    """pxb_scene_set_state_export: the step itself stores the packed state block into device memory (at a row offset) and into mapped pinned host
    memory; both equal pxb_scene_get_states after every step, on the environment path (fused into k_env_solve), on environments whose bodies
    are not contiguous in dynamic-body order (k_states_export) and on the device-wide path."""
    import torch
    if kind == "envs":
        sc = scenes.env_grid_stacks(n_envs=9, jitter=0.01)
    elif kind == "envs_ragged":
        sc = scenes.env_grid_stacks(n_envs=6, stacks_per_env=2, height=3, jitter=0.01)
        perm = np.concatenate([[0], 1 + np.random.RandomState(1).permutation(len(sc.actors) - 1)])    # interleave the environments' bodies
        sc = scenes.Scene(sc.header, sc.actors[perm].copy())
    else:
        sc = scenes.box_stacks(n_stacks=3, height=5, half_extent=0.25, spacing=1.0, jitter=0.01, solver=scenes.SOLVER_PGS if kind.startswith("pgs") else scenes.SOLVER_TGS)
    gpu = engine.Scene(sc, env_path=not kind.endswith("devicewide"))
    n = gpu.num_dynamic
    dev = torch.zeros((n + 7, 13), dtype=torch.float32, device="cuda")
    host = torch.zeros((n, 13), dtype=torch.float32).pin_memory()
    gpu.step()                                    # one step without export first: the launch sequence changes when it is switched on
    gpu.setStateExport([dev.data_ptr(), host.data_ptr()], 0)
    for t in range(12):
        if t == 6:
            gpu.setStateExport([dev.data_ptr()], 5)    # targets and row offset may change between steps
        gpu.step()
        st = gpu.getStates()
        assert gpu.uses_env_path == kind.startswith("envs")
        off = 0 if t < 6 else 5
        assert np.array_equal(dev[off:off + n].cpu().numpy(), st), f"device target, step {t}"
        if t < 6:
            assert np.array_equal(host.numpy(), st), f"pinned host target, step {t}"
    gpu.setStateExport(())
    before = dev.clone()
    gpu.step()
    assert torch.equal(dev, before), "export switched off"
    with pytest.raises(engine.PhysxB200Error):
        gpu.setStateExport([np.zeros(4, np.float32).ctypes.data])   # pageable host memory is refused


def test_stream_ordered_velocity_write_overlaps_and_matches_the_synchronous_one():
    """pxb_set_rigid_dynamic_data_async on velocities runs on the copy stream and is joined between the narrowphase and the solver (split step
    graph): same result as the synchronous write, bit for bit, step after step."""
    import torch
    sc = scenes.env_grid_stacks(n_envs=16, jitter=0.01)
    a, b = engine.Scene(sc), engine.Scene(sc)
    n = a.num_dynamic
    rng = np.random.RandomState(0)
    lin = torch.zeros((n, 3), dtype=torch.float32).pin_memory(); ang = torch.zeros((n, 3), dtype=torch.float32).pin_memory()
    for t in range(10):
        v = (0.05 * rng.standard_normal((n, 3))).astype(np.float32); w = (0.1 * rng.standard_normal((n, 3))).astype(np.float32)
        lin.numpy()[:] = v; ang.numpy()[:] = w
        engine._check(a._lib, a._lib.pxb_set_rigid_dynamic_data_async(a._h, lin.data_ptr(), engine.RD_LINEAR_VELOCITY, n))
        engine._check(a._lib, a._lib.pxb_set_rigid_dynamic_data_async(a._h, ang.data_ptr(), engine.RD_ANGULAR_VELOCITY, n))
        b.setRigidDynamicData(engine.RD_LINEAR_VELOCITY, v); b.setRigidDynamicData(engine.RD_ANGULAR_VELOCITY, w)
        if t == 4:   # a read between the write and the step is ordered after the write
            assert np.array_equal(a.getRigidDynamicData(engine.RD_LINEAR_VELOCITY), v)
        a.step(); b.step()
        assert np.array_equal(a.getStates(), b.getStates()), f"step {t}"


# ---- standalone broadphase object (pxb_bp_*: what the plugin shim's Bp::BroadPhase forwards to) ----
def _bp_groups(sc):
    dyn = (sc.actors["flags"] & scenes.ACTOR_DYNAMIC) != 0
    ids = np.arange(len(sc.actors), dtype=np.uint32)
    return np.where(dyn, ((ids + 1) << 3) | 2, 0).astype(np.uint32), dyn     # Bp::getFilterGroup_Dynamics / eSTATICS (BpFiltering.h:79-95)


@pytest.mark.parametrize("name", ["stacks_4x8_jitter", "tumble_12", "envs_4", "hulls_on_plane", "capsules_on_boxes"])
def test_broadphase_object_matches_abp_bit_exact(name):
    """pxb_bp_update / pxb_bp_fetch driven the way Bp::AABBManager drives a Bp::BroadPhase (created list on the first update, updated list of the
    dynamic objects afterwards, tight bounds + contact distance + filter groups): created / deleted pair lists identical to the reference's ABP
    (PxBroadPhase(eABP) + PxAABBManager, golden) at every step."""
    z, sc = util.load_golden(name)
    groups, dyn = _bp_groups(sc)
    n = len(sc.actors)
    dist = np.full(n, float(sc.header["contactOffset"]), np.float32)
    bp = engine.BroadPhase(n)
    for t in range(z["bounds"].shape[0]):
        if t == 0:
            bp.update(z["bounds"][t], dist, groups, created=np.arange(n))
        else:
            bp.update(z["bounds"][t], dist, groups, updated=np.nonzero(dyn)[0])
        created, deleted = bp.fetch()
        assert np.array_equal(created, util.golden_created(z, t)), f"created, step {t}"
        assert np.array_equal(deleted, util.golden_deleted(z, t)), f"deleted, step {t}"


def test_broadphase_object_removal_and_environment_filter():
    """Lost overlaps caused by a removal are not reported (BpBroadPhase.h:181-192), a re-added object finds its overlaps again, equal groups never
    pair, different environment ids never pair."""
    b = np.array([[0, 0, 0, 1, 1, 1], [0.5, 0, 0, 1.5, 1, 1], [0.9, 0, 0, 1.9, 1, 1], [5, 5, 5, 6, 6, 6]], np.float32)
    d = np.zeros(4, np.float32)
    g = np.array([(1 << 3) | 2, (2 << 3) | 2, (3 << 3) | 2, 0], np.uint32)
    bp = engine.BroadPhase(16)
    bp.update(b, d, g, created=[0, 1, 2, 3])
    c, dl = bp.fetch()
    assert c.tolist() == [[0, 1], [0, 2], [1, 2]] and len(dl) == 0
    bp.update(b, d, g, removed=[1])
    c, dl = bp.fetch()
    assert len(c) == 0 and len(dl) == 0
    bp.update(b, d, g, created=[1])
    c, dl = bp.fetch()
    assert c.tolist() == [[0, 1], [1, 2]] and len(dl) == 0
    b2 = b.copy(); b2[2, [0, 3]] += 5.0
    bp.update(b2, d, g, updated=[2])
    c, dl = bp.fetch()
    assert len(c) == 0 and dl.tolist() == [[0, 2], [1, 2]]
    # equal groups (shapes of one actor) and different environments
    bp2 = engine.BroadPhase(16)
    g2 = np.array([(1 << 3) | 2, (1 << 3) | 2, (2 << 3) | 2, (3 << 3) | 2], np.uint32)
    b3 = np.array([[0, 0, 0, 1, 1, 1]] * 4, np.float32)
    env = np.array([0, 0, 0, 1], np.uint32)
    bp2.update(b3, d, g2, envs=env, created=[0, 1, 2, 3])
    c, dl = bp2.fetch()
    assert c.tolist() == [[0, 2], [1, 2]]


# ---- a7: touch found / lost event lists ----
@pytest.mark.parametrize("kind", ["tumble_devicewide", "tumble_env", "envs_ragged", "fall_mixed"])
def test_touch_events_are_the_changes_of_the_touching_pair_set(kind):
    """pxb_scene_get_touch_found / _lost: exactly the pairs whose contact state changed in the step -- incl. touching pairs that left the
    broadphase -- on both paths and for the GJK-family pairs (their events are raised by k_narrowphase_gjk)."""
    sc = {"tumble_devicewide": scenes.tumbling_boxes(n=12, seed=7), "tumble_env": scenes.tumbling_boxes(n=12, seed=7), "envs_ragged": scenes.env_ragged(),
          "fall_mixed": scenes.falling_primitives(4, 3, 4, kinds=("sphere", "capsule", "box"))}[kind]
    gpu = engine.Scene(sc, max_pairs=32 * len(sc.actors), env_path=(kind != "tumble_devicewide"))
    prev, events = set(), 0
    for t in range(150):
        gpu.step()
        cur = {(int(a), int(b)) for (a, b), c in zip(gpu.getPairs(), gpu.getContacts()) if c[0] > 0}
        found = {(int(a), int(b)) for a, b in gpu.getTouchFound()}; lost = {(int(a), int(b)) for a, b in gpu.getTouchLost()}
        assert found == cur - prev, f"touch found, step {t}"
        assert lost == prev - cur, f"touch lost, step {t}"
        events += len(found) + len(lost)
        prev = cur
    assert events > 20


def test_touch_events_match_the_reference_contact_reports():
    """Teacher-forced from the golden states: the touch-found / touch-lost events equal the changes of the reference's own touching pair set
    (its contact reports carry every touching pair of the step: eNOTIFY_TOUCH_FOUND | eNOTIFY_TOUCH_PERSISTS)."""
    z, sc = util.load_golden("tumble_12")
    gpu = engine.Scene(sc, env_path=False)
    prev = set()
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        cur = set(util.golden_contact_counts(z, t))
        assert {(int(a), int(b)) for a, b in gpu.getTouchFound()} == cur - prev, f"touch found, step {t}"
        assert {(int(a), int(b)) for a, b in gpu.getTouchLost()} == prev - cur, f"touch lost, step {t}"
        prev = cur


# ---- a11: material table / combine modes ----
@pytest.mark.parametrize("solver", ["tgs", "pgs"])
@pytest.mark.parametrize("path", ["auto", "devicewide"])
def test_material_table_gpu_matches_oracle(oracle, solver, path):
    """pxb_scene_set_materials: per-pair combined friction / restitution (every PxCombineMode, eDISABLE_FRICTION) on both paths and both solvers,
    step by step against the oracle (itself pinned against the reference: test_material_table_matches_reference)."""
    sc = scenes.material_mix(solver=scenes.SOLVER_PGS if solver == "pgs" else scenes.SOLVER_TGS)
    gpu, cpu = engine.Scene(sc, env_path=(path == "auto")), oracle.OracleScene(sc)
    for t in range(90):
        gpu.step(); cpu.step()
        assert gpu.uses_env_path == (path == "auto")
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pair set, step {t}"
        cg, cc = gpu.getContacts(), cpu.getContacts()
        assert np.array_equal(cg[:, 0], cc[:, 0]) and np.abs(cg - cc).max(initial=0) < 1e-4, f"contacts / applied forces, step {t}"
        sg = gpu.getStates()
        assert np.abs(sg - cpu.getStates()).max() < TOL_STEP, f"state, step {t}"
        cpu.setStates(sg)


def test_material_table_gpu_matches_reference_golden():
    z, sc = util.load_golden("materials_mix")
    gpu = engine.Scene(sc, env_path=False)
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t])
        gpu.setConstraintOrder(util.golden_order(z, t))
        gpu.step()
        st, ref = gpu.getStates(), z["states"][t + 1]
        assert np.abs(st[:, :7] - ref[:, :7]).max() < 2e-5, f"pose, step {t}"
        assert np.abs(st[:, 7:10] - ref[:, 7:10]).max() < 2e-4 and np.abs(st[:, 10:] - ref[:, 10:]).max() < 2e-3, f"velocity, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"


def test_material_table_rejects_unsupported_entries():
    sc = scenes.material_mix()
    bad = sc.materials.copy(); bad["restitution"][1] = -100.0     # compliant contact
    with pytest.raises(engine.PhysxB200Error):
        engine.Scene(scenes.Scene(sc.header, sc.actors, materials=bad))


# ---- f3: tensor front end (ovphysx-style TensorBinding over DLPack / torch tensors) ----
def test_tensor_binding_reads_and_writes_device_and_host_tensors():
    import torch
    from physx_b200 import tensor_api as ta
    sc = scenes.env_grid_stacks(n_envs=8, jitter=0.01)
    gpu = engine.Scene(sc)
    for _ in range(3):
        gpu.step()
    st = gpu.getStates()
    n = gpu.num_dynamic
    pose = ta.TensorBinding(gpu, ta.TensorType.RIGID_BODY_POSE); vel = ta.TensorBinding(gpu, ta.TensorType.RIGID_BODY_VELOCITY)
    assert pose.shape == (n, 7) and vel.shape == (n, 6) and pose.count == n
    p = torch.empty((n, 7), dtype=torch.float32, device="cuda"); v = np.empty((n, 6), np.float32)
    pose.read(p); vel.read(v)                                   # CUDA tensor and host array
    assert np.array_equal(p.cpu().numpy(), st[:, :7]) and np.array_equal(v, st[:, 7:13])       # (p, q) and (lin, ang): the packed state's own order
    m = torch.empty(n, dtype=torch.float32, device="cuda"); ta.TensorBinding(gpu, ta.TensorType.RIGID_BODY_MASS).read(m)
    assert np.allclose(m.cpu().numpy(), sc.actors["mass"][1:], rtol=1e-6)
    # write through DLPack with indices and with a mask
    idx = torch.tensor([5, 1, 40], dtype=torch.int64)
    newv = torch.arange(18, dtype=torch.float32, device="cuda").reshape(3, 6)
    vel.write(torch.utils.dlpack.to_dlpack(newv) if False else newv, indices=idx)
    vel.read(v)
    assert np.array_equal(v[[5, 1, 40]], newv.cpu().numpy())
    mask = np.zeros(n, bool); mask[[2, 7]] = True
    newp = st[[2, 7], :7].copy(); newp[:, 1] += 1.0
    pose.write(newp, mask=mask)
    pose.read(p)
    assert np.array_equal(p.cpu().numpy()[[2, 7]], newp)
    with pytest.raises(ValueError):
        pose.read(torch.empty((n, 6), dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        ta.TensorBinding(gpu, ta.TensorType.RIGID_BODY_FORCE).read(p)
    # forces: same effect as PXB_RD_FORCE on a twin scene
    a, b = engine.Scene(sc), engine.Scene(sc)
    F = np.zeros((n, 3), np.float32); F[:, 0] = 3.0
    ta.TensorBinding(a, ta.TensorType.RIGID_BODY_FORCE).write(F); b.setForces(forces=F)
    a.step(); b.step()
    assert np.array_equal(a.getStates(), b.getStates())
    rep = ta.get_contact_report(a)
    assert len(rep["actor0"]) == a.num_constraints and rep["counts"].sum() == len(rep["positions"]) and np.all(rep["impulses"] >= 0)


# ---- actor removal (Bp::AABBManagerBase::removeBounds / removeDynamic) ----
@pytest.mark.parametrize("path", ["auto", "devicewide"])
def test_removed_actor_leaves_the_simulation_and_the_rest_continues(oracle, path):
    """pxb_scene_remove_actors: the middle box of a stack is taken out after 10 steps.  Its pairs are reported deleted and touch-lost in the next
    step, it is no longer integrated (state frozen, indices unchanged), and every other body continues exactly like a scene that never had
    it (oracle scene built without the actor, started from the same states)."""
    sc = scenes.box_stacks(n_stacks=2, height=3, half_extent=0.25, spacing=1.0, jitter=0.01)      # actors: plane 0 | stack 0: 1 2 3 | stack 1: 4 5 6
    gpu = engine.Scene(sc, env_path=(path == "auto"))
    for _ in range(10):
        gpu.step()
    st10 = gpu.getStates()
    gpu.removeActors([2])
    keep = [0, 1, 3, 4, 5, 6]
    cpu = oracle.OracleScene(scenes.Scene(sc.header, sc.actors[keep].copy()))
    dyn_keep = [0, 2, 3, 4, 5]
    cpu.setStates(st10[dyn_keep])
    remap = {a: i for i, a in enumerate(keep)}
    for t in range(40):
        gpu.step(); cpu.step()
        if t == 0:
            assert {(1, 2), (2, 3)} <= {(int(a), int(b)) for a, b in gpu.getDeletedPairs()}
            assert {(1, 2), (2, 3)} <= {(int(a), int(b)) for a, b in gpu.getTouchLost()}
        pg = gpu.getPairs()
        assert 2 not in pg
        assert {(remap[int(a)], remap[int(b)]) for a, b in pg} == {(int(a), int(b)) for a, b in cpu.getPairs()}, f"pair set, step {t}"
        sg = gpu.getStates()
        assert np.array_equal(sg[1], st10[1]), "the removed body keeps its last state"
        d = np.abs(sg[dyn_keep] - cpu.getStates())     # the oracle scene starts with cold manifolds / friction patches, the GPU scene keeps its warm ones
        assert d[:, :7].max() < 1e-4 and d[:, 7:].max() < 5e-3, f"state, step {t}"
    assert gpu.num_dynamic == 6 and sg[2, 1] < st10[2, 1] - 0.2, "the box above fell onto the bottom box"
    with pytest.raises(engine.PhysxB200Error):
        gpu.removeActors([99])


# ---- f3: PxDirectGPUAPI::copyContactData (PxGpuContactPair records + PxContactPatch / PxContact / force / PxFrictionPatch streams in device memory) ----
def _combine(a, b, mode):
    return {0: 0.5 * (a + b), 1: min(a, b), 2: a * b, 3: max(a, b)}[int(mode)]


def _check_contact_data(gpu, sc, host):
    pairs, con = gpu.getPairs(), gpu.getContacts()
    touching = con[:, 0] > 0
    rec = host["records"]
    assert host["total_pairs"] == int(touching.sum()) == len(rec)
    # the environment path keeps its pairs per environment, the host getters report them key-sorted: compare as sets keyed by the actor pair
    key = lambda a, b: (min(int(a), int(b)), max(int(a), int(b)))
    ref = {key(*pairs[i]): con[i] for i in np.nonzero(touching)[0]}
    dyn_of = {int(a): d for d, a in enumerate(np.nonzero(sc.actors["flags"] & 1)[0])}     # PxRigidDynamicGPUIndex = position among the dynamic actors
    mats = sc.materials
    for r, (rc, pt, fr, s0) in enumerate(zip(rec, host["patches"], host["friction"], host["start_indices"])):
        c = ref[key(rc["transformCacheRef0"], rc["transformCacheRef1"])]
        k = int(rc["nbContacts"])
        assert k == int(c[0]) == int(pt["nbContacts"]) and rc["nbPatches"] == 1 and pt["startContactIndex"] == 0
        assert np.array_equal(pt["normal"], c[1:4])                                        # normal (body1 -> body0), bit for bit
        pts = c[4:4 + 5 * k].reshape(k, 5)
        assert np.array_equal(host["points"][s0:s0 + k], pts[:, :4])                       # PxContact: point + separation
        assert np.array_equal(host["forces"][s0:s0 + k], pts[:, 4])                        # applied normal impulses
        assert np.array_equal(pt["massModification"], np.ones(4, np.float32)) and pt["damping"] == 0
        a0, a1 = int(rc["transformCacheRef0"]), int(rc["transformCacheRef1"])
        assert rc["actor0"] == a0 and rc["actor1"] == a1
        for a, node in ((a0, rc["nodeIndex0"]), (a1, rc["nodeIndex1"])):
            assert int(node) == (dyn_of[a] if a in dyn_of else 0xffffffff)
        if mats is not None and len(mats):
            m0, m1 = mats[sc.actors["materialIndex"][a0]], mats[sc.actors["materialIndex"][a1]]
            assert pt["materialIndex0"] == sc.actors["materialIndex"][a0] and pt["materialIndex1"] == sc.actors["materialIndex"][a1]
            b0, b1 = int(m0["bits"]), int(m1["bits"])
            rest = _combine(m0["restitution"], m1["restitution"], max((b0 >> 4) & 15, (b1 >> 4) & 15))
            assert abs(float(pt["restitution"]) - rest) < 1e-6
            if ((b0 | b1) >> 8) & 1:
                assert pt["staticFriction"] == 0 and pt["dynamicFriction"] == 0 and fr["anchorCount"] == 0
            else:
                fm = max(b0 & 15, b1 & 15)
                dynf = max(_combine(m0["dynamicFriction"], m1["dynamicFriction"], fm), 0.0)
                assert abs(float(pt["dynamicFriction"]) - dynf) < 1e-6
        # friction patch: anchors of the pair's friction patch in world space lie in the contact area, impulses are tangential and inside the friction cone
        na = int(fr["anchorCount"])
        assert na in (0, 1, 2)
        n = pt["normal"].astype(np.float64)
        lo, hi = pts[:, :3].min(0) - 0.05, pts[:, :3].max(0) + 0.05
        total_normal = float(pts[:, 4].sum())
        for j in range(na):
            assert np.all(fr["anchorPositions"][j] >= lo - 0.3) and np.all(fr["anchorPositions"][j] <= hi + 0.3)
            imp = fr["anchorImpulses"][j].astype(np.float64)
            assert abs(imp @ n) <= 1e-4 * (1.0 + np.linalg.norm(imp))
            assert np.linalg.norm(imp) <= float(pt["staticFriction"]) * total_normal * 1.001 + 1e-5
        for j in range(na, 2):
            assert not fr["anchorImpulses"][j].any() and not fr["anchorPositions"][j].any()
    return rec


@pytest.mark.gpu
@pytest.mark.parametrize("env_path", [True, False])
def test_copy_contact_data_matches_host_contacts(env_path):
    from physx_b200 import tensor_api as ta
    sc = scenes.material_mix()
    gpu = engine.Scene(sc, env_path=env_path)
    with pytest.raises(RuntimeError):                      # contact data has to be switched on before the step
        ta.GpuContactData(gpu, 16)
    gpu.enableContactData()
    for _ in range(25):
        gpu.step()
    cd = ta.GpuContactData(gpu, 4096)
    host = cd.to_host()
    assert host["total_pairs"] > 20
    _check_contact_data(gpu, sc, host)
    assert (host["friction"]["anchorCount"] > 0).any() and np.abs(host["friction"]["anchorImpulses"]).max() > 1e-4   # the sliding boxes carry friction impulses
    # fewer records than pairs: the count still reports all of them, only max_pairs records are written
    small = ta.GpuContactData(gpu, 5).to_host()
    assert small["total_pairs"] == host["total_pairs"] and len(small["records"]) == 5
    assert np.array_equal(small["records"]["transformCacheRef0"], host["records"]["transformCacheRef0"][:5])


@pytest.mark.gpu
def test_copy_contact_data_same_on_both_paths():
    from physx_b200 import tensor_api as ta
    sc = scenes.env_grid_stacks(n_envs=6, jitter=0.01)
    out = []
    for env_path in (True, False):
        gpu = engine.Scene(sc, env_path=env_path)
        gpu.enableContactData()
        for _ in range(12):
            gpu.step()
        h = ta.GpuContactData(gpu, 8192).to_host()
        _check_contact_data(gpu, sc, h)
        order = np.lexsort((h["records"]["transformCacheRef1"], h["records"]["transformCacheRef0"]))
        out.append((h, order))
    (a, oa), (b, ob) = out
    assert a["total_pairs"] == b["total_pairs"]
    for f in ("transformCacheRef0", "transformCacheRef1", "nbContacts", "nodeIndex0", "nodeIndex1"):
        assert np.array_equal(a["records"][f][oa], b["records"][f][ob])
    assert np.array_equal(a["patches"][oa], b["patches"][ob])
    assert np.array_equal(a["friction"][oa], b["friction"][ob])          # friction anchors and impulses bit-identical across the two paths


# ---- a1 with local poses: PxShape::setLocalPose + PxRigidBody::setCMassLocalPose (transform cache, actor-pose I/O) ----
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["local_poses_mix", "pgs_local_poses_mix"])
@pytest.mark.parametrize("env_path", [True, False])
def test_local_poses_gpu_matches_oracle_and_reference(oracle, name, env_path):
    """Shape frame, actor frame and centre-of-mass frame all different per body.  The engine integrates body frames, composes the shapes' world poses
    into the transform cache in the reference's operation order and reports actor poses: initial actor poses (through setCMassLocalPose and back)
    bit-identical to the reference; 30 free-running steps bit-identical to the oracle with TGS (1e-5 with PGS) and within 1e-5 (pose) of the reference."""
    z, sc = util.load_golden(name)
    gpu = engine.Scene(sc, env_path=env_path); cpu = oracle.OracleScene(sc)
    assert np.array_equal(gpu.getStates(), z["states"][0])
    pgs = name.startswith("pgs")
    for t in range(30):
        gpu.setConstraintOrder(util.golden_order(z, t)); gpu.step(); cpu.step(util.golden_order(z, t))
        a, b, ref = gpu.getStates(), cpu.getStates(), z["states"][t + 1]
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()), f"pairs, step {t}"
        if pgs:
            assert np.abs(a - b).max() < 1e-5, f"gpu vs oracle, step {t}"
        else:
            assert np.array_equal(a, b), f"gpu vs oracle, step {t}"
            assert np.array_equal(gpu.getContacts(), cpu.getContacts()), f"contacts, step {t}"
        assert np.abs(a[:, :7] - ref[:, :7]).max() < 1e-5 and np.abs(a[:, 7:] - ref[:, 7:]).max() < 5e-4, f"gpu vs reference, step {t}"


@pytest.mark.gpu
def test_local_poses_pose_io_is_in_actor_frames():
    import torch
    from physx_b200 import tensor_api as ta
    sc = scenes.local_pose_mix()
    gpu = engine.Scene(sc)
    for _ in range(5):
        gpu.step()
    st = gpu.getStates()
    n = gpu.num_dynamic
    pose = gpu.getRigidDynamicData(engine.RD_GLOBAL_POSE)                       # PxDirectGPUAPI pose wire format: q.xyzw, p.xyz
    assert np.array_equal(pose[:, 4:7], st[:, 0:3]) and np.array_equal(pose[:, 0:4], st[:, 3:7])
    t = torch.empty((n, 7), dtype=torch.float32, device="cuda"); ta.TensorBinding(gpu, ta.TensorType.RIGID_BODY_POSE).read(t)
    assert np.array_equal(t.cpu().numpy(), st[:, :7])
    # writing the actor poses back re-derives the body frames: one step later the result is the same to rounding (body2World = pose * body2Actor is not exactly invertible)
    ref = engine.Scene(sc)
    for _ in range(5):
        ref.step()
    gpu.setRigidDynamicData(engine.RD_GLOBAL_POSE, pose)
    back = gpu.getStates()
    assert np.abs(back[:, :7] - st[:, :7]).max() < 1e-6 and np.array_equal(back[:, 7:], st[:, 7:])
    gpu.step(); ref.step()
    assert np.abs(gpu.getStates() - ref.getStates()).max() < 1e-4
    # a moved actor really moves (shape and body frame follow)
    moved = pose.copy(); moved[3, 5] += 2.0
    gpu.setRigidDynamicData(engine.RD_GLOBAL_POSE, moved)
    assert abs(gpu.getStates()[3, 1] - (back[3, 1] + 2.0)) < 1e-5
    with pytest.raises(RuntimeError):      # the fused export reports body frames
        buf = torch.zeros((n, 13), dtype=torch.float32, device="cuda")
        gpu.setStateExport((buf.data_ptr(),))


# ---- BASELINE.json's full sizes through size-independent properties (configs 3, 4, 5; config 2: test_config2_full_size_properties) ----
def _fast_stats(sc, st, con):
    """kinetic energy (linear part), deepest separation and mean penetration of a large run, vectorised"""
    dyn = (sc.actors["flags"] & scenes.ACTOR_DYNAMIC) != 0
    mass = sc.actors["mass"][dyn].astype(np.float64)
    ke = float(0.5 * (mass * (st[:, 7:10].astype(np.float64) ** 2).sum(1)).sum())
    cnt = con[:, 0].astype(np.int64)
    seps = con[:, 4:24].reshape(-1, 4, 5)[:, :, 3]
    valid = np.arange(4)[None, :] < cnt[:, None]
    s = seps[valid].astype(np.float64)
    return dict(ke=ke, min_sep=float(s.min()) if len(s) else 0.0, mean_pen=float(np.clip(-s, 0, None).mean()) if len(s) else 0.0)


def _pair_set_properties(sc, gpu):
    pairs = gpu.getPairs()
    keys = pairs[:, 0].astype(np.int64) << 32 | pairs[:, 1]
    assert np.all(np.diff(keys) > 0), "pair list sorted and unique"
    dyn = (sc.actors["flags"] & scenes.ACTOR_DYNAMIC) != 0
    assert np.all(dyn[pairs[:, 0]] | dyn[pairs[:, 1]]), "every pair has a dynamic actor"
    return pairs


@pytest.mark.gpu
def test_config4_full_size_properties():
    """200 000 boxes in one island (BASELINE config 4), exact first-fit colouring: the run is bitwise deterministic, the pile neither explodes nor sinks
    (kinetic energy decays, penetration stays at the contact-offset scale), every touching pair is a constraint, and the box-box regeneration worklist
    (k_boxbox_generate) gives the same result as regeneration inside k_narrowphase."""
    import os
    sc = scenes.box_pile(100, 20, 100)
    assert sc.n_dynamic == 200000
    runs = []
    for phases in ("1", "0", "1"):
        os.environ["PXB_BOX_PHASES"] = phases
        try:
            gpu = engine.Scene(sc, max_pairs=16 * len(sc.actors))
        finally:
            del os.environ["PXB_BOX_PHASES"]
        assert not gpu.uses_env_path
        ke = []
        for t in range(24):
            gpu.step()
            if t in (3, 23):
                ke.append(_fast_stats(sc, gpu.getStates(), gpu.getContacts()))
        runs.append((gpu.getStates(), gpu.getPairs(), ke))
        if phases == "1" and len(runs) == 1:
            pairs = _pair_set_properties(sc, gpu)
            con = gpu.getContacts()
            assert gpu.num_constraints == int(np.count_nonzero(con[:, 0])) > 2.0e6
            assert gpu.num_partitions <= 64
    (s0, p0, k0), (s1, p1, k1), (s2, p2, k2) = runs
    assert np.isfinite(s0).all() and np.abs(np.linalg.norm(s0[:, 3:7], axis=1) - 1).max() < 1e-5
    assert np.array_equal(s0, s2) and np.array_equal(p0, p2), "bitwise deterministic"
    assert np.array_equal(s0, s1) and np.array_equal(p0, p1), "worklist regeneration == in-kernel regeneration"
    assert k0[1]["ke"] < k0[0]["ke"] * 1.5 and k0[1]["min_sep"] > -0.05 and k0[1]["mean_pen"] < 5e-3
    assert s0[:, 1].min() > 0.2 and s0[:, 1].max() < 11.0      # nobody fell through the floor or got shot out of the bin


@pytest.mark.gpu
def test_config3_full_size_properties():
    """1 048 576 spheres / capsules / convex hulls falling into a walled bin (BASELINE config 3): bitwise deterministic, no unsupported pair, finite,
    and the four-phase GJK family gives exactly what the single kernel gives (states and pair sets after 70 steps of the fall)."""
    import os
    sc = scenes.falling_primitives(128, 64, 128, kinds=("sphere", "capsule", "convex"))
    assert sc.n_dynamic == 1048576
    out = []
    for phases in ("1", "0"):
        os.environ["PXB_GJK_PHASES"] = phases
        try:
            gpu = engine.Scene(sc, max_pairs=16 * len(sc.actors))
        finally:
            del os.environ["PXB_GJK_PHASES"]
        for t in range(70):
            gpu.step()          # fetchResults raises on E_UNSUPPORTED_PAIR / capacity errors
        st = gpu.getStates()
        assert np.isfinite(st).all()
        out.append((st, _pair_set_properties(sc, gpu), gpu.num_constraints))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2] == out[1][2] > 5.0e5
    assert out[0][0][:, 1].min() > 0.0


@pytest.mark.gpu
def test_config5_shard_full_size_properties():
    """one GPU's shard of BASELINE config 5: 4096 envs x 128 boxes = 524 288 bodies.  No cross-environment pair, every environment's result equals the
    result of the same environment simulated alone (environments never interact), stacks stay standing."""
    sc = scenes.env_grid_stacks(n_envs=4096, stacks_per_env=16)
    assert sc.n_dynamic == 524288
    gpu = engine.Scene(sc)
    for _ in range(20):
        gpu.step()
    assert gpu.uses_env_path
    pairs = _pair_set_properties(sc, gpu)
    env = sc.actors["envId"]
    e0, e1 = env[pairs[:, 0]], env[pairs[:, 1]]
    assert np.all((e0 == e1) | (e0 == scenes.NO_ENV) | (e1 == scenes.NO_ENV)), "no cross-environment pair"
    st = gpu.getStates()
    assert np.isfinite(st).all()
    dyn_env = env[(sc.actors["flags"] & scenes.ACTOR_DYNAMIC) != 0]
    for e in (0, 1777, 4095):                      # the same environment alone (plus the shared ground plane): bit-identical trajectories
        keep = (env == e) | (env == scenes.NO_ENV)
        sub_actors = sc.actors[keep].copy()
        sub_actors["envId"][sub_actors["envId"] == e] = 0
        alone = engine.Scene(scenes.Scene(sc.header, sub_actors))
        for _ in range(20):
            alone.step()
        assert np.array_equal(alone.getStates(), st[dyn_env == e]), f"environment {e}"


# ---- f1: the default simulation filter shader on the device ----
@pytest.mark.gpu
@pytest.mark.parametrize("env_path", [True, False])
def test_default_filter_shader_gpu_matches_oracle_and_reference(oracle, env_path):
    """collision groups + groups masks (PxDefaultSimulationFilterShader): suppressed pairs stay in the pair lists and generate no contacts.  GPU == oracle bit for bit
    over 90 free-running steps (pairs, created / deleted, contacts, states); within 1e-4 of the reference's poses."""
    z, sc = util.load_golden("filter_groups_mix")
    gpu, cpu = engine.Scene(sc, env_path=env_path), oracle.OracleScene(sc)
    for t in range(90):
        gpu.setConstraintOrder(util.golden_order(z, t)); gpu.step(); cpu.step(util.golden_order(z, t))
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()) and np.array_equal(gpu.getCreatedPairs(), cpu.getCreatedPairs()) and np.array_equal(gpu.getDeletedPairs(), cpu.getDeletedPairs()), f"pairs, step {t}"
        assert np.array_equal(gpu.getContacts(), cpu.getContacts()), f"contacts, step {t}"
        assert np.array_equal(gpu.getStates(), cpu.getStates()), f"states, step {t}"
    assert np.abs(gpu.getStates()[:, :7] - z["states"][90][:, :7]).max() < 1e-4
    # the same scene without the filter section ends somewhere else (the ghosts rest on the solids)
    plain = engine.Scene(scenes.Scene(sc.header, sc.actors.copy()), env_path=env_path)
    for _ in range(90):
        plain.step()
    assert plain.getStates()[6, 1] > 0.7 > gpu.getStates()[6, 1]


@pytest.mark.gpu
def test_filter_shader_api_errors():
    sc = scenes.box_stacks(n_stacks=2, height=2)
    gpu = engine.Scene(sc)
    lib = gpu._lib
    fd = np.zeros((len(sc.actors), 4), np.uint32)
    assert lib.pxb_scene_set_filter_data(gpu._h, 0, len(fd), fd.ctypes.data) < 0          # before a filter shader is configured
    cfg = scenes.default_filter_config().reshape(1).copy()
    cfg["ops"][0][0] = 9
    assert lib.pxb_scene_set_filter_shader(gpu._h, cfg.ctypes.data) < 0                  # PxFilterOp out of range
    cfg["ops"][0][0] = 0
    assert lib.pxb_scene_set_filter_shader(gpu._h, cfg.ctypes.data) == 0
    assert lib.pxb_scene_set_filter_data(gpu._h, 0, len(fd), fd.ctypes.data) == 0
    ref = engine.Scene(sc)
    for _ in range(10):
        gpu.step(); ref.step()
    assert np.array_equal(gpu.getStates(), ref.getStates())                               # the extension's default state filters nothing


# ---- PxShape::setContactOffset / setRestOffset per shape ----
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["shape_offsets_mix", "pgs_shape_offsets_mix"])
@pytest.mark.parametrize("env_path", [True, False])
def test_shape_offsets_gpu_matches_oracle_and_reference(oracle, name, env_path):
    """Every shape with its own contact / rest offset: bounds inflated per shape in the broadphase, contact distance and rest distance per pair.  GPU vs oracle over 120
    steps (oracle re-synchronised to the GPU state every step): pair sets, created / deleted reports and contact counts identical, states within TOL_STEP (bit-identical for the
    first 40 steps with TGS); within 2e-4 (pose; PGS 2e-3) of the reference at step 40."""
    z, sc = util.load_golden(name)
    gpu, cpu = engine.Scene(sc, env_path=env_path), oracle.OracleScene(sc)
    pgs = name.startswith("pgs")
    for t in range(120):
        gpu.setConstraintOrder(util.golden_order(z, t)); gpu.step(); cpu.step(util.golden_order(z, t))
        assert np.array_equal(gpu.getPairs(), cpu.getPairs()) and np.array_equal(gpu.getCreatedPairs(), cpu.getCreatedPairs()) and np.array_equal(gpu.getDeletedPairs(), cpu.getDeletedPairs()), f"pairs, step {t}"
        a, b = gpu.getStates(), cpu.getStates()
        assert np.array_equal(gpu.getContacts()[:, 0], cpu.getContacts()[:, 0]), f"contact counts, step {t}"
        assert np.abs(a - b).max() < TOL_STEP, f"states, step {t}"      # tumbling primitives: the oracle is re-synchronised every step (as in test_gpu_matches_oracle_primitives)
        if t < 40 and not pgs:
            assert np.array_equal(a, b), f"bit-identical while nothing tumbles, step {t}"
        cpu.setStates(a)
        if t == 39:
            assert np.abs(a[:, :7] - z["states"][40][:, :7]).max() < (2e-3 if pgs else 2e-4)   # PGS: the block-solver difference
    lib = gpu._lib
    bad = np.array([[0.01, 0.02]], np.float32)      # contactOffset must exceed restOffset
    assert lib.pxb_scene_set_shape_offsets(gpu._h, 1, 1, bad.ctypes.data) < 0


# ---- PxSceneFlag::eENABLE_BODY_ACCELERATIONS: PxDirectGPUAPI acceleration getters, start / finish events ----
@pytest.mark.gpu
@pytest.mark.parametrize("env_path", [True, False])
def test_acceleration_getters_and_events(env_path):
    """eLINEAR_ACCELERATION / eANGULAR_ACCELERATION = (velocity - velocity the step started from) * (1 / dt), the expression of the reference's getter kernels
    (updateBodiesAndShapes.cu:1063-1106) and of its CPU path (NpSceneFetchResults.cpp:170-186): bit-exact against the velocities read through the same API, with a user
    velocity write between steps counted as the new start velocity (NpRigidDynamic.cpp:245-252).  The *_device_ev variants take PxDirectGPUAPI's start / finish events."""
    import torch
    sc = scenes.env_grid_stacks(n_envs=8)
    with pytest.raises(engine.PhysxB200Error):
        engine.Scene(sc, env_path=env_path).getRigidDynamicData(engine.RD_LINEAR_ACCELERATION)      # flag off: the reference reports an error as well
    gpu, plain = engine.Scene(sc, env_path=env_path, body_accelerations=True), engine.Scene(sc, env_path=env_path)
    nb = gpu.num_dynamic
    inv_dt = np.float32(1.0) / np.float32(gpu.dt)
    assert not gpu.getRigidDynamicData(engine.RD_LINEAR_ACCELERATION).any()                          # nothing stepped yet
    rng = np.random.RandomState(3)
    for t in range(6):
        if t == 3:                                                                                    # a velocity write is the velocity the next step starts from
            kick = rng.uniform(-0.5, 0.5, (nb, 3)).astype(np.float32)
            gpu.setRigidDynamicData(engine.RD_LINEAR_VELOCITY, kick); plain.setRigidDynamicData(engine.RD_LINEAR_VELOCITY, kick)
        l0, a0 = gpu.getRigidDynamicData(engine.RD_LINEAR_VELOCITY), gpu.getRigidDynamicData(engine.RD_ANGULAR_VELOCITY)
        gpu.step(); plain.step()
        l1, a1 = gpu.getRigidDynamicData(engine.RD_LINEAR_VELOCITY), gpu.getRigidDynamicData(engine.RD_ANGULAR_VELOCITY)
        assert np.array_equal(gpu.getRigidDynamicData(engine.RD_LINEAR_ACCELERATION), (l1 - l0) * inv_dt), f"step {t}"
        assert np.array_equal(gpu.getRigidDynamicData(engine.RD_ANGULAR_ACCELERATION), (a1 - a0) * inv_dt), f"step {t}"
        assert np.array_equal(gpu.getStates(), plain.getStates())                                    # the flag changes nothing else
    assert np.abs(gpu.getRigidDynamicData(engine.RD_LINEAR_ACCELERATION)[:, 2]).max() < 12.0         # resting stacks: far below free fall over one step
    idx = np.array([7, 2, 40], np.uint32)
    assert np.array_equal(gpu.getRigidDynamicData(engine.RD_LINEAR_ACCELERATION, idx), ((l1 - l0) * inv_dt)[idx])
    # start / finish events (CUevent arguments of PxDirectGPUAPI::getRigidDynamicData / setRigidDynamicData)
    other = torch.cuda.Stream()
    src = torch.zeros((nb, 3), dtype=torch.float32, device="cuda"); dst = torch.empty_like(src)
    start, finish, done = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    for e in (start, finish, done):
        e.record()                                                                                  # torch creates the cudaEvent_t lazily, at the first record
    with torch.cuda.stream(other):
        src.fill_(0.25); start.record(other)                                                         # producer on another stream: the set must wait for it
    gpu.setRigidDynamicDataDeviceEv(engine.RD_LINEAR_VELOCITY, src.data_ptr(), nb, start_event=start.cuda_event, finish_event=finish.cuda_event)
    gpu.getRigidDynamicDataDeviceEv(engine.RD_LINEAR_VELOCITY, dst.data_ptr(), nb, start_event=finish.cuda_event, finish_event=done.cuda_event)
    done.synchronize()
    assert bool((dst == 0.25).all())
    gpu.getRigidDynamicDataDeviceEv(engine.RD_ANGULAR_VELOCITY, dst.data_ptr(), nb)                  # no finish event: synchronous
    assert np.array_equal(dst.cpu().numpy(), gpu.getRigidDynamicData(engine.RD_ANGULAR_VELOCITY))


# ---- kinematic bodies (PxRigidBodyFlag::eKINEMATIC, PxRigidDynamic::setKinematicTarget) ----
@pytest.mark.gpu
@pytest.mark.parametrize("name,env_path", [("kinematic_mix", False), ("kinematic_envs_3", False), ("kinematic_envs_3", True), ("pgs_kinematic_mix", False), ("pgs_kinematic_mix", True)])
def test_kinematic_bodies_gpu_matches_oracle_and_reference(oracle, name, env_path):
    """Conveyor, lift, rotating paddle, a kinematic without target, kinematics crossing each other and the ground plane; targets before every step.  Teacher-forced from the
    reference's states: kinematic poses equal to the targets bit for bit, kinematic velocities 1e-6 relative (atan2f of the device library), the reference's created / deleted
    pairs and manifolds every step, dynamic bodies within pose 2e-5 / 2e-4 m/s / 2e-3 rad/s of the reference and within 1e-6 / 2e-5 of the oracle.  Free running for 40 steps:
    GPU within 1e-4 of the reference.  Both paths; on the environment path (no solver order can be given there) the canonical order differs from the reference's island order,
    so the dynamic bodies are compared with the oracle run in the same order, and with the device-wide path bit for bit."""
    z, sc = util.load_golden(name)
    kin = util.kinematic_indices(sc)
    gpu, cpu = engine.Scene(sc, env_path=env_path), oracle.OracleScene(sc)
    wide = engine.Scene(sc, env_path=False) if env_path else None
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t]); cpu.setStates(z["states"][t])
        i, p = util.golden_kin_targets(z, sc, t)
        if len(i):
            gpu.setKinematicTargets(i, p); cpu.setKinematicTargets(i, p)
        order = None if env_path else util.golden_order(z, t)
        gpu.setConstraintOrder(order); gpu.step(); cpu.step(order)
        assert gpu.uses_env_path == env_path
        st, ref, orc = gpu.getStates(), z["states"][t + 1], cpu.getStates()
        if env_path:
            wide.setStates(z["states"][t])
            if len(i):
                wide.setKinematicTargets(i, p)
            wide.step()
            assert np.array_equal(st, wide.getStates()), f"environment path == device-wide path, step {t}"

        assert np.array_equal(st[kin][:, :7], ref[kin][:, :7]), f"kinematic pose, step {t}"
        assert np.abs(st[kin][:, 7:] - ref[kin][:, 7:]).max() <= 1e-6 * max(1.0, np.abs(ref[kin][:, 7:]).max()), f"kinematic velocity, step {t}"
        assert np.array_equal(gpu.getCreatedPairs(), util.golden_created(z, t)) and np.array_equal(gpu.getDeletedPairs(), util.golden_deleted(z, t)), f"broadphase, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"
        if not env_path:    # (the environment path solves in canonical order: bodies squeezed between the paddle and the ground depend on the Gauss-Seidel order, 4e-3 in one step)
            assert np.abs(st[:, :7] - ref[:, :7]).max() < 2e-5 and np.abs(st[:, 7:10] - ref[:, 7:10]).max() < 2e-4 and np.abs(st[:, 10:] - ref[:, 10:]).max() < 2e-3, f"vs reference, step {t}"
        assert np.abs(st[:, :7] - orc[:, :7]).max() < 1e-6 and np.abs(st[:, 7:] - orc[:, 7:]).max() < 2e-5, f"vs oracle, step {t}"
    if env_path:
        return
    gpu = engine.Scene(sc)
    for t in range(40):
        i, p = util.golden_kin_targets(z, sc, t)
        if len(i):
            gpu.setKinematicTargets(i, p)
        gpu.setConstraintOrder(util.golden_order(z, t)); gpu.step()
        assert np.abs(gpu.getStates()[:, :7] - z["states"][t + 1][:, :7]).max() < (2e-3 if name.startswith("pgs") else 1e-4), f"free running, step {t}"


@pytest.mark.gpu
def test_kinematic_targets_api():
    """device-pointer variant == host variant (CUDA graph replay included: no constraint order given); errors: a non-kinematic body, an index out of range (reported by
    fetchResults for the device variant), scenes with sleeping, a kinematic flag on a static actor."""
    import torch
    z, sc = util.load_golden("kinematic_mix")
    a, b = engine.Scene(sc), engine.Scene(sc)
    for t in range(25):
        i, p = util.golden_kin_targets(z, sc, t)
        a.setKinematicTargets(i, p)
        di, dp = torch.from_numpy(i.astype(np.int32)).cuda(), torch.from_numpy(p).cuda()
        torch.cuda.synchronize()
        b.setKinematicTargetsDevice(di.data_ptr(), dp.data_ptr(), len(i))
        a.step(); b.step()
        assert np.array_equal(a.getStates(), b.getStates()), f"step {t}"
    kin = util.kinematic_indices(sc)
    assert np.abs(a.getStates()[kin[0], 0] - z["states"][25][kin[0], 0]) == 0 and a.getStates()[kin[0], 7] > 0.9       # the conveyor is where its targets put it, at 1 m/s
    a.step()                                                                                                              # no new target: it stands still
    assert not a.getStates()[kin, 7:].any()
    dyn = np.setdiff1d(np.arange(sc.n_dynamic), kin)
    with pytest.raises(engine.PhysxB200Error):
        a.setKinematicTargets(dyn[:1], p[:1])
    with pytest.raises(engine.PhysxB200Error):
        a.setKinematicTargets(np.array([10 ** 6], np.uint32), p[:1])
    bad = torch.tensor([int(dyn[0])], dtype=torch.int32).cuda(); torch.cuda.synchronize()
    a.setKinematicTargetsDevice(bad.data_ptr(), dp.data_ptr(), 1)
    with pytest.raises(engine.PhysxB200Error):
        a.step()
    a.step()                                                                                                              # reported once, the scene carries on
    with pytest.raises(engine.PhysxB200Error):
        engine.Scene(scenes.kinematic_mix(sleep_threshold=0.005))
    wrong = scenes.kinematic_mix(); wrong.actors["flags"][1] = scenes.ACTOR_KINEMATIC
    with pytest.raises(engine.PhysxB200Error):
        engine.Scene(wrong)


# ---- PxAggregate membership (members of an aggregate without self collisions never pair) ----
@pytest.mark.gpu
@pytest.mark.parametrize("name,env_path", [("aggregates_mix", False), ("aggregates_mix", True), ("aggregates_envs_3", False), ("aggregates_envs_3", True)])
def test_aggregates_gpu_matches_oracle_and_reference(oracle, name, env_path):
    """Teacher-forced from the reference's states: created / deleted pairs and manifolds of the reference every step on both paths (aggregates_mix has no environment
    ids and runs as one environment), states equal to the oracle's (1e-6 / 2e-5; device-wide path with the reference's solver order also within 2e-5 of the reference)."""
    z, sc = util.load_golden(name)
    gpu, cpu = engine.Scene(sc, env_path=env_path), oracle.OracleScene(sc)
    for t in range(z["states"].shape[0] - 1):
        gpu.setStates(z["states"][t]); cpu.setStates(z["states"][t])
        order = None if env_path else util.golden_order(z, t)
        gpu.setConstraintOrder(order); gpu.step(); cpu.step(order)
        assert gpu.uses_env_path == env_path
        st, ref, orc = gpu.getStates(), z["states"][t + 1], cpu.getStates()
        assert np.array_equal(gpu.getCreatedPairs(), util.golden_created(z, t)) and np.array_equal(gpu.getDeletedPairs(), util.golden_deleted(z, t)), f"broadphase, step {t}"
        assert util.contact_counts(gpu.getPairs(), gpu.getContacts()) == util.golden_contact_counts(z, t), f"manifolds, step {t}"
        assert np.abs(st[:, :7] - orc[:, :7]).max() < 1e-6 and np.abs(st[:, 7:] - orc[:, 7:]).max() < 2e-5, f"vs oracle, step {t}"
        if not env_path:
            assert np.abs(st[:, :7] - ref[:, :7]).max() < 2e-5, f"vs reference, step {t}"
    lib = gpu._lib
    bad = sc.actors[1:2].copy(); bad["aggregate"] = 0x40000000
    assert lib.pxb_scene_add_actors(gpu._h, bad.ctypes.data, 1) < 0


# ---- per-body pre-integration flags: eDISABLE_GRAVITY, eENABLE_GYROSCOPIC_FORCES ----
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["body_flags_mix", "pgs_body_flags_mix"])
@pytest.mark.parametrize("env_path", [True, False])
def test_body_flags_gpu_matches_oracle_and_reference(oracle, name, env_path):
    """28 steps of free flight bit-identical to the REFERENCE on both paths and both solvers (gyroscopic term, gravity switch); then, re-synchronised to the oracle every
    step, equal manifolds and states within TOL_STEP for the remaining steps (tumbling flat boxes: chaotic)."""
    z, sc = util.load_golden(name)
    gpu, cpu = engine.Scene(sc, env_path=env_path), oracle.OracleScene(sc)
    for t in range(z["states"].shape[0] - 1):
        order = None if env_path else util.golden_order(z, t)
        gpu.setConstraintOrder(order); gpu.step(); cpu.step(order)
        assert gpu.uses_env_path == env_path
        a, b = gpu.getStates(), cpu.getStates()
        if t < 28:
            assert np.array_equal(a, z["states"][t + 1]), f"free flight vs reference, step {t}"
        assert np.array_equal(gpu.getContacts()[:, 0], cpu.getContacts()[:, 0]), f"contact counts, step {t}"
        assert np.abs(a - b).max() < TOL_STEP, f"states, step {t}"
        cpu.setStates(a)


# ---- PxScene::setGravity between steps ----
@pytest.mark.gpu
@pytest.mark.parametrize("env_path", [True, False])
def test_set_gravity_between_steps(oracle, env_path):
    """The gravity vector is read by the next step's pre-integration (Sc::Scene::mGravity): GPU == oracle bit for bit through two changes (CUDA graphs are re-captured), and a
    free-falling body's acceleration getter shows the new vector."""
    sc = scenes.env_grid_stacks(n_envs=4, jitter=0.01)
    sc.actors["pos"][5, 1] += 3.0                                   # one box in free fall
    gpu, cpu = engine.Scene(sc, env_path=env_path, body_accelerations=True), oracle.OracleScene(sc)
    dyn5 = 4                                                        # actor 5 is dynamic body 4 (actor 0 is the ground plane)
    for t in range(12):
        if t == 4:
            gpu.setGravity([0.0, -3.0, 1.0]); cpu.setGravity([0.0, -3.0, 1.0])
        if t == 8:
            gpu.setGravity([0.0, -9.81, 0.0]); cpu.setGravity([0.0, -9.81, 0.0])
        gpu.step(); cpu.step()
        assert np.array_equal(gpu.getStates(), cpu.getStates()), f"step {t}"
        if t in (5, 6):
            acc = gpu.getRigidDynamicData(engine.RD_LINEAR_ACCELERATION)[dyn5]
            assert np.abs(acc - np.array([0.0, -3.0, 1.0], np.float32)).max() < 1e-4, acc
    gpu.simulate()
    with pytest.raises(engine.PhysxB200Error):
        gpu.setGravity([0, 0, 0])
    gpu.fetchResults(True)


# ---- PxRigidBody::setMass / setMassSpaceInertiaTensor between steps ----
@pytest.mark.gpu
@pytest.mark.parametrize("env_path", [True, False])
def test_set_mass_properties_between_steps(oracle, env_path):
    """Mass randomisation between steps: GPU == oracle bit for bit after the change (the heavier boxes sink differently into the stacks below), the tensor front end reads the new
    masses back, kinematic bodies and bad values are rejected."""
    sc = scenes.env_grid_stacks(n_envs=4, jitter=0.01)
    gpu, cpu = engine.Scene(sc, env_path=env_path), oracle.OracleScene(sc)
    rng = np.random.RandomState(5)
    idx = np.arange(0, sc.n_dynamic, 3, dtype=np.uint32)
    dyn = np.nonzero(sc.actors["flags"] & 1)[0]
    base = np.concatenate([sc.actors["mass"][dyn][idx, None], sc.actors["inertia"][dyn][idx]], axis=1)
    for t in range(16):
        if t in (3, 9):
            m = (base * rng.uniform(0.3, 4.0, (len(idx), 1))).astype(np.float32)
            gpu.setMassProperties(idx, m); cpu.setMassProperties(idx, m)
        gpu.step(); cpu.step()
        assert np.array_equal(gpu.getStates(), cpu.getStates()), f"step {t}"
    ref = engine.Scene(sc, env_path=env_path)
    for t in range(16):
        ref.step()
    assert np.abs(ref.getStates() - gpu.getStates()).max() > 1e-5                       # the change matters
    with pytest.raises(engine.PhysxB200Error):
        gpu.setMassProperties(idx[:1], np.array([[-1.0, 1, 1, 1]], np.float32))
    with pytest.raises(engine.PhysxB200Error):
        gpu.setMassProperties(np.array([10 ** 6], np.uint32), m[:1])
    z, ksc = util.load_golden("kinematic_mix")
    kin = util.kinematic_indices(ksc)
    with pytest.raises(engine.PhysxB200Error):
        engine.Scene(ksc).setMassProperties(kin[:1], m[:1])
