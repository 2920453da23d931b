"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference CPU SDK
(oracle/_ref/ref_harness, built from /root/reference by oracle/ref_build.mk) on small seeded scenes.
Run in the build container (needs oracle/_ref/):   python tests/golden/make_golden.py
Fixture = .npz with the scene bytes, per-step states, per-step solver constraint input order (island
manager order), per-step tight bounds + ABP created/deleted pairs, and per-step contact sets.
Reference version: 5.6.1.51c1f783."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from physx_b200 import scenes  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cooking import cook_hulls  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
LITE_FULL_STEPS = 4


def run_reference(scene, steps, threads=1, forces=None, lite=False, kin_targets=None):
    """lite: long-horizon fixtures keep the states, the island-manager order and the per-pair contact COUNTS of every step, but
    bounds / ABP events / contact points only for the first `LITE_FULL_STEPS` steps (fixture size)."""
    with tempfile.TemporaryDirectory() as d:
        sp = os.path.join(d, "s.bin")
        scene.save(sp)
        extra = []
        if forces is not None:   # (blocks, n_dyn, 6) force xyz + torque xyz: block t % blocks is applied before step t
            np.ascontiguousarray(forces, dtype="<f4").tofile(d + "/forces")
            extra = ["--forces", d + "/forces"]
        if kin_targets is not None:   # (steps, n_kinematic, 7) PxTransform rows for setKinematicTarget before each step (NaN row = no target)
            np.ascontiguousarray(kin_targets, dtype="<f4").tofile(d + "/kin")
            extra += ["--kin-targets", d + "/kin"]
        subprocess.run([HARNESS, "run", sp, "--steps", str(steps), "--threads", str(threads), "--states", d + "/st", "--order", d + "/ord",
                        "--bp", d + "/bp", "--contacts", d + "/con", "--sleep", d + "/sl"] + extra, check=True, capture_output=True)
        sl = np.fromfile(d + "/sl", dtype=[("wc", "<f4"), ("s", "<u4")]).reshape(steps, scene.n_dynamic)
        nd, na = scene.n_dynamic, len(scene.actors)
        states = np.fromfile(d + "/st", "<f4").reshape(steps + 1, nd, 13)
        ob = open(d + "/ord", "rb").read(); off = 0; order_flat = []; order_off = [0]
        for _ in range(steps):
            n, = struct.unpack_from("<I", ob, off); off += 4
            order_flat.append(np.frombuffer(ob, "<u4", n * 2, off).reshape(n, 2)); off += n * 8
            order_off.append(order_off[-1] + n)
        bb = open(d + "/bp", "rb").read(); off = 0; bounds = []; cr = []; cro = [0]; de = []; deo = [0]
        for _ in range(steps):
            n, = struct.unpack_from("<I", bb, off); off += 4
            bounds.append(np.frombuffer(bb, "<f4", n * 6, off).reshape(n, 6)); off += n * 24
            nc, ndl = struct.unpack_from("<II", bb, off); off += 8
            cr.append(np.frombuffer(bb, "<u4", nc * 2, off).reshape(nc, 2)); off += nc * 8
            de.append(np.frombuffer(bb, "<u4", ndl * 2, off).reshape(ndl, 2)); off += ndl * 8
            cro.append(cro[-1] + nc); deo.append(deo[-1] + ndl)
        cb = open(d + "/con", "rb").read(); off = 0; con_pairs = []; con_off = [0]; con_pts = []; pt_off = [0]
        for _ in range(steps):
            n, = struct.unpack_from("<I", cb, off); off += 4
            for _p in range(n):
                a0, a1, k = struct.unpack_from("<III", cb, off); off += 12
                con_pairs.append((a0, a1, k))
                con_pts.append(np.frombuffer(cb, "<f4", k * 10, off).reshape(k, 10)[:, :7]); off += k * 40
                pt_off.append(pt_off[-1] + k)
            con_off.append(con_off[-1] + n)
        if lite:
            k = LITE_FULL_STEPS
            bounds, cr, de = bounds[:k], cr[:k], de[:k]; cro, deo = cro[:k + 1], deo[:k + 1]
            con_pts = con_pts[:con_off[k]]; pt_off = pt_off[:con_off[k] + 1]
        return dict(kin_targets=(np.zeros((0, 0, 7), np.float32) if kin_targets is None else np.asarray(kin_targets, np.float32)), scene=np.frombuffer(scene.tobytes(), np.uint8), forces=(np.zeros((0, nd, 6), np.float32) if forces is None else np.asarray(forces, np.float32)), states=states, wake=sl["wc"].copy(), asleep=sl["s"].copy(),
                    order=np.concatenate(order_flat) if order_flat else np.zeros((0, 2), np.uint32), order_off=np.array(order_off),
                    bounds=np.stack(bounds), created=np.concatenate(cr), created_off=np.array(cro), deleted=np.concatenate(de) if de else np.zeros((0, 2), np.uint32),
                    deleted_off=np.array(deo), con_pairs=np.array(con_pairs, np.uint32).reshape(-1, 3), con_off=np.array(con_off),
                    con_pts=np.concatenate(con_pts).astype(np.float32) if con_pts else np.zeros((0, 7), np.float32), pt_off=np.array(pt_off))


def scene_statistics(scene, data, steps, keep_every=30):
    """Reduction of a reference run to the statistics SURVEY 8d asks for on configs 3 / 4 (tests/util.py: run_statistics computes the same
    quantities from an engine / oracle run)."""
    dyn = (scene.actors["flags"] & scenes.ACTOR_DYNAMIC) != 0
    mass = scene.actors["mass"][dyn].astype(np.float64); inertia = scene.actors["inertia"][dyn].astype(np.float64)
    st = data["states"].astype(np.float64)
    ke, mean_y, min_sep, mean_pen, n_touch, n_pts = [], [], [], [], [], []
    for t in range(steps):
        s = st[t + 1]
        # angular part in the body frame: w_body = q^-1 w q
        q = s[:, 3:7]; w = s[:, 10:13]
        qv, qw = q[:, :3], q[:, 3:4]
        wb = w + 2.0 * np.cross(-qv, np.cross(-qv, w) + qw * w)
        ke.append(float(0.5 * (mass * (s[:, 7:10] ** 2).sum(1)).sum() + 0.5 * (inertia * wb ** 2).sum()))
        mean_y.append(float(s[:, 1].mean()))
        a, b = data["con_off"][t], data["con_off"][t + 1]
        seps = data["con_pts"][data["pt_off"][a]:data["pt_off"][b], 6].astype(np.float64)
        n_touch.append(int(np.count_nonzero(data["con_pairs"][a:b, 2]))); n_pts.append(int(len(seps)))
        min_sep.append(float(seps.min()) if len(seps) else 0.0)
        mean_pen.append(float(np.clip(-seps, 0, None).mean()) if len(seps) else 0.0)
    keep = list(range(0, steps + 1, keep_every))
    return dict(scene=data["scene"], steps=np.int64(steps), ke=np.array(ke), mean_y=np.array(mean_y), min_sep=np.array(min_sep), mean_pen=np.array(mean_pen),
                n_touch=np.array(n_touch), n_pts=np.array(n_pts), kept_steps=np.array(keep), states=data["states"][keep].astype(np.float32))


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    cases = {
        # BASELINE config 1 at reduced size: 4 stacks x 8 boxes, jittered (every island has its own edge order)
        "stacks_4x8_jitter": (scenes.box_stacks(n_stacks=4, height=8, half_extent=0.25, spacing=1.0, jitter=0.01), 120),
        # SnippetHelloWorld-style exact stacking (zero gap, zero jitter), unit boxes
        "stacks_3x5_exact": (scenes.box_stacks(n_stacks=3, height=5, half_extent=0.5, spacing=4.0, jitter=0.0), 60),
        # boxes dropped from a height with initial spin: pairs are created and lost, manifolds rebuilt
        "tumble_12": (scenes.tumbling_boxes(n=12, seed=7), 150),
        # BASELINE config 2 shape at 4 envs
        "envs_4": (scenes.env_grid_stacks(n_envs=4, jitter=0.01), 30),
        # PGS solver (PxSolverType::ePGS) on the same shapes
        "pgs_stacks_3x6_jitter": (scenes.box_stacks(n_stacks=3, height=6, half_extent=0.25, spacing=1.0, jitter=0.01, solver=scenes.SOLVER_PGS), 100),
        "pgs_envs_4": (scenes.env_grid_stacks(n_envs=4, stacks_per_env=4, height=4, jitter=0.01, solver=scenes.SOLVER_PGS), 40),
        "pgs_spheres_capsules_12": (scenes.mixed_primitives(n=12, seed=3, kinds=("sphere", "capsule"), solver=scenes.SOLVER_PGS), 100),
    }
    # sleeping enabled (PxRigidDynamic::setSleepThreshold 0.005): resting stacks fall asleep; a box dropped on a sleeping stack wakes it
    cases["sleep_stacks_2x3"] = (scenes.box_stacks(n_stacks=2, height=3, half_extent=0.25, spacing=1.0, jitter=0.01, sleep_threshold=0.005), 60)
    drop = scenes.box_stacks(n_stacks=1, height=3, half_extent=0.25, spacing=1.0, jitter=0.0, sleep_threshold=0.005)
    drop.actors["pos"][3, 1] = 4.0
    cases["sleep_drop"] = (drop, 80)
    # PxRigidDynamicLockFlags (a18): boxes that may only move vertically / may not rotate inside jittered stacks (TGS and PGS),
    # spheres / capsules with assorted locks dropped in a column (teacher-forced comparison: chaotic)
    cases["lock_stacks"] = (scenes.locked_stacks(), 80)
    cases["pgs_lock_stacks"] = (scenes.locked_stacks(solver=scenes.SOLVER_PGS), 80)
    cases["lock_primitives"] = (scenes.locked_primitives(seed=4), 100)
    cases["pgs_lock_primitives"] = (scenes.locked_primitives(seed=3, solver=scenes.SOLVER_PGS), 100)
    # a10 (GJK family): capsules and spheres dropped onto static tilted / dynamic resting boxes -- capsule-box face, edge and corner contacts
    cases["capsules_on_boxes"] = (scenes.capsules_on_boxes(seed=3), 150)
    cases["hulls_on_plane"] = (scenes.hulls_on_plane(seed=6, cook=cook_hulls), 150)   # convex hulls (cooked by the reference): tight bounds + pcmContactPlaneConvex
    cases["hulls_and_spheres"] = (scenes.hulls_and_spheres(seed=11, cook=cook_hulls), 150)   # pcmContactSphereConvex: GJK with the hull support mapping
    cases["spheres_into_hulls"] = (scenes.spheres_into_hulls(seed=13, cook=cook_hulls), 60)  # ... through EPA
    cases["hulls_and_capsules"] = (scenes.hulls_and_capsules(seed=21, cook=cook_hulls), 150)            # pcmContactCapsuleConvex
    cases["capsules_into_hulls"] = (scenes.hulls_and_capsules(seed=23, speed=14.0, cook=cook_hulls), 60)   # ... through EPA (a seed whose hulls never come near each other)
    cases["hull_pile"] = (scenes.hull_pile(seed=32, cook=cook_hulls), 150)                                   # pcmContactConvexConvex: GJK / EPA + polygon clipping
    cases["box_hull_pile"] = (scenes.hull_pile(seed=33, kinds=("convex", "box"), cook=cook_hulls), 150)     # pcmContactBoxConvex
    # hulls of more than 32 vertices: the support mapping is hill climbing over Gu::BigConvexRawData (cube-map start sample + vertex adjacency)
    cases["big_hull_pile"] = (scenes.hull_pile(n=12, seed=35, kinds=("convex", "convex", "box", "convex", "sphere", "capsule"), hull_points=(40, 64), cook=cook_hulls), 150)
    # BASELINE config 3 at small size (spheres / capsules / library hulls falling into a walled bin): every hull pair type in one scene
    cases["config3_small"] = (scenes.falling_primitives(4, 3, 4, kinds=("sphere", "capsule", "convex")), 110)
    # a11: material table (every PxCombineMode, eDISABLE_FRICTION): sliding / stacked boxes and bouncing spheres, TGS and PGS
    cases["materials_mix"] = (scenes.material_mix(), 90)
    cases["pgs_materials_mix"] = (scenes.material_mix(solver=scenes.SOLVER_PGS), 90)
    # a1 with local poses: PxShape::setLocalPose + PxRigidBody::setCMassLocalPose per body (offset / rotated shape and centre-of-mass frames), TGS and PGS
    cases["local_poses_mix"] = (scenes.local_pose_mix(), 90)
    cases["pgs_local_poses_mix"] = (scenes.local_pose_mix(solver=scenes.SOLVER_PGS), 90)
    # f1: the default simulation filter shader (collision groups + groups masks): suppressed pairs stay broadphase pairs and generate no contacts
    cases["filter_groups_mix"] = (scenes.filter_groups_mix(), 90)
    cases["shape_offsets_mix"] = (scenes.shape_offsets_mix(), 120)   # PxShape::setContactOffset / setRestOffset per shape (TGS)
    cases["pgs_shape_offsets_mix"] = (scenes.shape_offsets_mix(solver=scenes.SOLVER_PGS), 120)
    # tumbling boxes next to spheres / capsules: energetic multi-body impacts with friction (teacher-forced comparison; pins the reference's reciprocal table in the edge clipping)
    cases["tumble_mixed_14"] = (scenes.mixed_primitives(n=14, seed=24, kinds=("box", "sphere", "capsule", "box")), 120)
    cases["capsules_into_boxes"] = (scenes.capsules_into_boxes(seed=3), 60)   # deep penetration: the EPA query
    # a19: PxDirectGPUAPI eFORCE / eTORQUE writes (= addForce / addTorque(eFORCE) before every step), a 7-step cycle of per-body forces
    forced = {"forces_stacks": scenes.box_stacks(n_stacks=3, height=4, half_extent=0.25, spacing=1.0, jitter=0.01),
              "pgs_forces_stacks": scenes.box_stacks(n_stacks=3, height=4, half_extent=0.25, spacing=1.0, jitter=0.01, solver=scenes.SOLVER_PGS),
              "forces_primitives": scenes.mixed_primitives(n=12, seed=3, kinds=("sphere", "capsule"))}
    # Long horizons (120 steps on configs 1 / 2 / 5, 300 on config 1 proper) and the dense pile of config 4 in the reference's own
    # first-fit order -- "lite" fixtures (states + solver order + contact counts every step)
    lite = {
        "stacks_10x10": (scenes.box_stacks(), 300),                                                   # BASELINE config 1 proper: 10 x 10 unit boxes
        "envs_4_long": (scenes.env_grid_stacks(n_envs=4, jitter=0.01), 120),                          # config 2 shape, 120 steps
        "envs_2x128": (scenes.env_grid_stacks(n_envs=2, stacks_per_env=16, jitter=0.01), 120),        # config 5 shape: 128 boxes per environment
        "pile_6x4x6": (scenes.box_pile(6, 4, 6), 60),                                                 # config 4 shape: one dense island in a walled bin
        "pgs_pile_6x4x6": (scenes.box_pile(6, 4, 6, solver=scenes.SOLVER_PGS), 60),
    }
    # statistics fixtures ("energy / penetration statistics only for configs 3-4"): per-step kinetic energy, mean height, deepest and
    # mean penetration, touching pairs and contact points of the reference on a pile large enough to be chaotic in detail; states kept every 30 steps
    stats = {"pile_12x6x12_stats": (scenes.box_pile(12, 6, 12), 120),
             "fall_6x5x6_stats": (scenes.falling_primitives(6, 5, 6, kinds=("sphere", "capsule", "convex")), 150)}
    only = sys.argv[1:]
    for name, (sc, steps) in lite.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        data = run_reference(sc, steps, lite=True)
        np.savez_compressed(os.path.join(out, name + ".npz"), **data)
        print(name, "bodies", sc.n_dynamic, "steps", steps, "bytes", os.path.getsize(os.path.join(out, name + ".npz")))
    for name, (sc, steps) in stats.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        data = run_reference(sc, steps)
        np.savez_compressed(os.path.join(out, name + ".npz"), **scene_statistics(sc, data, steps))
        print(name, "bodies", sc.n_dynamic, "steps", steps, "bytes", os.path.getsize(os.path.join(out, name + ".npz")))
    if only:
        cases = {k: v for k, v in cases.items() if any(k.startswith(o) for o in only)}
    for name, sc in forced.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        data = run_reference(sc, 80, forces=scenes.test_forces(sc.n_dynamic))
        np.savez_compressed(os.path.join(out, name + ".npz"), **data)
        print(name, "bodies", sc.n_dynamic, "steps", 80, "bytes", os.path.getsize(os.path.join(out, name + ".npz")))
    # kinematic bodies (PxRigidBodyFlag::eKINEMATIC + setKinematicTarget every step): conveyor, lift, rotating paddle, a kinematic without target, filtered kinematic pairs;
    # device-wide and with environment ids
    kinem = {"kinematic_mix": (scenes.kinematic_mix(), 120), "kinematic_envs_3": (scenes.kinematic_mix(n_envs=3), 90), "pgs_kinematic_mix": (scenes.kinematic_mix(solver=scenes.SOLVER_PGS), 120)}
    for name, (sc, steps) in kinem.items():
        if only and not any(name.startswith(o) for o in only):
            continue
        data = run_reference(sc, steps, kin_targets=scenes.kinematic_targets(sc, steps))
        np.savez_compressed(os.path.join(out, name + ".npz"), **data)
        print(name, "bodies", sc.n_dynamic, "steps", steps, "bytes", os.path.getsize(os.path.join(out, name + ".npz")))
    # PxAggregate membership with and without self collisions (the harness creates real PxAggregates; its standalone broadphase gives the members of an aggregate
    # without self collisions one filter group)
    # per-body pre-integration flags: PxActorFlag::eDISABLE_GRAVITY, PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES (TGS and PGS)
    cases["body_flags_mix"] = (scenes.body_flags_mix(), 100)
    cases["pgs_body_flags_mix"] = (scenes.body_flags_mix(solver=scenes.SOLVER_PGS), 100)
    cases["aggregates_mix"] = (scenes.aggregates_mix(), 100)
    cases["aggregates_envs_3"] = (scenes.aggregates_mix(n_envs=3), 80)
    if only:
        cases = {k: v for k, v in cases.items() if any(k.startswith(o) for o in only)}
    for name, (sc, steps) in cases.items():
        data = run_reference(sc, steps)
        np.savez_compressed(os.path.join(out, name + ".npz"), **data)
        print(name, "bodies", sc.n_dynamic, "steps", steps, "bytes", os.path.getsize(os.path.join(out, name + ".npz")))


if __name__ == "__main__":
    main()
