"""Host-side mirror of the reference scene interface for the rigid-body step hot path.

`Scene` wraps the C ABI of libphysx_b200.so (include/physx_b200.h) with the names a PhysX user knows:
`simulate(dt)` / `fetchResults(block)` (PxScene, physx/include/PxScene.h), `getRigidDynamicData` /
`setRigidDynamicData` (PxDirectGPUAPI, physx/include/PxDirectGPUAPI.h:311-463).  There is no CPU path:
importing works anywhere (so the symbol checks run on a CPU box) but creating a Scene without a CUDA
device raises `PhysxB200Error`.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import scenes as _scenes

_LIB_NAME = os.environ.get("PXB_LIB", "libphysx_b200.so")   # PXB_LIB: tooling hook to load an experimental build
_lib = None

# every symbol include/physx_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "pxb_scene_create", "pxb_scene_release", "pxb_last_error", "pxb_device_count", "pxb_scene_add_actors", "pxb_scene_set_convex_meshes",
    "pxb_scene_num_actors", "pxb_scene_num_dynamic", "pxb_scene_simulate", "pxb_scene_fetch_results",
    "pxb_scene_set_constraint_order", "pxb_get_rigid_dynamic_data", "pxb_set_rigid_dynamic_data",
    "pxb_get_rigid_dynamic_data_device", "pxb_set_rigid_dynamic_data_device", "pxb_scene_get_states",
    "pxb_scene_set_states", "pxb_scene_state_device_ptr", "pxb_scene_stream", "pxb_scene_compute_bounds",
    "pxb_scene_get_bounds", "pxb_scene_broadphase", "pxb_scene_num_pairs", "pxb_scene_get_pairs",
    "pxb_scene_num_created", "pxb_scene_num_deleted", "pxb_scene_get_created", "pxb_scene_get_deleted",
    "pxb_scene_get_contacts", "pxb_scene_last_num_partitions", "pxb_scene_last_num_constraints",
    "pxb_scene_last_num_launches", "pxb_scene_set_profiling", "pxb_scene_get_stage_times",
    "pxb_scene_get_states_device", "pxb_scene_uses_env_path", "pxb_scene_get_sleep_data", "pxb_get_rigid_dynamic_data_async", "pxb_set_rigid_dynamic_data_async", "pxb_scene_sync", "pxb_scatter_to_peers",
    "pxb_scene_set_state_export", "pxb_peer_signal", "pxb_peer_wait", "pxb_bp_create", "pxb_bp_release", "pxb_bp_update", "pxb_bp_fetch",
    "pxb_scene_set_materials", "pxb_scene_remove_actors", "pxb_tensor_read_device", "pxb_tensor_write_device", "pxb_scene_num_touch_found", "pxb_scene_num_touch_lost", "pxb_scene_get_touch_found", "pxb_scene_get_touch_lost", "pxb_scene_enable_contact_data", "pxb_scene_copy_contact_data", "pxb_scene_set_local_poses", "pxb_scene_set_filter_shader", "pxb_scene_set_filter_data", "pxb_scene_set_shape_offsets",
    "pxb_get_rigid_dynamic_data_device_ev", "pxb_set_rigid_dynamic_data_device_ev", "pxb_scene_set_kinematic_targets", "pxb_scene_set_kinematic_targets_device", "pxb_scene_set_gravity", "pxb_scene_set_mass_properties",
]

RD_GLOBAL_POSE, RD_LINEAR_VELOCITY, RD_ANGULAR_VELOCITY, RD_FORCE, RD_TORQUE = 0, 1, 2, 3, 4   # PxRigidDynamicGPUAPIRead/WriteType
RD_LINEAR_ACCELERATION, RD_ANGULAR_ACCELERATION = 5, 6   # read only; scenes created with body_accelerations=True (PxSceneFlag::eENABLE_BODY_ACCELERATIONS)


class PhysxB200Error(RuntimeError):
    pass


class SceneDesc(ctypes.Structure):
    _fields_ = [
        ("gravity", ctypes.c_float * 3), ("solverType", ctypes.c_uint32),
        ("bounceThresholdVelocity", ctypes.c_float), ("frictionOffsetThreshold", ctypes.c_float),
        ("frictionCorrelationDistance", ctypes.c_float), ("toleranceLength", ctypes.c_float),
        ("staticFriction", ctypes.c_float), ("dynamicFriction", ctypes.c_float), ("restitution", ctypes.c_float),
        ("contactOffset", ctypes.c_float), ("restOffset", ctypes.c_float),
        ("posIters", ctypes.c_uint32), ("velIters", ctypes.c_uint32),
        ("maxActors", ctypes.c_uint32), ("maxPairs", ctypes.c_uint32), ("device", ctypes.c_int32),
        ("reserved", ctypes.c_uint32 * 8),
    ]


def lib_path():
    # PXB_LIB: another build of the same library (A/B experiments with compile-time parameters; tools/gpu_run30.sh)
    return os.environ.get("PXB_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def load_library():
    """Loads libphysx_b200.so; raises loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PhysxB200Error(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(physx_b200 has no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, u32, i32, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int, ctypes.c_float
    lib.pxb_scene_create.argtypes = [ctypes.POINTER(SceneDesc), ctypes.POINTER(vp)]
    lib.pxb_scene_set_convex_meshes.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, u32]
    lib.pxb_scene_release.argtypes = [vp]
    lib.pxb_scene_release.restype = None
    lib.pxb_last_error.restype = ctypes.c_char_p
    lib.pxb_scene_add_actors.argtypes = [vp, vp, u32]
    lib.pxb_scene_set_materials.argtypes = [vp, vp, u32]
    lib.pxb_scene_remove_actors.argtypes = [vp, vp, u32]
    lib.pxb_tensor_read_device.argtypes = [vp, i32, vp, vp, u32]
    lib.pxb_tensor_write_device.argtypes = [vp, i32, vp, vp, u32]
    for f in ("pxb_scene_num_actors", "pxb_scene_num_dynamic", "pxb_scene_num_pairs", "pxb_scene_num_created",
              "pxb_scene_num_deleted", "pxb_scene_last_num_partitions", "pxb_scene_last_num_constraints",
              "pxb_scene_last_num_launches", "pxb_scene_num_touch_found", "pxb_scene_num_touch_lost"):
        getattr(lib, f).argtypes = [vp]
        getattr(lib, f).restype = u32
    lib.pxb_scene_simulate.argtypes = [vp, f32]
    lib.pxb_scene_fetch_results.argtypes = [vp, i32]
    lib.pxb_scene_set_constraint_order.argtypes = [vp, vp, u32]
    for f in ("pxb_get_rigid_dynamic_data", "pxb_set_rigid_dynamic_data", "pxb_get_rigid_dynamic_data_device",
              "pxb_set_rigid_dynamic_data_device"):
        getattr(lib, f).argtypes = [vp, vp, vp, i32, u32]
    lib.pxb_scene_set_kinematic_targets.argtypes = [vp, vp, vp, u32]
    lib.pxb_scene_set_gravity.argtypes = [vp, vp]
    lib.pxb_scene_set_mass_properties.argtypes = [vp, vp, vp, u32]
    lib.pxb_scene_set_kinematic_targets_device.argtypes = [vp, vp, vp, u32]
    for f in ("pxb_get_rigid_dynamic_data_device_ev", "pxb_set_rigid_dynamic_data_device_ev"):
        getattr(lib, f).argtypes = [vp, vp, vp, i32, u32, vp, vp]
    for f in ("pxb_scene_get_states", "pxb_scene_set_states", "pxb_scene_get_bounds", "pxb_scene_broadphase",
              "pxb_scene_get_pairs", "pxb_scene_get_created", "pxb_scene_get_deleted", "pxb_scene_get_contacts", "pxb_scene_get_touch_found", "pxb_scene_get_touch_lost"):
        getattr(lib, f).argtypes = [vp, vp]
    lib.pxb_scene_compute_bounds.argtypes = [vp]
    lib.pxb_scene_enable_contact_data.argtypes = [vp, i32]
    lib.pxb_scene_set_local_poses.argtypes = [vp, u32, u32, vp, vp]
    lib.pxb_scene_set_filter_shader.argtypes = [vp, vp]
    lib.pxb_scene_set_filter_data.argtypes = [vp, u32, u32, vp]
    lib.pxb_scene_set_shape_offsets.argtypes = [vp, u32, u32, vp]
    lib.pxb_scene_copy_contact_data.argtypes = [vp, vp, vp, u32]
    lib.pxb_scene_get_states_device.argtypes = [vp, vp]
    lib.pxb_scene_uses_env_path.argtypes = [vp]
    lib.pxb_scene_get_sleep_data.argtypes = [vp, vp, vp]
    lib.pxb_get_rigid_dynamic_data_async.argtypes = [vp, vp, i32, u32]
    lib.pxb_set_rigid_dynamic_data_async.argtypes = [vp, vp, i32, u32]
    lib.pxb_scene_sync.argtypes = [vp]
    lib.pxb_scatter_to_peers.argtypes = [vp, vp, vp, ctypes.c_size_t, vp, u32, u32]
    lib.pxb_scene_set_state_export.argtypes = [vp, vp, u32, u32]
    lib.pxb_peer_signal.argtypes = [vp, vp, vp, u32, u32]
    lib.pxb_peer_wait.argtypes = [vp, vp, vp, u32, u32]
    lib.pxb_bp_create.argtypes = [u32, u32, i32, ctypes.POINTER(vp)]
    lib.pxb_bp_release.argtypes = [vp]
    lib.pxb_bp_release.restype = None
    lib.pxb_bp_update.argtypes = [vp, vp, vp, vp, vp, u32, vp, vp, u32, vp, u32, vp, u32]
    lib.pxb_bp_fetch.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(u32), ctypes.POINTER(vp), ctypes.POINTER(u32)]
    lib.pxb_scene_set_profiling.argtypes = [vp, i32]
    lib.pxb_scene_get_stage_times.argtypes = [vp, vp]
    lib.pxb_scene_state_device_ptr.argtypes = [vp, i32]
    lib.pxb_scene_state_device_ptr.restype = vp
    lib.pxb_scene_stream.argtypes = [vp]
    lib.pxb_scene_stream.restype = vp
    _lib = lib
    return lib


def _check(lib, rc):
    if rc < 0:
        raise PhysxB200Error(f"physx_b200 error {rc}: {lib.pxb_last_error().decode()}")
    return rc


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Scene:
    """One simulation scene resident on one GPU (one PxScene <-> one device, ScScene.cpp:718)."""

    def __init__(self, scene: _scenes.Scene, device: int = 0, max_pairs: int = 0, max_actors: int = 0, env_path: bool = True, env_row_cap: int = 0, env_threads: int = 0,
                 body_accelerations: bool = False):
        lib = load_library()
        self._lib = lib
        h = scene.header
        d = SceneDesc()
        d.gravity[:] = [float(x) for x in h["gravity"]]
        d.solverType = int(h["solverType"])
        d.bounceThresholdVelocity = float(h["bounceThreshold"])
        d.frictionOffsetThreshold = float(h["frictionOffsetThreshold"])
        d.frictionCorrelationDistance = float(h["frictionCorrelationDistance"])
        d.toleranceLength = float(h["toleranceLength"])
        d.staticFriction, d.dynamicFriction, d.restitution = float(h["staticFriction"]), float(h["dynamicFriction"]), float(h["restitution"])
        d.contactOffset, d.restOffset = float(h["contactOffset"]), float(h["restOffset"])
        d.posIters, d.velIters = int(h["posIters"]), int(h["velIters"])
        d.maxActors = max(int(max_actors), len(scene.actors))
        d.maxPairs = int(max_pairs)
        d.device = int(device)
        relaxed = bool(int(h["reserved"][0]) & 1)   # scene header reserved[0] bit 0: PXB_FLAG_RELAXED_PARTITIONING (shared with the oracle)
        d.reserved[1] = (0 if env_path else 1) | (2 if relaxed else 0) | (4 if body_accelerations else 0)   # PXB_FLAG_NO_ENV_PATH | PXB_FLAG_RELAXED_PARTITIONING | PXB_FLAG_BODY_ACCELERATIONS
        d.reserved[2] = int(env_row_cap)
        d.reserved[4] = int(np.float32(h["sleepThreshold"]).view(np.uint32))   # sleep threshold as float bits (0 = sleeping off)
        d.reserved[3] = int(env_threads)
        self.dt = float(h["dt"])
        self.device_index = int(device)
        self._h = ctypes.c_void_p()
        _check(lib, lib.pxb_scene_create(ctypes.byref(d), ctypes.byref(self._h)))
        if scene.cooked:   # cooked convex hulls (reference cooking output carried by the scene) go in before the actors that use them
            _check(lib, lib.pxb_scene_set_convex_meshes(self._h, scene.cooked, len(scene.cooked), len(scene.hulls)))
        mats = getattr(scene, "materials", None)
        if mats is not None and len(mats):   # material table (PxMaterial per shape): before the actors that refer to it
            mm = np.ascontiguousarray(mats)
            _check(lib, lib.pxb_scene_set_materials(self._h, _ptr(mm), len(mm)))
        recs = np.ascontiguousarray(scene.actors)
        _check(lib, lib.pxb_scene_add_actors(self._h, _ptr(recs), len(recs)))
        self.num_actors = int(lib.pxb_scene_num_actors(self._h))
        self.num_dynamic = int(lib.pxb_scene_num_dynamic(self._h))
        fc = getattr(scene, "filter_config", None)
        if fc is not None:   # PxDefaultSimulationFilterShader state + PxFilterData per actor
            cfg = np.ascontiguousarray(fc).reshape(1)
            _check(lib, lib.pxb_scene_set_filter_shader(self._h, _ptr(cfg)))
            fd = np.ascontiguousarray(scene.filter_data, dtype=np.uint32)
            _check(lib, lib.pxb_scene_set_filter_data(self._h, 0, len(fd), _ptr(fd)))
        so = getattr(scene, "shape_offsets", None)
        if so is not None:   # PxShape::setContactOffset / setRestOffset per shape
            so = np.ascontiguousarray(so, dtype=np.float32)
            _check(lib, lib.pxb_scene_set_shape_offsets(self._h, 0, len(so), _ptr(so)))
        lp = getattr(scene, "local_poses", None)
        if lp is not None:   # PxShape::setLocalPose / PxRigidBody::setCMassLocalPose per actor
            self.setLocalPoses(0, np.concatenate([lp["shapeP"], lp["shapeQ"]], axis=1), np.concatenate([lp["bodyP"], lp["bodyQ"]], axis=1))

    def setLocalPoses(self, first_actor, shape2actor, body2actor):
        """(n, 7) arrays (p.xyz, q.xyzw) for actors first_actor .. first_actor + n - 1; actor poses are unchanged, body frames move with body2actor"""
        sa = np.ascontiguousarray(shape2actor, dtype=np.float32); ba = np.ascontiguousarray(body2actor, dtype=np.float32)
        assert sa.shape == ba.shape and sa.shape[1] == 7
        _check(self._lib, self._lib.pxb_scene_set_local_poses(self._h, int(first_actor), len(sa), _ptr(sa), _ptr(ba)))

    def release(self):
        if getattr(self, "_h", None):
            self._lib.pxb_scene_release(self._h)
            self._h = None

    __del__ = release

    def removeActors(self, actor_indices):
        """PxScene::removeActor for the listed actor indices: they leave the simulation at the next step; indices stay valid."""
        idx = np.ascontiguousarray(actor_indices, dtype=np.uint32)
        _check(self._lib, self._lib.pxb_scene_remove_actors(self._h, _ptr(idx), len(idx)))

    # ---- PxScene ----
    def simulate(self, dt: float | None = None):
        _check(self._lib, self._lib.pxb_scene_simulate(self._h, float(self.dt if dt is None else dt)))

    def fetchResults(self, block: bool = True):
        return _check(self._lib, self._lib.pxb_scene_fetch_results(self._h, 1 if block else 0)) == 0

    def step(self, dt: float | None = None):
        self.simulate(dt)
        self.fetchResults(True)

    def setConstraintOrder(self, pairs):
        """pairs: (n,2) actor indices in solver input order (island manager order); None/empty = canonical."""
        if pairs is None or len(pairs) == 0:
            _check(self._lib, self._lib.pxb_scene_set_constraint_order(self._h, None, 0))
            return
        p = np.ascontiguousarray(pairs, dtype=np.uint32)
        _check(self._lib, self._lib.pxb_scene_set_constraint_order(self._h, _ptr(p), len(p)))

    # ---- PxDirectGPUAPI ----
    def getRigidDynamicData(self, data_type: int, indices=None, nb: int | None = None):
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.uint32)
        n = self.num_dynamic if (idx is None and nb is None) else (len(idx) if idx is not None else int(nb))
        out = np.zeros((n, 7 if data_type == RD_GLOBAL_POSE else 3), np.float32)
        _check(self._lib, self._lib.pxb_get_rigid_dynamic_data(self._h, _ptr(out), _ptr(idx), data_type, n))
        return out

    def setRigidDynamicData(self, data_type: int, data, indices=None):
        d = np.ascontiguousarray(data, dtype=np.float32)
        idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.uint32)
        _check(self._lib, self._lib.pxb_set_rigid_dynamic_data(self._h, _ptr(d), _ptr(idx), data_type, len(d)))

    def setMassProperties(self, indices, mass_inertia):
        """PxRigidBody::setMass / setMassSpaceInertiaTensor for the dynamic bodies `indices`: (n, 4) rows (mass, inertia xyz)."""
        i = np.ascontiguousarray(indices, dtype=np.uint32); m = np.ascontiguousarray(mass_inertia, dtype=np.float32)
        assert m.shape == (len(i), 4)
        _check(self._lib, self._lib.pxb_scene_set_mass_properties(self._h, _ptr(i), _ptr(m), len(i)))

    def setGravity(self, g):
        """PxScene::setGravity: read by the next simulate."""
        v = np.ascontiguousarray(g, dtype=np.float32).reshape(3)
        _check(self._lib, self._lib.pxb_scene_set_gravity(self._h, _ptr(v)))

    def setKinematicTargets(self, indices, poses):
        """PxRigidDynamic::setKinematicTarget for the kinematic bodies `indices` (dynamic-body indices): (n, 7) PxTransform rows (q.xyzw, p.xyz), consumed by the next simulate."""
        i = np.ascontiguousarray(indices, dtype=np.uint32); p = np.ascontiguousarray(poses, dtype=np.float32)
        assert p.shape == (len(i), 7)
        _check(self._lib, self._lib.pxb_scene_set_kinematic_targets(self._h, _ptr(i), _ptr(p), len(i)))

    def setKinematicTargetsDevice(self, dev_indices: int, dev_poses: int, nb: int):
        _check(self._lib, self._lib.pxb_scene_set_kinematic_targets_device(self._h, dev_indices, dev_poses, nb))

    def setForces(self, forces=None, torques=None):
        """PxDirectGPUAPI::setRigidDynamicData(eFORCE / eTORQUE) for every dynamic body: applied by the next simulate only."""
        if forces is not None:
            self.setRigidDynamicData(RD_FORCE, forces)
        if torques is not None:
            self.setRigidDynamicData(RD_TORQUE, torques)

    def getRigidDynamicDataDevice(self, data_type: int, dev_ptr: int, nb: int, dev_indices: int = 0):
        _check(self._lib, self._lib.pxb_get_rigid_dynamic_data_device(self._h, dev_ptr, dev_indices or None, data_type, nb))

    def setRigidDynamicDataDevice(self, data_type: int, dev_ptr: int, nb: int, dev_indices: int = 0):
        _check(self._lib, self._lib.pxb_set_rigid_dynamic_data_device(self._h, dev_ptr, dev_indices or None, data_type, nb))

    def getRigidDynamicDataDeviceEv(self, data_type: int, dev_ptr: int, nb: int, dev_indices: int = 0, start_event: int = 0, finish_event: int = 0):
        """PxDirectGPUAPI::getRigidDynamicData with its startEvent / finishEvent arguments (cudaEvent_t handles; 0 = none, no finish event = synchronous)."""
        _check(self._lib, self._lib.pxb_get_rigid_dynamic_data_device_ev(self._h, dev_ptr, dev_indices or None, data_type, nb, start_event or None, finish_event or None))

    def setRigidDynamicDataDeviceEv(self, data_type: int, dev_ptr: int, nb: int, dev_indices: int = 0, start_event: int = 0, finish_event: int = 0):
        _check(self._lib, self._lib.pxb_set_rigid_dynamic_data_device_ev(self._h, dev_ptr, dev_indices or None, data_type, nb, start_event or None, finish_event or None))

    def getStatesDevice(self, dev_ptr: int):
        """13 floats per dynamic body (pos3 quat4 linVel3 angVel3) into a device buffer, async on the scene stream."""
        _check(self._lib, self._lib.pxb_scene_get_states_device(self._h, dev_ptr))

    def scatterToPeers(self, stream_ptr: int, src_ptr: int, nbytes: int, dst_ptrs, ctas: int = 32):
        """One kernel that stores a device block into up to 8 peer-mapped buffers (multi-GPU state exchange)."""
        arr = (ctypes.c_uint64 * len(dst_ptrs))(*[int(p) for p in dst_ptrs])
        _check(self._lib, self._lib.pxb_scatter_to_peers(self._h, stream_ptr or None, src_ptr, nbytes, arr, len(dst_ptrs), ctas))

    def setStateExport(self, dst_ptrs=(), row_offset: int = 0):
        """Fused state export: from the next simulate on, the step itself stores the packed [n_dyn, 13] state block at row `row_offset` of every
        buffer in `dst_ptrs` (own / peer-mapped device memory, or mapped pinned host memory).  Empty = off."""
        arr = (ctypes.c_void_p * max(1, len(dst_ptrs)))(*[int(p) for p in dst_ptrs])
        _check(self._lib, self._lib.pxb_scene_set_state_export(self._h, arr, len(dst_ptrs), int(row_offset)))

    def peerSignal(self, stream_ptr: int, flag_ptrs, value: int):
        arr = (ctypes.c_uint64 * max(1, len(flag_ptrs)))(*[int(p) for p in flag_ptrs])
        _check(self._lib, self._lib.pxb_peer_signal(self._h, stream_ptr or None, arr, len(flag_ptrs), int(value) & 0xFFFFFFFF))

    def peerWait(self, stream_ptr: int, flags_ptr: int, n: int, value: int):
        _check(self._lib, self._lib.pxb_peer_wait(self._h, stream_ptr or None, flags_ptr, int(n), int(value) & 0xFFFFFFFF))

    def stream(self) -> int:
        return int(self._lib.pxb_scene_stream(self._h) or 0)

    # ---- packed state / stage level ----
    def getStates(self):
        out = np.zeros((self.num_dynamic, _scenes.STATE_FLOATS), np.float32)
        _check(self._lib, self._lib.pxb_scene_get_states(self._h, _ptr(out)))
        return out

    def setStates(self, st):
        st = np.ascontiguousarray(st, dtype=np.float32)
        assert st.shape == (self.num_dynamic, _scenes.STATE_FLOATS)
        _check(self._lib, self._lib.pxb_scene_set_states(self._h, _ptr(st)))

    def computeBounds(self):
        _check(self._lib, self._lib.pxb_scene_compute_bounds(self._h))
        out = np.zeros((self.num_actors, 6), np.float32)
        _check(self._lib, self._lib.pxb_scene_get_bounds(self._h, _ptr(out)))
        return out

    def getBounds(self):
        out = np.zeros((self.num_actors, 6), np.float32)
        _check(self._lib, self._lib.pxb_scene_get_bounds(self._h, _ptr(out)))
        return out

    def broadphase(self, tight_bounds=None):
        b = None if tight_bounds is None else np.ascontiguousarray(tight_bounds, dtype=np.float32)
        _check(self._lib, self._lib.pxb_scene_broadphase(self._h, _ptr(b)))

    def _pairs(self, count_fn, get_fn):
        n = int(count_fn(self._h))
        out = np.zeros((n, 2), np.uint32)
        if n:
            _check(self._lib, get_fn(self._h, _ptr(out)))
        return out

    def getPairs(self):
        return self._pairs(self._lib.pxb_scene_num_pairs, self._lib.pxb_scene_get_pairs)

    def getCreatedPairs(self):
        return self._pairs(self._lib.pxb_scene_num_created, self._lib.pxb_scene_get_created)

    def getDeletedPairs(self):
        return self._pairs(self._lib.pxb_scene_num_deleted, self._lib.pxb_scene_get_deleted)

    def getTouchFound(self):
        """pairs that started producing contacts in the last step (touch-found events)"""
        return self._pairs(self._lib.pxb_scene_num_touch_found, self._lib.pxb_scene_get_touch_found)

    def getTouchLost(self):
        """pairs that stopped producing contacts in the last step, incl. touching pairs that left the broadphase"""
        return self._pairs(self._lib.pxb_scene_num_touch_lost, self._lib.pxb_scene_get_touch_lost)

    GPU_CONTACT_PAIR_DTYPE = np.dtype([("contactPatches", "<u8"), ("contactPoints", "<u8"), ("contactForces", "<u8"), ("frictionPatches", "<u8"), ("transformCacheRef0", "<u4"),
                                       ("transformCacheRef1", "<u4"), ("nodeIndex0", "<u8"), ("nodeIndex1", "<u8"), ("actor0", "<u8"), ("actor1", "<u8"), ("nbContacts", "<u2"),
                                       ("nbPatches", "<u2"), ("pad", "<u4")])   # PxGpuContactPair, PxContact.h:818-833

    def sync(self):
        """wait for everything queued on the scene stream"""
        _check(self._lib, self._lib.pxb_scene_sync(self._h))

    def enableContactData(self, on=True):
        """PxDirectGPUAPI::copyContactData needs the friction write-back of the step: switch it on BEFORE the step whose contacts are wanted"""
        _check(self._lib, self._lib.pxb_scene_enable_contact_data(self._h, 1 if on else 0))

    def copyContactData(self, data_ptr, count_ptr, max_pairs):
        """PxDirectGPUAPI::copyContactData: PxGpuContactPair records into device memory `data_ptr`, their number into the device word `count_ptr` (stream-ordered)"""
        _check(self._lib, self._lib.pxb_scene_copy_contact_data(self._h, ctypes.c_void_p(int(data_ptr)), ctypes.c_void_p(int(count_ptr)), int(max_pairs)))

    def getContacts(self):
        n = int(self._lib.pxb_scene_num_pairs(self._h))
        out = np.zeros((n, 24), np.float32)
        if n:
            _check(self._lib, self._lib.pxb_scene_get_contacts(self._h, _ptr(out)))
        return out

    STAGES = ("broadphase", "narrowphase", "colouring", "prep", "solve", "integrate", "step")

    def setProfiling(self, on: bool = True):
        _check(self._lib, self._lib.pxb_scene_set_profiling(self._h, 1 if on else 0))

    def getStageTimes(self):
        ms = np.zeros(7, np.float32)
        _check(self._lib, self._lib.pxb_scene_get_stage_times(self._h, _ptr(ms)))
        return dict(zip(self.STAGES, (float(x) for x in ms)))

    @property
    def num_partitions(self):
        return int(self._lib.pxb_scene_last_num_partitions(self._h))

    @property
    def num_constraints(self):
        return int(self._lib.pxb_scene_last_num_constraints(self._h))

    def getSleep(self):
        """(wake counters f32[n_dyn], asleep flags u32[n_dyn]) -- PxRigidDynamic::getWakeCounter / isSleeping"""
        w, a = np.zeros(self.num_dynamic, np.float32), np.zeros(self.num_dynamic, np.uint32)
        _check(self._lib, self._lib.pxb_scene_get_sleep_data(self._h, _ptr(w), _ptr(a)))
        return w, a

    @property
    def uses_env_path(self):
        return bool(self._lib.pxb_scene_uses_env_path(self._h))

    @property
    def num_launches(self):
        return int(self._lib.pxb_scene_last_num_launches(self._h))


class BroadPhase:
    """Standalone broadphase object (pxb_bp_*): what the plugin shim's Bp::BroadPhase forwards to.  Host-side mirror for the parity tests:
    update(bounds, dist, groups, envs, created, updated, removed) + fetch() -> (created pairs, deleted pairs)."""
    # Bp::BpFilter table for the default pair filtering modes (BpFiltering.cpp): statics never pair with statics, everything else may
    LUT_DEFAULT = np.ones((7, 7), np.uint8)
    LUT_DEFAULT[0, 0] = 0

    def __init__(self, max_objects: int, max_pairs: int = 0, device: int = 0):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        _check(self._lib, self._lib.pxb_bp_create(int(max_objects), int(max_pairs), int(device), ctypes.byref(self._h)))

    def release(self):
        if getattr(self, "_h", None):
            self._lib.pxb_bp_release(self._h)
            self._h = None

    __del__ = release

    def update(self, bounds6, dist, groups, envs=None, created=(), updated=(), removed=(), lut=None):
        b = np.ascontiguousarray(bounds6, np.float32); d = np.ascontiguousarray(dist, np.float32); g = np.ascontiguousarray(groups, np.uint32)
        e = None if envs is None else np.ascontiguousarray(envs, np.uint32)
        l = np.ascontiguousarray(self.LUT_DEFAULT if lut is None else lut, np.uint8)
        c, u, r = (np.ascontiguousarray(x, np.uint32) for x in (created, updated, removed))
        _check(self._lib, self._lib.pxb_bp_update(self._h, _ptr(b), _ptr(d), _ptr(g), _ptr(e), len(g), _ptr(l), _ptr(c), len(c), _ptr(u), len(u), _ptr(r), len(r)))

    def fetch(self):
        cp, dp, nc, nd = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(self._lib, self._lib.pxb_bp_fetch(self._h, ctypes.byref(cp), ctypes.byref(nc), ctypes.byref(dp), ctypes.byref(nd)))
        def arr(p, n):
            return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), (n.value, 2)).copy() if n.value else np.zeros((0, 2), np.uint32)
        return arr(cp, nc), arr(dp, nd)
