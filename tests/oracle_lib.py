"""ctypes binding of the CPU oracle restatement (oracle/libpxb_oracle.so) -- the checker, never the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-f", "oracle/Makefile"], cwd=ROOT, check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "libpxb_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        vp, u32 = ctypes.c_void_p, ctypes.c_uint32
        L.pxo_scene_create.restype = vp
        L.pxo_scene_create.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.pxo_scene_destroy.argtypes = [vp]
        L.pxo_scene_step.argtypes = [vp, vp, u32]
        for f in ("pxo_scene_num_dynamic", "pxo_scene_num_actors", "pxo_scene_num_pairs", "pxo_scene_num_created",
                  "pxo_scene_num_deleted", "pxo_scene_last_num_partitions", "pxo_scene_last_num_constraints", "pxo_scene_unsupported_pairs"):
            getattr(L, f).argtypes = [vp]
            getattr(L, f).restype = u32
        for f in ("pxo_scene_get_states", "pxo_scene_set_states", "pxo_scene_get_bounds", "pxo_scene_set_bounds",
                  "pxo_scene_get_pairs", "pxo_scene_get_created", "pxo_scene_get_deleted", "pxo_scene_get_contacts"):
            getattr(L, f).argtypes = [vp, vp]
        L.pxo_debug_epa_calls.restype = u32
        L.pxo_scene_get_sleep.argtypes = [vp, vp, vp]
        L.pxo_scene_set_forces.argtypes = [vp, vp, vp]
        L.pxo_scene_set_kinematic_targets.argtypes = [vp, vp, vp, u32]
        L.pxo_scene_set_gravity.argtypes = [vp, vp]
        L.pxo_scene_set_mass_properties.argtypes = [vp, vp, vp, u32]
        L.pxo_scene_compute_bounds.argtypes = [vp]
        L.pxo_scene_broadphase.argtypes = [vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def epa_calls():
    """number of EPA penetration queries run so far in this process (coverage check for the a10 tests)"""
    return int(lib().pxo_debug_epa_calls())


class OracleScene:
    def __init__(self, scene):
        self.L = lib()
        buf = scene.tobytes()
        self.h = ctypes.c_void_p(self.L.pxo_scene_create(buf, len(buf)))
        assert self.h, "oracle rejected the scene"
        self.num_dynamic = self.L.pxo_scene_num_dynamic(self.h)
        self.num_actors = self.L.pxo_scene_num_actors(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.pxo_scene_destroy(self.h)
            self.h = None

    @property
    def unsupported_pairs(self):
        return int(self.L.pxo_scene_unsupported_pairs(self.h))

    def getSleep(self):
        """(wake counters f32[n_dyn], asleep flags u32[n_dyn]) -- PxRigidDynamic::getWakeCounter / isSleeping"""
        w, a = np.zeros(self.num_dynamic, np.float32), np.zeros(self.num_dynamic, np.uint32)
        self.L.pxo_scene_get_sleep(self.h, _p(w), _p(a))
        return w, a

    def setForces(self, forces=None, torques=None):
        """PxDirectGPUAPI::setRigidDynamicData(eFORCE / eTORQUE): (n_dyn, 3) arrays applied at the next step only"""
        f = None if forces is None else np.ascontiguousarray(forces, dtype=np.float32)
        t = None if torques is None else np.ascontiguousarray(torques, dtype=np.float32)
        self.L.pxo_scene_set_forces(self.h, _p(f), _p(t))

    def setMassProperties(self, dyn_indices, mass_inertia):
        i = np.ascontiguousarray(dyn_indices, dtype=np.uint32); m = np.ascontiguousarray(mass_inertia, dtype=np.float32)
        self.L.pxo_scene_set_mass_properties(self.h, _p(i), _p(m), len(i))

    def setGravity(self, g):
        v = np.ascontiguousarray(g, dtype=np.float32).reshape(3)
        self.L.pxo_scene_set_gravity(self.h, _p(v))

    def setKinematicTargets(self, dyn_indices, poses):
        """PxRigidDynamic::setKinematicTarget: (n, 7) PxTransform rows (q.xyzw, p.xyz) for the kinematic bodies `dyn_indices`; consumed by the next step"""
        i = np.ascontiguousarray(dyn_indices, dtype=np.uint32); p = np.ascontiguousarray(poses, dtype=np.float32)
        assert self.L.pxo_scene_set_kinematic_targets(self.h, _p(i), _p(p), len(i)) == 0

    def step(self, order=None):
        if order is None or len(order) == 0:
            self.L.pxo_scene_step(self.h, None, 0)
        else:
            o = np.ascontiguousarray(order, dtype=np.uint32)
            self.L.pxo_scene_step(self.h, _p(o), len(o))

    def getStates(self):
        out = np.zeros((self.num_dynamic, 13), np.float32)
        self.L.pxo_scene_get_states(self.h, _p(out))
        return out

    def setStates(self, st):
        st = np.ascontiguousarray(st, dtype=np.float32)
        self.L.pxo_scene_set_states(self.h, _p(st))

    def computeBounds(self):
        self.L.pxo_scene_compute_bounds(self.h)
        out = np.zeros((self.num_actors, 6), np.float32)
        self.L.pxo_scene_get_bounds(self.h, _p(out))
        return out

    def setBounds(self, b):
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float32)
        self.L.pxo_scene_set_bounds(self.h, _p(b))

    def broadphase(self):
        self.L.pxo_scene_broadphase(self.h)

    def _pairs(self, nfn, gfn):
        n = nfn(self.h)
        out = np.zeros((n, 2), np.uint32)
        if n:
            gfn(self.h, _p(out))
        return out

    def getPairs(self):
        return self._pairs(self.L.pxo_scene_num_pairs, self.L.pxo_scene_get_pairs)

    def getCreatedPairs(self):
        return self._pairs(self.L.pxo_scene_num_created, self.L.pxo_scene_get_created)

    def getDeletedPairs(self):
        return self._pairs(self.L.pxo_scene_num_deleted, self.L.pxo_scene_get_deleted)

    def getContacts(self):
        n = self.L.pxo_scene_num_pairs(self.h)
        out = np.zeros((n, 24), np.float32)
        if n:
            self.L.pxo_scene_get_contacts(self.h, _p(out))
        return out

    @property
    def num_partitions(self):
        return self.L.pxo_scene_last_num_partitions(self.h)

    @property
    def num_constraints(self):
        return self.L.pxo_scene_last_num_constraints(self.h)
