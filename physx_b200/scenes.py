"""Synthetic scene builders for the BASELINE.json configs and the scene file
format shared with the oracle tools (oracle/scene_format.h).

A scene is plain data: a header + one record per actor (one shape per actor, identity local pose).
Mass properties are computed here in float32 (box: m = rho*8*hx*hy*hz, I = m/3*(hy^2+hz^2, ...),
the closed forms PxRigidBodyExt::updateMassAndInertia evaluates for a box,
physx/source/physxextensions/src/ExtRigidBodyExt.cpp) and set explicitly on both sides so the
reference and this engine start from bit-identical inputs.
"""
from __future__ import annotations

import numpy as np

GEOM_SPHERE, GEOM_PLANE, GEOM_CAPSULE, GEOM_BOX, GEOM_CONVEX = 0, 1, 2, 3, 5
ACTOR_DYNAMIC = 1
SOLVER_PGS, SOLVER_TGS = 0, 1
SCENE_MAGIC = 0x314E4353
NO_ENV = 0xFFFFFFFF

HEADER_DTYPE = np.dtype([
    ("magic", "<u4"), ("nActors", "<u4"), ("nHulls", "<u4"), ("solverType", "<u4"),
    ("gravity", "<f4", 3), ("dt", "<f4"), ("posIters", "<u4"), ("velIters", "<u4"),
    ("staticFriction", "<f4"), ("dynamicFriction", "<f4"), ("restitution", "<f4"),
    ("contactOffset", "<f4"), ("restOffset", "<f4"), ("sleepThreshold", "<f4"),
    ("bounceThreshold", "<f4"), ("frictionOffsetThreshold", "<f4"),
    ("frictionCorrelationDistance", "<f4"), ("toleranceLength", "<f4"), ("reserved", "<u4", 4),
])

ACTOR_DTYPE = np.dtype([
    ("flags", "<u4"), ("geomType", "<u4"), ("envId", "<u4"), ("hullIdx", "<u4"),
    ("pos", "<f4", 3), ("quat", "<f4", 4), ("dims", "<f4", 4),
    ("linVel", "<f4", 3), ("angVel", "<f4", 3), ("mass", "<f4"), ("inertia", "<f4", 3),
    ("linDamping", "<f4"), ("angDamping", "<f4"), ("maxLinVel", "<f4"), ("maxAngVel", "<f4"),
    ("maxDepenetrationVel", "<f4"), ("materialIndex", "<u4"), ("aggregate", "<u4"),
])
MATERIAL_DTYPE = np.dtype([("staticFriction", "<f4"), ("dynamicFriction", "<f4"), ("restitution", "<f4"), ("bits", "<u4")])
COMBINE_AVERAGE, COMBINE_MIN, COMBINE_MULTIPLY, COMBINE_MAX = 0, 1, 2, 3
MATERIAL_DISABLE_FRICTION = 1 << 8


def make_materials(entries):
    """entries: (staticFriction, dynamicFriction, restitution, frictionCombineMode, restitutionCombineMode[, disableFriction]) tuples"""
    m = np.zeros(len(entries), MATERIAL_DTYPE)
    for i, e in enumerate(entries):
        m[i] = (e[0], e[1], e[2], int(e[3]) | int(e[4]) << 4 | (MATERIAL_DISABLE_FRICTION if len(e) > 5 and e[5] else 0))
    return m

assert HEADER_DTYPE.itemsize == 96 and ACTOR_DTYPE.itemsize == 128

STATE_FLOATS = 13  # pos3 quat4 linVel3 angVel3

def normalize_quat_f32(q):
    """PxQuat::getNormalized in float32, iterated to its fixed point.  The reference normalises every
    pose handed to createRigidStatic/Dynamic (physx/source/physx/src/NpPhysics.cpp); scene files carry
    quaternions that normalisation leaves unchanged so both sides start from identical bits."""
    q = np.asarray(q, dtype=np.float32)

    def settle(q):
        for _ in range(8):
            m = np.float32(np.sqrt(np.float32(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])))
            q2 = (q * (np.float32(1.0) / m)).astype(np.float32)
            if (q2 == q).all():
                return q, True
            q = q2
        return q, False

    q, ok = settle(q)
    if ok:
        return q
    # normalisation can settle into a 2-cycle one ulp apart: re-solve one component so that the float32 norm is exactly 1 (then 1/m = 1 and
    # the quaternion is a fixed point), trying the components from the smallest to the largest and a window of ulps around the exact solution
    def norm_is_one(c):
        return np.float32(np.sqrt(np.float32(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]))) == np.float32(1.0)

    for k in np.argsort(np.where(q == 0, np.inf, np.abs(q))):
        if q[k] == 0:
            continue
        rest = float(np.sum(np.square(q.astype(np.float64)))) - float(q[k]) ** 2
        base = q.copy()
        base[k] = np.float32(np.sign(q[k]) * np.sqrt(max(0.0, 1.0 - rest)))
        for off in range(0, 4096):
            for sgn in (1, -1):
                c = base.copy()
                c[k:k + 1].view(np.int32)[0] += sgn * off
                if norm_is_one(c):
                    c2, ok = settle(c)
                    if ok:
                        return c2
    raise ValueError("no float32 normalisation fixed point found")


# rotation taking local +X (PxPlaneGeometry normal) to world +Y: 90 deg about Z
PLANE_UP_Y_QUAT = normalize_quat_f32([0.0, 0.0, np.sqrt(0.5), np.sqrt(0.5)])


def default_header(solver=SOLVER_TGS, pos_iters=4, vel_iters=1, sleep_threshold=0.0, relaxed_partitioning=False):
    h = np.zeros((), dtype=HEADER_DTYPE)
    h["magic"] = SCENE_MAGIC
    h["solverType"] = solver
    h["gravity"] = (0.0, -9.81, 0.0)
    h["dt"] = np.float32(1.0 / 60.0)
    h["posIters"], h["velIters"] = pos_iters, vel_iters
    h["staticFriction"], h["dynamicFriction"], h["restitution"] = 0.5, 0.5, 0.6
    h["contactOffset"], h["restOffset"] = 0.02, 0.0
    h["sleepThreshold"] = sleep_threshold
    h["bounceThreshold"] = 2.0          # 0.2 * toleranceSpeed (PxSceneDesc.h)
    h["frictionOffsetThreshold"] = 0.04
    h["frictionCorrelationDistance"] = 0.025
    h["toleranceLength"] = 1.0
    h["reserved"][0] = 1 if relaxed_partitioning else 0   # bit 0: PXB_FLAG_RELAXED_PARTITIONING (engine and oracle; ignored by the reference)
    return h


def _new_actors(n):
    a = np.zeros(n, dtype=ACTOR_DTYPE)
    a["quat"][:, 3] = 1.0
    a["envId"] = NO_ENV
    a["maxLinVel"] = 1.0e16   # PX_MAX_F32-ish: PxRigidDynamic default is 1e32^0.5
    a["maxAngVel"] = 100.0    # PxRigidDynamic default
    a["maxDepenetrationVel"] = 1.0e32  # NpRigidDynamic default (PX_MAX_F32 in practice)
    a["angDamping"] = 0.05    # PxRigidDynamic default
    return a


def set_box(a, idx, half_extents, density=10.0):
    he = np.broadcast_to(np.asarray(half_extents, dtype=np.float32), (np.size(idx) if np.ndim(idx) else 1, 3)) \
        if np.ndim(half_extents) == 1 else np.asarray(half_extents, dtype=np.float32)
    a["geomType"][idx] = GEOM_BOX
    a["flags"][idx] = ACTOR_DYNAMIC
    a["dims"][idx, :3] = he
    hx, hy, hz = he[..., 0], he[..., 1], he[..., 2]
    m = (np.float32(density) * np.float32(8.0) * hx * hy * hz).astype(np.float32)
    third = np.float32(1.0 / 3.0)
    a["mass"][idx] = m
    a["inertia"][idx, 0] = m * third * (hy * hy + hz * hz)
    a["inertia"][idx, 1] = m * third * (hx * hx + hz * hz)
    a["inertia"][idx, 2] = m * third * (hx * hx + hy * hy)


def add_ground_plane(actors):
    p = _new_actors(1)
    p["geomType"] = GEOM_PLANE
    p["quat"][0] = PLANE_UP_Y_QUAT
    return np.concatenate([p, actors])


COOKED_MAGIC = 0x43485850   # "PXHC": header.reserved[1] when the cooked-hull section is present (oracle/scene_format.h)
COOKED_HDR_DTYPE = np.dtype([
    ("nVerts", "<u4"), ("nPolys", "<u4"), ("nEdges", "<u4"), ("nIdx", "<u4"),
    ("centerOfMass", "<f4", 3), ("boundsCenter", "<f4", 3), ("boundsExtents", "<f4", 3),
    ("internalRadius", "<f4"), ("internalExtents", "<f4", 3),
    ("unitMass", "<f4"), ("unitInertiaDiag", "<f4", 3), ("unitCom", "<f4", 3), ("reserved", "<u4", 1),
])
COOKED_POLY_DTYPE = np.dtype([("plane", "<f4", 4), ("vref", "<u4"), ("nbVerts", "<u4"), ("minIndex", "<u4"), ("pad", "<u4")])
assert COOKED_HDR_DTYPE.itemsize == 100 and COOKED_POLY_DTYPE.itemsize == 32


def parse_cooked(buf, n_hulls, off=0):
    """cooked-hull section -> list of dicts (hdr, verts, polys, vertexRefs, facesByEdges [, samples, valencies, adjacentVerts]); returns (list, end offset)"""
    out = []
    for _ in range(n_hulls):
        hdr = np.frombuffer(buf, COOKED_HDR_DTYPE, 1, off)[0].copy(); off += COOKED_HDR_DTYPE.itemsize
        nv, npoly, ne, ni = int(hdr["nVerts"]), int(hdr["nPolys"]), int(hdr["nEdges"]), int(hdr["nIdx"])
        verts = np.frombuffer(buf, "<f4", nv * 3, off).reshape(nv, 3).copy(); off += nv * 12
        polys = np.frombuffer(buf, COOKED_POLY_DTYPE, npoly, off).copy(); off += npoly * COOKED_POLY_DTYPE.itemsize
        refs = np.frombuffer(buf, np.uint8, ni, off).copy(); off += (ni + 3) // 4 * 4
        fbe = np.frombuffer(buf, np.uint8, 2 * ne, off).copy(); off += (2 * ne + 3) // 4 * 4
        rec = dict(hdr=hdr, verts=verts, polys=polys, vertexRefs=refs, facesByEdges=fbe)
        big = int(hdr["reserved"][0])   # hill-climbing data of hulls with more than 32 vertices (Gu::BigConvexRawData): subdiv | nAdj << 16
        if big:
            subdiv, n_adj = big & 0xffff, big >> 16
            ns = 6 * subdiv * subdiv
            rec["samples"] = np.frombuffer(buf, np.uint8, ns, off).copy(); off += (ns + 3) // 4 * 4
            rec["valencies"] = np.frombuffer(buf, "<u2", 2 * nv, off).reshape(nv, 2).copy(); off += 4 * nv
            rec["adjacentVerts"] = np.frombuffer(buf, np.uint8, n_adj, off).copy(); off += (n_adj + 3) // 4 * 4
        out.append(rec)
    return out, off


# Local poses (PxgShapeSim.shape2Actor + PxsBodyCore.body2Actor): per actor the shape's pose in the actor frame (PxShape::setLocalPose)
# and the centre-of-mass frame in the actor frame (PxRigidBody::setCMassLocalPose).  ActorRec.pos / quat stay the ACTOR pose (PxRigidActor::getGlobalPose).
LOCAL_POSE_DTYPE = np.dtype([("shapeP", "<f4", 3), ("shapeQ", "<f4", 4), ("bodyP", "<f4", 3), ("bodyQ", "<f4", 4), ("pad", "<f4", 2)])
LOCAL_POSE_MAGIC = 0x504c5850   # "PXLP": header.reserved[3] when the local-pose section (one record per actor, after the material table) is present
assert LOCAL_POSE_DTYPE.itemsize == 64


# f1: PxDefaultSimulationFilterShader state (oracle/scene_format.h PxbFilterShaderConfig) + PxFilterData per actor
FILTER_CONFIG_DTYPE = np.dtype([("collisionTable", "<u4", 32), ("ops", "<u4", 3), ("filterBool", "<u4"), ("constants", "<u4", 4)])
FLAG_FILTER_SECTION = 2
FLAG_SHAPE_OFFSETS = 4   # header.reserved[0] bit 2: per-shape (contactOffset, restOffset) section
FILTER_AND, FILTER_OR, FILTER_XOR, FILTER_NAND, FILTER_NOR, FILTER_NXOR, FILTER_SWAP_AND = range(7)
assert FILTER_CONFIG_DTYPE.itemsize == 160


def default_filter_config():
    """the extension's initial state: every group collides with every group, ops AND / AND / AND, constants 0, filterBool false"""
    c = np.zeros((), FILTER_CONFIG_DTYPE)
    c["collisionTable"][:] = 0xffffffff
    return c


def set_group_collision_flag(cfg, g1, g2, enable):   # PxSetGroupCollisionFlag
    for a, b in ((g1, g2), (g2, g1)):
        if enable:
            cfg["collisionTable"][a] |= np.uint32(1 << b)
        else:
            cfg["collisionTable"][a] &= np.uint32(~(1 << b) & 0xffffffff)


def identity_local_poses(n):
    lp = np.zeros(n, LOCAL_POSE_DTYPE)
    lp["shapeQ"][:, 3] = 1.0
    lp["bodyQ"][:, 3] = 1.0
    return lp


class Scene:
    def __init__(self, header, actors, hulls=(), cooked=b"", materials=None, local_poses=None, filter_config=None, filter_data=None, shape_offsets=None):
        self.header = header.copy()
        self.actors = actors
        # PxShape::setContactOffset / setRestOffset per shape: (n, 2) float32, or None = the header's values for every shape
        self.shape_offsets = None if shape_offsets is None else np.ascontiguousarray(shape_offsets, dtype="<f4").reshape(len(actors), 2)
        self.header["reserved"][0] = (int(self.header["reserved"][0]) & ~FLAG_SHAPE_OFFSETS) | (FLAG_SHAPE_OFFSETS if self.shape_offsets is not None else 0)
        # default simulation filter shader: global state + PxFilterData (word0..3) per actor
        self.filter_config = None if filter_config is None else np.asarray(filter_config, FILTER_CONFIG_DTYPE).reshape(())
        self.filter_data = None if filter_data is None else np.ascontiguousarray(filter_data, dtype="<u4").reshape(len(actors), 4)
        assert (self.filter_config is None) == (self.filter_data is None)
        self.header["reserved"][0] = (int(self.header["reserved"][0]) & ~FLAG_FILTER_SECTION) | (FLAG_FILTER_SECTION if self.filter_config is not None else 0)
        self.local_poses = None if local_poses is None else np.asarray(local_poses, LOCAL_POSE_DTYPE)
        assert self.local_poses is None or len(self.local_poses) == len(actors)
        self.header["reserved"][3] = LOCAL_POSE_MAGIC if self.local_poses is not None else 0
        self.materials = np.zeros(0, MATERIAL_DTYPE) if materials is None else np.asarray(materials, MATERIAL_DTYPE)   # material table (empty: the header's material)
        self.header["reserved"][2] = len(self.materials)
        self.hulls = list(hulls)
        self.cooked = bytes(cooked)   # cooked-hull section (reference cooking output, see cook_hulls); empty = not cooked
        self.header["nActors"] = len(actors)
        self.header["nHulls"] = len(self.hulls)

    @property
    def n_dynamic(self):
        return int(np.count_nonzero(self.actors["flags"] & ACTOR_DYNAMIC))

    def tobytes(self):
        h = self.header.copy()
        h["reserved"][1] = COOKED_MAGIC if self.cooked else 0
        out = [h.tobytes(), self.actors.tobytes(), self.materials.tobytes()]
        if self.local_poses is not None:
            out.append(self.local_poses.tobytes())
        if self.filter_config is not None:
            out.append(self.filter_config.tobytes()); out.append(self.filter_data.tobytes())
        if self.shape_offsets is not None:
            out.append(self.shape_offsets.tobytes())
        for hl in self.hulls:
            hl = np.asarray(hl, dtype="<f4").reshape(-1, 3)
            out.append(np.uint32(len(hl)).tobytes())
            out.append(hl.tobytes())
        out.append(self.cooked)
        return b"".join(out)

    def save(self, path):
        with open(path, "wb") as f:
            f.write(self.tobytes())

    @staticmethod
    def load(path):
        buf = open(path, "rb").read()
        h = np.frombuffer(buf, dtype=HEADER_DTYPE, count=1)[0].copy()
        off = HEADER_DTYPE.itemsize
        a = np.frombuffer(buf, dtype=ACTOR_DTYPE, count=int(h["nActors"]), offset=off).copy()
        off += a.nbytes
        mats = np.frombuffer(buf, dtype=MATERIAL_DTYPE, count=int(h["reserved"][2]), offset=off).copy()
        off += mats.nbytes
        lp = None
        if int(h["reserved"][3]) == LOCAL_POSE_MAGIC:
            lp = np.frombuffer(buf, dtype=LOCAL_POSE_DTYPE, count=int(h["nActors"]), offset=off).copy()
            off += lp.nbytes
        fc = fd = None
        if int(h["reserved"][0]) & FLAG_FILTER_SECTION:
            fc = np.frombuffer(buf, dtype=FILTER_CONFIG_DTYPE, count=1, offset=off)[0].copy(); off += FILTER_CONFIG_DTYPE.itemsize
            fd = np.frombuffer(buf, dtype="<u4", count=4 * int(h["nActors"]), offset=off).reshape(-1, 4).copy(); off += fd.nbytes
        so = None
        if int(h["reserved"][0]) & FLAG_SHAPE_OFFSETS:
            so = np.frombuffer(buf, dtype="<f4", count=2 * int(h["nActors"]), offset=off).reshape(-1, 2).copy(); off += so.nbytes
        hulls = []
        for _ in range(int(h["nHulls"])):
            nv = int(np.frombuffer(buf, "<u4", 1, off)[0]); off += 4
            hulls.append(np.frombuffer(buf, "<f4", nv * 3, off).reshape(nv, 3).copy()); off += nv * 12
        cooked = buf[off:] if int(h["reserved"][1]) == COOKED_MAGIC else b""
        return Scene(h, a, hulls, cooked, mats, lp, fc, fd, so)

    def cooked_hulls(self):
        return parse_cooked(self.cooked, len(self.hulls))[0] if self.cooked else []


# Convex cooking is the host SDK's job (PxCreateConvexMesh); the test infrastructure cooks with the unmodified reference
# (tests/golden/cooking.py: cook_hulls) and scene fixtures / physx_b200/data/hull_library.npz carry the cooked bytes.  The builders below that
# create new hull point clouds take that function as `cook`; without it they return the scene uncooked (hull clouds only).


def random_hull_points(rng, n_points=16, radius=0.2):
    """points on a jittered sphere (12-20 vertex hulls, r ~ 0.2), centred on their mean"""
    p = rng.normal(size=(n_points, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    p *= radius * rng.uniform(0.75, 1.0, (n_points, 1))
    p -= p.mean(axis=0, keepdims=True)
    return p.astype(np.float32)


def set_convex(a, idx, hull_idx):
    a["geomType"][idx] = GEOM_CONVEX
    a["flags"][idx] = ACTOR_DYNAMIC
    a["hullIdx"][idx] = hull_idx
    a["mass"][idx] = 1.0          # placeholders until cook_hulls fills them from the cooked mass information
    a["inertia"][idx] = 0.01


def hulls_on_plane(n=8, n_hulls=3, seed=6, cook=None, **hdr):
    """Convex hulls (12-20 vertices) with random orientations and spins dropped onto the ground plane, spaced so that they never touch
    each other: exercises convex bounds and pcmContactPlaneConvex."""
    rng = np.random.RandomState(seed)
    hulls = [random_hull_points(rng, int(rng.randint(12, 21)), 0.25) for _ in range(n_hulls)]
    a = _new_actors(n)
    for i in range(n):
        set_convex(a, i, i % n_hulls)
        a["pos"][i] = (1.5 * (i % 4), 0.6 + 0.25 * (i // 4), 1.5 * (i // 4))
        a["angVel"][i] = rng.uniform(-2, 2, 3)
    a["quat"] = random_unit_quats(rng, n)
    sc = Scene(default_header(**hdr), add_ground_plane(a), hulls)
    return cook(sc) if cook else sc


def box_stacks(n_stacks=10, height=10, half_extent=0.5, spacing=4.0, jitter=0.0, seed=1234, **hdr):
    """BASELINE config 1: SnippetHelloWorld-style stacks on a ground plane."""
    rng = np.random.RandomState(seed)
    n = n_stacks * height
    a = _new_actors(n)
    he = np.float32(half_extent)
    k = 0
    for s in range(n_stacks):
        for j in range(height):
            jx, jz = (rng.uniform(-jitter, jitter, 2) if jitter > 0 else (0.0, 0.0))
            a["pos"][k] = (s * spacing + jx, he + 2 * he * j, jz)
            k += 1
    set_box(a, np.arange(n), np.array([he, he, he], dtype=np.float32))
    return Scene(default_header(**hdr), add_ground_plane(a))


def env_grid_stacks(n_envs=4096, stacks_per_env=8, height=8, half_extent=0.25, env_pitch=8.0,
                    stack_spacing=1.0, jitter=0.01, seed=1234, **hdr):
    """BASELINE config 2 / 5: E independent envs on a square grid, each `stacks_per_env` stacks of
    `height` boxes, +-jitter seeded offsets in the ground plane, one shared ground plane."""
    rng = np.random.RandomState(seed)
    side = int(np.ceil(np.sqrt(n_envs)))
    per_env = stacks_per_env * height
    n = n_envs * per_env
    a = _new_actors(n)
    he = np.float32(half_extent)
    env = np.repeat(np.arange(n_envs), per_env)
    local = np.tile(np.arange(per_env), n_envs)
    stack = local // height
    level = local % height
    ssx = int(np.ceil(np.sqrt(stacks_per_env)))
    ex = (env % side).astype(np.float32) * np.float32(env_pitch)
    ez = (env // side).astype(np.float32) * np.float32(env_pitch)
    jit = rng.uniform(-jitter, jitter, size=(n, 2)).astype(np.float32) if jitter > 0 else np.zeros((n, 2), np.float32)
    a["pos"][:, 0] = ex + (stack % ssx).astype(np.float32) * np.float32(stack_spacing) + jit[:, 0]
    a["pos"][:, 1] = he + np.float32(2.0) * he * level.astype(np.float32)
    a["pos"][:, 2] = ez + (stack // ssx).astype(np.float32) * np.float32(stack_spacing) + jit[:, 1]
    a["envId"] = env.astype(np.uint32)
    set_box(a, np.arange(n), np.array([he, he, he], dtype=np.float32))
    return Scene(default_header(**hdr), add_ground_plane(a))


def random_unit_quats(rng, n):
    q = rng.normal(size=(n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    return np.stack([normalize_quat_f32(x) for x in q]).astype(np.float32)


def tumbling_boxes(n=12, seed=7, half_extent=0.25, **hdr):
    """Boxes with random orientations and spins dropped from a loose column: exercises pair creation /
    deletion, manifold invalidation and regeneration, edge clipping and non-axis-aligned contacts."""
    rng = np.random.RandomState(seed)
    a = _new_actors(n)
    he = np.array([half_extent, half_extent * 0.8, half_extent * 1.3], dtype=np.float32)
    a["pos"][:, 0] = rng.uniform(-0.3, 0.3, n)
    a["pos"][:, 1] = 0.6 + 0.75 * np.arange(n)
    a["pos"][:, 2] = rng.uniform(-0.3, 0.3, n)
    a["quat"] = random_unit_quats(rng, n)
    a["angVel"] = rng.uniform(-2, 2, (n, 3))
    set_box(a, np.arange(n), he)
    return Scene(default_header(**hdr), add_ground_plane(a))


def set_sphere(a, idx, radius, density=10.0):
    r = np.float32(radius)
    a["geomType"][idx] = GEOM_SPHERE
    a["flags"][idx] = ACTOR_DYNAMIC
    a["dims"][idx, 0] = r
    m = np.float32(density * 4.0 / 3.0 * np.pi) * r * r * r
    a["mass"][idx] = m
    a["inertia"][idx] = (np.float32(0.4) * m * r * r).astype(np.float32)[..., None] if np.ndim(m) else np.float32(0.4) * m * r * r


def set_capsule(a, idx, radius, half_height, density=10.0):
    """Capsule along local X (PxCapsuleGeometry); solid cylinder + two hemispheres."""
    r, h = np.float32(radius), np.float32(half_height)
    a["geomType"][idx] = GEOM_CAPSULE
    a["flags"][idx] = ACTOR_DYNAMIC
    a["dims"][idx, 0] = r
    a["dims"][idx, 1] = h
    mc = np.float32(density * np.pi) * r * r * (2 * h)
    ms = np.float32(density * 4.0 / 3.0 * np.pi) * r * r * r
    a["mass"][idx] = mc + ms
    ix = mc * r * r * np.float32(0.5) + ms * r * r * np.float32(0.4)
    iy = mc * (r * r * np.float32(0.25) + h * h * np.float32(1.0 / 3.0)) + ms * (r * r * np.float32(0.4) + h * h + np.float32(0.75) * h * r)
    a["inertia"][idx, 0] = ix
    a["inertia"][idx, 1] = iy
    a["inertia"][idx, 2] = iy


def mixed_primitives(n=18, seed=3, kinds=("sphere", "capsule", "box"), spread=0.5, **hdr):
    """Spheres, capsules and boxes dropped in a loose column onto the ground plane (BASELINE config 3 shape
    without the convex hulls): exercises every primitive pair of the sphere family."""
    rng = np.random.RandomState(seed)
    a = _new_actors(n)
    a["pos"][:, 0] = rng.uniform(-spread, spread, n)
    a["pos"][:, 1] = 0.5 + 0.45 * np.arange(n)
    a["pos"][:, 2] = rng.uniform(-spread, spread, n)
    a["quat"] = random_unit_quats(rng, n)
    for i in range(n):
        k = kinds[i % len(kinds)]
        if k == "sphere":
            set_sphere(a, i, rng.uniform(0.1, 0.2))
        elif k == "capsule":
            set_capsule(a, i, rng.uniform(0.08, 0.15), rng.uniform(0.1, 0.3))
        else:
            set_box(a, np.array([i]), np.array([rng.uniform(0.12, 0.25), rng.uniform(0.12, 0.25), rng.uniform(0.12, 0.25)], dtype=np.float32))
    return Scene(default_header(**hdr), add_ground_plane(a))


def env_ragged(n_envs=12, max_bodies=20, seed=5, static_blocks=True, env_pitch=6.0, **hdr):
    """Environments of different sizes (including an empty one) with tumbling boxes / spheres / capsules, an optional
    static box inside every other environment and two env-less statics (ground plane + a wall plane): exercises the
    environment path's ragged lists, pair creation/deletion, static env members and several shared statics."""
    rng = np.random.RandomState(seed)
    recs = []
    for e in range(n_envs):
        nb = 0 if e == 3 else int(rng.randint(1, max_bodies + 1))
        a = _new_actors(nb + (1 if (static_blocks and e % 2 == 0) else 0))
        ex, ez = (e % 4) * env_pitch, (e // 4) * env_pitch
        for i in range(nb):
            # capsule-box pairs need the GJK/EPA family (not built yet): even environments hold boxes + spheres, odd ones capsules + spheres
            kind = rng.randint(0, 2) if e % 2 == 0 else rng.randint(1, 3)
            if kind == 0:
                set_box(a, np.array([i]), np.array([rng.uniform(0.12, 0.3), rng.uniform(0.12, 0.3), rng.uniform(0.12, 0.3)], dtype=np.float32))
            elif kind == 1:
                set_sphere(a, i, rng.uniform(0.1, 0.25))
            else:
                set_capsule(a, i, rng.uniform(0.08, 0.15), rng.uniform(0.1, 0.3))
            a["pos"][i] = (ex + rng.uniform(-0.4, 0.4), 0.5 + 0.5 * i, ez + rng.uniform(-0.4, 0.4))
            a["angVel"][i] = rng.uniform(-2, 2, 3)
        if nb:
            a["quat"][:nb] = random_unit_quats(rng, nb)
        if len(a) > nb:   # static box (flags = 0) that belongs to the environment
            a["geomType"][nb] = GEOM_BOX
            a["dims"][nb, :3] = (0.5, 0.2, 0.5)
            a["pos"][nb] = (ex + 0.2, 0.2, ez)
        a["envId"] = e
        recs.append(a)
    wall = _new_actors(1)   # env-less wall plane x >= -1.5 (normal +X = identity pose)
    wall["geomType"] = GEOM_PLANE
    wall["pos"][0] = (-1.5, 0.0, 0.0)
    actors = np.concatenate(recs + [wall])
    return Scene(default_header(**hdr), add_ground_plane(actors))


def add_bin(actors, half_size, wall_height=4.0, thickness=0.5):
    """Ground plane + four static box walls around [-half_size, half_size]^2 (BASELINE configs 3 / 4: "falling into a bin")."""
    w = _new_actors(4)
    w["geomType"] = GEOM_BOX
    t, h, L = np.float32(thickness), np.float32(wall_height), np.float32(half_size + thickness)
    w["dims"][0, :3] = (t, h, L); w["pos"][0] = (-half_size - t, h, 0)
    w["dims"][1, :3] = (t, h, L); w["pos"][1] = (half_size + t, h, 0)
    w["dims"][2, :3] = (L, h, t); w["pos"][2] = (0, h, -half_size - t)
    w["dims"][3, :3] = (L, h, t); w["pos"][3] = (0, h, half_size + t)
    return add_ground_plane(np.concatenate([w, actors]))


def box_pile(nx=100, ny=20, nz=100, half_extent=0.25, gap=0.001, seed=1, **hdr):
    """BASELINE config 4: nx*ny*nz boxes in a lattice with small gaps settling in a walled bin -- one giant island with high
    contact density (solver partitioning / colouring stress).  Device-wide path (no environment ids)."""
    rng = np.random.RandomState(seed)
    n = nx * ny * nz
    a = _new_actors(n)
    he = np.float32(half_extent)
    pitch = np.float32(2 * half_extent + gap)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    jit = rng.uniform(-gap * 0.4, gap * 0.4, size=(n, 2)).astype(np.float32)
    a["pos"][:, 0] = (ix.ravel() - (nx - 1) / 2).astype(np.float32) * pitch + jit[:, 0]
    a["pos"][:, 1] = he + iy.ravel().astype(np.float32) * pitch + np.float32(gap)
    a["pos"][:, 2] = (iz.ravel() - (nz - 1) / 2).astype(np.float32) * pitch + jit[:, 1]
    set_box(a, np.arange(n), np.array([he, he, he], dtype=np.float32))
    return Scene(default_header(**hdr), add_bin(a, half_size=float(max(nx, nz) * pitch / 2 + 0.5)))


def load_hull_library():
    """16 cooked convex hulls (12-20 input points, r ~ 0.2; cooked once by the reference's PxCreateConvexMesh with tests/golden/cooking.py
    and committed as physx_b200/data/hull_library.npz) -> (point clouds, cooked bytes, unit masses, unit inertias).
    Lets full-size scenes use hulls on the GPU box, where the reference (and its cooking) does not exist."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hull_library.npz"), allow_pickle=True)
    return [np.asarray(c, np.float32) for c in z["clouds"]], z["cooked"].tobytes(), z["unitMass"], z["unitInertiaDiag"]


def falling_primitives(nx=128, ny=64, nz=128, pitch=0.6, seed=2, kinds=("sphere", "box"), **hdr):
    """BASELINE config 3 shape (broadphase + narrowphase stress): nx*ny*nz mixed primitives with random orientations dropped
    from a lattice into a walled bin.  (Config 3 is spheres / capsules / convex hulls; boxes stand in for the hulls until the
    hull half of a10 lands: kinds=("sphere", "capsule", "box") runs every primitive pair type incl. capsule-box through GJK / EPA.)"""
    rng = np.random.RandomState(seed)
    n = nx * ny * nz
    a = _new_actors(n)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    a["pos"][:, 0] = (ix.ravel() - (nx - 1) / 2).astype(np.float32) * np.float32(pitch)
    a["pos"][:, 1] = np.float32(0.5) + iy.ravel().astype(np.float32) * np.float32(pitch)
    a["pos"][:, 2] = (iz.ravel() - (nz - 1) / 2).astype(np.float32) * np.float32(pitch)
    q = rng.normal(size=(n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    a["quat"] = q   # unit to float rounding; the engine does not depend on the normalisation fixed point
    kind = np.arange(n) % len(kinds)
    for ki, k in enumerate(kinds):
        idx = np.nonzero(kind == ki)[0]
        if k == "sphere":
            set_sphere(a, idx, rng.uniform(0.1, 0.2, len(idx)).astype(np.float32))
        elif k == "capsule":
            set_capsule(a, idx, rng.uniform(0.08, 0.15, len(idx)).astype(np.float32), rng.uniform(0.1, 0.3, len(idx)).astype(np.float32))
        elif k == "convex":   # the 16 library hulls, round robin; mass / inertia from the cooked mass information at density 10
            clouds, cooked, unit_mass, unit_inertia = load_hull_library()
            hi = (np.arange(len(idx)) % len(clouds)).astype(np.uint32)
            a["geomType"][idx] = GEOM_CONVEX; a["flags"][idx] = ACTOR_DYNAMIC; a["hullIdx"][idx] = hi
            a["mass"][idx] = np.float32(10.0) * unit_mass[hi]; a["inertia"][idx] = np.float32(10.0) * unit_inertia[hi]
        else:
            set_box(a, idx, rng.uniform(0.1, 0.2, (len(idx), 3)).astype(np.float32))
    if "convex" in kinds:
        clouds, cooked, _, _ = load_hull_library()
        return Scene(default_header(**hdr), add_bin(a, half_size=float(max(nx, nz) * pitch / 2 + 1.0)), clouds, cooked)
    return Scene(default_header(**hdr), add_bin(a, half_size=float(max(nx, nz) * pitch / 2 + 1.0)))


def env_piles(n_envs=3, nx=5, ny=4, nz=5, half_extent=0.25, gap=0.001, env_pitch=12.0, seed=4, **hdr):
    """Environments that each hold a dense lattice pile of boxes (no walls): many more pairs and constraints per environment than
    bodies, so the environment path runs with oversize constraint lists / rows streamed through global scratch."""
    rng = np.random.RandomState(seed)
    per = nx * ny * nz
    a = _new_actors(n_envs * per)
    he = np.float32(half_extent)
    pitch = np.float32(2 * half_extent + gap)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    for e in range(n_envs):
        sl = slice(e * per, (e + 1) * per)
        jit = rng.uniform(-gap * 0.4, gap * 0.4, size=(per, 2)).astype(np.float32)
        a["pos"][sl, 0] = np.float32(e * env_pitch) + ix.ravel().astype(np.float32) * pitch + jit[:, 0]
        a["pos"][sl, 1] = he + iy.ravel().astype(np.float32) * pitch + np.float32(gap)
        a["pos"][sl, 2] = iz.ravel().astype(np.float32) * pitch + jit[:, 1]
        a["envId"][sl] = e
    set_box(a, np.arange(len(a)), np.array([he, he, he], dtype=np.float32))
    return Scene(default_header(**hdr), add_ground_plane(a))


LOCK_LINEAR_X, LOCK_LINEAR_Y, LOCK_LINEAR_Z, LOCK_ANGULAR_X, LOCK_ANGULAR_Y, LOCK_ANGULAR_Z = 1, 2, 4, 8, 16, 32   # PxRigidDynamicLockFlag


def set_lock_flags(actors, idx, lock):
    """PxRigidDynamic::setRigidDynamicLockFlags: the six lock bits live in bits 8..13 of the record's flags."""
    actors["flags"][idx] = (actors["flags"][idx] & np.uint32(0xFF)) | (np.uint32(lock) << np.uint32(8))


def locked_primitives(n=12, seed=3, kinds=("sphere", "capsule"), **hdr):
    """Spheres / capsules dropped in a column with assorted PxRigidDynamicLockFlags (planar motion, no rotation, one free axis ...)."""
    sc = mixed_primitives(n=n, seed=seed, kinds=kinds, **hdr)
    sc.actors["angVel"][1:] = np.random.RandomState(seed).uniform(-2, 2, (n, 3)).astype(np.float32)
    combos = [LOCK_LINEAR_Z | LOCK_ANGULAR_X | LOCK_ANGULAR_Y, LOCK_ANGULAR_X | LOCK_ANGULAR_Y | LOCK_ANGULAR_Z, LOCK_LINEAR_X, 0,
              LOCK_LINEAR_X | LOCK_LINEAR_Z, LOCK_ANGULAR_Z, LOCK_LINEAR_Y | LOCK_ANGULAR_Y, 0]
    for i in range(n):
        set_lock_flags(sc.actors, 1 + i, combos[i % len(combos)])
    return sc


def locked_stacks(**hdr):
    """Jittered box stacks in which some boxes may only move vertically / may not rotate."""
    sc = box_stacks(n_stacks=3, height=5, half_extent=0.25, spacing=1.0, jitter=0.02, **hdr)
    for i, lock in ((2, LOCK_LINEAR_X | LOCK_LINEAR_Z), (4, LOCK_ANGULAR_X | LOCK_ANGULAR_Y | LOCK_ANGULAR_Z), (7, LOCK_LINEAR_X), (9, LOCK_ANGULAR_Y), (12, LOCK_LINEAR_X | LOCK_LINEAR_Z | LOCK_ANGULAR_X | LOCK_ANGULAR_Z)):
        set_lock_flags(sc.actors, i, lock)
    return sc


def capsules_on_boxes(n_boxes=6, per_box=3, seed=9, **hdr):
    """Capsules (and a few spheres) dropped onto a row of boxes that never touch each other: every other box is static and tilted, the
    rest are dynamic and rest flat on the ground.  Exercises the capsule-box GJK path (face, edge and corner contacts, manifold recycling)
    without tumbling box-box pairs."""
    rng = np.random.RandomState(seed)
    nb = n_boxes
    a = _new_actors(nb + nb * per_box)
    for b in range(nb):
        he = np.array([rng.uniform(0.25, 0.4), rng.uniform(0.12, 0.25), rng.uniform(0.25, 0.4)], dtype=np.float32)
        set_box(a, np.array([b]), he)
        a["pos"][b] = (1.6 * b, he[1], 0.0)
        if b % 2 == 1:   # static, tilted
            a["flags"][b] = 0
            a["mass"][b] = 0; a["inertia"][b] = 0
            ang = rng.uniform(-0.35, 0.35)
            a["quat"][b] = normalize_quat_f32([0.0, 0.0, np.sin(ang / 2), np.cos(ang / 2)])
            a["pos"][b, 1] = 0.45
        for k in range(per_box):
            i = nb + b * per_box + k
            if k == per_box - 1 and b % 3 == 0:
                set_sphere(a, i, rng.uniform(0.1, 0.18))
            else:
                set_capsule(a, i, rng.uniform(0.08, 0.15), rng.uniform(0.1, 0.3))
            a["pos"][i] = (1.6 * b + rng.uniform(-0.2, 0.2), 1.0 + 0.5 * k, rng.uniform(-0.2, 0.2))
            a["quat"][i] = random_unit_quats(rng, 1)[0]
            a["angVel"][i] = rng.uniform(-2, 2, 3)
    return Scene(default_header(**hdr), add_ground_plane(a))


def capsules_into_boxes(seed=2, speed=14.0, **hdr):
    """capsules_on_boxes with the capsules fired downwards at `speed` m/s and a few spawned overlapping their box: the core segment ends up
    inside the box, which sends pcmContactCapsuleBox through the EPA penetration query."""
    sc = capsules_on_boxes(seed=seed, **hdr)
    cap = np.nonzero(sc.actors["geomType"] == GEOM_CAPSULE)[0]
    sc.actors["linVel"][cap, 1] = -np.float32(speed)
    for j, i in enumerate(cap[::4]):
        b = 1 + (i - 1 - 6) // 3          # the box under this capsule (capsules_on_boxes layout: 6 boxes, 3 droppers per box)
        sc.actors["pos"][i] = sc.actors["pos"][b] + np.array([0.05 * j, 0.1, -0.03 * j], dtype=np.float32)
    return sc


def test_forces(n_dynamic, seed=1, blocks=7):
    """(blocks, n_dynamic, 6) force xyz + torque xyz cycle for the eFORCE / eTORQUE tests: every third body and one block stay force-free."""
    rng = np.random.RandomState(seed)
    f = np.zeros((blocks, n_dynamic, 6), np.float32)
    f[:, :, :3] = rng.uniform(-3, 3, (blocks, n_dynamic, 3))
    f[:, :, 3:] = rng.uniform(-0.3, 0.3, (blocks, n_dynamic, 3))
    f[:, ::3] = 0
    f[2] = 0
    return f


def hulls_and_spheres(n=10, n_hulls=3, seed=11, cook=None, **hdr):
    """Spheres dropped onto convex hulls that rest on the ground plane (hulls spaced apart so that no hull touches another hull):
    exercises pcmContactSphereConvex (GJK with the hull support mapping) next to plane-convex and sphere-plane."""
    rng = np.random.RandomState(seed)
    hulls = [random_hull_points(rng, int(rng.randint(12, 21)), 0.3) for _ in range(n_hulls)]
    nh = n // 2
    a = _new_actors(n)
    for i in range(nh):
        set_convex(a, i, i % n_hulls)
        a["pos"][i] = (1.6 * i, 0.35, 0.0)
    a["quat"][:nh] = random_unit_quats(rng, nh)
    for k in range(n - nh):
        i = nh + k
        set_sphere(a, i, rng.uniform(0.08, 0.16))
        a["pos"][i] = (1.6 * (k % nh) + rng.uniform(-0.12, 0.12), 0.9 + 0.4 * (k // nh), rng.uniform(-0.12, 0.12))
        a["linVel"][i] = (0.0, -rng.uniform(0.0, 3.0), 0.0)
    sc = Scene(default_header(**hdr), add_ground_plane(a), hulls)
    return cook(sc) if cook else sc


def spheres_into_hulls(seed=11, speed=14.0, cook=None, **hdr):
    """hulls_and_spheres with the spheres fired downwards and a few spawned inside a hull: pcmContactSphereConvex through EPA."""
    sc = hulls_and_spheres(seed=seed, n=12, cook=cook, **hdr)
    sph = np.nonzero(sc.actors["geomType"] == GEOM_SPHERE)[0]
    sc.actors["linVel"][sph, 1] = -np.float32(speed)
    inside = sph[::3]
    sc.actors["pos"][inside] = sc.actors["pos"][1 + (np.arange(len(inside)) % 6)] + np.array([0.02, 0.05, 0.01], np.float32)
    return sc


def hulls_and_capsules(n=12, n_hulls=3, seed=21, speed=0.0, cook=None, **hdr):
    """Capsules dropped onto convex hulls resting on the ground plane (hulls spaced apart): pcmContactCapsuleConvex -- GJK over the hull
    support mapping, face (ray) and edge-edge contacts against the witness polygon.  speed > 0 fires the capsules downwards (EPA)."""
    rng = np.random.RandomState(seed)
    hulls = [random_hull_points(rng, int(rng.randint(12, 21)), 0.3) for _ in range(n_hulls)]
    nh = n // 2
    a = _new_actors(n)
    for i in range(nh):
        set_convex(a, i, i % n_hulls)
        a["pos"][i] = (1.6 * i, 0.35, 0.0)
    a["quat"] = random_unit_quats(rng, n)
    for k in range(n - nh):
        i = nh + k
        set_capsule(a, i, rng.uniform(0.06, 0.12), rng.uniform(0.1, 0.25))
        a["pos"][i] = (1.6 * (k % nh) + rng.uniform(-0.1, 0.1), 0.9 + 0.45 * (k // nh), rng.uniform(-0.1, 0.1))
        a["angVel"][i] = rng.uniform(-2, 2, 3)
        a["linVel"][i] = (0.0, -speed, 0.0)
    sc = Scene(default_header(**hdr), add_ground_plane(a), hulls)
    return cook(sc) if cook else sc


def hull_pile(n=10, n_hulls=3, seed=31, kinds=("convex",), spread=0.35, cook=None, hull_points=(12, 21), **hdr):
    """Convex hulls (optionally mixed with boxes / spheres / capsules) dropped in a loose column onto the ground plane: hull-hull and
    box-hull contacts (GJK / EPA point + polygon clipping of the witness faces), BASELINE config 3's pair types at small size."""
    rng = np.random.RandomState(seed)
    hulls = [random_hull_points(rng, int(rng.randint(*hull_points)), 0.25) for _ in range(n_hulls)]
    a = _new_actors(n)
    a["pos"][:, 0] = rng.uniform(-spread, spread, n)
    a["pos"][:, 1] = 0.5 + 0.5 * np.arange(n)
    a["pos"][:, 2] = rng.uniform(-spread, spread, n)
    a["quat"] = random_unit_quats(rng, n)
    for i in range(n):
        k = kinds[i % len(kinds)]
        if k == "convex":
            set_convex(a, i, i % n_hulls)
        elif k == "box":
            set_box(a, np.array([i]), np.array([rng.uniform(0.12, 0.25), rng.uniform(0.12, 0.25), rng.uniform(0.12, 0.25)], dtype=np.float32))
        elif k == "sphere":
            set_sphere(a, i, rng.uniform(0.1, 0.2))
        else:
            set_capsule(a, i, rng.uniform(0.08, 0.15), rng.uniform(0.1, 0.3))
    sc = Scene(default_header(**hdr), add_ground_plane(a), hulls)
    return cook(sc) if cook else sc


def env_hulls(n_envs=6, per_env=7, seed=8, env_pitch=6.0, **hdr):
    """Environments that each hold a small heap of library hulls, spheres, capsules and boxes (environment ids set, shared ground plane):
    every pair type of a10 on the environment path."""
    rng = np.random.RandomState(seed)
    clouds, cooked, unit_mass, unit_inertia = load_hull_library()
    a = _new_actors(n_envs * per_env)
    for e in range(n_envs):
        ex, ez = (e % 3) * env_pitch, (e // 3) * env_pitch
        for k in range(per_env):
            i = e * per_env + k
            kind = (e + k) % 4
            if kind in (0, 2):
                hi = int(rng.randint(0, len(clouds)))
                set_convex(a, i, hi)
                a["mass"][i] = np.float32(10.0) * unit_mass[hi]; a["inertia"][i] = np.float32(10.0) * unit_inertia[hi]
            elif kind == 1:
                set_sphere(a, i, rng.uniform(0.1, 0.18))
            else:
                set_capsule(a, i, rng.uniform(0.07, 0.12), rng.uniform(0.1, 0.25)) if k % 2 else set_box(a, np.array([i]), np.array([0.15, 0.12, 0.2], dtype=np.float32))
            a["pos"][i] = (ex + rng.uniform(-0.25, 0.25), 0.4 + 0.45 * k, ez + rng.uniform(-0.25, 0.25))
            a["envId"][i] = e
    a["quat"] = random_unit_quats(rng, len(a))
    return Scene(default_header(**hdr), add_ground_plane(a), clouds, cooked)


def material_mix(n_boxes=12, n_spheres=8, seed=5, **hdr):
    """a11: a material table with every combine mode and eDISABLE_FRICTION.  Boxes slide on the ground plane with an initial lateral velocity
    (combined friction decides where they stop), pairs of stacked boxes carry different materials, spheres are dropped from 1.5 m (combined
    restitution decides the bounce).  Bodies are far enough apart to be independent islands."""
    rng = np.random.RandomState(seed)
    mats = make_materials([(0.5, 0.5, 0.6, COMBINE_AVERAGE, COMBINE_AVERAGE), (0.9, 0.8, 0.1, COMBINE_MAX, COMBINE_MIN), (0.2, 0.1, 0.0, COMBINE_MULTIPLY, COMBINE_MULTIPLY),
                           (0.6, 0.4, 0.3, COMBINE_MIN, COMBINE_MAX, True), (0.3, 0.7, 0.9, COMBINE_AVERAGE, COMBINE_MAX)])
    n = 2 * n_boxes + n_spheres
    a = _new_actors(n)
    he = np.float32(0.25)
    for i in range(n_boxes):          # lower box slides, upper box rides on it
        x = np.float32(3.0 * i)
        a["pos"][2 * i] = (x, he, 0.0); a["pos"][2 * i + 1] = (x, 3 * he, 0.0)
        a["linVel"][2 * i] = (1.5, 0.0, 0.8); a["linVel"][2 * i + 1] = (1.5, 0.0, 0.8)
        a["materialIndex"][2 * i] = i % len(mats); a["materialIndex"][2 * i + 1] = (i // 2 + 1) % len(mats)
    set_box(a, np.arange(2 * n_boxes), np.array([he, he, he], dtype=np.float32))
    for k in range(n_spheres):
        j = 2 * n_boxes + k
        a["pos"][j] = (np.float32(3.0 * k), 1.5, 4.0)
        a["materialIndex"][j] = (k * 2 + 1) % len(mats)
    set_sphere(a, np.arange(2 * n_boxes, n), np.float32(0.2))
    sc = Scene(default_header(**hdr), add_ground_plane(a), materials=mats)
    sc.actors["materialIndex"][0] = 4     # the plane has its own material
    return sc


def _rand_quat(rng, n=None):
    q = rng.normal(size=(4,) if n is None else (n, 4))
    return (q / np.linalg.norm(q, axis=-1, keepdims=True)).astype(np.float32)


def _qmul(a, b):   # float64 helpers for BUILDING scenes (inputs, not parity arithmetic)
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + bw * ax + ay * bz - by * az, aw * by + bw * ay + az * bx - bz * ax, aw * bz + bw * az + ax * by - bx * ay, aw * bw - ax * bx - ay * by - az * bz])


def _qrot(q, v):
    qv, qw = np.asarray(q[:3], np.float64), float(q[3])
    return v + 2.0 * np.cross(qv, np.cross(qv, v) + qw * v)


def local_pose_mix(n_stacks=4, height=3, n_loose=10, seed=9, **hdr):
    """a1 with local poses (PxShape::setLocalPose, PxRigidBody::setCMassLocalPose): stacks of boxes whose SHAPES line up while the actor frames and the
    centre-of-mass frames are offset and rotated per body (so the stacks lean and topple: pairs come and go), spheres / capsules with offset shapes
    dropped next to them, one static box with a shape offset.  Inertia tensors are given in the body (centre-of-mass) frame."""
    rng = np.random.RandomState(seed)
    nb = n_stacks * height
    n = nb + n_loose + 1
    a = _new_actors(n)
    lp = identity_local_poses(n)
    he = np.float32(0.25)
    set_box(a, np.arange(nb), np.array([he, he * 0.8, he * 1.2], dtype=np.float32))
    kinds = []
    for k in range(n_loose):
        j = nb + k
        if k % 2:
            set_capsule(a, j, 0.12, 0.2); kinds.append("capsule")
        else:
            set_sphere(a, j, 0.18); kinds.append("sphere")
    shape_world = []
    for i in range(nb):
        s, l = divmod(i, height)
        shape_world.append((np.array([1.5 * s, he * 0.8 * (2 * l + 1) + 0.002 * l, 0.0]), np.array([0, 0, 0, 1.0])))
    for k in range(n_loose):
        shape_world.append((np.array([1.5 * (k % n_stacks) + 0.6, 1.0 + 0.45 * (k // n_stacks), 0.5 * ((k % 3) - 1)]), _rand_quat(rng).astype(np.float64)))
    for i, (wp, wq) in enumerate(shape_world):
        # shape2Actor: offset up to 15 cm, any rotation (every third body keeps an identity shape pose, every fourth an identity CoM pose)
        sp = rng.uniform(-0.15, 0.15, 3) if i % 3 else np.zeros(3)
        sq = _rand_quat(rng).astype(np.float64) if i % 3 else np.array([0, 0, 0, 1.0])
        # actor pose = shape world pose * shape2Actor^-1
        sq_inv = np.array([-sq[0], -sq[1], -sq[2], sq[3]])
        aq = _qmul(wq, sq_inv)
        ap = wp - _qrot(aq, sp)
        a["pos"][i] = ap.astype(np.float32); a["quat"][i] = (aq / np.linalg.norm(aq)).astype(np.float32)
        lp["shapeP"][i] = sp.astype(np.float32); lp["shapeQ"][i] = sq.astype(np.float32)
        if i % 4:
            lp["bodyP"][i] = (sp + rng.uniform(-0.06, 0.06, 3)).astype(np.float32)   # centre of mass near the shape centre, not at it
            lp["bodyQ"][i] = _rand_quat(rng) if i % 2 else np.array([0, 0, 0, 1], np.float32)
            a["inertia"][i] *= np.array([1.0, 1.3, 0.8], np.float32)                   # distinct principal moments (body frame)
        a["angVel"][i] = rng.uniform(-0.5, 0.5, 3).astype(np.float32) if i >= nb else 0.0
    # static box with an offset shape: a ledge some loose bodies land on
    j = n - 1
    set_box(a, np.arange(j, j + 1), np.array([0.6, 0.1, 0.6], dtype=np.float32))
    a["flags"][j] = 0; a["mass"][j] = 0; a["inertia"][j] = 0
    a["pos"][j] = (2.0, 0.3, 1.2)
    lp["shapeP"][j] = (0.3, 0.1, -0.2); lp["shapeQ"][j] = _rand_quat(np.random.RandomState(seed + 1)) * np.float32(1.0)
    lp["shapeQ"][j] = np.array([0.0, 0.0, 0.08715574, 0.9961947], np.float32)   # 10 degrees about z
    actors = add_ground_plane(a)
    lp = np.concatenate([identity_local_poses(1), lp])
    return Scene(default_header(**hdr), actors, local_poses=lp)


def filter_groups_mix(n=6, seed=17, **hdr):
    """f1: PxDefaultSimulationFilterShader.  Collision groups (PxSetGroupCollisionFlag): "ghost" boxes (group 2) are dropped onto resting "solid" boxes (group 1)
    with the 1-2 flag cleared -- they fall through the solids and land on the ground (group 0).  Groups masks (PxSetFilterOps AND / AND / AND, constants all ones,
    PxSetFilterBool(true): a pair collides only when the two masks share a bit): spheres with mask 1 are dropped onto spheres with mask 2 and pass through them,
    both kinds collide with the boxes and the ground (mask 3)."""
    rng = np.random.RandomState(seed)
    nb = 2 * n
    a = _new_actors(nb + 2 * n)
    he = np.float32(0.25)
    set_box(a, np.arange(nb), np.array([he, he, he], dtype=np.float32))
    set_sphere(a, np.arange(nb, nb + 2 * n), np.float32(0.15))
    fd = np.zeros((nb + 2 * n + 1, 4), np.uint32)   # row 0: the ground plane
    fd[0] = (0, 0, 3, 0)
    for i in range(n):
        x = np.float32(1.2 * i)
        a["pos"][i] = (x, he, 0.0)                                   # solid, resting
        a["pos"][n + i] = (x + np.float32(rng.uniform(-0.05, 0.05)), 1.0 + 0.1 * i, np.float32(rng.uniform(-0.05, 0.05)))   # ghost, dropped on it
        fd[1 + i] = (1, 0, 3, 0); fd[1 + n + i] = (2, 0, 3, 0)
        a["pos"][nb + i] = (x, 0.15, 1.5)                            # sphere with mask 2 resting on the ground
        a["pos"][nb + n + i] = (x + np.float32(rng.uniform(-0.02, 0.02)), 0.9 + 0.15 * i, 1.5 + np.float32(rng.uniform(-0.02, 0.02)))   # sphere with mask 1 dropped on it
        fd[1 + nb + i] = (0, 0, 2, 0); fd[1 + nb + n + i] = (0, 0, 1, 0)
    # two more solids stacked on the first two solids: same group, they do collide
    cfg = default_filter_config()
    set_group_collision_flag(cfg, 1, 2, False)
    cfg["ops"][:] = (FILTER_AND, FILTER_AND, FILTER_AND); cfg["filterBool"] = 1; cfg["constants"][:] = 0xffffffff
    return Scene(default_header(**hdr), add_ground_plane(a), filter_config=cfg, filter_data=fd)


def shape_offsets_mix(seed=29, **hdr):
    """PxShape::setContactOffset / setRestOffset per shape: stacks and loose primitives whose shapes carry different contact offsets (0.01 .. 0.06: pairs enter the
    narrowphase at different distances) and rest offsets (-0.01 .. 0.03: bodies come to rest hovering above / sunk into each other by the sum of the two rest offsets)."""
    rng = np.random.RandomState(seed)
    sc = mixed_primitives(n=14, seed=seed, kinds=("box", "sphere", "capsule", "box"), **hdr)
    n = len(sc.actors)
    so = np.zeros((n, 2), np.float32)
    so[:, 0] = rng.uniform(0.03, 0.06, n); so[:, 1] = rng.uniform(-0.01, 0.025, n)
    so[0] = (0.02, 0.01)   # the ground plane
    return Scene(sc.header, sc.actors, shape_offsets=so)


ACTOR_KINEMATIC = 2   # with ACTOR_DYNAMIC: PxRigidBodyFlag::eKINEMATIC (oracle/scene_format.h)


def kinematic_mix(n_envs=0, env_pitch=12.0, **hdr):
    """Kinematic bodies (PxRigidBodyFlag::eKINEMATIC, moved with setKinematicTarget): a conveyor platform that carries three boxes sideways, a lift that moves a
    box and a sphere up and down, a paddle that rotates through a row of boxes and capsules standing on the ground, a kinematic that never gets a target (stands
    still, carries a box) and one that passes through another kinematic and through the ground plane (no pairs: kinematic-kinematic and kinematic-static are
    filtered).  n_envs > 0: the same group once per environment (environment ids, config-2 style); the ground plane is shared."""
    groups = max(1, n_envs)
    per = 5 + 3 + 2 + 6 + 1
    a = _new_actors(groups * per)
    he = np.float32(0.25)
    for g in range(groups):
        o = g * per
        ox = np.float32(env_pitch * g)
        k = np.arange(o, o + 5)
        a["pos"][k[0]] = (ox + 0.0, 1.0, 0.0); a["dims"][k[0], :3] = (1.5, 0.1, 1.0)         # conveyor
        a["pos"][k[1]] = (ox + 0.0, 0.6, 4.0); a["dims"][k[1], :3] = (0.8, 0.1, 0.8)         # lift
        a["pos"][k[2]] = (ox + 4.0, 0.3, 0.0); a["dims"][k[2], :3] = (1.2, 0.3, 0.1)         # paddle (sweeps about the y axis)
        a["pos"][k[3]] = (ox + 4.0, 0.8, 4.0); a["dims"][k[3], :3] = (0.6, 0.1, 0.6)         # no target
        a["pos"][k[4]] = (ox + 4.0, 0.6, 4.0); a["dims"][k[4], :3] = (0.2, 0.2, 0.2)         # crosses the one above and the ground
        b = np.arange(o + 5, o + 8)                                                            # boxes on the conveyor
        for i, j in enumerate(b):
            a["pos"][j] = (ox - 0.8 + 0.8 * i, 1.1 + he, -0.3 + 0.3 * i)
        lift = np.arange(o + 8, o + 10)
        a["pos"][lift[0]] = (ox - 0.3, 0.7 + he, 4.0); a["pos"][lift[1]] = (ox + 0.4, 0.7 + 0.2, 4.2)
        row = np.arange(o + 10, o + 16)
        for i, j in enumerate(row):
            a["pos"][j] = (ox + 4.0 + 0.9 * np.cos(i), he if i % 2 == 0 else 0.2, 0.9 * np.sin(i) + (0.45 if i % 2 else -0.45))
        top = o + 16
        a["pos"][top] = (ox + 4.0, 0.9 + he, 4.0)
        boxes = np.concatenate([b, lift[:1], row[0::2], [top]])
        dims_k = a["dims"][k, :3].copy()
        set_box(a, k, dims_k); set_box(a, boxes, np.array([he, he, he], dtype=np.float32))
        set_sphere(a, lift[1:], np.float32(0.2))
        set_capsule(a, row[1::2], np.float32(0.2), np.float32(0.25))
        a["flags"][k] = ACTOR_DYNAMIC | ACTOR_KINEMATIC
        if n_envs:
            a["envId"][o:o + per] = g
    return Scene(default_header(**hdr), add_ground_plane(a))


def kinematic_targets(scene, steps):
    """(steps, n_kinematic, 7) PxTransform (q.xyzw, p.xyz) handed to setKinematicTarget before each step of kinematic_mix; NaN rows = no target that step.
    Conveyor: 1 m/s along x, stops after 2/3 of the run; lift: sinusoidal; paddle: 1.5 rad/s about y; fourth: never; fifth: sinks through the ground and rises again."""
    kin = np.nonzero((scene.actors["flags"] & ACTOR_KINEMATIC) != 0)[0]
    dt = float(scene.header["dt"])
    out = np.full((steps, len(kin), 7), np.nan, np.float64)
    for n, ai in enumerate(kin):
        p0 = scene.actors["pos"][ai].astype(np.float64); role = n % 5
        for t in range(steps):
            time = (t + 1) * dt
            q = np.array([0.0, 0.0, 0.0, 1.0]); p = p0.copy()
            if role == 0:
                if t >= (2 * steps) // 3:
                    continue
                p[0] += 1.0 * time
            elif role == 1:
                p[1] += 0.4 * np.sin(3.0 * time)
            elif role == 2:
                ang = 1.5 * time
                q = np.array([0.0, np.sin(ang / 2), 0.0, np.cos(ang / 2)])
            elif role == 3:
                continue
            else:
                p[1] += 0.8 * np.sin(2.0 * time) - 0.3
            out[t, n] = np.concatenate([q, p])
    return out.astype(np.float32)


AGGREGATE_SELF_COLLISION = 0x80000000   # ActorRec.aggregate bit 31


def aggregates_mix(n_envs=0, env_pitch=8.0, seed=41, **hdr):
    """PxAggregate membership (first half).  Per group: aggregate A = four boxes stacked with 40 % overlap and NO self collisions (they fall through
    each other and come to rest side by side / on the free boxes), aggregate B = a column with 2 cm overlaps WITH self collisions (the overlaps are pushed apart, the column stands), three free boxes
    and a sphere that collide with every member.  n_envs > 0: one group per environment."""
    rng = np.random.RandomState(seed)
    groups = max(1, n_envs)
    per = 4 + 4 + 4
    a = _new_actors(groups * per)
    he = np.float32(0.25)
    for g in range(groups):
        o = g * per; ox = np.float32(env_pitch * g)
        for k in range(4):
            a["pos"][o + k] = (ox + 0.05 * k, he + 0.3 * k, 0.03 * k)
            a["pos"][o + 4 + k] = (ox + 2.0 + 0.05 * k, he + 0.48 * k, -0.03 * k)
        a["aggregate"][o:o + 4] = 2 * g + 1
        a["aggregate"][o + 4:o + 8] = (2 * g + 2) | AGGREGATE_SELF_COLLISION
        for k in range(3):
            a["pos"][o + 8 + k] = (ox + 0.1 + 0.9 * k + rng.uniform(-0.05, 0.05), 2.2 + 0.6 * k, rng.uniform(-0.1, 0.1))
        a["pos"][o + 11] = (ox + 0.3, 4.2, 0.1)
        if n_envs:
            a["envId"][o:o + per] = g
    idx = np.arange(groups * per)
    set_box(a, idx[(idx % per) != 11], np.array([he, he, he], dtype=np.float32))
    set_sphere(a, idx[(idx % per) == 11], np.float32(0.3))
    return Scene(default_header(**hdr), add_ground_plane(a))


ACTOR_DISABLE_GRAVITY, ACTOR_GYROSCOPIC = 4, 8   # PxActorFlag::eDISABLE_GRAVITY, PxRigidBodyFlag::eENABLE_GYROSCOPIC_FORCES (oracle/scene_format.h)


def body_flags_mix(n=16, seed=53, **hdr):
    """Per-body flags of the pre-integration stage (a12): flat boxes and capsules (three different principal inertias) thrown up spinning about a non-principal axis --
    every other one with eENABLE_GYROSCOPIC_FORCES -- and a row of bodies with eDISABLE_GRAVITY that float until something lands on them (some drift sideways)."""
    rng = np.random.RandomState(seed)
    a = _new_actors(n)
    q = random_unit_quats(rng, n)
    for i in range(n):
        a["pos"][i] = (1.2 * (i % 8), 1.0 + 1.3 * (i // 8), 0.0)
        a["quat"][i] = q[i]
    flat = np.arange(0, n, 2); caps = np.arange(1, n, 2)
    set_box(a, flat, np.array([0.1, 0.25, 0.4], dtype=np.float32))
    set_capsule(a, caps, np.float32(0.12), np.float32(0.3))
    a["angVel"][:] = rng.uniform(-6, 6, (n, 3)); a["linVel"][:, 1] = rng.uniform(0.5, 2.5, n)
    a["angDamping"][:] = 0.0
    a["flags"][np.arange(0, n, 4)] |= ACTOR_GYROSCOPIC; a["flags"][np.arange(1, n, 4)] |= ACTOR_GYROSCOPIC
    upper = np.arange(n // 2, n)
    a["flags"][upper[::2]] |= ACTOR_DISABLE_GRAVITY
    a["linVel"][upper[::2], 1] = 0.0; a["linVel"][upper[::4], 0] = 0.4
    return Scene(default_header(**hdr), add_ground_plane(a))
