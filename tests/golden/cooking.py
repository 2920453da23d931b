"""Convex cooking for the test fixtures (TEST INFRASTRUCTURE): runs the UNMODIFIED reference cooking (oracle/_ref/ref_harness cook =
PxCreateConvexMesh) over a scene's hull point clouds.  Needs the build container (/root/reference); fixtures carry the cooked bytes so that the
GPU box never cooks.  `python tests/golden/cooking.py` regenerates physx_b200/data/hull_library.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from physx_b200 import scenes  # noqa: E402


def cook_hulls(scene, density=10.0):
    """Runs the reference's convex cooking (oracle/_ref/ref_harness cook: PxCreateConvexMesh, unmodified) over the scene's hull point clouds,
    attaches the cooked section and fills mass / inertia of the convex actors from the cooked mass information (mass = density * volume,
    diagonal of the inertia tensor, centre-of-mass frame = actor frame: both sides get the same explicit values).  Needs the build container
    (/root/reference); fixtures under tests/golden carry the cooked bytes so that the GPU box never cooks."""
    import os, subprocess, tempfile
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    with tempfile.TemporaryDirectory() as d:
        scene.cooked = b""
        scene.save(d + "/s.bin")
        subprocess.run([harness, "cook", d + "/s.bin", d + "/c.bin"], check=True, capture_output=True)
        scene.cooked = open(d + "/c.bin", "rb").read()
    ch = scene.cooked_hulls()
    for i in np.nonzero(scene.actors["geomType"] == scenes.GEOM_CONVEX)[0]:
        hdr = ch[int(scene.actors["hullIdx"][i])]["hdr"]
        if scene.actors["flags"][i] & scenes.ACTOR_DYNAMIC:
            scene.actors["mass"][i] = np.float32(density) * hdr["unitMass"]
            scene.actors["inertia"][i] = np.float32(density) * hdr["unitInertiaDiag"]
    return scene




def make_hull_library(path=os.path.join(ROOT, "physx_b200", "data", "hull_library.npz")):
    """16 hulls of 12-20 input points, r ~ 0.2, cooked by the reference"""
    rng = np.random.RandomState(1234)
    hulls = [scenes.random_hull_points(rng, int(rng.randint(12, 21)), 0.2) for _ in range(16)]
    a = scenes._new_actors(16)
    for i in range(16):
        scenes.set_convex(a, i, i)
    sc = cook_hulls(scenes.Scene(scenes.default_header(), scenes.add_ground_plane(a), hulls))
    ch = sc.cooked_hulls()
    np.savez_compressed(path, cooked=np.frombuffer(sc.cooked, np.uint8), clouds=np.array(hulls, dtype=object), unitMass=np.array([h["hdr"]["unitMass"] for h in ch], np.float32),
                        unitInertiaDiag=np.stack([h["hdr"]["unitInertiaDiag"] for h in ch]).astype(np.float32), nVerts=np.array([h["hdr"]["nVerts"] for h in ch]))


if __name__ == "__main__":
    make_hull_library()
