import numpy as np

from physx_b200 import scenes


def test_scene_roundtrip(tmp_path):
    s = scenes.env_grid_stacks(n_envs=4)
    p = tmp_path / "s.bin"
    s.save(str(p))
    t = scenes.Scene.load(str(p))
    assert np.array_equal(s.actors, t.actors) and s.header == t.header
    assert s.n_dynamic == 4 * 64 and len(s.actors) == 257


def test_config2_shape():
    s = scenes.env_grid_stacks(n_envs=4096)
    assert s.n_dynamic == 262144
    a = s.actors[1:]
    assert a["envId"].max() == 4095 and (a["geomType"] == scenes.GEOM_BOX).all()
    # mass/inertia closed forms for a box of density 10
    assert np.allclose(a["mass"], 10 * 8 * 0.25 ** 3)
    assert np.allclose(a["inertia"][:, 0], a["mass"] / 3 * (2 * 0.25 ** 2))


def test_quaternions_are_normalisation_fixed_points():
    q = scenes.PLANE_UP_Y_QUAT
    assert np.array_equal(scenes.normalize_quat_f32(q), q)
    t = scenes.tumbling_boxes(n=5)
    for qq in t.actors["quat"]:
        assert np.array_equal(scenes.normalize_quat_f32(qq), qq)


def test_big_hull_fixture_carries_consistent_hill_climbing_data(tmp_path):
    """Hulls of more than 32 vertices carry Gu::BigConvexRawData in the cooked section (oracle/scene_format.h): cube-map samples and
    neighbours are vertex indices, valencies tile the adjacency list, adjacency is symmetric and matches the hull's edge count, and the
    section survives a save / load round trip."""
    import util
    z, sc = util.load_golden("big_hull_pile")
    big = [h for h in sc.cooked_hulls() if "samples" in h]
    assert len(big) >= 2
    for h in big:
        nv = int(h["hdr"]["nVerts"]); subdiv = int(h["hdr"]["reserved"][0]) & 0xffff; n_adj = int(h["hdr"]["reserved"][0]) >> 16
        assert 32 < nv <= 64 and int(h["hdr"]["nPolys"]) <= 64          # the reference's GPU-compatible limits
        assert len(h["samples"]) == 6 * subdiv * subdiv and h["samples"].max() < nv
        cnt, off = h["valencies"][:, 0].astype(int), h["valencies"][:, 1].astype(int)
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(cnt)[:-1]])) and cnt.sum() == n_adj == 2 * int(h["hdr"]["nEdges"])
        nb = [set(h["adjacentVerts"][off[i]:off[i] + cnt[i]].tolist()) for i in range(nv)]
        assert all(i in nb[j] for i in range(nv) for j in nb[i]) and all(i not in nb[i] for i in range(nv))
    p = tmp_path / "big.bin"
    sc.save(str(p))
    assert scenes.Scene.load(str(p)).cooked == sc.cooked
