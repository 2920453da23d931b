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
