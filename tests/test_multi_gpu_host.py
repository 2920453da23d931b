"""CPU tests (gloo, world_size 2) of the env-sharding and state all-gather host logic (physx_b200/multi_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from physx_b200 import multi_gpu, scenes


def test_env_ranges_partition_exactly():
    for n_envs in (1, 7, 8, 4096, 32768):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = multi_gpu.env_range(n_envs, world, r)
                got.extend(range(lo, hi))
            assert got == list(range(n_envs))


def test_shard_scene_keeps_shared_actors_and_order():
    sc = scenes.env_grid_stacks(n_envs=6)
    seen = []
    for r in range(4):
        sh = multi_gpu.shard_scene(sc, 4, r)
        lo, hi = multi_gpu.env_range(6, 4, r)
        assert sh.actors[0]["geomType"] == scenes.GEOM_PLANE           # the shared ground plane is in every shard
        env = sh.actors["envId"][1:]
        assert np.all(np.diff(env.astype(np.int64)) >= 0)
        assert env.min() == 0 and env.max() == hi - lo - 1              # environment ids are rebased to 0 on every rank
        seen.extend((np.unique(env) + lo).tolist())
        assert sh.n_dynamic == len(env)
        # same bodies as the global scene's slice, in the same order
        glob = sc.actors[(sc.actors["envId"] >= lo) & (sc.actors["envId"] < hi)]
        assert np.array_equal(glob["pos"], sh.actors["pos"][1:])
    assert sorted(seen) == list(range(6))


def test_shard_scene_carries_the_cooked_hulls():
    """A hull scene shards with its cooked section (the shard used to drop it, so every convex actor was refused)."""
    sc = scenes.env_hulls()
    assert len(sc.cooked) > 0
    for r in range(2):
        sh = multi_gpu.shard_scene(sc, 2, r)
        assert sh.cooked == sc.cooked and len(sh.hulls) == len(sc.hulls)
        assert sh.actors["hullIdx"][sh.actors["geomType"] == scenes.GEOM_CONVEX].max() < len(sh.hulls)
        rt = scenes.Scene.load(_save_tmp(sh))
        assert rt.cooked == sc.cooked and np.array_equal(rt.actors, sh.actors)


def _save_tmp(sc):
    import tempfile
    f = tempfile.NamedTemporaryFile(suffix=".bin", delete=False)
    f.write(sc.tobytes())
    f.close()
    return f.name


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, counts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_local = counts[rank]
        for cols in (7, 3):
            g = multi_gpu.StateGather(dist, n_local, cols, torch.device("cpu"))

            def fill(view, cols=cols):
                view.copy_(torch.arange(n_local * cols, dtype=torch.float32).reshape(n_local, cols) + 1000.0 * rank)
            out = g(fill)
            exp = torch.cat([torch.arange(c * cols, dtype=torch.float32).reshape(c, cols) + 1000.0 * r for r, c in enumerate(counts)])
            assert g.layout == multi_gpu.gather_layout(counts)
            assert torch.equal(out, exp), f"rank {rank} cols {cols}"
        # double-buffered gather: buffers alternate, every step's global tensor is complete and the previous one stays intact
        pg = multi_gpu.PipelinedStateGather(dist, n_local, 13, torch.device("cpu"))
        prev = None
        for step in range(3):
            b = pg.step(lambda view, step=step: view.fill_(100.0 * step + rank))
            assert b == step % 2
            exp = torch.cat([torch.full((c, 13), 100.0 * step + r) for r, c in enumerate(counts)])
            assert torch.equal(pg.latest(), exp), f"rank {rank} step {step}"
            if prev is not None:
                assert torch.equal(pg.bufs[(step - 1) % 2].global_tensor, prev)
            prev = exp
        pg.wait()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("counts", [(5, 5), (4, 7)])
def test_state_all_gather_world2_gloo(counts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_scene_carries_per_actor_sections_and_flags():
    """local poses, PxFilterData, per-shape offsets are sliced with the actors; kinematic flags and aggregate ids are fields of the actor records"""
    from physx_b200 import scenes
    sc = scenes.kinematic_mix(n_envs=4)
    n = len(sc.actors)
    lp = scenes.identity_local_poses(n); lp["shapeP"][:, 0] = np.arange(n)
    fd = np.arange(4 * n, dtype=np.uint32).reshape(n, 4)
    so = np.stack([0.02 + 0.001 * np.arange(n), np.zeros(n)], axis=1).astype(np.float32)
    sc.actors["aggregate"][1:4] = 7
    full = scenes.Scene(sc.header, sc.actors, local_poses=lp, filter_config=scenes.default_filter_config(), filter_data=fd, shape_offsets=so)
    seen = 0
    for rank in range(2):
        sh = multi_gpu.shard_scene(full, 2, rank)
        keep = np.nonzero((full.actors["envId"] == scenes.NO_ENV) | ((full.actors["envId"] >= 2 * rank) & (full.actors["envId"] < 2 * rank + 2)))[0]
        assert len(sh.actors) == len(keep)
        assert np.array_equal(sh.local_poses["shapeP"][:, 0], lp["shapeP"][keep, 0]) and np.array_equal(sh.filter_data, fd[keep]) and np.array_equal(sh.shape_offsets, so[keep])
        assert np.array_equal(sh.actors["flags"], full.actors["flags"][keep]) and np.array_equal(sh.actors["aggregate"], full.actors["aggregate"][keep])
        assert np.count_nonzero(sh.actors["flags"] & scenes.ACTOR_KINEMATIC) == 10 and sh.actors["envId"][sh.actors["envId"] != scenes.NO_ENV].max() == 1
        rt = scenes.Scene.load  # round trip through the scene format keeps the sections
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".bin") as f:
            f.write(sh.tobytes()); f.flush()
            back = rt(f.name)
        assert np.array_equal(back.shape_offsets, sh.shape_offsets) and np.array_equal(back.filter_data, sh.filter_data)
        seen += int((sh.actors["aggregate"] == 7).sum())
    assert seen == 3
