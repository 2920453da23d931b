"""Env-partitioned multi-GPU stepping (BASELINE config 5).

Environments never interact (environment-id filter in the broadphase), so the path shards by environment:
rank r owns envs [r*E/G, (r+1)*E/G) in its own scene on its own GPU -- one process per GPU, no physics
coupling, no halo exchange.  The only collective is an NCCL all-gather of the per-env body-state tensors the
Direct GPU API exposes (pose 7 floats, linear and angular velocity 3 floats each), so that every rank (the
learner) sees the global [E*B, 7|3|3] tensors.  The reference has no multi-GPU mode (one PxScene <-> one
device, physx/source/simulationcontroller/src/ScScene.cpp:718), so this module is new work.

`torch.distributed` is plumbing only: the gather operates directly on the device buffers the engine's
get-kernels fill (no staging copy through the host).
"""
from __future__ import annotations

import numpy as np

from . import scenes as _scenes

STATE_COLS = {0: 7, 1: 3, 2: 3}  # RD_GLOBAL_POSE, RD_LINEAR_VELOCITY, RD_ANGULAR_VELOCITY


def env_range(n_envs: int, world_size: int, rank: int):
    """Contiguous env shard of `rank`; remainders go to the lowest ranks."""
    base, rem = divmod(n_envs, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_scene(scene: _scenes.Scene, world_size: int, rank: int) -> _scenes.Scene:
    """Sub-scene with the actors of this rank's environments plus every env-less (shared) actor, e.g. the
    ground plane.  Actor order is preserved, so local dynamic-body order = global order restricted to the shard.
    Environment ids are rebased to start at 0 on every rank (the engine builds one CTA / warp per id up to the largest one),
    and the convex hulls travel with the shard (point clouds and the cooked section; hull indices are unchanged), as do the per-actor
    sections of the scene format.  Aggregate ids stay global (an aggregate lives inside one environment)."""
    env = scene.actors["envId"]
    n_envs = int(env[env != _scenes.NO_ENV].max()) + 1 if np.any(env != _scenes.NO_ENV) else 0
    lo, hi = env_range(n_envs, world_size, rank)
    keep = (env == _scenes.NO_ENV) | ((env >= lo) & (env < hi))
    actors = scene.actors[keep].copy()
    own = actors["envId"] != _scenes.NO_ENV
    actors["envId"][own] -= np.uint32(lo)
    # per-actor sections travel with their actors (local poses, PxFilterData, per-shape offsets); kinematic flags and aggregate ids are fields of the actor records
    rows = lambda x: None if x is None else x[keep].copy()
    return _scenes.Scene(scene.header, actors, scene.hulls, scene.cooked, scene.materials, local_poses=rows(scene.local_poses), filter_config=scene.filter_config,
                         filter_data=rows(scene.filter_data), shape_offsets=rows(scene.shape_offsets))


def gather_layout(counts):
    """Row offsets of every rank's block in the gathered tensor."""
    off = np.concatenate([[0], np.cumsum(counts)])
    return [(int(off[i]), int(off[i + 1])) for i in range(len(counts))]


class StateGather:
    """All-gathers one Direct-GPU-API state tensor across ranks.

    `local_fill(dst_tensor)` must write this rank's [n_local, cols] float32 block into `dst_tensor`
    (a view of the send region); on the GPU that is pxb_get_rigid_dynamic_data_device writing straight
    into the NCCL send buffer.  Equal per-rank counts use all_gather_into_tensor (one NCCL call over
    NVLink/NVSwitch); ragged counts fall back to all_gather with per-rank tensors."""

    def __init__(self, dist, n_local: int, cols: int, device):
        import torch
        self.dist, self.torch = dist, torch
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(allc, cnt)
        self.counts = [int(c.item()) for c in allc]
        self.layout = gather_layout(self.counts)
        self.equal = len(set(self.counts)) == 1
        self.cols = cols
        self.global_tensor = torch.zeros((sum(self.counts), cols), dtype=torch.float32, device=device)
        lo, hi = self.layout[self.rank]
        self.local_view = self.global_tensor[lo:hi]

    def __call__(self, local_fill):
        local_fill(self.local_view)
        if self.equal:
            # in place: the send region is this rank's slice of the receive tensor (NCCL in-place all-gather)
            src = self.local_view if self.global_tensor.is_cuda else self.local_view.clone()
            self.dist.all_gather_into_tensor(self.global_tensor, src)
        else:
            # ragged shards: pad every block to the largest one, gather, then compact into the global tensor
            mx = max(self.counts)
            if not hasattr(self, "_padded"):
                self._padded = self.torch.zeros((self.world * mx, self.cols), dtype=self.torch.float32, device=self.global_tensor.device)
                self._send = self.torch.zeros((mx, self.cols), dtype=self.torch.float32, device=self.global_tensor.device)
            self._send[:self.counts[self.rank]].copy_(self.local_view)
            self.dist.all_gather_into_tensor(self._padded, self._send)
            for r, (lo, hi) in enumerate(self.layout):
                if r != self.rank:
                    self.global_tensor[lo:hi].copy_(self._padded[r * mx:r * mx + (hi - lo)])
        return self.global_tensor


class PipelinedStateGather:
    """Double-buffered all-gather of the packed state that overlaps the collective of step N with the kernels of
    step N+1 ("overlappable with the next step's BP/NP").

    Per step:  `pack(view)` enqueues the engine's pack kernel on the scene stream into this step's buffer, the
    collective runs on a dedicated communication stream ordered after it by an event, and the scene stream only
    waits for the collective that last used the buffer it is about to overwrite (two steps back).  On CPU tensors
    (gloo tests) the same rotation runs synchronously."""

    def __init__(self, dist, n_local: int, cols: int, device, scene_stream=None):
        import torch
        self.torch, self.dist = torch, dist
        self.bufs = [StateGather(dist, n_local, cols, device) for _ in range(2)]
        self.cuda = device.type == "cuda"
        self.scene_stream = scene_stream
        self.comm_stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.done = [None, None]
        self.k = 0

    def step(self, pack):
        """pack(view): fill this rank's block (stream-ordered on the scene stream).  Returns the buffer index used."""
        b = self.k & 1
        g = self.bufs[b]
        if not self.cuda:
            g(pack)
        else:
            t = self.torch
            if self.done[b] is not None:
                self.scene_stream.wait_event(self.done[b])      # gather k-2 read/wrote this buffer
            pack(g.local_view)
            ready = t.cuda.Event()
            ready.record(self.scene_stream)
            self.comm_stream.wait_event(ready)
            with t.cuda.stream(self.comm_stream):
                g(lambda view: None)
                self.done[b] = t.cuda.Event()
                self.done[b].record(self.comm_stream)
        self.k += 1
        return b

    def pre_step(self, consumer_stream=None):
        """Nothing to prepare: the exchange is enqueued by step() after the simulate call."""

    def latest(self):
        """Global tensor of the most recent step (caller must `wait()` or order its stream after `done`)."""
        return self.bufs[(self.k - 1) & 1].global_tensor

    def wait(self, stream=None):
        """Orders `stream` (default: the scene stream) after every outstanding collective."""
        if self.cuda:
            for ev in self.done:
                if ev is not None:
                    (stream or self.scene_stream).wait_event(ev)


class PeerStateGather:
    """All-gather of the packed state over NVLink PEER MEMORY instead of an NCCL kernel: every rank's global tensor lives in
    symmetric memory (torch.distributed._symmetric_memory), the engine's pack kernel writes this rank's block into its own copy
    and the block is then pushed into every peer's copy with device-to-device peer copies (copy engines: no SM is taken away
    from the next step's kernels, which the collective overlaps), followed by one symmetric-memory barrier.  Double-buffered
    like PipelinedStateGather; same interface.  Raises at construction when symmetric memory is not available (callers fall
    back to NCCL)."""

    def __init__(self, dist, n_local: int, cols: int, device, scene_stream, scene=None, mode: str = "kernel"):
        import torch
        import torch.distributed._symmetric_memory as symm
        self.scene, self.mode = scene, (mode if scene is not None else "copy")
        self.torch, self.dist = torch, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(allc, cnt)
        self.counts = [int(c.item()) for c in allc]
        self.layout = gather_layout(self.counts)
        total = sum(self.counts)
        self.scene_stream = scene_stream
        self.comm_stream = torch.cuda.Stream(device=device)
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(max(1, self.world - 1))]   # one per peer: concurrent copy engines
        self.bufs, self.handles, self.peer_views = [], [], []
        lo, hi = self.layout[self.rank]
        for _ in range(2):
            t = symm.empty((total, cols), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, dist.group.WORLD)
            self.bufs.append(t)
            self.handles.append(h)
            self.peer_views.append([h.get_buffer(p, (total, cols), torch.float32)[lo:hi] for p in range(self.world)])
        self.row_bytes = cols * 4
        if (n_local * self.row_bytes) % 16 or (lo * self.row_bytes) % 16 or self.world > 9:
            self.mode = "copy"   # the scatter kernel moves 16-byte words to at most 8 peers
        self.done = [None, None]
        self.k = 0
        dist.barrier()

    def step(self, pack):
        t = self.torch
        b = self.k & 1
        lo, hi = self.layout[self.rank]
        local = self.bufs[b][lo:hi]
        if self.done[b] is not None:
            self.scene_stream.wait_event(self.done[b])          # exchange k-2 used this buffer
        pack(local)                                             # engine pack kernel, stream-ordered on the scene stream
        ready = t.cuda.Event()
        ready.record(self.scene_stream)
        if self.mode == "kernel":
            # ONE launch: the engine's scatter kernel stores this rank's block into every peer's buffer (P2P stores over NVLink)
            self.comm_stream.wait_event(ready)
            ptrs = [v.data_ptr() for p, v in enumerate(self.peer_views[b]) if p != self.rank]
            self.scene.scatterToPeers(self.comm_stream.cuda_stream, local.data_ptr(), local.numel() * 4, ptrs)
            with t.cuda.stream(self.comm_stream):
                self.handles[b].barrier(channel=b)
                self.done[b] = t.cuda.Event()
                self.done[b].record(self.comm_stream)
            self.k += 1
            return b
        j = 0
        for p in range(self.world):                             # pushes to different peers run concurrently, one stream (copy engine) each
            if p == self.rank:
                continue
            cs = self.copy_streams[j]; j += 1
            cs.wait_event(ready)
            with t.cuda.stream(cs):
                self.peer_views[b][p].copy_(local, non_blocking=True)   # peer D2D copy over NVLink
                ev = t.cuda.Event(); ev.record(cs)
            self.comm_stream.wait_event(ev)
        self.comm_stream.wait_event(ready)
        with t.cuda.stream(self.comm_stream):
            self.handles[b].barrier(channel=b)                  # every rank's pushes into this buffer have landed
            self.done[b] = t.cuda.Event()
            self.done[b].record(self.comm_stream)
        self.k += 1
        return b

    def pre_step(self, consumer_stream=None):
        """Consumer release: a peer may push step k into the buffer this rank last read for step k-2 only after every rank has
        ENTERED step k-1, i.e. finished with it.  One symmetric-memory barrier on the communication stream, before this step's pushes."""
        t = self.torch
        ev = t.cuda.Event()
        ev.record(consumer_stream if consumer_stream is not None else self.scene_stream)
        self.comm_stream.wait_event(ev)
        with t.cuda.stream(self.comm_stream):
            self.handles[self.k & 1].barrier(channel=2 + (self.k & 1))

    def latest(self):
        return self.bufs[(self.k - 1) & 1]

    def wait(self, stream=None):
        for ev in self.done:
            if ev is not None:
                (stream or self.scene_stream).wait_event(ev)


class FusedStateGather:
    """The all-gather FUSED INTO THE STEP: every rank's global [sum(n), 13] tensor lives in symmetric memory, and the engine's integration
    epilogue (pxb_scene_set_state_export) stores each environment's packed block straight into this rank's rows of EVERY rank's tensor with P2P
    stores over NVLink -- no pack kernel, no copies, no collective call; the transfer overlaps the solve environment by environment.
    Synchronisation is two flags per rank in a symmetric signal pad, written with pxb_peer_signal (one tiny kernel) and awaited with
    pxb_peer_wait:
      * data[r]    = k  once rank r's block of step k has landed everywhere (raised right after r's step on its scene stream);
      * release[r] = k  once rank r has ENTERED step k, i.e. is done reading the tensor of step k-1.
    Buffers rotate over `nbuf` slots; a rank may start writing step k (slot k % nbuf) only when every release flag has reached k - nbuf + 1, so
    a peer can never overwrite a tensor this rank is still reading (the consumer-release handshake the copy-based exchange lacked).
    Usage per step:  pre_step() -> scene.simulate() -> step()  [-> wait() / latest() by the consumer]."""

    def __init__(self, dist, n_local: int, cols: int, device, scene_stream, scene, nbuf: int = 2):
        import torch
        import torch.distributed._symmetric_memory as symm
        assert cols == 13, "the fused export writes the packed 13-float state record"
        self.torch, self.dist, self.scene, self.scene_stream, self.nbuf = torch, dist, scene, scene_stream, nbuf
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 9:
            raise RuntimeError("at most 9 export targets")
        cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(allc, cnt)
        self.counts = [int(c.item()) for c in allc]
        self.layout = gather_layout(self.counts)
        total = sum(self.counts)
        self.bufs, self.handles, self.ptrs = [], [], []
        for _ in range(nbuf):
            t = symm.empty((total, cols), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, dist.group.WORLD)
            self.bufs.append(t)
            self.handles.append(h)
            self.ptrs.append([h.get_buffer(p, (total, cols), torch.float32).data_ptr() for p in range(self.world)])
        self.flags = symm.empty((2 * self.world,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.fh = symm.rendezvous(self.flags, dist.group.WORLD)
        peer_flags = [self.fh.get_buffer(p, (2 * self.world,), torch.int32).data_ptr() for p in range(self.world)]
        self.data_flag_ptrs = [f + 4 * self.rank for f in peer_flags]                      # my slot in everybody's pad
        self.release_flag_ptrs = [f + 4 * (self.world + self.rank) for f in peer_flags]
        self.k = 0
        torch.cuda.synchronize(device)
        dist.barrier()

    def pre_step(self, consumer_stream=None):
        """Before scene.simulate(): announce that this rank is done with the previous tensor, wait until the slot of this step is free on every
        rank, and point the step's export at it."""
        k = self.k + 1
        st = self.scene_stream
        if consumer_stream is not None:
            ev = self.torch.cuda.Event()
            ev.record(consumer_stream)
            st.wait_event(ev)
        self.scene.peerSignal(st.cuda_stream, self.release_flag_ptrs, k)
        self.scene.peerWait(st.cuda_stream, self.flags.data_ptr() + 4 * self.world, self.world, k - self.nbuf + 1)
        self.scene.setStateExport(self.ptrs[k % self.nbuf], self.layout[self.rank][0])

    def step(self, pack=None):
        """After scene.simulate(): raise this rank's data flag on every rank (the step's P2P stores are complete at its kernel boundary)."""
        self.k += 1
        self.scene.peerSignal(self.scene_stream.cuda_stream, self.data_flag_ptrs, self.k)
        return self.k % self.nbuf

    def close(self):
        """Stops exporting (the scene keeps stepping without writing into the exchange buffers)."""
        self.scene.setStateExport(())

    def latest(self):
        return self.bufs[self.k % self.nbuf]

    def wait(self, stream=None):
        """Orders `stream` (default: the scene stream) after the arrival of every rank's block of the latest step."""
        st = stream or self.scene_stream
        self.scene.peerWait(st.cuda_stream, self.flags.data_ptr(), self.world, self.k)


class GraphPeerGather:
    """The per-step all-gather as ONE CUDA-graph replay per step and no pack kernel:
      * the step's integration epilogue writes this rank's packed block into its rows of its OWN symmetric-memory tensor
        (pxb_scene_set_state_export with one local target: nothing crosses NVLink from the solve kernel);
      * the pushes into the peers (one copy-engine peer copy per peer, concurrent) and the closing symmetric-memory barrier are captured once
        per buffer into a CUDA graph on the communication stream and replayed: the host enqueues a step's exchange with three calls instead
        of ~10 per peer (at N = 8 the Python-side enqueue of the copy-based exchange took about as long as the step itself).
    Consumer release: graph launches are serialised on the communication stream, so the pushes of exchange k+2 (same buffer as k) start only
    after the barrier of exchange k+1 has passed, i.e. after EVERY rank has finished step k+1 -- and a rank starts step k+1 only after its
    consumer is done with the tensor of step k (program / stream order).  No peer can overwrite a tensor that is still being read."""

    def __init__(self, dist, n_local: int, cols: int, device, scene_stream, scene):
        import torch
        import torch.distributed._symmetric_memory as symm
        assert cols == 13
        self.torch, self.dist, self.scene, self.scene_stream = torch, dist, scene, scene_stream
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        cnt = torch.tensor([n_local], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(allc, cnt)
        self.counts = [int(c.item()) for c in allc]
        self.layout = gather_layout(self.counts)
        total = sum(self.counts)
        lo, hi = self.layout[self.rank]
        self.comm_stream = torch.cuda.Stream(device=device)
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(max(1, self.world - 1))]
        self.bufs, self.handles, self.peer_views, self.graphs = [], [], [], [None, None]
        for _ in range(2):
            t = symm.empty((total, cols), dtype=torch.float32, device=device)
            h = symm.rendezvous(t, dist.group.WORLD)
            self.bufs.append(t)
            self.handles.append(h)
            self.peer_views.append([h.get_buffer(p, (total, cols), torch.float32)[lo:hi] for p in range(self.world)])
        self.done = [None, None]
        self.k = 0
        dist.barrier()
        for b in range(2):      # one eager exchange per buffer (lazy initialisation inside the barrier / copy paths), then capture
            with torch.cuda.stream(self.comm_stream):
                self._exchange(b)
        torch.cuda.synchronize(device)
        dist.barrier()
        for b in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.comm_stream, capture_error_mode="thread_local"):
                self._exchange(b)
            self.graphs[b] = g
        torch.cuda.synchronize(device)
        dist.barrier()

    def _exchange(self, b):
        """pushes of this rank's block into every peer's copy of buffer b (fork over the copy streams) + barrier; runs on / is captured from the communication stream"""
        t = self.torch
        lo, hi = self.layout[self.rank]
        local = self.bufs[b][lo:hi]
        j = 0
        for p in range(self.world):
            if p == self.rank:
                continue
            cs = self.copy_streams[j]; j += 1
            cs.wait_stream(self.comm_stream)
            with t.cuda.stream(cs):
                self.peer_views[b][p].copy_(local, non_blocking=True)
            self.comm_stream.wait_stream(cs)
        self.handles[b].barrier(channel=b)

    def pre_step(self, consumer_stream=None):
        b = self.k & 1
        if consumer_stream is not None:
            self.scene_stream.wait_stream(consumer_stream)
        if self.done[b] is not None:
            self.scene_stream.wait_event(self.done[b])          # exchange k-2 has finished reading this rank's rows of buffer b
        self.scene.setStateExport([self.bufs[b].data_ptr()], self.layout[self.rank][0])

    def step(self, pack=None):
        t = self.torch
        b = self.k & 1
        ready = t.cuda.Event()
        ready.record(self.scene_stream)
        self.comm_stream.wait_event(ready)
        with t.cuda.stream(self.comm_stream):
            self.graphs[b].replay()
            self.done[b] = t.cuda.Event()
            self.done[b].record(self.comm_stream)
        self.k += 1
        return b

    def close(self):
        self.scene.setStateExport(())

    def latest(self):
        return self.bufs[(self.k - 1) & 1]

    def wait(self, stream=None):
        for ev in self.done:
            if ev is not None:
                (stream or self.scene_stream).wait_event(ev)


def make_state_gather(dist, n_local: int, cols: int, device, scene_stream, kind: str = "auto", scene=None):
    """kind: 'fused' (the step's integration epilogue stores into every rank's symmetric-memory tensor: FusedStateGather), 'peer-copy'
    (concurrent copy-engine peer copies into symmetric memory), 'peer' (one scatter kernel with P2P stores), 'nccl' (all_gather_into_tensor)
    'graph' (GraphPeerGather: local export + one CUDA-graph replay of the peer pushes and the barrier) or 'auto' (peer-copy when symmetric
    memory is available, else NCCL)."""
    # measured at N = 2 (config 2 per GPU): copy engines 0.3115 ms/step, NCCL 0.3161, fused export 0.3418 (its P2P stores and flag kernels sit on
    # the step's critical path; the copy engines do not) -> 'auto' stays with the copy engines, 'fused' is opt-in
    if kind == "fused" and device.type == "cuda" and scene is not None and cols == 13:
        try:
            g = FusedStateGather(dist, n_local, cols, device, scene_stream, scene)
            return g, "P2P stores from the step's integration epilogue into every rank's symmetric-memory tensor (fused export) + per-rank flags"
        except Exception as e:  # pragma: no cover - depends on the platform
            if kind == "fused":
                raise
            kind = "auto-copy"
    # measured (config 2 per GPU): N = 2: copy engines 0.3157 ms/step, graph 0.3257; N = 8: copy engines 0.3281, graph 0.4168 (the captured
    # peer copies no longer run concurrently on the copy engines) -> 'auto' stays with the eager copy-engine exchange, 'graph' is opt-in
    if kind == "graph" and device.type == "cuda" and scene is not None and cols == 13:
        try:
            g = GraphPeerGather(dist, n_local, cols, device, scene_stream, scene)
            return g, "state export into the local symmetric-memory tensor by the step itself + one CUDA-graph replay per step (concurrent copy-engine peer pushes + barrier)"
        except Exception as e:  # pragma: no cover - depends on the platform
            if kind == "graph":
                raise
            kind = "auto-copy"
    if kind in ("auto", "auto-copy", "peer", "peer-copy") and device.type == "cuda":
        try:
            mode = "kernel" if kind == "peer" else "copy"   # auto: copy engines (measured faster at N=8: they take no SM from the next step)
            g = PeerStateGather(dist, n_local, cols, device, scene_stream, scene=scene, mode=mode)
            return g, ("one scatter kernel with P2P stores into every peer's symmetric-memory buffer" if g.mode == "kernel" else "peer-memory copies (symmetric memory, copy engines)") + " + barrier"
        except Exception as e:  # pragma: no cover - depends on the platform
            if kind in ("peer", "peer-copy"):
                raise
            reason = f" (peer memory unavailable: {type(e).__name__})"
    else:
        reason = ""
    return PipelinedStateGather(dist, n_local, cols, device, scene_stream=scene_stream), "NCCL all_gather_into_tensor" + reason


class EnvShardedScene:
    """One rank's scene of an env-partitioned job + the state all-gather."""

    def __init__(self, full_scene: _scenes.Scene, dist=None, device_index: int = 0):
        import torch
        from . import engine
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.local_scene_desc = shard_scene(full_scene, self.world, self.rank) if self.world > 1 else full_scene
        self.scene = engine.Scene(self.local_scene_desc, device=device_index)
        self.device = torch.device("cuda", device_index)
        self.stream = torch.cuda.ExternalStream(self.scene.stream(), device=self.device)
        self.gathers = {}
        if dist is not None and self.world > 1:
            for t, cols in STATE_COLS.items():
                self.gathers[t] = StateGather(dist, self.scene.num_dynamic, cols, self.device)

    def step(self):
        self.scene.step()

    def all_gather_state(self, data_type: int):
        """Global [sum(n_dyn), cols] tensor of one state type; identical on every rank afterwards."""
        import torch
        if not self.gathers:
            out = torch.empty((self.scene.num_dynamic, STATE_COLS[data_type]), dtype=torch.float32, device=self.device)
            self.scene.getRigidDynamicDataDevice(data_type, out.data_ptr(), self.scene.num_dynamic)
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            return out
        g = self.gathers[data_type]

        def fill(view):
            # the engine's gather kernel writes straight into the NCCL send region on the scene stream
            self.scene.getRigidDynamicDataDevice(data_type, view.data_ptr(), self.scene.num_dynamic)
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        return g(fill)
