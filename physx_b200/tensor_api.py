"""Tensor front end in the style of the reference's ovphysx bindings (ovphysx/python/ovphysx/api.py `TensorBinding.read /
write`, types.py `TensorType`; wire formats of ovphysx/include/ovphysx/ovphysx.h).

    view = TensorBinding(scene, TensorType.RIGID_BODY_POSE)        # [N, 7] = (p.xyz, q.xyzw)
    view.read(t)                                                   # t: anything that speaks DLPack (torch / cupy / jax ...) or a numpy array
    view.write(t, indices=idx)                                     # or mask=

CUDA tensors never leave the device: the engine's tensor kernels (pxb_tensor_read_device / pxb_tensor_write_device) read / write the caller's buffer
directly, stream-ordered on the scene stream, and the caller's current stream is ordered around the call.  Host tensors go through a device
staging tensor.  torch is used for DLPack import and stream plumbing only."""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from . import engine as _engine


class TensorType(IntEnum):       # values of ovphysx TensorType (types.py:24-110), rigid-body subset
    RIGID_BODY_POSE = 1          # [N, 7] world pose (px,py,pz,qx,qy,qz,qw)
    RIGID_BODY_VELOCITY = 2      # [N, 6] linear + angular velocity
    RIGID_BODY_MASS = 3          # [N]    read-only here
    RIGID_BODY_INV_MASS = 7      # [N]    read-only
    RIGID_BODY_FORCE = 50        # [N, 3] write-only, applied by the next step
    RIGID_BODY_WRENCH = 51       # [N, 9] write-only: force, torque, application point (world)


_COLS = {TensorType.RIGID_BODY_POSE: 7, TensorType.RIGID_BODY_VELOCITY: 6, TensorType.RIGID_BODY_MASS: 1, TensorType.RIGID_BODY_INV_MASS: 1,
         TensorType.RIGID_BODY_FORCE: 3, TensorType.RIGID_BODY_WRENCH: 9}
_READABLE = {TensorType.RIGID_BODY_POSE, TensorType.RIGID_BODY_VELOCITY, TensorType.RIGID_BODY_MASS, TensorType.RIGID_BODY_INV_MASS}
_WRITABLE = {TensorType.RIGID_BODY_POSE, TensorType.RIGID_BODY_VELOCITY, TensorType.RIGID_BODY_FORCE, TensorType.RIGID_BODY_WRENCH}


class TensorBinding:
    """One tensor view of a scene's dynamic bodies (ovphysx `TensorBinding`): `shape`, `count`, `read(tensor)`, `write(tensor, indices=, mask=)`."""

    def __init__(self, scene: _engine.Scene, tensor_type: int):
        self.scene, self.tensor_type = scene, TensorType(tensor_type)
        c = _COLS[self.tensor_type]
        self.shape = (scene.num_dynamic,) if c == 1 else (scene.num_dynamic, c)

    @property
    def count(self) -> int:
        return self.scene.num_dynamic

    @property
    def ndim(self) -> int:
        return len(self.shape)

    # ---- helpers
    def _torch(self):
        import torch
        return torch

    def _as_torch(self, obj):
        torch = self._torch()
        if isinstance(obj, torch.Tensor):
            return obj
        if isinstance(obj, np.ndarray):
            return torch.from_numpy(obj)
        return torch.from_dlpack(obj)       # any DLPack producer

    def _device(self):
        return self._torch().device("cuda", self.scene.device_index)

    def _scene_stream(self):
        torch = self._torch()
        return torch.cuda.ExternalStream(self.scene.stream(), device=self._device())

    def _check(self, t, rows):
        cols = _COLS[self.tensor_type]
        want = (rows,) if cols == 1 else (rows, cols)
        if tuple(t.shape) != want or str(t.dtype) != "torch.float32" or not t.is_contiguous():
            raise ValueError(f"{self.tensor_type.name}: expected a contiguous float32 tensor of shape {want}, got {tuple(t.shape)} {t.dtype}")

    def _indices(self, indices, mask):
        torch = self._torch()
        if mask is not None:
            m = self._as_torch(mask).to(self._device())
            indices = torch.nonzero(m.reshape(-1), as_tuple=False).reshape(-1)
        if indices is None:
            return None
        return self._as_torch(indices).to(device=self._device(), dtype=torch.int32).contiguous()

    # ---- ovphysx API
    def read(self, tensor) -> None:
        """Fills `tensor` ([N, cols] float32; CUDA or host) with the current values of every dynamic body."""
        if self.tensor_type not in _READABLE:
            raise ValueError(f"{self.tensor_type.name} is write-only")
        torch = self._torch()
        t = self._as_torch(tensor)
        self._check(t, self.count)
        dev = t if t.is_cuda else torch.empty(t.shape, dtype=torch.float32, device=self._device())
        st = self._scene_stream()
        st.wait_stream(torch.cuda.current_stream(self._device()))
        _engine._check(self.scene._lib, self.scene._lib.pxb_tensor_read_device(self.scene._h, int(self.tensor_type), dev.data_ptr(), None, self.count))
        torch.cuda.current_stream(self._device()).wait_stream(st)
        if not t.is_cuda:
            t.copy_(dev)

    def write(self, tensor, indices=None, mask=None) -> None:
        """Writes rows of `tensor` into the simulation: all bodies, the bodies listed in `indices`, or those selected by the boolean `mask`
        (rows of `tensor` then correspond to the selected bodies in ascending order)."""
        if self.tensor_type not in _WRITABLE:
            raise ValueError(f"{self.tensor_type.name} is read-only")
        torch = self._torch()
        idx = self._indices(indices, mask)
        rows = self.count if idx is None else int(idx.numel())
        t = self._as_torch(tensor)
        self._check(t, rows)
        dev = t if t.is_cuda else t.to(self._device())
        st = self._scene_stream()
        st.wait_stream(torch.cuda.current_stream(self._device()))
        _engine._check(self.scene._lib, self.scene._lib.pxb_tensor_write_device(self.scene._h, int(self.tensor_type), dev.data_ptr(), idx.data_ptr() if idx is not None else None, rows))
        torch.cuda.current_stream(self._device()).wait_stream(st)   # staging / index tensors are freed in current-stream order, i.e. after the kernel
                                                                    # (no record_stream on the scene's stream: it may be destroyed before the tensors are)


def get_contact_report(scene: _engine.Scene) -> dict:
    """Contact data of the last step in the spirit of ovphysx `PhysX.get_contact_report` / PxDirectGPUAPI::copyContactData (PxGpuContactPair):
    per touching pair the two actor indices, the contact normal (body1 -> body0), and per contact point position, separation and applied
    normal impulse; flat arrays with `start_indices` / `counts` per pair."""
    pairs, con = scene.getPairs(), scene.getContacts()
    cnt = con[:, 0].astype(np.int32)
    keep = cnt > 0
    pts = np.concatenate([con[i, 4:4 + 5 * k].reshape(k, 5) for i, k in enumerate(cnt) if k]) if keep.any() else np.zeros((0, 5), np.float32)
    counts = cnt[keep]
    return {"actor0": pairs[keep, 0], "actor1": pairs[keep, 1], "normals": con[keep, 1:4], "counts": counts,
            "start_indices": np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32) if len(counts) else np.zeros(0, np.int32),
            "positions": pts[:, :3], "separations": pts[:, 3], "impulses": pts[:, 4]}


class GpuContactData:
    """PxDirectGPUAPI::copyContactData through the C ABI (pxb_scene_copy_contact_data): the PxGpuContactPair records stay in a device tensor
    (`records`, uint8 [max_pairs, 80]; `count`, int32 [1]); `to_host()` follows the record pointers into the scene-owned PxContactPatch / PxContact /
    force / PxFrictionPatch streams and returns numpy views of everything (for tests and debugging: a learner reads the device memory directly).

        scene.enableContactData()                # before the step
        scene.step()
        cd = GpuContactData(scene, max_pairs)    # after fetchResults
        host = cd.to_host()
    """

    PATCH_DTYPE = np.dtype([("massModification", "<f4", 4), ("normal", "<f4", 3), ("restitution", "<f4"), ("dynamicFriction", "<f4"), ("staticFriction", "<f4"), ("damping", "<f4"),
                            ("startContactIndex", "<u2"), ("nbContacts", "u1"), ("materialFlags", "u1"), ("internalFlags", "<u2"), ("materialIndex0", "<u2"), ("materialIndex1", "<u2"),
                            ("pad", "<u2", 5)])   # PxContactPatch, PxContact.h:56-137
    FRICTION_DTYPE = np.dtype([("anchorPositions", "<f4", (2, 3)), ("anchorImpulses", "<f4", (2, 3)), ("anchorCount", "<u4")])   # PxFrictionPatch, PxContact.h:635-658

    def __init__(self, scene: _engine.Scene, max_pairs: int):
        import torch
        assert self.PATCH_DTYPE.itemsize == 64 and self.FRICTION_DTYPE.itemsize == 52 and _engine.Scene.GPU_CONTACT_PAIR_DTYPE.itemsize == 80
        dev = torch.device("cuda", scene.device_index)
        self.scene, self.max_pairs = scene, int(max_pairs)
        self.records = torch.zeros((self.max_pairs, 80), dtype=torch.uint8, device=dev)
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        scene.copyContactData(self.records.data_ptr(), self.count.data_ptr(), self.max_pairs)
        scene.sync()

    @staticmethod
    def _read(ptr: int, nbytes: int) -> bytes:
        from cuda.bindings import runtime as cudart
        buf = np.empty(max(nbytes, 1), np.uint8)
        if nbytes:
            err, = cudart.cudaMemcpy(buf.ctypes.data, int(ptr), nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            if int(err) != 0:
                raise RuntimeError(f"cudaMemcpy failed: {err}")
        return buf[:nbytes].tobytes()

    def to_host(self) -> dict:
        n_total = int(self.count.cpu()[0])
        n = min(n_total, self.max_pairs)
        rec = self.records[:n].cpu().numpy().reshape(-1).view(_engine.Scene.GPU_CONTACT_PAIR_DTYPE).copy()
        out = {"total_pairs": n_total, "records": rec}
        if n == 0:
            return out
        nc = int(rec["nbContacts"].astype(np.int64).sum())
        # the streams are compact and in record order, so one copy per stream starting at the first record's pointer covers all records
        out["patches"] = np.frombuffer(self._read(rec["contactPatches"][0], 64 * n), self.PATCH_DTYPE)
        out["points"] = np.frombuffer(self._read(rec["contactPoints"][0], 16 * nc), np.float32).reshape(nc, 4)
        out["forces"] = np.frombuffer(self._read(rec["contactForces"][0], 4 * nc), np.float32)
        out["friction"] = np.frombuffer(self._read(rec["frictionPatches"][0], 52 * n), self.FRICTION_DTYPE)
        out["start_indices"] = ((rec["contactPoints"] - rec["contactPoints"][0]) // 16).astype(np.int64)
        return out
