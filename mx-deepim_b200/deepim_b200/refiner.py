"""PoseRefiner: the batched, device-resident replacement of the per-instance test loop in
deepim/core/tester.py:284-485 (pred_eval's hot loop: predict -> RT_transform -> render ->
update_data_batch -> predict ...).  Also lifts the reference's batch=1 / single-GPU limitation
(tester.py:83): any number of instances, sharded over ranks.

Two device contexts on two CUDA streams are used round-robin, so the H2D copy of batch k+1 overlaps
the kernels of batch k (instances are independent; nothing else is shared but read-only weights)."""
from __future__ import annotations

import numpy as np
import torch

from . import _capi as capi
from . import sharding, synth
from .context import Context


class PoseRefiner:
    def __init__(self, meshes, weights, K=synth.K_LINEMOD, device=0, max_batch=16, n_iter=4,
                 pixel_means_rgb=synth.PIXEL_MEANS_RGB, znear=synth.ZNEAR, zfar=synth.ZFAR, precision="fp16",
                 n_slots=2):
        self.K = np.asarray(K, np.float32)
        self.n_iter, self.means, self.zn, self.zf = n_iter, np.asarray(pixel_means_rgb, np.float64), znear, zfar
        self.precision = capi.precision_id(precision)
        mv = max(len(m.verts) for m in meshes)
        mf = max(len(m.faces) for m in meshes)
        self.max_batch = max_batch
        self.slots = []
        for _ in range(n_slots):
            ctx = Context(device, max_batch=max_batch, max_classes=len(meshes), max_verts=mv, max_faces=mf)
            for i, m in enumerate(meshes):
                ctx.upload_mesh(i, m)
            ctx.load_weights(weights)
            self.slots.append({
                "ctx": ctx, "stream": torch.cuda.Stream(device=ctx.device), "busy": False, "n": 0,
                "poses": torch.empty((n_iter, max_batch, 3, 4), dtype=torch.float64).pin_memory(),
                "se3": torch.empty((n_iter, max_batch, 7), dtype=torch.float32).pin_memory(),
                "status": torch.zeros((min(n_iter, 8) * max_batch,), dtype=torch.int32).pin_memory(),
                "img": None, "cls": None, "pose": None,
            })
        self.ctx = self.slots[0]["ctx"]
        self._next = 0

    # ------------------------------------------------------------------ pipelined submit / result
    def _pinned(self, slot, key, arr, dtype):
        t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))
        if t.dtype != dtype:
            t = t.to(dtype)
        if t.is_pinned() and t.is_contiguous():
            return t
        buf = slot[key]
        if buf is None or buf.shape[1:] != t.shape[1:] or buf.shape[0] < t.shape[0]:
            buf = torch.empty((self.max_batch,) + tuple(t.shape[1:]), dtype=dtype).pin_memory()
            slot[key] = buf
        buf[: t.shape[0]].copy_(t)
        return buf[: t.shape[0]]

    def submit(self, images_bgr_u8, cls_idx, poses_init):
        """Enqueue one batch (<= max_batch instances, host arrays; pinned torch tensors are used in place).
        Returns a ticket for result().  At most len(slots) batches may be in flight."""
        i = self._next
        slot = self.slots[i]
        if slot["busy"]:
            raise RuntimeError("PoseRefiner: slot still in flight; call result() first")
        n = len(cls_idx)
        if n > self.max_batch:
            raise ValueError("batch larger than max_batch")
        img = self._pinned(slot, "img", images_bgr_u8, torch.uint8)
        cls = self._pinned(slot, "cls", cls_idx, torch.int32)
        pose = self._pinned(slot, "pose", poses_init, torch.float64)
        with torch.cuda.stream(slot["stream"]):
            slot["ctx"].refine_host(img, cls, pose, self.K, self.n_iter, self.zn, self.zf, self.means, self.precision,
                                    poses_out=slot["poses"], se3_out=slot["se3"], sync=False)
            slot["ctx"].refine_status(n, self.n_iter, out=slot["status"], sync=False)
        slot["busy"], slot["n"] = True, n
        self._next = (i + 1) % len(self.slots)
        return i

    def result(self, ticket, strict=True):
        """Block until the batch is done; returns poses [n_iter, n, 3, 4] float64 (numpy copy).
        The per-iteration device status is checked: an instance whose object left the view frustum (empty rendered mask;
        the reference crashes in ZoomMask there) or whose class index is invalid raises DeepIMError when strict, else the
        flags are left in `self.last_status` ([n_iter, n] int32) for the caller."""
        slot = self.slots[ticket]
        slot["stream"].synchronize()
        slot["busy"] = False
        n, B = slot["n"], self.max_batch
        ni = min(self.n_iter, 8)
        self.last_status = slot["status"][: ni * n].view(ni, n).numpy().copy()
        if strict and self.last_status.any():
            bad = sorted(set(np.nonzero(self.last_status)[1].tolist()))
            raise capi.DeepIMError("PoseRefiner: instances %s of the batch have a non-zero device status (bit 0: empty rendered "
                                   "mask -- object left the frustum; bit 1: bad class index): %s"
                                   % (bad, self.last_status[:, bad].tolist()))
        # refine_host packs outputs densely as [n_iter, n, ...] at the start of the pinned buffer
        p = slot["poses"].view(-1)[: self.n_iter * n * 12].view(self.n_iter, n, 3, 4)
        return p.numpy().copy()

    # ------------------------------------------------------------------------------ convenience
    def refine(self, images_bgr_u8, cls_idx, poses_init, dist=None):
        """images_bgr_u8 [N,H,W,3] uint8 (cv2 layout), cls_idx [N] int, poses_init [N,3,4] float64 (host).
        Returns poses [n_iter,N,3,4] float64 (host).  With torch.distributed initialised each rank
        processes its contiguous slice and the poses are all-gathered."""
        n = len(cls_idx)
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.is_initialized() else (0, 1)
        lo, hi = sharding.shard_range(n, rank, world)
        out = np.zeros((self.n_iter, hi - lo, 3, 4), np.float64)
        pending = []
        for a, b in sharding.chunks(lo, hi, self.max_batch):
            if len(pending) == len(self.slots):
                t, (pa, pb) = pending.pop(0)
                out[:, pa - lo:pb - lo] = self.result(t)
            pending.append((self.submit(images_bgr_u8[a:b], cls_idx[a:b], poses_init[a:b]), (a, b)))
        for t, (pa, pb) in pending:
            out[:, pa - lo:pb - lo] = self.result(t)
        return sharding.gather_results(out, n, axis=1, dist=dist, device=self.ctx.device if world > 1 else None)

    def close(self):
        for s in self.slots:
            s["ctx"].close()
