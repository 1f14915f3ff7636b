"""PoseRefiner: the batched, device-resident replacement of the per-instance test loop in
deepim/core/tester.py:284-485 (pred_eval's hot loop: predict -> RT_transform -> render ->
update_data_batch -> predict ...).  Also fixes the reference's batch=1 / single-GPU limitation
(tester.py:83): any number of instances, sharded over ranks."""
from __future__ import annotations

import numpy as np
import torch

from . import _capi as capi
from . import sharding, synth
from .context import Context


class PoseRefiner:
    def __init__(self, meshes, weights, K=synth.K_LINEMOD, device=0, max_batch=16, n_iter=4,
                 pixel_means_rgb=synth.PIXEL_MEANS_RGB, znear=synth.ZNEAR, zfar=synth.ZFAR, precision="bf16"):
        self.K = np.asarray(K, np.float32)
        self.n_iter, self.means, self.zn, self.zf = n_iter, np.asarray(pixel_means_rgb, np.float64), znear, zfar
        self.precision = capi.PREC_BF16X3 if precision == "bf16x3" else capi.PREC_BF16
        mv = max(len(m.verts) for m in meshes)
        mf = max(len(m.faces) for m in meshes)
        self.ctx = Context(device, max_batch=max_batch, max_classes=len(meshes), max_verts=mv, max_faces=mf)
        for i, m in enumerate(meshes):
            self.ctx.upload_mesh(i, m)
        self.ctx.load_weights(weights)
        self.max_batch = max_batch

    def refine(self, images_bgr_u8, cls_idx, poses_init, dist=None):
        """images_bgr_u8 [N,H,W,3] uint8 (cv2 layout), cls_idx [N] int, poses_init [N,3,4] float64 (host).
        Returns poses [n_iter,N,3,4] float64 (host).  With torch.distributed initialised each rank
        processes its contiguous slice and the poses are all-gathered."""
        n = len(cls_idx)
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None and dist.is_initialized() else (0, 1)
        lo, hi = sharding.shard_range(n, rank, world)
        out = np.zeros((self.n_iter, hi - lo, 3, 4), np.float64)
        cls32 = np.ascontiguousarray(cls_idx, np.int32)
        poses64 = np.ascontiguousarray(poses_init, np.float64)
        for a, b in sharding.chunks(lo, hi, self.max_batch):
            p, _ = self.ctx.refine_host(np.ascontiguousarray(images_bgr_u8[a:b]), cls32[a:b], poses64[a:b], self.K,
                                        self.n_iter, self.zn, self.zf, self.means, self.precision)
            out[:, a - lo:b - lo] = p
        return sharding.gather_results(out, n, axis=1, dist=dist, device=self.ctx.device if world > 1 else None)

    def close(self):
        self.ctx.close()
