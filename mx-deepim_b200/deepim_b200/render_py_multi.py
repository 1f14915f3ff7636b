"""Render_Py -- call-compatible stand-in for lib/render_glumpy/render_py_multi.py:Render_Py backed by
the CUDA rasteriser (dim_render) instead of glumpy/OpenGL.

    rm = Render_Py(meshes, classes, K, width=640, height=480, zNear=0.25, zFar=6.0)
    bgr, depth = rm.render(cls_idx, R, t, r_type="mat", K=K)     # (480,640,3) f32 BGR*255, (480,640) f32 m

The reference takes a model directory and loads textured.obj / texture_map.png with glumpy
(l.54-80); disk formats are out of scope here, so meshes are passed as
deepim_b200.synth.Mesh objects (verts, uvs, faces, texture already flipped as l.76 does)."""
import numpy as np
import torch

from .context import Context


class Render_Py:
    def __init__(self, meshes, classes, K, width=640, height=480, zNear=0.25, zFar=6.0, device=0, ctx=None):
        self.width, self.height, self.zNear, self.zFar = width, height, zNear, zFar
        self.K = np.asarray(K, np.float32)
        self.classes = list(classes)
        self._own = ctx is None
        self.ctx = ctx or Context(device, max_batch=1, height=height, width=width, max_classes=len(meshes),
                                  max_verts=max(len(m.verts) for m in meshes),
                                  max_faces=max(len(m.faces) for m in meshes))
        if self._own:
            for i, m in enumerate(meshes):
                self.ctx.upload_mesh(i, m)

    def render(self, cls_idx, r, t, r_type="quat", K=None):
        if r_type == "quat":
            w, x, y, z = [float(v) for v in r]
            n = w * w + x * x + y * y + z * z
            s = 2.0 / n
            R = np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                          [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                          [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]])
        elif r_type == "mat":
            R = np.asarray(r)
        else:
            raise ValueError("unknown r_type %r" % (r_type,))
        pose = np.hstack([R, np.asarray(t).reshape(3, 1)]).astype(np.float32)[None]
        dev = self.ctx.device
        out = self.ctx.render(torch.tensor([cls_idx], dtype=torch.int32, device=dev), torch.from_numpy(pose).to(dev),
                              self.K if K is None else np.asarray(K, np.float32), self.zNear, self.zFar,
                              trunc_u8=False, want=("bgr", "depth"))
        return out["bgr"][0].cpu().numpy(), out["depth"][0, 0].cpu().numpy()

    def __del__(self):
        if getattr(self, "_own", False):
            try:
                self.ctx.close()
            except Exception:
                pass
