"""Context: Python owner of a dim_ctx (one per CUDA device) and thin tensor-level wrappers.

PyTorch is used only for device memory and streams (torch.Tensor.data_ptr, torch.cuda.current_stream);
all compute goes through libdeepim_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi as capi
from ._capi import check, farr, lib

WEIGHT_ORDER = ["flow_conv1", "conv2", "conv3", "conv3_1", "conv4", "conv4_1", "conv5", "conv5_1", "conv6",
                "conv6_1", "fc6", "fc7", "rot", "trans"]




def _p(t):
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("expected a contiguous CUDA tensor")
    return C.c_void_p(t.data_ptr())


def _chk(t, dtype, shape=None, name="tensor"):
    if t.dtype != dtype:
        raise TypeError("%s: expected dtype %s, got %s" % (name, dtype, t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t


class Context:
    def __init__(self, device=0, max_batch=16, height=480, width=640, max_classes=16, max_verts=60000,
                 max_faces=120000):
        if not torch.cuda.is_available():
            raise capi.DeepIMError("deepim_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.H, self.W, self.max_batch = height, width, max_batch
        h = C.c_void_p()
        check(lib.dim_ctx_create(device, max_batch, height, width, max_classes, max_verts, max_faces, C.byref(h)))
        self._h = h
        self.num_classes = 0

    def _stream(self):
        """torch's current stream OF THIS CONTEXT'S DEVICE (a process may hold contexts on several GPUs)"""
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            lib.dim_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------ setup
    def upload_mesh(self, cls_idx, mesh):
        v = np.ascontiguousarray(mesh.verts, np.float32)
        uv = np.ascontiguousarray(mesh.uvs, np.float32)
        f = np.ascontiguousarray(mesh.faces, np.int32)
        tex = np.ascontiguousarray(mesh.tex, np.uint8)
        check(lib.dim_mesh_upload(self._h, cls_idx, v.ctypes.data, uv.ctypes.data, len(v), f.ctypes.data, len(f),
                                  tex.ctypes.data, tex.shape[0], tex.shape[1]))
        self.num_classes = max(self.num_classes, cls_idx + 1)
        if getattr(mesh, "normals", None) is not None:
            self.upload_normals(cls_idx, mesh.normals)

    def upload_normals(self, cls_idx, normals):
        n = np.ascontiguousarray(normals, np.float32)
        check(lib.dim_mesh_upload_normals(self._h, cls_idx, n.ctypes.data, len(n)))

    def load_weights(self, weights: dict):
        """weights: name_weight / name_bias float32 arrays with MXNet layouts
        (deepim/symbols/deepIM_flownet.py:63-116,716-717)."""
        keep = []
        W = (C.c_void_p * 14)()
        Bv = (C.c_void_p * 14)()
        for i, n in enumerate(WEIGHT_ORDER):
            w = np.ascontiguousarray(weights[n + "_weight"], np.float32)
            b = np.ascontiguousarray(weights[n + "_bias"], np.float32)
            keep += [w, b]
            W[i], Bv[i] = w.ctypes.data, b.ctypes.data
        check(lib.dim_net_load(self._h, W, Bv))

    def _new(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------------------ render
    def render(self, cls_idx, pose, K, znear=0.25, zfar=6.0, pixel_means_rgb=(0, 0, 0), trunc_u8=True,
               want=("image", "depth", "mask")):
        """cls_idx int32[B], pose float32[B,3,4] (CUDA). Returns dict of CUDA tensors + bbox int32[B,4]."""
        B = pose.shape[0]
        _chk(pose, torch.float32, (B, 3, 4), "pose")
        _chk(cls_idx, torch.int32, (B,), "cls_idx")
        out = {
            "image": self._new((B, 3, self.H, self.W)) if "image" in want else None,
            "depth": self._new((B, 1, self.H, self.W)) if "depth" in want else None,
            "mask": self._new((B, 1, self.H, self.W)) if "mask" in want else None,
            "bgr": self._new((B, self.H, self.W, 3)) if "bgr" in want else None,
            "bbox": self._new((B, 4), torch.int32),
        }
        K9 = farr(np.asarray(K, np.float32).reshape(9), 9)
        means = farr(pixel_means_rgb, 3, C.c_double)
        check(lib.dim_render(self._h, _p(cls_idx), _p(pose), B, K9, znear, zfar, means, int(trunc_u8), _p(out["image"]),
                             _p(out["depth"]), _p(out["mask"]), _p(out["bgr"]), _p(out["bbox"]), self._stream()))
        return out

    def render_lit(self, cls_idx, pose, K, light_position, light_intensity, brightness_ratio=0.7, znear=0.25, zfar=6.0,
                   pixel_means_rgb=(0, 0, 0), want=("bgr", "depth")):
        """Lambert-lit render (render_py_light_modelnet_multi.py:131-175).  light_position / light_intensity float32[B,3]
        CUDA tensors (position in the GL camera frame)."""
        B = pose.shape[0]
        _chk(pose, torch.float32, (B, 3, 4), "pose")
        _chk(cls_idx, torch.int32, (B,), "cls_idx")
        _chk(light_position, torch.float32, (B, 3), "light_position")
        _chk(light_intensity, torch.float32, (B, 3), "light_intensity")
        out = {
            "image": self._new((B, 3, self.H, self.W)) if "image" in want else None,
            "depth": self._new((B, 1, self.H, self.W)) if "depth" in want else None,
            "mask": self._new((B, 1, self.H, self.W)) if "mask" in want else None,
            "bgr": self._new((B, self.H, self.W, 3)) if "bgr" in want else None,
            "bbox": self._new((B, 4), torch.int32),
        }
        check(lib.dim_render_lit(self._h, _p(cls_idx), _p(pose), B, farr(np.asarray(K, np.float32).reshape(9), 9), znear, zfar,
                                 farr(pixel_means_rgb, 3, C.c_double), _p(light_position), _p(light_intensity),
                                 float(np.float32(brightness_ratio)), _p(out["image"]), _p(out["depth"]), _p(out["mask"]),
                                 _p(out["bgr"]), _p(out["bbox"]), self._stream()))
        return out

    # -------------------------------------------------------------------------------- zoom
    def zoom_mask(self, mask_observed, mask_gt_observed, mask_rendered, src_pose, K):
        B = mask_observed.shape[0]
        shp = (B, 1, self.H, self.W)
        for n, t in (("mask_observed", mask_observed), ("mask_gt_observed", mask_gt_observed),
                     ("mask_rendered", mask_rendered)):
            _chk(t, torch.float32, shp, n)
        _chk(src_pose, torch.float32, (B, 3, 4), "src_pose")
        zo, zg, zr = self._new(shp), self._new(shp), self._new(shp)
        zf = self._new((B, 4))
        bbox = self._new((B, 8), torch.int32)
        status = self._new((B,), torch.int32)
        K9 = farr(np.asarray(K, np.float32).reshape(9), 9)
        check(lib.dim_zoom_mask_fwd(self._h, _p(mask_observed), _p(mask_gt_observed), _p(mask_rendered), _p(src_pose), B,
                                    K9, _p(zo), _p(zg), _p(zr), _p(zf), _p(bbox), _p(status), self._stream()))
        return zo, zg, zr, zf, bbox, status

    def zoom_image_with_factor(self, zoom_factor, image_observed, image_rendered, pixel_means_rgb):
        B = image_observed.shape[0]
        shp = (B, 3, self.H, self.W)
        _chk(image_observed, torch.float32, shp, "image_observed")
        _chk(image_rendered, torch.float32, shp, "image_rendered")
        _chk(zoom_factor, torch.float32, (B, 4), "zoom_factor")
        zo, zr = self._new(shp), self._new(shp)
        check(lib.dim_zoom_image_with_factor_fwd(self._h, _p(zoom_factor), _p(image_observed), _p(image_rendered), B,
                                                 farr(pixel_means_rgb, 3), _p(zo), _p(zr), self._stream()))
        return zo, zr

    def zoom_image(self, image_observed, image_rendered, src_pose, K, pixel_means_rgb):
        """ZoomImage (zoom_image.py:26-107): boxes from sum_c(image + mean) > 0.01."""
        B = image_observed.shape[0]
        shp = (B, 3, self.H, self.W)
        _chk(image_observed, torch.float32, shp, "image_observed")
        _chk(image_rendered, torch.float32, shp, "image_rendered")
        _chk(src_pose, torch.float32, (B, 3, 4), "src_pose")
        zo, zr, zf = self._new(shp), self._new(shp), self._new((B, 4))
        bbox, status = self._new((B, 8), torch.int32), self._new((B,), torch.int32)
        check(lib.dim_zoom_image_fwd(self._h, _p(image_observed), _p(image_rendered), _p(src_pose), B,
                                     farr(np.asarray(K, np.float32).reshape(9), 9), farr(pixel_means_rgb, 3), _p(zo), _p(zr),
                                     _p(zf), _p(bbox), _p(status), self._stream()))
        return zo, zr, zf, bbox, status

    def group_picker(self, data, group_idx, group_num, backward=False, channels=None):
        """GroupPicker (group_picker.py:22-56).  forward: data [B,C,...] -> [B,C/group_num,...];
        backward: data = out_grad [B,C/group_num,...] -> [B,channels,...] (zero outside the picked group)."""
        B = data.shape[0]
        Ctot = channels if backward else data.shape[1]
        n = int(np.prod(data.shape[2:])) if data.dim() > 2 else 1
        gi = group_idx.reshape(-1).to(torch.float32).contiguous()
        _chk(data, torch.float32, tuple(data.shape), "data")
        out = self._new((B, Ctot if backward else Ctot // group_num) + tuple(data.shape[2:]))
        check(lib.dim_group_picker(self._h, _p(data), _p(gi), B, Ctot, group_num, n, int(backward), _p(out), self._stream()))
        return out

    def zoom_mask_with_factor(self, zoom_factor, mask, b_inv_zoom):
        B = mask.shape[0]
        _chk(mask, torch.float32, (B, 1, self.H, self.W), "mask")
        out = self._new(mask.shape)
        check(lib.dim_zoom_mask_with_factor_fwd(self._h, _p(zoom_factor), _p(mask), B, int(b_inv_zoom), _p(out),
                                                self._stream()))
        return out

    def zoom_flow(self, zoom_factor, flow, flow_weights=None, b_inv_zoom=False):
        B = flow.shape[0]
        _chk(flow, torch.float32, (B, 2, self.H, self.W), "flow")
        out = self._new(flow.shape)
        outw = None
        fwc = 1
        if not b_inv_zoom and flow_weights is not None:
            fwc = flow_weights.shape[1]  # 1, or 2 when tiled like batch_updater_py_multi.py:293-296
            _chk(flow_weights, torch.float32, (B, fwc, self.H, self.W), "flow_weights")
            outw = self._new(flow_weights.shape)
        check(lib.dim_zoom_flow_fwd(self._h, _p(zoom_factor), _p(flow), _p(flow_weights), fwc, B, int(b_inv_zoom), _p(out),
                                    _p(outw), self._stream()))
        return out, outw

    def zoom_depth(self, zoom_factor, depth_observed, depth_rendered):
        B = depth_observed.shape[0]
        zo, zr = self._new(depth_observed.shape), self._new(depth_rendered.shape)
        check(lib.dim_zoom_depth_fwd(self._h, _p(zoom_factor), _p(depth_observed), _p(depth_rendered), B, _p(zo), _p(zr),
                                     self._stream()))
        return zo, zr

    def zoom_trans(self, zoom_factor, trans, b_inv_zoom):
        B = trans.shape[0]
        _chk(trans, torch.float32, (B, 3), "trans_delta")
        out = self._new((B, 3))
        check(lib.dim_zoom_trans_fwd(self._h, _p(zoom_factor), _p(trans), B, int(b_inv_zoom), _p(out), self._stream()))
        return out

    def zoom_trans_backward(self, zoom_factor, out_grad, b_inv_zoom, b_zoom_grad):
        B = out_grad.shape[0]
        out = self._new((B, 3))
        check(lib.dim_zoom_trans_bwd(self._h, _p(zoom_factor), _p(out_grad), B, int(b_inv_zoom), int(b_zoom_grad),
                                     _p(out), self._stream()))
        return out

    def update_mask_box(self, bbox4):
        B = bbox4.shape[0]
        _chk(bbox4, torch.int32, (B, 4), "bbox")
        out = self._new((B, 1, self.H, self.W))
        check(lib.dim_update_mask_box(self._h, _p(bbox4), B, _p(out), self._stream()))
        return out

    # ---------------------------------------------------------------------------- geometry
    def se3_compose(self, pose_src, se3, T_means=(0, 0, 0), T_stds=(1, 1, 1), rot_coord="camera"):
        B = pose_src.shape[0]
        _chk(pose_src, torch.float64, (B, 3, 4), "pose_src")
        _chk(se3, torch.float32, (B, 7), "se3")
        out = self._new((B, 3, 4), torch.float64)
        check(lib.dim_se3_compose(self._h, _p(pose_src), _p(se3), B, farr(T_means, 3, C.c_double),
                                  farr(T_stds, 3, C.c_double), capi.ROT_COORD[rot_coord.lower()], _p(out), self._stream()))
        return out

    def flow(self, depth_src, depth_tgt, KT, Kinv):
        B = depth_src.shape[0]
        shp = (B, 1, self.H, self.W)
        _chk(depth_src, torch.float32, shp, "depth_src")
        _chk(depth_tgt, torch.float32, shp, "depth_tgt")
        _chk(KT, torch.float32, (B, 3, 4), "KT")
        fl, va = self._new((B, 2, self.H, self.W)), self._new(shp)
        check(lib.dim_flow_fwd(self._h, _p(depth_src), _p(depth_tgt), _p(KT), farr(np.asarray(Kinv, np.float32).reshape(9), 9),
                               B, _p(fl), _p(va), self._stream()))
        return fl, va

    def transform3d(self, point_cloud, rotation, translation, pose_src, T_means, T_stds, rot_coord="model"):
        B, _, N = point_cloud.shape
        out = self._new(point_cloud.shape)
        check(lib.dim_transform3d_fwd(self._h, _p(point_cloud), _p(rotation), _p(translation), _p(pose_src), B, N,
                                      farr(T_means, 3), farr(T_stds, 3), capi.ROT_COORD[rot_coord.lower()], _p(out),
                                      self._stream()))
        return out

    def transform3d_backward(self, out_grad, point_cloud, rotation, translation, pose_src, T_means, T_stds,
                             rot_coord="model"):
        B, _, N = point_cloud.shape
        rg, tg = self._new((B, 4)), self._new((B, 3))
        check(lib.dim_transform3d_bwd(self._h, _p(out_grad), _p(point_cloud), _p(rotation), _p(translation),
                                      _p(pose_src), B, N, farr(T_means, 3), farr(T_stds, 3),
                                      capi.ROT_COORD[rot_coord.lower()], _p(rg), _p(tg), self._stream()))
        return rg, tg

    def train_update(self, cls_idx, src_pose, rot_est, trans_est, tgt_pose, depth_gt_observed, K,
                     pixel_means_rgb=(103.939, 116.779, 123.68), T_means=(0, 0, 0), T_stds=(1, 1, 1),
                     rot_coord="camera", znear=0.25, zfar=6.0, want_flow=True):
        """batchUpdaterPyMulti.forward on the device (lib/pair_matching/batch_updater_py_multi.py:91-328)."""
        B = src_pose.shape[0]
        for n, t, shp in (("src_pose", src_pose, (B, 3, 4)), ("tgt_pose", tgt_pose, (B, 3, 4)), ("rot_est", rot_est, (B, 4)),
                          ("trans_est", trans_est, (B, 3))):
            _chk(t, torch.float32, shp, n)
        _chk(cls_idx, torch.int32, (B,), "cls_idx")
        out = {
            "image_rendered": self._new((B, 3, self.H, self.W)), "depth_rendered": self._new((B, 1, self.H, self.W)),
            "mask_rendered": self._new((B, 1, self.H, self.W)), "src_pose": self._new((B, 3, 4)),
            "rot": self._new((B, 4)), "trans": self._new((B, 3)),
            "flow": self._new((B, 2, self.H, self.W)) if want_flow else None,
            "flow_weights": self._new((B, 2, self.H, self.W)) if want_flow else None,
        }
        if want_flow:
            _chk(depth_gt_observed, torch.float32, (B, 1, self.H, self.W), "depth_gt_observed")
        check(lib.dim_train_update(self._h, _p(cls_idx), _p(src_pose), _p(rot_est), _p(trans_est), _p(tgt_pose),
                                   _p(depth_gt_observed) if want_flow else None, B,
                                   farr(np.asarray(K, np.float64).reshape(9), 9, C.c_double), znear, zfar,
                                   farr(pixel_means_rgb, 3, C.c_double), farr(T_means, 3, C.c_double),
                                   farr(T_stds, 3, C.c_double), capi.ROT_COORD[rot_coord.lower()],
                                   _p(out["image_rendered"]), _p(out["depth_rendered"]), _p(out["mask_rendered"]),
                                   _p(out["src_pose"]), _p(out["rot"]), _p(out["trans"]), _p(out["flow"]),
                                   _p(out["flow_weights"]), self._stream()))
        return out

    def transform_image_u8(self, bgr_u8, pixel_means_rgb):
        B = bgr_u8.shape[0]
        _chk(bgr_u8, torch.uint8, (B, self.H, self.W, 3), "bgr_u8")
        out = self._new((B, 3, self.H, self.W))
        check(lib.dim_transform_image_u8(self._h, _p(bgr_u8), B, farr(pixel_means_rgb, 3, C.c_double), _p(out), self._stream()))
        return out

    # --------------------------------------------------------------------------------- net
    def net_forward(self, zoom_image_observed, zoom_image_rendered, zoom_mask_observed, zoom_mask_rendered,
                    precision=capi.PREC_BF16X3):
        B = zoom_image_observed.shape[0]
        rot, trans = self._new((B, 4)), self._new((B, 3))
        check(lib.dim_net_fwd(self._h, _p(zoom_image_observed), _p(zoom_image_rendered), _p(zoom_mask_observed),
                              _p(zoom_mask_rendered), B, precision, _p(rot), _p(trans), self._stream()))
        return rot, trans

    def debug_activation(self, idx, B, lo=False, fp16=False):
        """16-bit NHWC activation buffer feeding conv layer idx (10 = fc6 input) as float32 numpy
        [B, rows, cols, C] including the zero border (fp16=True: the last pass ran in DIM_PREC_FP16)."""
        g = (C.c_int32 * 8)()
        check(lib.dim_debug_layer_geometry(self._h, idx, g))
        rows, cols, ch = g[0], g[1], g[2]
        n = B * rows * cols * ch
        buf = np.empty(n, np.uint16)
        torch.cuda.synchronize()
        check(lib.dim_debug_activation(self._h, idx, int(lo), buf.ctypes.data, n * 2))
        f = buf.view(np.float16).astype(np.float32) if fp16 else (buf.astype(np.uint32) << 16).view(np.float32)
        return f.reshape(B, rows, cols, ch), tuple(g)

    # ------------------------------------------------------------------------------ refine
    def get_config(self) -> dict:
        """dim_train_get_config: loss weights / normalisers / pose parameterisation in effect on this context"""
        cfg = capi.TrainConfig()
        check(lib.dim_train_get_config(self._h, C.byref(cfg)))
        return {"lw_flow": cfg.lw_flow, "lw_mask": cfg.lw_mask, "lw_pm": cfg.lw_pm, "num_3d_sample": cfg.num_3d_sample,
                "normalize_3d_point": cfg.normalize_3d_point, "normalize_flow": cfg.normalize_flow,
                "trans_means": list(cfg.trans_means), "trans_stds": list(cfg.trans_stds),
                "rot_coord": "CAMERA" if cfg.rot_coord == 1 else "MODEL"}

    def set_config(self, **fields):
        """dim_train_set_config: override some of the yaml-level constants (see get_config for the names).  trans_means /
        trans_stds / rot_coord also drive refine(); the rest applies to the training step."""
        cfg = capi.TrainConfig()
        check(lib.dim_train_get_config(self._h, C.byref(cfg)))
        for k, v in fields.items():
            if k in ("trans_means", "trans_stds"):
                setattr(cfg, k, (C.c_float * 3)(*[float(x) for x in v]))
            elif k == "rot_coord":
                cfg.rot_coord = capi.ROT_COORD[v.lower()] if isinstance(v, str) else int(v)
            elif hasattr(cfg, k):
                setattr(cfg, k, float(v))
            else:
                raise ValueError("unknown config field %r" % k)
        check(lib.dim_train_set_config(self._h, C.byref(cfg)))

    def refine(self, image_observed, cls_idx, pose_init, K, n_iter=4, znear=0.25, zfar=6.0,
               pixel_means_rgb=(103.939, 116.779, 123.68), precision=capi.PREC_FP16, pose_override=None, out=None):
        """Device-resident fused loop.  image_observed f32[B,3,H,W], cls_idx i32[B], pose_init f64[B,3,4].
        out = the dict returned by an earlier call with the same shapes: results are written into those tensors again
        (same device addresses -> the library replays its CUDA graph of the chain instead of re-enqueuing ~90 launches)."""
        B = image_observed.shape[0]
        _chk(image_observed, torch.float32, (B, 3, self.H, self.W), "image_observed")
        _chk(cls_idx, torch.int32, (B,), "cls_idx")
        _chk(pose_init, torch.float64, (B, 3, 4), "pose_init")
        if out is not None:
            poses, se3, zf, bbox = out["poses"], out["se3"], out["zoom_factor"], out["bbox"]
            _chk(poses, torch.float64, (n_iter, B, 3, 4), "out['poses']")
            _chk(se3, torch.float32, (n_iter, B, 7), "out['se3']")
            _chk(zf, torch.float32, (n_iter, B, 4), "out['zoom_factor']")
            _chk(bbox, torch.int32, (n_iter, B, 8), "out['bbox']")
        else:
            poses = self._new((n_iter, B, 3, 4), torch.float64)
            se3 = self._new((n_iter, B, 7))
            zf = self._new((n_iter, B, 4))
            bbox = self._new((n_iter, B, 8), torch.int32)
        if pose_override is not None:
            _chk(pose_override, torch.float64, (n_iter, B, 3, 4), "pose_override")
        check(lib.dim_refine(self._h, _p(image_observed), _p(cls_idx), _p(pose_init), B, n_iter,
                             farr(np.asarray(K, np.float32).reshape(9), 9), znear, zfar,
                             farr(pixel_means_rgb, 3, C.c_double), precision, _p(pose_override), _p(poses), _p(se3),
                             _p(zf), _p(bbox), self._stream()))
        return {"poses": poses, "se3": se3, "zoom_factor": zf, "bbox": bbox}

    def refine_host(self, image_observed_u8, cls_idx, pose_init, K, n_iter=4, znear=0.25, zfar=6.0,
                    pixel_means_rgb=(103.939, 116.779, 123.68), precision=capi.PREC_FP16, poses_out=None,
                    se3_out=None, sync=True):
        """Host-buffer entry (what a tester loop calls): uint8 BGR HWC images (pinned torch tensors or
        numpy), host poses in / out.  sync=False only enqueues on the current torch stream (outputs must
        then be pinned and are valid after the stream is synchronised)."""
        def hptr(a):
            return C.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else C.c_void_p(a.ctypes.data)
        B = image_observed_u8.shape[0]
        if poses_out is None:
            poses_out = np.empty((n_iter, B, 3, 4), np.float64)
        if se3_out is None:
            se3_out = np.empty((n_iter, B, 7), np.float32)
        fn = lib.dim_refine_host if sync else lib.dim_refine_host_async
        check(fn(self._h, hptr(image_observed_u8), hptr(cls_idx), hptr(pose_init), B, n_iter,
                                  farr(np.asarray(K, np.float32).reshape(9), 9), znear, zfar,
                 farr(pixel_means_rgb, 3, C.c_double), precision, hptr(poses_out), hptr(se3_out), self._stream()))
        return poses_out, se3_out


def _refine_status(self, B, n_iter, out=None, sync=True):
    """Per-iteration status of the last refine / refine_host call: int32 [min(n_iter,8), B]; 0 = ok, bit 0 = empty rendered
    mask in that iteration (pose meaningless; the reference crashes there), bit 1 = bad class index.  `out`: pinned int32
    tensor for an asynchronous copy on the current stream (sync=False)."""
    n = min(int(n_iter), 8)
    if out is None:
        out = torch.empty((n, B), dtype=torch.int32).pin_memory()
    check(lib.dim_refine_status(self._h, B, n_iter, C.c_void_p(out.data_ptr()), self._stream()))
    if sync:
        torch.cuda.current_stream(self.device).synchronize()
    return out


Context.refine_status = _refine_status


def _profile_enable(self, on=True):
    check(lib.dim_profile_enable(self._h, int(on)))


def _profile_read(self):
    """-> (dict stage -> ms accumulated since last read, iterations)"""
    ms = (C.c_float * 4)()
    n = C.c_int32()
    check(lib.dim_profile_read(self._h, ms, C.byref(n)))
    return {"render": ms[0], "zoom": ms[1], "conv": ms[2], "head": ms[3]}, n.value


Context.profile_enable = _profile_enable
Context.profile_read = _profile_read


def launch_count(reset=False):
    return int(lib.dim_launch_count(int(reset)))
