"""Golden fixtures from the reference's UNMODIFIED custom operators, executed under the numpy-backed `mxnet` stand-in
(tests/golden/mx_shim.py).  Runs only where /root/reference exists:

    python tests/golden/make_golden_mx.py        # rewrites tests/golden/ref_mx_*.npz

What is executed as written (imported by file path from /root/reference, never copied):
    deepim/operator_py/zoom_mask.py, zoom_image_with_factor.py, zoom_mask_with_factor.py, zoom_flow.py, zoom_trans.py,
    zoom_depth.py, zoom_image.py, transform3d.py        (forward, and backward where the graph uses it)
    lib/pair_matching/data_pair.py:update_data_batch    (inter-iteration blob update incl. the end-exclusive box mask)
Not pinned: group_picker.py (dead code: `output_shape[1] /= group_num` on an int array raises under true division, the
file's own `from __future__ import division`), flow_updater.py (imported by the symbol file but never instantiated; the
train loop labels flow with lib/flow_c's gpu_flow -> tests/golden/ref_flow.npz).

The 8 "integer zoom bbox indices" of the parity spec are LOCAL variables of ZoomMaskOperator.forward
(obj_real_start_x ... obj_rendered_end_y, zoom_mask.py:55-82); they are captured by a sys.settrace hook on the frame of
`forward` (one instance per call), i.e. observed, not re-derived.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import mx_cases as C  # noqa: E402
import mx_shim  # noqa: E402


def load_ops():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; fixtures are committed, nothing to do")
    mx_shim.install()
    sys.path.insert(0, REF)                      # transform3d.py: from lib.pair_matching.RT_transform import ...
    mods = {}
    for name in ("zoom_mask", "zoom_image_with_factor", "zoom_mask_with_factor", "zoom_flow", "zoom_trans", "zoom_depth",
                 "zoom_image", "transform3d"):
        spec = importlib.util.spec_from_file_location("ref_op_" + name, os.path.join(REF, "deepim/operator_py", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def forward_locals(op_type, inputs, names, **attrs):
    """run the op once and return (outputs, {name: value}) with the named locals of its `forward` frame at return"""
    seen = {}

    def tracer(frame, event, arg):
        if event == "call" and frame.f_code.co_name == "forward":
            def local(frame, event, arg):
                if event == "return":
                    for n in names:
                        if n in frame.f_locals:
                            seen[n] = frame.f_locals[n]
                return local
            return local
        return None

    sys.settrace(tracer)
    try:
        outs, op, io = mx_shim.run_op(op_type, inputs, **attrs)
    finally:
        sys.settrace(None)
    return outs, seen


BBOX_LOCALS = ["obj_real_start_x", "obj_real_end_x", "obj_real_start_y", "obj_real_end_y",
               "obj_rendered_start_x", "obj_rendered_end_x", "obj_rendered_start_y", "obj_rendered_end_y"]


def zoom_mask_with_bbox(c, H, W):
    B = c["mo"].shape[0]
    attrs = dict(K=C.k_attr(c["K"]), height=H, width=W)
    outs, _, _ = mx_shim.run_op("ZoomMask", [c["mo"], c["mo"], c["mr"], c["pose"]], **attrs)
    bbox = np.zeros((B, 8), np.int32)
    for b in range(B):  # one instance per call: the bbox locals of the loop body survive until `forward` returns
        o1, loc = forward_locals("ZoomMask", [c["mo"][b:b + 1], c["mo"][b:b + 1], c["mr"][b:b + 1], c["pose"][b:b + 1]],
                                 BBOX_LOCALS, **attrs)
        bbox[b] = [int(loc[n]) for n in BBOX_LOCALS]
        assert np.array_equal(o1[3][0], outs[3][b])
    return outs, bbox


def gen_zoom(tag, seed, B, H, W):
    c = C.zoom_case(seed, B, H, W)
    out = {}
    (zmo, zmg, zmr, zf), bbox = zoom_mask_with_bbox(c, H, W)
    out.update(zm_obs=zmo.astype(np.uint8), zm_gt=zmg.astype(np.uint8), zm_ren=zmr.astype(np.uint8), zoom_factor=zf, bbox=bbox)
    means_attr = C.vec_attr(C.PIXEL_MEANS_RGB[::-1])      # the attr is reversed again inside the Prop (zoom_image_with_factor.py:79-81)
    (zio, zir), _, _ = mx_shim.run_op("ZoomImageWithFactor", [zf, c["img_o"], c["img_r"]], height=H, width=W, pixel_means=means_attr)
    out.update(zio=zio, zir=zir)
    for inv in (False, True):
        (zm,), _, _ = mx_shim.run_op("ZoomMaskWithFactor", [zf, c["depth"]], height=H, width=W, b_inv_zoom=str(inv))
        out["zmwf_inv%d" % inv] = zm.astype(np.uint8)
    (zfl, zfw), _, _ = mx_shim.run_op("ZoomFlow", [zf, c["flow"], c["fw"]], height=H, width=W, b_inv_zoom="False")
    out.update(zflow=zfl, zflow_w=zfw.astype(np.int8))
    (zfl_inv,), _, _ = mx_shim.run_op("ZoomFlow", [zf, c["flow"]], height=H, width=W, b_inv_zoom="True")
    out.update(zflow_inv=zfl_inv)
    (zd_o, zd_r), _, _ = mx_shim.run_op("ZoomDepth", [zf, c["depth"], c["depth"]], height=H, width=W)
    assert np.array_equal(zd_o, zd_r)
    out.update(zdepth=zd_o)
    (zi_o, zi_r, zi_f), _, _ = mx_shim.run_op("ZoomImage", [c["img_o"], c["img_r"], c["pose"]], K=C.k_attr(c["K"]), height=H, width=W,
                                              pixel_means=means_attr)
    # ZoomImage = ZoomMask's centre / crop rule on boxes taken from the images + ZoomImageWithFactor's sampling: the factor and
    # a digest of the planes pin it (the planes themselves would duplicate zio / zir)
    out.update(zimg_factor=zi_f, zimg_o_sha=np.frombuffer(hashlib.sha256(zi_o.tobytes()).digest(), np.uint8),
               zimg_r_sha=np.frombuffer(hashlib.sha256(zi_r.tobytes()).digest(), np.uint8))
    # ZoomTrans forward / backward, all flag combinations the graph uses (symbol:222, 456, 722)
    rng = np.random.default_rng(seed + 1)
    tr = (rng.normal(size=(B, 3)) * 0.05).astype(np.float32)
    for inv in (False, True):
        for zg in (False, True):
            (zt,), op, io = mx_shim.run_op("ZoomTrans", [zf, tr], b_inv_zoom=str(inv), b_zoom_grad=str(zg))
            g = mx_shim.run_op_backward(op, io, [tr[::-1].copy()])
            out["ztrans_inv%d" % inv] = zt
            out["ztrans_bwd_inv%d_zg%d" % (inv, zg)] = g[1]
    out["trans_in"] = tr
    np.savez_compressed(os.path.join(HERE, "ref_mx_zoom_%s.npz" % tag), seed=seed, B=B, H=H, W=W, **out)
    print("ref_mx_zoom_%s.npz" % tag, {k: v.shape for k, v in out.items() if hasattr(v, "shape")}.keys())
    return c, out


def gen_full_masks(seed, B):
    """480 x 640: bbox ints, zoom_factor, zoomed masks (bit-packed)"""
    H, W = C.FULL
    c = C.zoom_case(seed, B, H, W)
    (zmo, zmg, zmr, zf), bbox = zoom_mask_with_bbox(c, H, W)
    np.savez_compressed(os.path.join(HERE, "ref_mx_zoom_full.npz"), seed=seed, B=B, H=H, W=W, zoom_factor=zf, bbox=bbox,
                        zm_obs=np.packbits(zmo.astype(np.uint8)), zm_ren=np.packbits(zmr.astype(np.uint8)))
    print("ref_mx_zoom_full.npz bbox", bbox.tolist())


def gen_t3d(seed):
    c = C.t3d_case(seed)
    out = {}
    for coord in ("MODEL", "CAMERA"):
        (fw,), op, io = mx_shim.run_op("Transform3D", [c["pts"], c["q"], c["t"], c["pose_src"]], T_means=C.vec_attr(c["T_means"]),
                                       T_stds=C.vec_attr(c["T_stds"]), rot_coord=coord, b_project_2d="False")
        g = mx_shim.run_op_backward(op, io, [c["og"]])
        out["fwd_" + coord], out["rot_grad_" + coord], out["trans_grad_" + coord] = fw, g[1], g[2]
    np.savez_compressed(os.path.join(HERE, "ref_mx_transform3d.npz"), seed=seed, **out)
    print("ref_mx_transform3d.npz", {k: float(np.abs(v).max()) for k, v in out.items()})


def gen_update_data_batch(seed):
    """lib/pair_matching/data_pair.py:66-129 with TEST.UPDATE_MASK = box_rendered (yaml:118): feed one rendered mask / image /
    pose through the reference function and store the blobs it writes."""
    mx_shim.install()
    sys.path.insert(0, REF)
    from lib.pair_matching import data_pair
    H, W = C.SMALL
    c = C.zoom_case(seed, 2, H, W)
    rng = np.random.default_rng(seed)
    means = np.array([123.68, 116.779, 103.939])                # config.network.PIXEL_MEANS (yaml:57-59), indexed as B,G,R (a2)
    cfg = types.SimpleNamespace(network=types.SimpleNamespace(PIXEL_MEANS=means), TEST=types.SimpleNamespace(UPDATE_MASK="box_rendered"))
    outs = {}
    for b in range(2):
        mask_ren = (c["mr"][b, 0] > 0.2).astype(np.float32)
        img_bgr = rng.integers(0, 256, size=(H, W, 3)).astype(np.uint8)
        pose = c["pose"][b].astype(np.float64)
        names = ["image_observed", "image_rendered", "src_pose", "class_index", "mask_observed", "mask_rendered"]  # loader.py:32-38
        shapes = [(1, 3, H, W), (1, 3, H, W), (1, 3, 4), (1,), (1, 1, H, W), (1, 1, H, W)]
        db = types.SimpleNamespace(data=[[None] * 6], provide_data=[[(n, s) for n, s in zip(names, shapes)]])
        pkg = [{"image_rendered": img_bgr, "src_pose": pose, "mask_rendered": mask_ren, "mask_observed": mask_ren}]
        data_pair.update_data_batch(cfg, db, pkg)
        outs["img_bgr_%d" % b] = img_bgr
        outs["image_rendered_%d" % b] = np.array(db.data[0][1]._a)
        outs["src_pose_%d" % b] = np.array(db.data[0][2]._a)
        outs["mask_observed_%d" % b] = np.array(db.data[0][4]._a).astype(np.uint8)
        outs["mask_rendered_%d" % b] = np.array(db.data[0][5]._a).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_mx_update_data_batch.npz"), seed=seed, **outs)
    print("ref_mx_update_data_batch.npz", list(outs))


def main():
    load_ops()
    gen_zoom("small", 101, 3, *C.SMALL)
    gen_full_masks(202, 6)
    gen_t3d(303)
    gen_update_data_batch(404)


if __name__ == "__main__":
    main()
