"""The oracle (oracle/, test infrastructure) against fixtures produced by the reference's UNMODIFIED custom operators
(tests/golden/make_golden_mx.py: deepim/operator_py/*.py and lib/pair_matching/data_pair.py executed under the numpy-backed
mxnet stand-in tests/golden/mx_shim.py).  This pins the reference-held half of the zoom ops -- bbox extraction, the mixed
float32 / float64 zoom-factor arithmetic, inverse-zoom affines, the 0.2 / 0.3 / 0.45 thresholds, round, the +-mean order,
ZoomTrans, Transform3D forward / backward, update_data_batch's end-exclusive box -- to the reference's own source.  The
GridGenerator / BilinearSampler arithmetic underneath is MXNet's (not in /root/reference): there the stand-in follows the
oracle's formula, so planes are expected bit-identical and asserted to <= 1 ulp.  CPU only; reads nothing outside the repo."""
import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import mx_cases as C  # noqa: E402


def load(name):
    return np.load(os.path.join(HERE, "golden", name))


def ulp_diff(a, b):
    """max distance in units of float32 spacing at max(|a|,|b|)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    sp = np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / sp)) if a.size else 0.0


@pytest.fixture(scope="module")
def small():
    g = load("ref_mx_zoom_small.npz")
    return g, C.zoom_case(int(g["seed"]), int(g["B"]), int(g["H"]), int(g["W"]))


def test_zoom_mask_bbox_ints_and_zoom_factor_bit_exact_full_frame():
    """480 x 640, 6 instances (ragged masks, depth-like rendered mask, one object cut by the frame): the 8 integer zoom bbox
    indices observed inside the reference's ZoomMaskOperator.forward, zoom_factor and the rounded zoomed masks, bit for bit."""
    g = load("ref_mx_zoom_full.npz")
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    c = C.zoom_case(int(g["seed"]), B, H, W)
    zo, zg, zr, zf, bbox = O.zoom_mask(c["mo"], c["mo"], c["mr"], c["pose"], c["K"])
    assert np.array_equal(bbox, g["bbox"])
    assert np.array_equal(zf, g["zoom_factor"])                        # float32 bit patterns
    assert np.array_equal(np.packbits(zo.astype(np.uint8)), g["zm_obs"])
    assert np.array_equal(np.packbits(zr.astype(np.uint8)), g["zm_ren"])
    assert np.array_equal(zo, zg)


def test_zoom_ops_small_frame(small):
    g, c = small
    zo, zg, zr, zf, bbox = O.zoom_mask(c["mo"], c["mo"], c["mr"], c["pose"], c["K"])
    assert np.array_equal(bbox, g["bbox"]) and np.array_equal(zf, g["zoom_factor"])
    assert np.array_equal(zo, g["zm_obs"]) and np.array_equal(zg, g["zm_gt"]) and np.array_equal(zr, g["zm_ren"])
    zio, zir = O.zoom_image_with_factor(zf, c["img_o"], c["img_r"], C.PIXEL_MEANS_RGB)
    assert ulp_diff(zio, g["zio"]) <= 1 and ulp_diff(zir, g["zir"]) <= 1
    for inv in (False, True):
        assert np.array_equal(O.zoom_mask_with_factor(zf, c["depth"], inv), g["zmwf_inv%d" % inv])
    fl, fw = O.zoom_flow(zf, c["flow"], c["fw"], False)
    assert ulp_diff(fl, g["zflow"]) <= 1 and np.array_equal(fw, g["zflow_w"].astype(np.float32))
    fl_inv, _ = O.zoom_flow(zf, c["flow"], None, True)
    assert ulp_diff(fl_inv, g["zflow_inv"]) <= 1
    assert ulp_diff(O.zoom_depth(zf, c["depth"]), g["zdepth"]) <= 1
    zi_o, zi_r, zi_f, _ = O.zoom_image(c["img_o"], c["img_r"], c["pose"], c["K"], C.PIXEL_MEANS_RGB)
    assert np.array_equal(zi_f, g["zimg_factor"])
    assert hashlib.sha256(zi_o.tobytes()).digest() == g["zimg_o_sha"].tobytes()
    assert hashlib.sha256(zi_r.tobytes()).digest() == g["zimg_r_sha"].tobytes()


def test_zoom_trans_forward_backward(small):
    g, _ = small
    zf, tr = g["zoom_factor"], g["trans_in"]
    for inv in (False, True):
        assert np.array_equal(O.zoom_trans(zf, tr, inv), g["ztrans_inv%d" % inv])
        # backward: b_zoom_grad=True scales like the forward, False passes the gradient through (zoom_trans.py:48-74)
        assert np.array_equal(O.zoom_trans(zf, tr[::-1].copy(), inv), g["ztrans_bwd_inv%d_zg1" % inv])
        assert np.array_equal(tr[::-1], g["ztrans_bwd_inv%d_zg0" % inv])


def test_transform3d_forward_backward_against_the_reference_operator():
    """Reference tolerances are 1e-4 (forward, transform3d.py:407) and 5e-3 relative (finite-difference gradient check,
    l.421-539); against the operator itself the oracle is held to float32 rounding.  Instance 3 carries an un-normalised
    quaternion: identity rotation forward, zero rotation gradient backward (l.188-189, 221-222)."""
    g = load("ref_mx_transform3d.npz")
    c = C.t3d_case(int(g["seed"]))
    for coord in ("MODEL", "CAMERA"):
        fw = O.transform3d_forward(c["pts"], c["q"], c["t"], c["pose_src"], c["T_means"], c["T_stds"], coord.lower())
        assert np.abs(fw - g["fwd_" + coord]).max() < 2e-6
        rg, tg = O.transform3d_backward(c["og"], c["pts"], c["q"], c["t"], c["pose_src"], c["T_means"], c["T_stds"], coord.lower())
        assert np.abs(rg - g["rot_grad_" + coord]).max() < 2e-5 * max(1.0, np.abs(g["rot_grad_" + coord]).max())
        assert np.abs(tg - g["trans_grad_" + coord]).max() < 2e-5 * max(1.0, np.abs(g["trans_grad_" + coord]).max())
        assert not g["rot_grad_" + coord][3].any() and not rg[3].any()


def test_update_data_batch_box_mask_and_image_transform():
    """data_pair.py:66-129 with UPDATE_MASK = box_rendered: mask_observed := end-EXCLUSIVE rectangle of the new rendered
    mask, image_rendered := transform(bgr), src_pose float32."""
    g = load("ref_mx_update_data_batch.npz")
    from deepim_b200 import synth
    for b in range(2):
        mr = g["mask_rendered_%d" % b][0, 0].astype(np.float32)
        bb = O.mask_bbox(mr, 0.5)
        assert np.array_equal(O.box_mask(bb, *mr.shape), g["mask_observed_%d" % b][0, 0].astype(np.float32))
        assert np.array_equal(synth.transform_image(g["img_bgr_%d" % b]), g["image_rendered_%d" % b][0])
        assert g["src_pose_%d" % b].dtype == np.float32
