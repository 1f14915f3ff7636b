"""Rendered-pair synthesis on the CUDA rasteriser (second half): what toolkit/LM6d_1_gen_rendered_pose.py
and toolkit/LM6d_2_gen_rendered.py do for the LM6d_refine layout -- perturbed "rendered" poses around every observed pose,
their renders, depth maps and pose files, and the observed / rendered pair lists -- with the glumpy renderer replaced by
dim_render, batched.

  gen_rendered_poses   LM6d_1_gen_rendered_pose.py:77-125: euler-angle noise N(0, 15 deg) on the three sxyz angles,
                       translation noise N(0, (1, 1, 5) cm); redrawn while the rotation distance to the observed pose exceeds
                       45 deg or the projected centre is within 16 px of the image border
  write_rendered_set   LM6d_2_gen_rendered.py:60-150: <root>/data/rendered/<cls>/<prefix>_<k>-color.png (uint8 BGR),
                       -depth.png (uint16, metres x 1000), -pose.txt (class-index header + 3x4), and the pair list
                       <root>/image_set/<name>_<cls>.txt with lines "<observed index> <cls>/<prefix>_<k>"

The euler helpers are the static-xyz convention of lib/pair_matching/RT_transform.py:240-360 (R = Rz(ak) Ry(aj) Rx(ai)); they
are pinned to the live reference functions in tests/golden/ref_euler.npz."""
from __future__ import annotations

import math
import os

import numpy as np

from . import lm6d_io

_EPS4 = np.finfo(float).eps * 4.0


def euler2mat(ai, aj, ak):
    """static x-y-z euler angles -> rotation matrix (RT_transform.euler2mat, axes='sxyz')"""
    si, sj, sk = math.sin(ai), math.sin(aj), math.sin(ak)
    ci, cj, ck = math.cos(ai), math.cos(aj), math.cos(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([[cj * ck, sj * sc - cs, sj * cc + ss],
                     [cj * sk, sj * ss + cc, sj * cs - sc],
                     [-sj, cj * si, cj * ci]])


def mat2euler(M):
    """rotation matrix -> static x-y-z euler angles (RT_transform.mat2euler, axes='sxyz')"""
    M = np.asarray(M, np.float64)
    cy = math.sqrt(M[0, 0] * M[0, 0] + M[1, 0] * M[1, 0])
    if cy > _EPS4:
        return math.atan2(M[2, 1], M[2, 2]), math.atan2(-M[2, 0], cy), math.atan2(M[1, 0], M[0, 0])
    return math.atan2(-M[1, 2], M[1, 1]), math.atan2(-M[2, 0], cy), 0.0


def rot_dist_deg(Ra, Rb):
    """rotation part of calc_rt_dist_m (RT_transform.py:162-173): angle of Ra Rb^T in degrees"""
    c = (np.trace(np.asarray(Ra) @ np.asarray(Rb).T) - 1.0) / 2.0
    return math.degrees(math.acos(min(1.0, max(-1.0, c))))


def gen_rendered_poses(poses_observed, K, n_per_observed=10, angle_std=15.0, angle_max=45.0, xyz_std=(0.01, 0.01, 0.05),
                       width=640, height=480, margin=16, seed=2333, max_draws=10000):
    """[N,3,4] observed poses -> [N, n_per_observed, 3, 4] rendered poses.  Rejection rule and noise model as in the
    reference script; the random stream is numpy's legacy RandomState(seed) drawn in the script's order (three angles, then
    x, y, z).  (The script's own sequence is not reproducible beyond its first draw: scipy's logm inside calc_rt_dist_m pulls
    from the same global generator.)"""
    rs = np.random.RandomState(seed)
    K = np.asarray(K, np.float64).reshape(3, 3)
    P = np.asarray(poses_observed, np.float64).reshape(-1, 3, 4)
    out = np.zeros((len(P), n_per_observed, 3, 4))
    for a, src in enumerate(P):
        e = np.array(mat2euler(src[:, :3]))
        for k in range(n_per_observed):
            for draw in range(max_draws):
                te = e + rs.normal(0, angle_std / 180.0 * math.pi, 3)
                tt = src[:, 3] + np.array([rs.normal(0, s, 1)[0] for s in xyz_std])
                R = euler2mat(te[0], te[1], te[2])
                c = K @ tt
                cx, cy = c[0] / c[2], c[1] / c[2]
                if rot_dist_deg(R, src[:, :3]) <= angle_max and margin < cx < width - margin and margin < cy < height - margin:
                    break
            else:
                raise RuntimeError("no admissible rendered pose after %d draws (observed pose %d)" % (max_draws, a))
            out[a, k, :, :3], out[a, k, :, 3] = R, tt
    return out


def write_rendered_set(root, ctx, cls_name, cls_idx, mesh_slot, observed_indices, poses_rendered, K, set_name="train",
                       znear=0.25, zfar=6.0, batch=16, first_only_in_set=False):
    """Render every pose of poses_rendered [N, n, 3, 4] (class `mesh_slot` of `ctx`) and write the rendered half of an
    LM6d_refine directory + the pair list.  observed_indices[i] is the observed index "<video>/<prefix>" of row i.
    first_only_in_set: list only rendered pose 0 of every observed frame (the reference's validation pairs)."""
    import torch
    cv2 = lm6d_io._cv2()
    N, n = poses_rendered.shape[:2]
    d = os.path.join(root, "data", "rendered", cls_name)
    os.makedirs(d, exist_ok=True)
    os.makedirs(os.path.join(root, "image_set"), exist_ok=True)
    flat = poses_rendered.reshape(N * n, 3, 4)
    names = ["%s_%d" % (observed_indices[i].split("/")[1], k) for i in range(N) for k in range(n)]
    dev = ctx.device
    for lo in range(0, N * n, batch):
        hi = min(N * n, lo + batch)
        cls = torch.full((hi - lo,), mesh_slot, dtype=torch.int32, device=dev)
        r = ctx.render(cls, torch.from_numpy(flat[lo:hi].astype(np.float32)).to(dev), K, znear, zfar, trunc_u8=False,
                       want=("bgr", "depth"))
        bgr, depth = r["bgr"].cpu().numpy(), r["depth"].cpu().numpy()[:, 0]
        for j in range(hi - lo):
            base = os.path.join(d, names[lo + j])
            cv2.imwrite(base + "-color.png", bgr[j].astype(np.uint8))                 # rgb_gl.astype('uint8') (l.103)
            cv2.imwrite(base + "-depth.png", (depth[j] * lm6d_io.DEPTH_FACTOR).astype(np.uint16))   # truncation like l.105
            lm6d_io.write_pose(base + "-pose.txt", cls_idx, flat[lo + j])
    lines = ["%s %s/%s" % (observed_indices[i], cls_name, names[i * n + k]) for i in range(N)
             for k in range(1 if first_only_in_set else n)]
    with open(os.path.join(root, "image_set", "%s_%s.txt" % (set_name, cls_name)), "w") as f:
        f.write("\n".join(lines) + "\n")
    return lines
