"""Generate golden fixtures from the LIVE reference (runs only where /root/reference exists).

    python tests/golden/make_golden.py        # rewrites tests/golden/ref_*.npz

What can be imported from the reference in this container:
  lib/pair_matching/RT_transform.py  (needs a 3-line numpy-2 shim for np.float/np.maximum_sctype)
  lib/pair_matching/flow.py:calc_flow, lib/utils/projection.py, lib/utils/pose_error.py:add/adi,
  lib/utils/image.py:transform
MXNet / glumpy are not installable offline, so the zoom ops, the network and the renderer have no
reference-generated fixtures (parity unpinned there).
Nothing from the reference is copied: this script only *calls* it and stores inputs/outputs.
"""
import doctest
import os
import math
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; fixtures are committed, nothing to do")
    # numpy-2 shim (RT_transform.py:236-237 uses np.float / np.maximum_sctype at import time)
    np.float = float
    np.int = int
    np.maximum_sctype = lambda t: np.float64
    sys.path.insert(0, REF)
    from lib.pair_matching import RT_transform as RT
    from lib.pair_matching.flow import calc_flow
    from lib.utils import pose_error
    from lib.utils import projection
    return RT, calc_flow, pose_error, projection


def rand_rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def main():
    RT, calc_flow, pose_error, projection = import_reference()
    res = doctest.testmod(RT)
    print("RT_transform doctests:", res)
    assert res.failed == 0 and res.attempted >= 20

    rng = np.random.default_rng(1234)
    N = 64
    pose_src = np.zeros((N, 3, 4))
    quat = rng.normal(size=(N, 4)) * 0.2 + np.array([1.0, 0, 0, 0])
    quat[:8] = rng.normal(size=(8, 4))  # big rotations too
    trans = rng.normal(size=(N, 3)) * np.array([0.05, 0.05, 0.2])
    T_means = np.array([0.01, -0.02, 0.03])
    T_stds = np.array([0.5, 0.7, 1.3])
    out = {}
    for k in range(N):
        pose_src[k, :, :3] = rand_rot(rng)
        pose_src[k, :, 3] = [rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), rng.uniform(0.4, 1.5)]
    for coord in ("MODEL", "CAMERA", "CAMERA_NEW"):
        po = np.zeros((N, 3, 4))
        po_norm = np.zeros((N, 3, 4))
        Rd = np.zeros((N, 3, 3))
        Td = np.zeros((N, 3))
        for k in range(N):
            po[k] = RT.RT_transform(pose_src[k], quat[k], trans[k], np.zeros(3), np.ones(3), coord)
            po_norm[k] = RT.RT_transform(pose_src[k], quat[k], trans[k], T_means, T_stds, coord)
            r, t = RT.calc_RT_delta(pose_src[k], po_norm[k], T_means, T_stds, coord, "MATRIX")
            Rd[k], Td[k] = r, t
        out["pose_out_" + coord] = po
        out["pose_out_norm_" + coord] = po_norm
        out["R_delta_" + coord] = Rd
        out["T_delta_" + coord] = Td
    # train-time labels: calc_RT_delta(..., "QUAT") (mat2quat, eigh) and K . calc_se3 (batch_updater_py_multi.py:239-259)
    Kmat = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
    pose_tgt32 = out["pose_out_CAMERA"].astype(np.float32)
    quat_lbl, trans_lbl, KT = np.zeros((N, 4)), np.zeros((N, 3)), np.zeros((N, 3, 4))
    for k in range(N):
        q, t = RT.calc_RT_delta(pose_src[k], pose_tgt32[k], np.zeros(3), np.ones(3), "CAMERA", "QUAT")
        quat_lbl[k], trans_lbl[k] = q, t
        rm, tt = RT.calc_se3(pose_src[k], pose_tgt32[k])
        se3_m = np.zeros([3, 4]); se3_m[:, :3] = rm; se3_m[:, 3] = tt
        KT[k] = np.dot(Kmat, se3_m)
    out["label_quat"], out["label_trans"], out["label_KT"], out["label_K"] = quat_lbl, trans_lbl, KT, Kmat
    # quat2mat known answers from the docstring examples (RT_transform.py:397-404)
    out["quat2mat_in"] = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0]])
    out["quat2mat_out"] = np.stack([RT.quat2mat(q) for q in out["quat2mat_in"]])
    np.savez_compressed(os.path.join(HERE, "ref_se3.npz"), pose_src=pose_src, quat=quat, trans=trans,
                        T_means=T_means, T_stds=T_stds, **out)

    # ---- calc_flow (lib/pair_matching/flow.py:12-63) on a small synthetic depth pair
    H, W = 60, 80
    K = np.array([[572.4114 / 8, 0, 325.2611 / 8], [0, 573.57043 / 8, 242.04899 / 8], [0, 0, 1]])
    yy, xx = np.mgrid[0:H, 0:W]
    depth_src = np.zeros((H, W), np.float32)
    blob = ((xx - 40) ** 2 / 400.0 + (yy - 30) ** 2 / 250.0) < 1
    depth_src[blob] = (0.8 + 0.002 * (xx - 40) + 0.001 * (yy - 30))[blob]
    ps = np.zeros((3, 4)); ps[:, :3] = np.eye(3); ps[:, 3] = [0, 0, 0]
    pt = np.zeros((3, 4)); pt[:, :3] = rand_rot(np.random.default_rng(5)) * 0 + np.eye(3)
    ang = 0.05
    pt[:, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    pt[:, 3] = [0.01, -0.005, 0.02]
    # target depth = source surface seen from the target pose: splat
    X = projection.backproject_camera(depth_src, K)
    T = projection.se3_mul(pt, projection.se3_inverse(ps))
    Xp = T[:, :3] @ X + T[:, 3:4]
    depth_tgt = np.zeros((H, W), np.float32)
    uvw = K @ Xp
    for k in range(X.shape[1]):
        if depth_src.flat[k] == 0:
            continue
        u, v = int(round(uvw[0, k] / uvw[2, k])), int(round(uvw[1, k] / uvw[2, k]))
        if 0 <= u < W and 0 <= v < H:
            depth_tgt[v, u] = uvw[2, k]
    flow, visible, _ = calc_flow(depth_src, ps, pt, K, depth_tgt)
    np.savez_compressed(os.path.join(HERE, "ref_flow.npz"), depth_src=depth_src, depth_tgt=depth_tgt,
                        pose_src=ps, pose_tgt=pt, K=K, flow=flow.astype(np.float32),
                        visible=visible.astype(np.float32))

    # ---- ADD / ADI (lib/utils/pose_error.py:72-108)
    pts = rng.normal(size=(500, 3)) * 0.03
    Re, Rg = rand_rot(rng), rand_rot(rng)
    te, tg = rng.normal(size=(3, 1)) * 0.01 + [[0], [0], [0.8]], np.array([[0.0], [0.0], [0.8]])
    np.savez_compressed(os.path.join(HERE, "ref_pose_error.npz"), pts=pts, R_est=Re, t_est=te, R_gt=Rg, t_gt=tg,
                        add=pose_error.add(Re, te, Rg, tg, pts), adi=pose_error.adi(Re, te, Rg, tg, pts))

    # ---- rotation / translation distance + average 2D re-projection error: the inputs of LM6D_REFINE.evaluate_pose
    # (5 cm 5 deg, lib/dataset/LM6D_REFINE.py:278-371) and evaluate_pose_arp_2d (Proj. 2D, l.514-)
    rng2 = np.random.default_rng(77)
    M = 48
    Kc = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1.0]])
    pts2 = rng2.normal(size=(300, 3)) * 0.04
    est, gt = np.zeros((M, 3, 4)), np.zeros((M, 3, 4))
    rd, td, arp, re_deg = np.zeros(M), np.zeros(M), np.zeros(M), np.zeros(M)
    for k in range(M):
        Rg = rand_rot(rng2)
        ang = rng2.normal(0, [0.5, 4.0, 30.0][k % 3], size=3) * np.pi / 180.0      # tiny / moderate / large perturbations
        cx, cy, cz = np.cos(ang); sx, sy, sz = np.sin(ang)
        Rd = (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
              @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
        if k == M - 1:
            Rd = np.diag([-1.0, -1.0, 1.0])                                             # exactly 180 degrees
        if k == M - 2:
            Rd = np.eye(3)                                                              # identical rotation
        gt[k, :, :3], gt[k, :, 3] = Rg, [rng2.uniform(-.1, .1), rng2.uniform(-.1, .1), rng2.uniform(.5, 1.1)]
        est[k, :, :3], est[k, :, 3] = Rd @ Rg, gt[k, :, 3] + rng2.normal(0, [0.002, 0.02, 0.06][k % 3], size=3)
        rd[k], td[k] = RT.calc_rt_dist_m(est[k], gt[k])
        re_deg[k] = pose_error.re(est[k, :, :3], gt[k, :, :3])
        arp[k] = pose_error.arp_2d(est[k, :, :3], est[k, :, 3], gt[k, :, :3], gt[k, :, 3], pts2, Kc)
    np.savez_compressed(os.path.join(HERE, "ref_pose_eval.npz"), pts=pts2, K=Kc, poses_est=est, poses_gt=gt, rot_deg=np.real(rd),
                        trans_m=td, re_deg=np.real(re_deg), arp_2d=arp)

    # ---- static-xyz euler helpers (RT_transform.py:240-360, 512-560) + the rendered-pose sampling loop of
    # toolkit/LM6d_1_gen_rendered_pose.py:77-125 written with the reference's own helpers and its np.random.seed(2333) stream
    rng3 = np.random.default_rng(99)
    ang = rng3.uniform(-math.pi, math.pi, size=(40, 3))
    ang[:, 1] = rng3.uniform(-math.pi / 2, math.pi / 2, size=40)
    ang[0] = [0.3, math.pi / 2, -0.2]                       # gimbal case
    mats = np.stack([RT.euler2mat(a[0], a[1], a[2]) for a in ang])
    back = np.stack([np.array(RT.mat2euler(m)) for m in mats])
    quats = np.stack([RT.euler2quat(a[0], a[1], a[2]) for a in ang])
    Kl = np.array([[572.4114, 0, 325.2611], [0, 573.57043, 242.04899], [0, 0, 1]])
    obs = np.zeros((3, 3, 4))
    for k in range(3):
        obs[k, :, :3] = rand_rot(rng3)
        obs[k, :, 3] = [rng3.uniform(-0.1, 0.1), rng3.uniform(-0.08, 0.08), rng3.uniform(0.7, 1.1)]
    np.random.seed(2333)
    angle_std, angle_max, x_std, y_std, z_std = [15.0, 45.0, 0.01, 0.01, 0.05]
    ren = np.zeros((3, 4, 3, 4))
    for a in range(3):
        src_pose_m = obs[a]
        src_euler = np.squeeze(RT.mat2euler(src_pose_m[:3, :3]))
        src_trans = src_pose_m[:, 3]
        for k in range(4):
            while True:
                tgt_euler = src_euler + np.random.normal(0, angle_std / 180 * math.pi, 3)
                tgt_trans = src_trans + np.array([np.random.normal(0, x_std, 1)[0], np.random.normal(0, y_std, 1)[0],
                                                  np.random.normal(0, z_std, 1)[0]])
                tgt_pose_m = np.hstack((RT.euler2mat(tgt_euler[0], tgt_euler[1], tgt_euler[2]), tgt_trans.reshape((3, 1))))
                r_dist, t_dist = RT.calc_rt_dist_m(tgt_pose_m, src_pose_m)
                c = np.matmul(Kl, tgt_trans.reshape(3, 1))
                if not (r_dist > angle_max or not (16 < c[0] / c[2] < (640 - 16) and 16 < c[1] / c[2] < (480 - 16))):
                    break
            ren[a, k] = tgt_pose_m
    np.savez_compressed(os.path.join(HERE, "ref_euler.npz"), angles=ang, mats=mats, back=back, quats=quats, K=Kl,
                        poses_observed=obs, poses_rendered=ren)

    # ---- image.transform (lib/utils/image.py:583-594)
    try:
        from lib.utils.image import transform
        im = rng.integers(0, 256, size=(6, 8, 3)).astype(np.uint8)
        pm = np.array([123.68, 116.779, 103.939])
        np.savez_compressed(os.path.join(HERE, "ref_transform.npz"), im=im, pixel_means=pm,
                            out=transform(im, pm).astype(np.float32))
    except Exception as e:  # cv2 / other imports of image.py may be missing
        print("image.transform not importable:", e)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
