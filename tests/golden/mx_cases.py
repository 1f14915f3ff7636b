"""Seeded inputs of the reference-pinned operator fixtures (shared by make_golden_mx.py, which feeds them to the LIVE
reference operators, and by the tests, which feed the same arrays to the oracle / the CUDA library).

Geometry: a SMALL frame (96 x 128, K scaled by 1/5) keeps the committed planes small; one FULL-SIZE case (480 x 640) pins
the 8 integer zoom bbox indices + zoom_factor + the zoomed masks at the reference's own resolution."""
import numpy as np

K_FULL = np.array([[572.4114, 0.0, 325.2611], [0.0, 573.57043, 242.04899], [0.0, 0.0, 1.0]], np.float64)  # config.py:58
SMALL = (96, 128)
FULL = (480, 640)
PIXEL_MEANS_RGB = np.array([103.939, 116.779, 123.68], np.float32)  # the zoom ops' reversed pixel_means attr (yaml:57-59)


def K_for(H, W):
    s = H / 480.0
    K = K_FULL.copy()
    K[:2] *= s
    return K.astype(np.float32)


def k_attr(K):
    """the `K` string attribute as the symbol file writes it: a bracketed, space separated list (np.fromstring(K[1:-1]))"""
    return "[" + " ".join(repr(float(v)) for v in np.asarray(K, np.float32).reshape(-1)) + "]"


def vec_attr(v):
    return "[" + " ".join(repr(float(x)) for x in np.asarray(v).reshape(-1)) + "]"


def _blob_mask(rng, H, W, cx, cy, rx, ry):
    yy, xx = np.mgrid[0:H, 0:W]
    ang = np.arctan2(yy - cy, xx - cx)
    r = 1.0 + 0.25 * np.sin(3 * ang + rng.uniform(0, 6)) + 0.15 * np.cos(5 * ang + rng.uniform(0, 6))
    return ((((xx - cx) / (rx * r)) ** 2 + ((yy - cy) / (ry * r)) ** 2) <= 1.0)


def zoom_case(seed, B, H, W):
    """ragged object masks, a rendered 'mask' given as a depth-like image (re-thresholded at 0.2 by the ops), poses whose
    projected centre lies inside the object, images = textured noise inside the masks minus the means, flow, weights, depth"""
    rng = np.random.default_rng(seed)
    K = K_for(H, W)
    s = H / 480.0
    mo = np.zeros((B, 1, H, W), np.float32)
    mr = np.zeros((B, 1, H, W), np.float32)
    pose = np.zeros((B, 3, 4), np.float32)
    img_o = np.zeros((B, 3, H, W), np.float32)
    img_r = np.zeros((B, 3, H, W), np.float32)
    for b in range(B):
        z = rng.uniform(0.55, 1.1)
        cx, cy = rng.uniform(0.3 * W, 0.7 * W), rng.uniform(0.3 * H, 0.7 * H)
        rx, ry = rng.uniform(30, 110) * s / z * 0.8, rng.uniform(30, 90) * s / z * 0.8
        m_o = _blob_mask(rng, H, W, cx, cy, rx, ry)
        dx, dy = rng.uniform(-12, 12) * s, rng.uniform(-12, 12) * s
        m_r = _blob_mask(rng, H, W, cx + dx, cy + dy, rx * rng.uniform(0.85, 1.15), ry * rng.uniform(0.85, 1.15))
        if b == B - 1:  # object cut by the frame border
            m_r[:, : W // 3] = False
            m_o[: H // 4] = False
        mo[b, 0] = m_o
        mr[b, 0] = np.where(m_r, rng.uniform(0.25, 1.2, size=(H, W)), rng.uniform(0.0, 0.15, size=(H, W)) * (rng.uniform(size=(H, W)) > 0.97))
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pose[b, :, :3] = R * np.sign(np.linalg.det(R))
        pose[b, :, 3] = [(cx + dx - K[0, 2]) / K[0, 0] * z, (cy + dy - K[1, 2]) / K[1, 1] * z, z]
        tex = rng.integers(0, 256, size=(3, H, W)).astype(np.float32)
        img_o[b] = np.where(m_o[None], tex, 0.0) - PIXEL_MEANS_RGB[:, None, None]
        img_r[b] = np.where(m_r[None], np.roll(tex, 3, axis=2), 0.0) - PIXEL_MEANS_RGB[:, None, None]
    flow = (rng.normal(size=(B, 2, H, W)) * 6).astype(np.float32)
    fw = (rng.uniform(size=(B, 1, H, W)) > 0.4).astype(np.float32)
    depth = np.where(mr > 0.2, mr + 0.4, 0.0).astype(np.float32)
    return dict(K=K, mo=mo, mr=mr, pose=pose, img_o=img_o.astype(np.float32), img_r=img_r.astype(np.float32), flow=flow,
                fw=np.tile(fw, (1, 2, 1, 1)), fw1=fw, depth=depth)


def t3d_case(seed, B=4, N=300):
    """Transform3D inputs in the style of the reference's self-test (transform3d.py:311-372): points ~ object scale, a
    quaternion near identity (normalised: the op returns identity / zero gradient otherwise -- instance 3 is left
    un-normalised to pin exactly that quirk), small translation deltas, T_means / T_stds non-trivial."""
    rng = np.random.default_rng(seed)
    pts = (rng.normal(size=(B, 3, N)) * 0.05).astype(np.float32)
    q = rng.normal(size=(B, 4)) * 0.15 + np.array([1.0, 0, 0, 0])
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[B - 1] *= 1.05                                     # |q|^2 - 1 = 0.1: quat2mat_forward returns identity
    t = (rng.normal(size=(B, 3)) * 0.05).astype(np.float32)
    pose_src = np.zeros((B, 3, 4), np.float32)
    for b in range(B):
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pose_src[b, :, :3] = R * np.sign(np.linalg.det(R))
        pose_src[b, :, 3] = [rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.5, 1.2)]
    og = rng.normal(size=(B, 3, N)).astype(np.float32)
    return dict(pts=pts, q=q.astype(np.float32), t=t, pose_src=pose_src, og=og,
                T_means=np.array([0.01, -0.02, 0.03], np.float32), T_stds=np.array([0.5, 0.7, 1.3], np.float32))
