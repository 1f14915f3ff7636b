"""RT_transform -- the compose half of lib/pair_matching/RT_transform.py (RT_transform l.127-151) on
the device (dim_se3_compose, float64), batched.  Same argument meaning: pose_src (3,4) or (B,3,4),
r = quaternion (w,x,y,z, normalised inside), t = (dx, dy, dz), T_means / T_stds, rot_coord."""
import numpy as np
import torch

from .operator_py.base import get_context


def RT_transform(pose_src, r, t, T_means, T_stds, rot_coord="MODEL", ctx=None):
    c = get_context(ctx)
    ps = np.asarray(pose_src, np.float64)
    single = ps.ndim == 2
    ps = ps.reshape(-1, 3, 4)
    r = np.asarray(r, np.float32).reshape(-1, 4)
    if r.shape[1] != 4:
        raise Exception("Unknown r shape: {}".format(r.shape))  # euler input is not used by the shipped config
    se3 = np.concatenate([r, np.asarray(t, np.float32).reshape(-1, 3)], axis=1)
    out = c.se3_compose(torch.from_numpy(ps).to(c.device), torch.from_numpy(se3).to(c.device), T_means, T_stds,
                        rot_coord).cpu().numpy()
    return out[0] if single else out
