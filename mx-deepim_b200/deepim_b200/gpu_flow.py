"""gpu_flow -- numpy-in / numpy-out twin of lib/flow_c/gpu_flow.pyx:gpu_flow (l.24-41):
    flow, valid = gpu_flow(depth_src, depth_tgt, KT, Kinv, device_id)
depth_* (B,1,H,W) float32, KT (B,3,4), Kinv (3,3).  The reference mallocs/frees six device buffers
per call (gpu_flow_kernel.cu:87-147); here the context owns everything."""
import numpy as np
import torch

from .operator_py.base import get_context


def gpu_flow(depth_src, depth_tgt, KT, Kinv, device_id=0, ctx=None):
    c = get_context(ctx)
    d = c.device
    fl, va = c.flow(torch.from_numpy(np.ascontiguousarray(depth_src, np.float32)).to(d),
                    torch.from_numpy(np.ascontiguousarray(depth_tgt, np.float32)).to(d),
                    torch.from_numpy(np.ascontiguousarray(KT, np.float32)).to(d), np.asarray(Kinv, np.float32))
    return fl.cpu().numpy(), va.cpu().numpy()
