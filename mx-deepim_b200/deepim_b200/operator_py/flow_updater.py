"""FlowUpdater -- mirrors the op surface of deepim/operator_py/flow_updater.py (prop l.111-125:
depth_src, depth_tgt, pose_src, pose_tgt -> flow [B,2,H,W], flow_weights [B,2,H,W] (the validity plane tiled to both
channels, l.98-99); attrs K, thresh, batch_size, height, width, wh_rep).  The reference registers it but never puts it
in the graph.  NUMERICS FOLLOW lib/flow_c (gpu_flow), not flow_updater.py: the device path is the
reprojection-flow kernel dim_flow_fwd -- sub-pixel flow h_proj - h, depth test at the rounded target pixel with the
fixed 3e-3 threshold (gpu_flow_kernel.cu:45-58) -- where flow_updater.py rounds the projection to integer pixels and
takes `thresh` as an attribute; any other thresh is refused.  KT = K . T_src->tgt is composed on the host (calc_se3,
RT_transform.py:176-187), one small D2H read of the two pose blobs (the reference does the same with asnumpy())."""
import numpy as np
import torch

from .base import CustomOp, CustomOpProp, parse_bool, parse_vec, register


class FlowUpdaterOperator(CustomOp):
    def __init__(self, ctx, K, thresh, wh_rep):
        self.ctx, self.K, self.thresh, self.wh_rep = ctx, K, thresh, wh_rep
        if abs(thresh - 3e-3) > 1e-9:
            raise NotImplementedError("the flow kernel hard-codes thresh=3e-3 like lib/flow_c/gpu_flow_kernel.cu:53")

    def forward(self, is_train, req, in_data, out_data, aux):
        ps = in_data[2].double().cpu().numpy()
        pt = in_data[3].double().cpu().numpy()
        B = ps.shape[0]
        KT = np.zeros((B, 3, 4), np.float32)
        for b in range(B):
            R = pt[b, :, :3] @ ps[b, :, :3].T
            KT[b] = (self.K.astype(np.float64) @ np.hstack([R, (pt[b, :, 3] - R @ ps[b, :, 3])[:, None]])).astype(np.float32)
        Kinv = np.linalg.inv(self.K.astype(np.float64)).astype(np.float32)
        flow, valid = self.ctx.flow(in_data[0], in_data[1], torch.from_numpy(KT).to(in_data[0].device), Kinv)
        if self.wh_rep:  # standard (dw, dh) channel order
            flow = flow.flip(1).contiguous()
        self.assign(out_data[0], req[0], flow)
        self.assign(out_data[1], req[1], valid.expand(-1, 2, -1, -1).contiguous())   # mx.nd.tile(valid, (1,2,1,1)), l.99


@register("FlowUpdater")
class FlowUpdaterProp(CustomOpProp):
    def __init__(self, K, thresh="3e-3", batch_size="1", height="480", width="640", wh_rep="False"):
        super().__init__(False)
        self.K = parse_vec(K, 9).reshape(3, 3)
        self.thresh, self.wh_rep = float(thresh), parse_bool(wh_rep)
        self.batch_size, self.height, self.width = int(batch_size), int(height), int(width)

    def list_arguments(self):
        return ["depth_src", "depth_tgt", "pose_src", "pose_tgt"]

    def list_outputs(self):
        return ["flow", "flow_weights"]

    def infer_shape(self, in_shape):
        b, _, h, w = in_shape[0]
        return in_shape, [[b, 2, h, w], [b, 2, h, w]], []

    def create_operator(self, ctx, shapes, dtypes):
        return FlowUpdaterOperator(ctx, self.K, self.thresh, self.wh_rep)
