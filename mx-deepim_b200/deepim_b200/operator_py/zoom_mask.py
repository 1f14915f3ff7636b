"""ZoomMask -- mirrors deepim/operator_py/zoom_mask.py (forward l.29-112, prop l.121-150).
Device path: dim_zoom_mask_fwd (mask bbox reduction + zoom factor + 3 bilinear gathers, no host sync)."""
from .base import CustomOp, CustomOpProp, parse_vec, register


class ZoomMaskOperator(CustomOp):
    def __init__(self, ctx, K, height, width):
        self.ctx, self.K, self.height, self.width = ctx, K, height, width

    def forward(self, is_train, req, in_data, out_data, aux):
        zo, zg, zr, zf, bbox, status = self.ctx.zoom_mask(in_data[0], in_data[1], in_data[2], in_data[3], self.K)
        if int(status.sum().item()) != 0:  # the reference dies on np.min of an empty array (zoom_mask.py:53)
            raise ValueError("ZoomMask: mask_gt_observed has no valid pixel")
        self.bbox = bbox  # the 8 integer zoom bbox indices per instance
        for dst, r, src in zip(out_data, req, (zo, zg, zr, zf)):
            self.assign(dst, r, src)


@register("ZoomMask")
class ZoomMaskProp(CustomOpProp):
    def __init__(self, K, width="640", height="480"):
        super().__init__(True)
        self.K = parse_vec(K, 9).reshape(3, 3)
        self.height, self.width = int(height), int(width)

    def list_arguments(self):
        return ["mask_observed", "mask_gt_observed", "mask_rendered", "src_pose"]

    def list_outputs(self):
        return ["zoom_mask_observed", "zoom_mask_gt_observed", "zoom_mask_rendered", "zoom_factor"]

    def infer_shape(self, in_shape):
        batch_size = in_shape[0][0]
        return in_shape, list(in_shape[:-1]) + [[batch_size, 4]], []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomMaskOperator(ctx, self.K, self.height, self.width)
