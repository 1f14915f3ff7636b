"""ZoomDepth -- mirrors deepim/operator_py/zoom_depth.py (forward l.24-44)."""
from .base import CustomOp, CustomOpProp, register


class ZoomDepthOperator(CustomOp):
    def __init__(self, ctx, height, width):
        self.ctx = ctx

    def forward(self, is_train, req, in_data, out_data, aux):
        zo, zr = self.ctx.zoom_depth(in_data[0], in_data[1], in_data[2])
        self.assign(out_data[0], req[0], zo)
        self.assign(out_data[1], req[1], zr)


@register("ZoomDepth")
class ZoomDepthProp(CustomOpProp):
    def __init__(self, width="640", height="480"):
        super().__init__(True)
        self.height, self.width = int(height), int(width)

    def list_arguments(self):
        return ["zoom_factor", "depth_observed", "depth_rendered"]

    def list_outputs(self):
        return ["zoom_depth_observed", "zoom_depth_rendered"]

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[1], in_shape[2]], []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomDepthOperator(ctx, self.height, self.width)
