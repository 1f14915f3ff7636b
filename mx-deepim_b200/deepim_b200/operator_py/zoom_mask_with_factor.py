"""ZoomMaskWithFactor -- mirrors deepim/operator_py/zoom_mask_with_factor.py (forward l.29-64)."""
from .base import CustomOp, CustomOpProp, parse_bool, register


class ZoomMaskWithFactorOperator(CustomOp):
    def __init__(self, ctx, height, width, b_inv_zoom):
        self.ctx, self.b_inv_zoom = ctx, b_inv_zoom

    def forward(self, is_train, req, in_data, out_data, aux):
        self.assign(out_data[0], req[0], self.ctx.zoom_mask_with_factor(in_data[0], in_data[1], self.b_inv_zoom))


@register("ZoomMaskWithFactor")
class ZoomMaskWithFactorProp(CustomOpProp):
    def __init__(self, width="640", height="480", b_inv_zoom="False"):
        super().__init__(True)
        self.height, self.width, self.b_inv_zoom = int(height), int(width), parse_bool(b_inv_zoom)

    def list_arguments(self):
        return ["zoom_factor", "mask"]

    def list_outputs(self):
        return ["zoom_mask"]

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[1]], []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomMaskWithFactorOperator(ctx, self.height, self.width, self.b_inv_zoom)
