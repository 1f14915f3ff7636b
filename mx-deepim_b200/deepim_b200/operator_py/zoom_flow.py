"""ZoomFlow -- mirrors deepim/operator_py/zoom_flow.py (forward l.28-71, prop l.81-99): 3 inputs /
2 outputs when zooming in (flow + flow_weights), 2 / 1 for the inverse zoom."""
from .base import CustomOp, CustomOpProp, parse_bool, register


class ZoomFlowOperator(CustomOp):
    def __init__(self, ctx, height, width, b_inv_zoom):
        self.ctx, self.b_inv_zoom = ctx, b_inv_zoom

    def forward(self, is_train, req, in_data, out_data, aux):
        fw = in_data[2] if not self.b_inv_zoom else None
        zfl, zfw = self.ctx.zoom_flow(in_data[0], in_data[1], fw, self.b_inv_zoom)
        self.assign(out_data[0], req[0], zfl)
        if not self.b_inv_zoom:
            self.assign(out_data[1], req[1], zfw)


@register("ZoomFlow")
class ZoomFlowProp(CustomOpProp):
    def __init__(self, width="640", height="480", b_inv_zoom="False"):
        super().__init__(True)
        self.height, self.width, self.b_inv_zoom = int(height), int(width), parse_bool(b_inv_zoom)

    def list_arguments(self):
        return ["zoom_factor", "flow"] + ([] if self.b_inv_zoom else ["flow_weights"])

    def list_outputs(self):
        return ["zoom_flow"] + ([] if self.b_inv_zoom else ["zoom_flow_weights"])

    def infer_shape(self, in_shape):
        return in_shape, list(in_shape[1:]), []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomFlowOperator(ctx, self.height, self.width, self.b_inv_zoom)
