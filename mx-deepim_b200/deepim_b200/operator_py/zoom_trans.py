"""ZoomTrans -- mirrors deepim/operator_py/zoom_trans.py (forward l.22-46, backward l.48-74):
zoom_factor[:,0] scales both x and y (quirk App.B-8)."""
from .base import CustomOp, CustomOpProp, parse_bool, register


class ZoomTransOperator(CustomOp):
    def __init__(self, ctx, b_inv_zoom, b_zoom_grad):
        self.ctx, self.b_inv_zoom, self.b_zoom_grad = ctx, b_inv_zoom, b_zoom_grad

    def forward(self, is_train, req, in_data, out_data, aux):
        self.assign(out_data[0], req[0], self.ctx.zoom_trans(in_data[0], in_data[1], self.b_inv_zoom))

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        self.assign(in_grad[0], req[0], 0)
        self.assign(in_grad[1], req[1],
                    self.ctx.zoom_trans_backward(in_data[0], out_grad[0], self.b_inv_zoom, self.b_zoom_grad))


@register("ZoomTrans")
class ZoomTransProp(CustomOpProp):
    def __init__(self, b_inv_zoom="False", b_zoom_grad="False"):
        super().__init__(True)
        self.b_inv_zoom, self.b_zoom_grad = parse_bool(b_inv_zoom), parse_bool(b_zoom_grad)

    def list_arguments(self):
        return ["zoom_factor", "trans_delta"]

    def list_outputs(self):
        return ["zoom_trans_delta"]

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[1]], []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomTransOperator(ctx, self.b_inv_zoom, self.b_zoom_grad)
