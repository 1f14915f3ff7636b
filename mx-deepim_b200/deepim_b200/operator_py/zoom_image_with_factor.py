"""ZoomImageWithFactor -- mirrors deepim/operator_py/zoom_image_with_factor.py (forward l.31-65,
prop l.73-104; pixel_means attr is reversed as in l.79-81; high_light_center is a debug aid and is
not supported on the device path)."""
from .base import CustomOp, CustomOpProp, parse_bool, parse_vec, register


class ZoomImageWithFactorOperator(CustomOp):
    def __init__(self, ctx, height, width, pixel_means, high_light_center):
        self.ctx, self.height, self.width, self.pixel_means = ctx, height, width, pixel_means
        if high_light_center:
            raise NotImplementedError("high_light_center (debug overlay) is not implemented")

    def forward(self, is_train, req, in_data, out_data, aux):
        zo, zr = self.ctx.zoom_image_with_factor(in_data[0], in_data[1], in_data[2], self.pixel_means)
        self.assign(out_data[0], req[0], zo)
        self.assign(out_data[1], req[1], zr)


@register("ZoomImageWithFactor")
class ZoomImageWithFactorProp(CustomOpProp):
    def __init__(self, width="640", height="480", pixel_means="[0 0 0]", high_light_center="False"):
        super().__init__(True)
        self.height, self.width = int(height), int(width)
        self.pixel_means = parse_vec(pixel_means, 3)[::-1].copy()
        self.high_light_center = parse_bool(high_light_center)

    def list_arguments(self):
        return ["zoom_factor", "image_observed", "image_rendered"]

    def list_outputs(self):
        return ["zoom_image_observed", "zoom_image_rendered"]

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[1], in_shape[2]], []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomImageWithFactorOperator(ctx, self.height, self.width, self.pixel_means, self.high_light_center)
