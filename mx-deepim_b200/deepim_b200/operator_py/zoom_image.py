"""ZoomImage -- mirrors deepim/operator_py/zoom_image.py (forward l.26-107, prop l.117-150): the zoom front end
of the INPUT_MASK: False graphs; boxes come from sum_c(image + pixel_mean) > 0.01, centre = projected src_pose
translation.  Device path: dim_zoom_image_fwd."""
from .base import CustomOp, CustomOpProp, parse_vec, register


class ZoomImageOperator(CustomOp):
    def __init__(self, ctx, K, height, width, pixel_means):
        self.ctx, self.K, self.height, self.width, self.pixel_means = ctx, K, height, width, pixel_means

    def forward(self, is_train, req, in_data, out_data, aux):
        zo, zr, zf, bbox, status = self.ctx.zoom_image(in_data[0], in_data[1], in_data[2], self.K, self.pixel_means)
        if int(status.sum().item()) != 0:  # the reference dies on np.min of an empty array (zoom_image.py:49)
            raise ValueError("ZoomImage: image_observed has no valid pixel")
        self.bbox = bbox
        for dst, r, src in zip(out_data, req, (zo, zr, zf)):
            self.assign(dst, r, src)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        for g, r in zip(in_grad, req):
            self.assign(g, r, 0)


@register("ZoomImage")
class ZoomImageProp(CustomOpProp):
    def __init__(self, K, width="640", height="480", pixel_means="[0 0 0]"):
        super().__init__(True)
        self.K = parse_vec(K, 9).reshape(3, 3)
        self.height, self.width = int(height), int(width)
        self.pixel_means = parse_vec(pixel_means, 3)[::-1].copy()  # reversed like zoom_image.py:123-125

    def list_arguments(self):
        return ["image_observed", "image_rendered", "src_pose"]

    def list_outputs(self):
        return ["zoom_image_observed", "zoom_image_rendered", "zoom_factor"]

    def infer_shape(self, in_shape):
        batch_size = in_shape[0][0]
        return in_shape, [in_shape[0], in_shape[1], [batch_size, 4]], []

    def create_operator(self, ctx, shapes, dtypes):
        return ZoomImageOperator(ctx, self.K, self.height, self.width, self.pixel_means)
