"""Transform3D -- mirrors deepim/operator_py/transform3d.py (forward l.34-97, backward l.99-151,
prop l.286-297): point cloud transformed by the composed pose; analytic grads wrt quat / trans."""
from .base import CustomOp, CustomOpProp, parse_bool, parse_vec, register


class transform3dOperator(CustomOp):
    def __init__(self, ctx, T_means, T_stds, rot_coord="MODEL", projection_2d=False):
        assert not projection_2d, "NOT_IMPLEMENTED"  # as in the reference (transform3d.py:32)
        self.ctx, self.T_means, self.T_stds, self.rot_coord = ctx, T_means, T_stds, rot_coord
        if rot_coord.lower() == "naive":
            raise NotImplementedError("rot_coord NAIVE is not used by the shipped configs")

    def forward(self, is_train, req, in_data, out_data, aux):
        pc = in_data[0]
        B = pc.shape[0]
        out = self.ctx.transform3d(pc.reshape(B, 3, -1).contiguous(), in_data[1], in_data[2], in_data[3],
                                   self.T_means, self.T_stds, self.rot_coord)
        self.assign(out_data[0], req[0], out.reshape(pc.shape))

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        pc = in_data[0]
        B = pc.shape[0]
        rg, tg = self.ctx.transform3d_backward(out_grad[0].reshape(B, 3, -1).contiguous(),
                                               pc.reshape(B, 3, -1).contiguous(), in_data[1], in_data[2], in_data[3],
                                               self.T_means, self.T_stds, self.rot_coord)
        self.assign(in_grad[0], req[0], 0)
        self.assign(in_grad[1], req[1], rg)
        self.assign(in_grad[2], req[2], tg)
        self.assign(in_grad[3], req[3], 0)


@register("Transform3D")
class transform3dProp(CustomOpProp):
    def __init__(self, T_means=None, T_stds=None, rot_coord="MODEL", b_project_2d="False"):
        super().__init__(True)
        self.T_means, self.T_stds = parse_vec(T_means, 3), parse_vec(T_stds, 3)
        self._b_project_2d = parse_bool(b_project_2d)
        self.rot_coord = rot_coord

    def list_arguments(self):
        return ["point_cloud", "rotation", "translation", "pose_src"]

    def list_outputs(self):
        return ["transformed_3d_points"]

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[0]], []

    def create_operator(self, ctx, shapes, dtypes):
        return transform3dOperator(ctx, self.T_means, self.T_stds, self.rot_coord, self._b_project_2d)
