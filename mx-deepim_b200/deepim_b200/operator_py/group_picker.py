"""GroupPicker -- mirrors deepim/operator_py/group_picker.py (forward l.22-40, backward l.42-56, prop l.61-83):
picks the channel group `group_idx[b]` of every instance (per-class regressors); backward scatters the gradient
into that group.  Device path: dim_group_picker (no host read of group_idx)."""
import numpy as np

from .base import CustomOp, CustomOpProp, register


class GroupPickerOperator(CustomOp):
    def __init__(self, ctx, group_num):
        self.ctx, self.group_num = ctx, group_num

    def forward(self, is_train, req, in_data, out_data, aux):
        self.assign(out_data[0], req[0], self.ctx.group_picker(in_data[0], in_data[1], self.group_num, backward=False))

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        self.assign(in_grad[0], req[0], self.ctx.group_picker(out_grad[0], in_data[1], self.group_num, backward=True,
                                                              channels=in_data[0].shape[1]))
        self.assign(in_grad[1], req[1], 0)


@register("GroupPicker")
class GroupPickerProp(CustomOpProp):
    def __init__(self, group_num):
        super().__init__(True)
        self.group_num = int(group_num)

    def list_arguments(self):
        return ["input_data", "group_idx"]

    def list_outputs(self):
        return ["picked_data"]

    def infer_shape(self, in_shape):
        out = list(np.copy(in_shape[0]))
        out[1] = out[1] // self.group_num
        return in_shape, [out], []

    def create_operator(self, ctx, shapes, dtypes):
        return GroupPickerOperator(ctx, self.group_num)
