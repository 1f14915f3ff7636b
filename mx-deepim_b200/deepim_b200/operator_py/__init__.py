"""Host-side mirror of deepim/operator_py/: the same registered op names, argument / output lists,
string-typed attributes and forward/backward protocol as the reference's mx.operator.CustomOp classes
(SURVEY 8(b)), executing through libdeepim_b200.so on torch CUDA tensors.

    op = create("ZoomMask", K="[572.4114 0 325.2611 0 573.57043 242.04899 0 0 1]", height="480", width="640")
    op.forward(is_train=False, req=["write"]*4, in_data=[...], out_data=[...], aux=[])

`in_data` / `out_data` are lists of contiguous float32 CUDA tensors (the borrowed-NDArray handles of
the MXNet protocol); `assign(dst, req, src)` honours req in {null, write, inplace, add}.
An MXNet adapter only needs to wrap NDArrays as torch tensors via DLPack.
"""
from .base import CustomOp, CustomOpProp, REGISTRY, create, register, set_default_context  # noqa: F401
from . import zoom_mask, zoom_image_with_factor, zoom_mask_with_factor, zoom_flow, zoom_trans  # noqa: F401
from . import zoom_depth, transform3d, flow_updater, zoom_image, group_picker  # noqa: F401
