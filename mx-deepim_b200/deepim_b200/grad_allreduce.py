"""Gradient all-reduce for the data-parallel training step (training row).

The reference pushes/pulls every parameter through MXNet's kvstore='device' after each of the 4 inner
iterations (deepim/core/module.py:616-635; 57.75 M parameters = 231 MB fp32, P2P disabled by env).  The
B200 equivalent is one process per GPU and an NCCL all-reduce (sum: rescale_grad = 1.0, train.py:302)
over NVLink/NVSwitch on a flat gradient buffer cut into buckets in reverse layer order, so that a bucket
can be reduced as soon as the backward pass has produced it.  This module owns the flat buffer and the
bucketing; the backward kernels that fill it are the next round's work.
"""
from __future__ import annotations

import numpy as np
import torch

# (name, shape) of every trainable tensor of the training graph, forward order
# (deepim/symbols/deepIM_flownet.py:63-167,716-717; decoder shapes SURVEY 8(a) row a10)
PARAM_SHAPES = [
    ("flow_conv1", (64, 8, 7, 7)), ("conv2", (128, 64, 5, 5)), ("conv3", (256, 128, 5, 5)), ("conv3_1", (256, 256, 3, 3)),
    ("conv4", (512, 256, 3, 3)), ("conv4_1", (512, 512, 3, 3)), ("conv5", (512, 512, 3, 3)), ("conv5_1", (512, 512, 3, 3)),
    ("conv6", (1024, 512, 3, 3)), ("conv6_1", (1024, 1024, 3, 3)), ("fc6", (256, 81920)), ("fc7", (256, 256)),
    ("rot", (4, 256)), ("trans", (3, 256)),
    ("Convolution1", (2, 1024, 3, 3)), ("deconv5", (1024, 512, 4, 4)), ("upsample_flow6to5", (2, 2, 4, 4)),
    ("Convolution2", (2, 1026, 3, 3)), ("deconv4", (1026, 256, 4, 4)), ("upsample_flow5to4", (2, 2, 4, 4)),
    ("Convolution3", (2, 770, 3, 3)), ("mask_conv3", (1, 770, 3, 3)),
]


def param_table(with_bias=True):
    out = []
    for name, shp in PARAM_SHAPES:
        out.append((name + "_weight", int(np.prod(shp))))
        if with_bias:
            nb = shp[1] if name.startswith("deconv") or name.startswith("upsample") else shp[0]
            out.append((name + "_bias", nb))
    return out


class GradBuckets:
    """Flat fp32 gradient buffer + reverse-order buckets of ~bucket_mb each."""

    def __init__(self, device, bucket_mb=32.0, table=None, dtype=torch.float32):
        self.table = table or param_table()
        self.offsets, off = {}, 0
        for name, n in self.table:
            self.offsets[name] = (off, n)
            off += n
        self.numel = off
        self.flat = torch.zeros(off, dtype=dtype, device=device)
        # buckets are formed walking the parameters backwards (the order backward produces them)
        limit = int(bucket_mb * (1 << 20) / self.flat.element_size())
        self.buckets, hi = [], off
        lo = off
        for name, n in reversed(self.table):
            lo -= n
            if hi - lo >= limit:
                self.buckets.append((lo, hi))
                hi = lo
        if hi > 0:
            self.buckets.append((0, hi))

    def view(self, name):
        off, n = self.offsets[name]
        return self.flat[off:off + n]

    def allreduce(self, dist, async_op=False):
        """Sum-all-reduce every bucket (NCCL on GPUs, gloo in tests).  Returns the work handles."""
        works = []
        for lo, hi in self.buckets:
            works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=async_op))
        return works
