"""One small training step (B=1: forward, backward, update, forward) -- used under compute-sanitizer."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mx-deepim_b200"))
from deepim_b200 import synth  # noqa: E402
from deepim_b200.context import Context  # noqa: E402
from deepim_b200.trainer import Trainer, fit_batch, make_device_batch  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
meshes = [synth.make_cube()]
ctx = Context(0, max_batch=B, max_classes=1, max_verts=len(meshes[0].verts), max_faces=len(meshes[0].faces))
ctx.upload_mesh(0, meshes[0])
tr = Trainer(ctx, synth.make_train_weights(0))
batch, cls, tgt, depth = make_device_batch(ctx, meshes, B, 3, synth.K_LINEMOD, synth.PIXEL_MEANS_RGB)
objs = fit_batch(tr, batch, cls, tgt, depth, synth.K_LINEMOD, n_inner=2)
torch.cuda.synchronize()
print("objectives", [float(v) for v in objs.cpu()], "finite", bool(torch.isfinite(objs).all()))
