"""GPU: batched pred_eval over an LM6d_refine-style directory: disk formats -> PoseRefiner ->
device ADD / ADI -> accuracy / AUC tables."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import lm6d_fixture  # noqa: E402
from oracle import oracle as O  # noqa: E402
from deepim_b200 import lm6d_io, synth  # noqa: E402


def test_evaluate_dataset_directory(tmp_path):
    classes, meshes = lm6d_fixture.build(str(tmp_path), n_per_class=2)
    ds = lm6d_io.LM6DRefine(str(tmp_path), classes, "val")
    w = synth.make_weights(0)
    res, poses, gt = lm6d_io.evaluate(ds, w, synth.K_LINEMOD, n_iter=4, max_batch=3)   # 4 pairs in chunks of 3 + 1
    assert poses.shape == (4, 4, 3, 4) and gt.shape == (4, 3, 4)
    # ground truth / initial poses came through the pose files; errors equal the CPU metric on the same poses
    for ci, c in enumerate(classes):
        pts = ds.points(c)
        for it in range(4):
            for k, j in enumerate(range(2 * ci, 2 * ci + 2)):
                f = O.adi_metric if c == "glue" else O.add_metric
                e = f(poses[it, j][:, :3], poses[it, j][:, 3], gt[j][:, :3], gt[j][:, 3], pts)
                assert abs(res["classes"][ci]["errors"][it][k] - e) < 1e-10
        assert len(res["classes"][ci]["0.10"]) == 4 and 0.0 <= res["classes"][ci]["auc"][0] <= 100.0
    assert len(res["mean"]["auc"]) == 4
    # the meshes loaded from disk (un-rolled OBJ) drive the same refinement as the in-memory ones
    from deepim_b200.refiner import PoseRefiner
    ref = PoseRefiner([meshes[c] for c in classes], w, max_batch=4, n_iter=4)
    imgs = np.stack([ds.load_pair(c, p)["image_observed"] for c in classes for p in ds.pairs(c)])
    init = np.stack([ds.load_pair(c, p)["pose_rendered"] for c in classes for p in ds.pairs(c)])
    direct = ref.refine(imgs, np.array([0, 0, 1, 1], np.int32), init)
    ref.close()
    assert np.abs(direct - poses).max() < 1e-9
