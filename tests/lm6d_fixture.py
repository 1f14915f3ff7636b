"""Builds a tiny LM6d_refine-style directory from synthetic meshes with the CPU oracle renderer (test infrastructure)."""
import os

import numpy as np

from deepim_b200 import lm6d_io, synth
from oracle import oracle as O

K = synth.K_LINEMOD


def build(root, n_per_class=2, seed=21):
    classes = ["cube", "glue"]  # 'glue' is in the reference's symmetric list -> scored with ADI
    meshes = {"cube": synth.make_cube(), "glue": synth.make_blob(nlat=24, nlon=48, tex_size=128)}
    os.makedirs(os.path.join(root, "models"), exist_ok=True)
    os.makedirs(os.path.join(root, "image_set"), exist_ok=True)
    with open(os.path.join(root, "models", "models_info.txt"), "w") as f:
        for i, c in enumerate(classes):
            f.write("%d diameter %.6f min_x 0\n" % (i + 1, meshes[c].diameter * 1000.0))
    for ci, c in enumerate(classes):
        m = meshes[c]
        d = os.path.join(root, "models", c)
        lm6d_io.write_textured_obj(m, os.path.join(d, "textured.obj"), os.path.join(d, "texture_map.png"))
        np.savetxt(os.path.join(d, "points.xyz"), m.verts[:: max(1, len(m.verts) // 400)].astype(np.float64))
        obs, ini = synth.sample_pose_pairs(n_per_class, seed + ci)
        lines = []
        for k in range(n_per_class):
            oi, ri = "%02d/%06d" % (ci + 1, k), "%s/%02d_%06d_0" % (c, ci + 1, k)
            for sub in ("observed/%02d" % (ci + 1), "gt_observed/%s" % c, "rendered/%s" % c):
                os.makedirs(os.path.join(root, "data", sub), exist_ok=True)
            r = O.render(m, obs[k], K)
            import cv2
            cv2.imwrite(os.path.join(root, "data", "observed", oi + "-color.png"), synth.composite_observed(r["bgr"], r["mask"], k))
            lm6d_io.write_depth(os.path.join(root, "data", "observed", oi + "-depth.png"), r["depth"])
            cv2.imwrite(os.path.join(root, "data", "observed", oi + "-label.png"), (r["mask"] * (ci + 1)).astype(np.uint8))
            lm6d_io.write_pose(os.path.join(root, "data", "gt_observed", c, oi.split("/")[1] + "-pose.txt"), ci + 1, obs[k])
            rr = O.render(m, ini[k], K)
            cv2.imwrite(os.path.join(root, "data", "rendered", ri + "-color.png"), rr["bgr"].astype(np.uint8))
            lm6d_io.write_depth(os.path.join(root, "data", "rendered", ri + "-depth.png"), rr["depth"])
            cv2.imwrite(os.path.join(root, "data", "rendered", ri + "-label.png"), (rr["mask"] * (ci + 1)).astype(np.uint8))
            lm6d_io.write_pose(os.path.join(root, "data", "rendered", ri + "-pose.txt"), ci + 1, ini[k])
            lines.append("%s %s" % (oi, ri))
        with open(os.path.join(root, "image_set", "val_%s.txt" % c), "w") as f:
            f.write("\n".join(lines) + "\n")
    return classes, meshes
