#!/usr/bin/env python
"""train_synth -- trains the network with THIS repo's training step (dim_train_forward_backward + SGD, row a10) on synthetic
rendered pairs of the C2 mesh and checks that the 4-iteration refinement then REDUCES the pose error: ADD before / after,
accuracy at 0.02 / 0.05 / 0.10 d on held-out pairs, evaluated with the inference path (DIM_PREC_FP16 and the other modes).
Recipe = seed + hyper-parameters below (no checkpoint is committed: ~230 MB).  Writes one JSON document.

    python tools/train_synth.py --steps 300 --batch 16 --lr 1e-4 --out gpurun_out/train_synth.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mx-deepim_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def add_err(pts, p, q):
    return float(np.linalg.norm((pts @ p[:, :3].T + p[:, 3]) - (pts @ q[:, :3].T + q[:, 3]), axis=1).mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300, help="data batches; each = 4 inner updates (module.py:1131-1137)")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--momentum", type=float, default=0.975)
    ap.add_argument("--wd", type=float, default=5e-4)
    ap.add_argument("--eval-every", type=int, default=100)
    ap.add_argument("--eval-n", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/train_synth.json")
    ap.add_argument("--save", default="", help="optional .npz path for the trained inference weights")
    ap.add_argument("--mask", default="box_gt", choices=["box_gt", "box_rendered"],
                    help="mask_observed convention during training: the reference's training config (box of the GT mask) or the "
                         "test-time convention (box of the rendered mask, refreshed every inner iteration)")
    ap.add_argument("--overfit", action="store_true", help="train on ONE fixed batch and evaluate on exactly those pairs (sanity: "
                    "train graph and test graph agree on every convention)")
    ap.add_argument("--clean-eval", action="store_true", help="evaluate on clean renders (black background) like the training pairs")
    args = ap.parse_args()
    import torch
    from deepim_b200 import _capi as capi
    from deepim_b200 import synth, trainer
    from deepim_b200.context import Context

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    K, means = synth.K_LINEMOD, synth.PIXEL_MEANS_RGB
    mesh = synth.make_blob()
    B = args.batch
    ctx = Context(0, max_batch=max(B, 16), max_classes=1, max_verts=len(mesh.verts), max_faces=len(mesh.faces))
    ctx.upload_mesh(0, mesh)
    tr = trainer.Trainer(ctx, synth.make_train_weights(args.seed), lr=args.lr, momentum=args.momentum, wd=args.wd)
    pts = mesh.verts.astype(np.float64)

    # held-out evaluation pairs: observed = render composited over noise (what the refiner sees at test time)
    if args.overfit:
        args.eval_n = B
    obs, ini = synth.sample_pose_pairs(args.eval_n, (10_000 + 1) if args.overfit else 900001)
    ev = []
    for a in range(0, args.eval_n, 16):
        n = min(16, args.eval_n - a)
        cls = torch.zeros(n, dtype=torch.int32, device=dev)
        r = ctx.render(cls, torch.from_numpy(obs[a:a + n].astype(np.float32)).to(dev), K, want=("bgr", "mask"))
        g = torch.Generator(device=dev); g.manual_seed(a)
        bg = torch.randint(0, 256, r["bgr"].shape, generator=g, device=dev, dtype=torch.int32).to(torch.uint8)
        if args.overfit or args.clean_eval:
            bg = torch.zeros_like(bg)
        u8 = torch.where(r["mask"].permute(0, 2, 3, 1) > 0, r["bgr"].to(torch.uint8), bg).contiguous()
        ev.append((ctx.transform_image_u8(u8, means), cls, torch.from_numpy(ini[a:a + n]).to(dev), a, n))

    def evaluate(prec_name="fp16"):
        prec = capi.precision_id(prec_name)
        errs = np.zeros((5, args.eval_n))
        for img, cls, pose0, a, n in ev:
            res = ctx.refine(img, cls, pose0, K, 4, pixel_means_rgb=means, precision=prec)
            poses = res["poses"].cpu().numpy()
            for b in range(n):
                errs[0, a + b] = add_err(pts, ini[a + b], obs[a + b])
                for it in range(4):
                    errs[1 + it, a + b] = add_err(pts, poses[it, b], obs[a + b])
        d = mesh.diameter
        return {"add_mean_m": [round(float(e.mean()), 6) for e in errs],
                "add_median_m": [round(float(np.median(e)), 6) for e in errs],
                "acc_pct_at_0.02_0.05_0.10_d": [[round(100.0 * float((e < f * d).mean()), 2) for f in (0.02, 0.05, 0.10)] for e in errs],
                "rows": "initial pose, then after iteration 1..4"}

    log = {"recipe": {k: getattr(args, k) for k in ("steps", "batch", "lr", "momentum", "wd", "seed", "mask", "overfit")},
           "workload": "C2 mesh (%d verts), training pairs from synth.sample_pose_pairs(seed = step), 4 inner updates per batch" % len(mesh.verts),
           "diameter_m": float(mesh.diameter), "evals": [], "loss": []}
    log["evals"].append({"step": 0, **evaluate()})
    print("step 0", log["evals"][-1]["add_mean_m"], flush=True)
    t0 = time.time()
    for step in range(1, args.steps + 1):
        batch, cls, tgt, depth_gt = trainer.make_device_batch(ctx, [mesh], B, 10_000 + (1 if args.overfit else step), K, means,
                                                              init_mask=args.mask)
        objs = trainer.fit_batch(tr, batch, cls, tgt, depth_gt, K, n_inner=4,
                                 update_mask="box_rendered" if args.mask == "box_rendered" else "fixed")
        if step % 10 == 0 or step == 1:
            o = [round(float(v), 5) for v in objs.cpu().numpy()]
            log["loss"].append({"step": step, "objective_per_inner_iteration": o})
            if not np.isfinite(o).all():
                print("non-finite objective at step", step, o, flush=True)
                break
        if step % args.eval_every == 0 or step == args.steps:
            torch.cuda.synchronize()
            log["evals"].append({"step": step, "wall_s": round(time.time() - t0, 1), **evaluate()})
            print("step", step, log["loss"][-1]["objective_per_inner_iteration"], log["evals"][-1]["add_mean_m"],
                  log["evals"][-1]["acc_pct_at_0.02_0.05_0.10_d"][-1], flush=True)
    log["final_by_precision"] = {p: evaluate(p) for p in ("fp16", "bf16x3", "bf16")}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(log, open(args.out, "w"), indent=1)
    if args.save:
        w = tr.get_params()
        np.savez(args.save, **{k: v for k, v in w.items()})
    ctx.close()


if __name__ == "__main__":
    main()
