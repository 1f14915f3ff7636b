#!/usr/bin/env python
"""conv_lab -- A/B of conv-tower kernel variants inside ONE process (same box, same thermal state), C2 workload.

For every variant (dim_debug_set_option "pair_mask"): per-layer device times of the conv tower (single stream, CUDA events
around each layer, median over several forwards) and, for the variants listed in --full, the multi-stream throughput of the
whole refinement (the bench's `value` arm, a few steps).  Variants are visited round-robin `--rounds` times.

    python tools/conv_lab.py --out gpurun_out/lab.json
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mx-deepim_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/lab.json")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--step-batches", type=int, default=32)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--masks", default="2,8194,8192,0")
    ap.add_argument("--full", default="2,8194", help="masks that also get the multi-stream throughput pass")
    args = ap.parse_args()
    import torch
    import bench
    from deepim_b200 import _capi as capi
    from deepim_b200 import synth
    from deepim_b200._capi import check, lib
    from deepim_b200.refiner import PoseRefiner

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    K, means, B, N_ITER = synth.K_LINEMOD, synth.PIXEL_MEANS_RGB, 16, 4
    prec = capi.precision_id(args.precision)
    mesh = synth.make_blob()
    refiner = PoseRefiner([mesh], synth.make_weights(0), K, device=0, max_batch=B, n_iter=N_ITER, precision=args.precision, n_slots=4)
    ctxs = [s["ctx"] for s in refiner.slots]
    streams = [s["stream"] for s in refiner.slots]
    sets = bench.make_inputs(ctxs[0], synth, mesh, B, 3, 1000, dev, torch)

    outs = {}

    def batch(i, k):
        s = sets[k % len(sets)]
        outs[i] = ctxs[i].refine(s["img_dev"], s["cls_dev"], s["pose_dev"], K, N_ITER, pixel_means_rgb=means, precision=prec,
                                 out=outs.get(i))
        return outs[i]

    def set_opt(key, v):
        for c in ctxs:
            check(lib.dim_debug_set_option(c._h, key, int(v)))

    def set_mask(m):
        # mask bits 1..9: conv layer on the CTA-pair kernel; bit 11 (2048): no CUDA graph; bit 13 (8192): conv1 on the
        # stacked-filter-rows kernel (else rolling strips)
        set_opt(b"pair_mask", m & 0x3FE)
        set_opt(b"conv1_stack", (m >> 13) & 1)
        set_opt(b"graph", 0 if (m >> 11) & 1 else 1)

    def layer_times(n=7):
        check(lib.dim_debug_layer_profile(ctxs[0]._h, 1, None))
        ms = (C.c_float * 10)()
        acc = []
        for k in range(n):
            batch(0, k)
            check(lib.dim_debug_layer_profile(ctxs[0]._h, 1, ms))
            acc.append([ms[i] * 1e3 for i in range(10)])
        check(lib.dim_debug_layer_profile(ctxs[0]._h, 0, None))
        return [round(float(v), 1) for v in np.median(np.array(acc), axis=0)]

    def throughput(steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for k in range(steps * args.step_batches):
            i = k % len(streams)
            with torch.cuda.stream(streams[i]):
                batch(i, k)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        return B * steps * args.step_batches / (e0.elapsed_time(e1) / 1e3)

    masks = [int(m) for m in args.masks.split(",")]
    full = set(int(m) for m in args.full.split(",") if m != "")
    res = {str(m): {"layers_us": [], "refinements_per_s": []} for m in masks}
    for k in range(3):
        batch(0, k)
    throughput(2)
    for r in range(args.rounds):
        for m in masks:
            set_mask(m)
            batch(0, 0)
            torch.cuda.synchronize()
            res[str(m)]["layers_us"].append(layer_times())
            if m in full:
                throughput(1)
                res[str(m)]["refinements_per_s"].append(round(throughput(args.steps), 1))
            print(m, res[str(m)]["layers_us"][-1], res[str(m)]["refinements_per_s"][-1:], flush=True)
    set_mask(0)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "precision": args.precision,
               "note": "layers_us: conv1, conv2, conv3, conv3_1, conv4, conv4_1, conv5, conv5_1, conv6, conv6_1 (single stream, B=16)",
               "results": res}, open(args.out, "w"), indent=1)
    refiner.close()


if __name__ == "__main__":
    main()
