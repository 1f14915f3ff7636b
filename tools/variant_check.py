#!/usr/bin/env python
"""variant_check -- kernel variants of the conv tower must reproduce the default kernels: same accumulation order per output
element, so conv outputs are compared BITWISE (activation buffers + regressed se3) for every precision mode and two batch
sizes; also exercises the CUDA-graph replay of the refinement chain against the eager chain.  Exits non-zero on any mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mx-deepim_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from deepim_b200 import _capi as capi  # noqa: E402
from deepim_b200 import synth  # noqa: E402
from deepim_b200._capi import check, lib  # noqa: E402
from deepim_b200.context import Context  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    fails = 0
    w = synth.make_weights(0)
    for B in (3, 16):
        ctx = Context(0, max_batch=B, max_classes=1, max_verts=6000, max_faces=11000)
        ctx.upload_mesh(0, synth.make_blob())
        ctx.load_weights(w)
        g = torch.Generator(device=dev); g.manual_seed(B)
        zio = (torch.rand((B, 3, 480, 640), generator=g, device=dev) * 255 - 110).contiguous()
        zir = (torch.rand((B, 3, 480, 640), generator=g, device=dev) * 255 - 110).contiguous()
        zmo = (torch.rand((B, 1, 480, 640), generator=g, device=dev) > 0.5).float().contiguous()
        zmr = (torch.rand((B, 1, 480, 640), generator=g, device=dev) > 0.5).float().contiguous()
        for prec_name in ("fp16", "bf16", "bf16x3"):
            prec = capi.precision_id(prec_name)

            def run(opts):
                for k, v in opts.items():
                    check(lib.dim_debug_set_option(ctx._h, k, v))
                rot, trans = ctx.net_forward(zio, zir, zmo, zmr, prec)
                torch.cuda.synchronize()
                acts = [ctx.debug_activation(i, B)[0].copy() for i in (1, 2, 3, 10)]
                return rot.cpu().numpy(), trans.cpu().numpy(), acts

            base = run({b"pair_mask": 0, b"conv1_stack": 0})
            for name, opts in (("conv2_pair", {b"pair_mask": 2}), ("conv2,3 pair", {b"pair_mask": 6}),
                               ("stack", {b"conv1_stack": 1, b"pair_mask": 0}), ("stack+conv2 pair", {b"conv1_stack": 1, b"pair_mask": 2})):
                got = run(opts)
                ok = all(np.array_equal(a, b) for a, b in zip(base[2], got[2])) and np.array_equal(base[0], got[0]) and np.array_equal(base[1], got[1])
                d = max(float(np.abs(a - b).max()) for a, b in zip(base[2], got[2]))
                print("B=%d %-6s %-18s %s  max|dact|=%.3g  max|drot|=%.3g" % (B, prec_name, name, "bitwise-equal" if ok else "DIFFERENT", d,
                                                                              float(np.abs(base[0] - got[0]).max())), flush=True)
                fails += 0 if ok else 1
            run({b"pair_mask": 2, b"conv1_stack": 1})
        # CUDA graph replay vs eager chain
        K, means = synth.K_LINEMOD, synth.PIXEL_MEANS_RGB
        obs, ini = synth.sample_pose_pairs(B, 5)
        cls = torch.zeros(B, dtype=torch.int32, device=dev)
        r = ctx.render(cls, torch.from_numpy(obs.astype(np.float32)).to(dev), K, want=("image",))
        img = r["image"].contiguous()
        pose = torch.from_numpy(ini).to(dev)
        check(lib.dim_debug_set_option(ctx._h, b"graph", 0))
        eager = ctx.refine(img, cls, pose, K, 4, pixel_means_rgb=means)
        check(lib.dim_debug_set_option(ctx._h, b"graph", 1))
        out = ctx.refine(img, cls, pose, K, 4, pixel_means_rgb=means)   # legacy default stream: never captured, must just work
        torch.cuda.synchronize()
        fails += 0 if torch.equal(eager["poses"], out["poses"]) else 1
        out = None
        side = torch.cuda.Stream(device=dev)
        for it in range(4):  # 1st call eager warm-up, 2nd captures + launches, 3rd / 4th replay
            with torch.cuda.stream(side):
                out = ctx.refine(img, cls, pose, K, 4, pixel_means_rgb=means, out=out)
            torch.cuda.synchronize()
            ok = all(torch.equal(eager[k], out[k]) for k in ("poses", "se3", "zoom_factor", "bbox"))
            print("B=%d graph call %d: %s" % (B, it, "bitwise-equal to eager" if ok else "DIFFERENT"), flush=True)
            fails += 0 if ok else 1
        ctx.close()
    print("variant_check:", "OK" if fails == 0 else "%d FAILURES" % fails)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
