"""Golden fixture from the reference's UNMODIFIED reprojection-flow CUDA kernel (lib/flow_c/gpu_flow_kernel.cu), compiled by
oracle/build_ref.py into oracle/_ref/libgpu_flow_ref.so.  Runs on a GPU box (the .so travels with the snapshot; the
reference tree itself is not needed):

    python tests/golden/make_golden_flow_cuda.py      # writes gpurun_out/ref_flow_cuda.npz -> copy to tests/golden/

Cases: (a) the 60 x 80 depth pair of ref_flow.npz (also pinned to the reference's numpy twin calc_flow); (b) two 480 x 640
instances whose depth maps come from this repo's rasteriser at the poses of synth.sample_pose_pairs(2, 31) -- a rendered
source / target pair as the train loop labels it (batch_updater_py_multi.py:262-300)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mx-deepim_b200"))


def ref_flow(lib, depth_src, depth_tgt, KT, Kinv):
    B, _, H, W = depth_src.shape
    flow = np.zeros((B, 2, H, W), np.float32)
    valid = np.zeros((B, 1, H, W), np.float32)
    fn = getattr(lib, "_Z5_flowPfS_S_S_S_S_iiii")
    fn.restype = None
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    ds, dt = np.ascontiguousarray(depth_src, np.float32), np.ascontiguousarray(depth_tgt, np.float32)
    kt, ki = np.ascontiguousarray(KT, np.float32), np.ascontiguousarray(Kinv, np.float32)
    fn(p(flow), p(valid), p(ds), p(dt), p(kt), p(ki), C.c_int(B), C.c_int(H), C.c_int(W), C.c_int(0))
    return flow, valid


def main():
    import torch
    from deepim_b200 import synth
    from deepim_b200.context import Context
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgpu_flow_ref.so"))
    out = {}
    # (a) small case
    f = np.load(os.path.join(HERE, "ref_flow.npz"))
    Rs, ts, Rt, tt = f["pose_src"][:, :3], f["pose_src"][:, 3], f["pose_tgt"][:, :3], f["pose_tgt"][:, 3]
    T = np.hstack([Rt @ Rs.T, (tt - Rt @ Rs.T @ ts)[:, None]])
    KTs = (f["K"] @ T).astype(np.float32)[None]
    Kinv_s = np.linalg.inv(f["K"]).astype(np.float32)
    fl, va = ref_flow(lib, f["depth_src"][None, None], f["depth_tgt"][None, None], KTs, Kinv_s)
    out.update(small_KT=KTs, small_Kinv=Kinv_s, small_flow=fl, small_valid=va.astype(np.uint8))
    # (b) rendered 480 x 640 pair
    B = 2
    K = synth.K_LINEMOD
    meshes = [synth.make_cube(), synth.make_blob()]
    ctx = Context(0, max_batch=B, max_classes=2, max_verts=6000, max_faces=11000)
    for i, m in enumerate(meshes):
        ctx.upload_mesh(i, m)
    obs, ini = synth.sample_pose_pairs(B, 31)
    dev = torch.device("cuda", 0)
    cls = torch.tensor([0, 1], dtype=torch.int32, device=dev)
    d_src = ctx.render(cls, torch.from_numpy(ini.astype(np.float32)).to(dev), K, want=("depth",))["depth"].cpu().numpy()
    d_tgt = ctx.render(cls, torch.from_numpy(obs.astype(np.float32)).to(dev), K, want=("depth",))["depth"].cpu().numpy()
    K64 = np.asarray(K, np.float64).reshape(3, 3)
    KT = np.zeros((B, 3, 4), np.float32)
    for b in range(B):
        Rrel = obs[b, :, :3] @ ini[b, :, :3].T
        KT[b] = (K64 @ np.hstack([Rrel, (obs[b, :, 3] - Rrel @ ini[b, :, 3])[:, None]])).astype(np.float32)
    Kinv = np.linalg.inv(K64).astype(np.float32)
    fl, va = ref_flow(lib, d_src, d_tgt, KT, Kinv)
    out.update(depth_src=d_src, depth_tgt=d_tgt, KT=KT, Kinv=Kinv, flow=fl, valid=va.astype(np.uint8))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_flow_cuda.npz"), **out)
    print("ref_flow_cuda.npz: small valid", int(out["small_valid"].sum()), "full valid", int(out["valid"].sum()),
          "max |flow|", float(np.abs(fl).max()))
    ctx.close()


if __name__ == "__main__":
    main()
