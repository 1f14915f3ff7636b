"""NCCL all-reduce of the 231 MB fp32 training gradient (run under torchrun on N GPUs):
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mx-deepim_b200"))
import torch, torch.distributed as dist
from deepim_b200 import trainer
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
res = {}
for mb in (8, 32, 231):
    class GB:  # the flat gradient vector + the buckets Trainer.allreduce_overlapped walks (trainer.make_buckets)
        pass
    gb = GB()
    gb.buckets, _ = trainer.make_buckets(float(mb))
    gb.numel = sum(n for _, n in trainer.tensor_sizes())
    gb.flat = torch.empty(gb.numel, device=torch.device("cuda", lr))
    gb.allreduce = lambda d: [d.all_reduce(gb.flat[lo:hi], op=d.ReduceOp.SUM) for lo, hi in gb.buckets]
    gb.flat.normal_()
    for _ in range(3):
        gb.allreduce(dist)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gb.allreduce(dist)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    nbytes = gb.numel * 4
    res["bucket_%dMB" % mb] = {"ms": round(ms, 4), "buckets": len(gb.buckets), "algbw_GBs": round(nbytes / ms / 1e6, 1),
                               "busbw_GBs": round(nbytes / ms / 1e6 * 2 * (world - 1) / world, 1)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "grad_bytes": nbytes, "results": res}))
dist.destroy_process_group()
