"""CPU (gloo, world_size 2): the host side of the data-parallel training step -- gradient buckets of the flat parameter
vector (reverse tensor order, readiness index per bucket) and the bucketed sum-all-reduce that replaces kvstore push/pull
(deepim/core/module.py:616-635).  The device-side events that gate each bucket are exercised on the GPU box
(tools/train_bench.py under torchrun)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mx-deepim_b200"))


def test_buckets_cover_the_flat_vector_in_backward_order():
    from deepim_b200 import trainer
    sizes = trainer.tensor_sizes()
    n = sum(m for _, m in sizes)
    assert n == 57749164
    offs, off = {}, 0
    for idx, m in sizes:
        offs[idx] = off
        off += m
    for mb in (8.0, 32.0, 1000.0):
        buckets, first = trainer.make_buckets(mb)
        assert buckets[0][1] == n and buckets[-1][0] == 0
        for (lo, hi), (lo2, hi2) in zip(buckets[:-1], buckets[1:]):
            assert hi2 == lo and lo2 < hi2                          # contiguous, walking backwards
        for (lo, hi), f in zip(buckets, first):
            assert lo == offs[f]                                    # every bucket starts at a tensor boundary ...
        assert first == sorted(first, reverse=True)                 # ... and readiness indices decrease (backward order)
        assert all(hi - lo >= mb * (1 << 20) / 4 for lo, hi in buckets[:-1])
    assert len(trainer.make_buckets(1000.0)[0]) == 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepim_b200 import trainer
    sizes = [(i, m) for i, m in enumerate((1000, 17, 4096, 3, 50000, 9))]      # a small stand-in table
    buckets, first = trainer.make_buckets(0.05, sizes)                         # 0.05 MB = 13107 floats
    n = sum(m for _, m in sizes)
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    for lo, hi in buckets:                                                     # what Trainer.allreduce does per bucket
        dist.all_reduce(g[lo:hi], op=dist.ReduceOp.SUM)
    out[rank] = (g.numpy().copy(), buckets, first)
    dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    port = 29600 + (os.getpid() % 200)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g0, buckets, first = out[0]
    g1 = out[1][0]
    n = len(g0)
    assert np.array_equal(g0, g1) and np.array_equal(g0, np.arange(n, dtype=np.float32) * 3)   # rescale_grad = 1: plain sum
    assert buckets[0][1] == n and buckets[-1][0] == 0 and first[-1] == 0
