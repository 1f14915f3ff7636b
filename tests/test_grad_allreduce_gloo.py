"""CPU, world_size 2 over gloo: bucketed gradient all-reduce of the training row (sum semantics,
rescale_grad = 1.0), bucket coverage and ordering."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_bucket_layout_covers_every_parameter_once():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mx-deepim_b200"))
    from deepim_b200.grad_allreduce import GradBuckets, param_table
    tab = param_table()
    total = sum(n for _, n in tab)
    assert 57.5e6 < total < 58.0e6          # 57.75 M parameters (231 MB fp32)
    gb = GradBuckets("cpu", bucket_mb=32.0)
    cover = np.zeros(gb.numel, np.int8)
    for lo, hi in gb.buckets:
        cover[lo:hi] += 1
    assert cover.min() == 1 and cover.max() == 1
    assert gb.buckets[0][1] == gb.numel     # first bucket = last layers (backward order)
    assert all(gb.buckets[i][0] == gb.buckets[i + 1][1] for i in range(len(gb.buckets) - 1))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mx-deepim_b200"))
    from deepim_b200.grad_allreduce import GradBuckets
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    small = [("a_weight", 1000), ("a_bias", 10), ("b_weight", 5000), ("b_bias", 7), ("c_weight", 333)]
    gb = GradBuckets("cpu", bucket_mb=0.01, table=small)
    g = torch.Generator().manual_seed(rank)
    gb.flat.copy_(torch.randn(gb.numel, generator=g))
    mine = gb.flat.clone()
    for w in gb.allreduce(dist, async_op=True):
        w.wait()
    q.put((rank, mine.numpy(), gb.flat.numpy().copy(), len(gb.buckets)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucketed_sum():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (m, s, nb) for r, m, s, nb in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = res[0][0] + res[1][0]
    assert res[0][2] > 1
    for r in range(world):
        np.testing.assert_allclose(res[r][1], expect, rtol=0, atol=1e-6)
