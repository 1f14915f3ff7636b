"""CPU, world_size 2 over gloo: the N>1 inference path is 'shard instances, no data-path collective,
all-gather the poses'.  The device work is replaced by a deterministic stub so only the host-side
sharding / gather logic (deepim_b200/sharding.py) is exercised."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "mx-deepim_b200"))
    from deepim_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n, rank, world)
    # stub "refine": pose[it, i] = i*100 + it (so ordering errors are visible)
    local = np.zeros((4, hi - lo, 3, 4))
    for it in range(4):
        for i in range(lo, hi):
            local[it, i - lo] = i * 100 + it
    full = sharding.gather_results(local, n, axis=1, dist=dist)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 16])
def test_two_rank_shard_and_gather(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.zeros((4, n, 3, 4))
    for it in range(4):
        for i in range(n):
            expect[it, i] = i * 100 + it
    for r in range(world):
        assert np.array_equal(res[r], expect)
