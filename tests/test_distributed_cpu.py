"""world_size-2 gloo tests (CPU) of the batch sharding and of the final gather used by bench.py --gpus N."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cafe_mpc_b200.distributed import gather_records, shard_range


def test_shard_ranges_cover_the_batch():
    for B in (1, 7, 4096, 4097):
        for G in (1, 2, 4, 8):
            got = []
            for r in range(G):
                lo, hi = shard_range(B, G, r)
                assert 0 <= lo <= hi <= B
                got += list(range(lo, hi))
            assert got == list(range(B))


def _worker(rank, world, port, B, rec, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(B, world, rank)
    local = torch.arange(lo * rec, hi * rec, dtype=torch.float64).reshape(hi - lo, rec)
    out = gather_records(local, B, world, rank)
    if rank == 0:
        q.put(out.numpy())
    dist.destroy_process_group()


def test_gather_to_rank0_world2_ragged():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    B, rec = 5, 3  # ragged: shards of 3 and 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, B, rec, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = q.get(timeout=120)
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(out, np.arange(B * rec, dtype=np.float64).reshape(B, rec))
