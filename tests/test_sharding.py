"""CPU: the clip-sharding logic of the multi-GPU path (world_size 2, gloo): every clip is
owned by exactly one rank and shard-wise oracle features concatenate to the whole batch."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speechrecognitionproject_b200.sharding import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1000, 262144):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank: int, world: int, port: int, n_clips: int, q):
    import oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b0, b1 = shard_range(n_clips, rank, world)
    x = oracle.synthetic_corpus(b1 - b0, config_index=4, start=b0)       # each rank generates only its shard
    feats = np.stack([oracle.mfcc_ref(c) for c in x]) if b1 > b0 else np.zeros((0, 39, 51), np.float32)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([b1 - b0]))
    # max-over-ranks timing reduction, as bench.py does it
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, b0, b1, feats, [int(s) for s in sizes], float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_concatenate():
    import oracle
    world, n_clips = 2, 5
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[4] for r in res] == [[3, 2], [3, 2]] and all(r[5] == 2.0 for r in res)
    whole = np.stack([oracle.mfcc_ref(c) for c in oracle.synthetic_corpus(n_clips, config_index=4)])
    np.testing.assert_array_equal(np.concatenate([r[3] for r in res]), whole)
