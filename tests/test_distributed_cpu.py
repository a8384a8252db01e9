"""CPU: the N>1 path of the sweep driver (sharding + final gather) with world_size 2 over gloo."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from pttools_b200 import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, cols = 37, 5
    sol_type = (np.arange(n) * 7) % 3
    calls = []

    def compute(idx):
        calls.append(idx)
        return np.stack([idx * 10.0 + c for c in range(cols)], axis=1)

    out = sweep.distributed_table(compute, n, cols, sol_type=sol_type, device="cpu")
    mine = calls[0]
    res = {"rank": rank, "mine": mine.tolist(), "out": None if out is None else out.numpy().tolist()}
    q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_reassembles_the_table():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = {r["rank"]: r for r in results}
    n, cols = 37, 5
    # shards are disjoint, cover everything, and are balanced within each solution type
    all_idx = sorted(by_rank[0]["mine"] + by_rank[1]["mine"])
    assert all_idx == list(range(n))
    sol_type = (np.arange(n) * 7) % 3
    for t in range(3):
        c0 = sum(1 for i in by_rank[0]["mine"] if sol_type[i] == t)
        c1 = sum(1 for i in by_rank[1]["mine"] if sol_type[i] == t)
        assert abs(c0 - c1) <= 1
    assert by_rank[1]["out"] is None
    want = np.stack([np.arange(n) * 10.0 + c for c in range(cols)], axis=1)
    np.testing.assert_array_equal(np.array(by_rank[0]["out"]), want)


def test_shard_indices_properties():
    sys.path.insert(0, ROOT)
    from pttools_b200 import sweep
    n = 1001
    st = np.random.default_rng(0).integers(-1, 3, n)
    parts = [sweep.shard_indices(n, r, 8, st) for r in range(8)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(n))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
    for t in (-1, 0, 1, 2):
        counts = [int((st[p] == t).sum()) for p in parts]
        assert max(counts) - min(counts) <= 1
    # without types: plain cyclic deal
    assert sweep.shard_indices(10, 1, 4).tolist() == [1, 5, 9]
