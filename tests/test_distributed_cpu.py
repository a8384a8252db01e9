"""CPU: the N>1 path of the sweep driver (whole-row sharding + block-wise delivery to the root) with world_size 2 over
gloo."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grid():
    v_walls = np.linspace(0.05, 0.95, 11)
    alpha_ns = np.linspace(0.01, 0.4, 7)
    # the reference's parameter order: params[i_alpha, i_vw] (analysis/parallel.py:126-130), v_wall fastest
    vw, an = np.meshgrid(v_walls, alpha_ns)
    return vw.ravel(), an.ravel()


def _worker(rank, world, port, q, scramble):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from pttools_b200 import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vw, an = _grid()
    cols, block = 5, 8
    cost = 1.0 + (an > 0.3)
    shards = sweep.shard_rows(vw, cost, world)
    calls = []

    def compute(idx, out, on_chunk):
        calls.append(np.asarray(idx))
        # ranges inside blocks of `block` rows, optionally out of order (the retry folding delivers chunks late)
        ranges = [(a, min(a + 4, len(idx))) for a in range(0, len(idx), 4)]
        if scramble:
            ranges = ranges[1::2] + ranges[0::2]
        for a, b in ranges:
            out[a:b] = torch.as_tensor(np.stack([idx[a:b] * 10.0 + c for c in range(cols)], axis=1))
            on_chunk(a, b, None)

    out = sweep.distributed_table(compute, shards, cols, block, device="cpu")
    q.put({"rank": rank, "mine": calls[0].tolist(), "out": None if out is None else out.numpy().tolist()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scramble", [False, True])
def test_two_rank_sweep_reassembles_the_table_in_caller_order(scramble):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + (50 if scramble else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, scramble)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = {r["rank"]: r for r in results}
    vw, an = _grid()
    n, cols = vw.size, 5
    all_idx = sorted(by_rank[0]["mine"] + by_rank[1]["mine"])
    assert all_idx == list(range(n))
    # whole v_wall rows per rank
    for r in (0, 1):
        rows = set(vw[by_rank[r]["mine"]])
        other = set(vw[by_rank[1 - r]["mine"]])
        assert not rows & other
    assert by_rank[1]["out"] is None
    want = np.stack([np.arange(n) * 10.0 + c for c in range(cols)], axis=1)
    np.testing.assert_array_equal(np.array(by_rank[0]["out"]), want)


def test_shard_rows_properties():
    sys.path.insert(0, ROOT)
    from pttools_b200 import sweep
    rng = np.random.default_rng(0)
    v_walls = np.linspace(0.05, 0.95, 1000)
    vw = np.repeat(v_walls, 20)
    codes = rng.integers(-1, 3, vw.size)
    cost = sweep.point_cost(codes)
    parts = sweep.shard_rows(vw, cost, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(vw.size))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.01                   # cost-balanced
    for p in parts:
        assert np.all(np.diff(vw[p]) >= 0)                      # ordered by v_wall: groups of K6' stay contiguous
        for v in np.unique(vw[p]):
            assert np.count_nonzero(vw[p] == v) == 20           # whole rows
    # one rank: the identity up to the v_wall ordering
    one = sweep.shard_rows(vw[::-1].copy(), None, 1)[0]
    assert np.all(np.diff(vw[::-1][one]) >= 0) and one.size == vw.size
