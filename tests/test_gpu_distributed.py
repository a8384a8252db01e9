"""GPU, 2 ranks over NCCL (skipped below 2 GPUs): the product's sharded sweep (sweep.sweep_omgw0_distributed: whole v_wall
rows per rank, blocks sent to rank 0 while the rest is computed) against the same sweep on one GPU."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(r_star=0.1, nt=400, n_z_lookup=500, n_xi=1500)
Y = np.logspace(-1, 3, 80)


def _grid():
    v_walls = np.linspace(0.2, 0.9, 12)
    alpha_ns = np.linspace(0.02, 0.45, 25)
    vw, an = np.meshgrid(v_walls, alpha_ns)          # the reference's order: alpha-major, v_wall fastest
    return vw.ravel(), an.ravel()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from pttools_b200 import sweep
    from pttools_b200.models import BagModel
    vw, an = _grid()
    model = BagModel(a_s=1.1, a_b=1, V_s=1)
    out = sweep.sweep_omgw0_distributed(model, vw, an, y=Y, chunk=64, **KW)          # table shared through CUDA IPC
    out = None if out is None else out.cpu().numpy()
    shared = sweep.SharedTable(vw.size, Y.size)                                        # the same, table reused
    for _ in range(2):
        out2 = sweep.sweep_omgw0_distributed(model, vw, an, y=Y, chunk=64, shared=shared, **KW)
    out2 = None if out2 is None else out2.cpu().numpy()
    out3 = sweep.sweep_omgw0_distributed(model, vw, an, y=Y, chunk=64, transport="sendrecv", **KW)   # NCCL messages
    out3 = None if out3 is None else out3.cpu().numpy()
    shard, host = sweep.sweep_omgw0_distributed(model, vw, an, y=Y, chunk=64, mem="host", **KW)
    torch.cuda.synchronize()
    dist.barrier()
    shared.close()
    q.put({"rank": rank, "out": out, "out2": out2, "out3": out3, "shard": np.asarray(shard), "host": np.asarray(host)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_sweep_equals_single_gpu_sweep():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pttools_b200 import sweep
    from pttools_b200.models import BagModel
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=600)
        res[r["rank"]] = r
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    vw, an = _grid()
    one = sweep.sweep_spectra(BagModel(a_s=1.1, a_b=1, V_s=1), vw, an, y=Y, want_pow_v=False, **KW).numpy()
    assert res[1]["out"] is None and res[1]["out2"] is None and res[1]["out3"] is None
    for key in ("out", "out2", "out3"):
        np.testing.assert_allclose(res[0][key], one.omgw0, rtol=1e-10, equal_nan=True)
    assert np.isfinite(one.omgw0).all(axis=1).sum() > 150
    # host-side form: every rank holds its shard; together they are the table
    full = np.full_like(one.omgw0, -1.0)
    for r in (0, 1):
        full[res[r]["shard"]] = res[r]["host"]
        assert len(set(vw[res[r]["shard"]]) & set(vw[res[1 - r]["shard"]])) == 0       # whole v_wall rows per rank
    np.testing.assert_allclose(full, one.omgw0, rtol=1e-10, equal_nan=True)
