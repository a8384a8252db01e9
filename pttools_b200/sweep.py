"""Sweep-level drivers: whole (v_wall, alpha_n) grids through the device pipeline in chunks, optionally sharded
over the GPUs of one box (one process per GPU, torch.distributed/NCCL only for the final gather of the result table).

This is the performance seam B of the reference (pttools/analysis/parallel.py:89-187 -> speedup/parallel.py:67-196:
one future per parameter point, pickled objects back); here a sweep is a handful of kernel launches per chunk and
the results are struct-of-arrays tables.
"""
from __future__ import annotations

import dataclasses
import typing as tp

import numpy as np

from . import engine

CS0 = engine.CS0


@dataclasses.dataclass
class SweepResult:
    """Result table of a sweep (rows in the order of the input points)."""
    v_wall: tp.Any
    alpha_n: tp.Any
    sol_type: tp.Any
    status: tp.Any
    alpha_plus: tp.Any
    y: np.ndarray
    pow_gw: tp.Any = None
    pow_v: tp.Any = None
    omgw0: tp.Any = None
    scalars: tp.Dict[str, tp.Any] = dataclasses.field(default_factory=dict)

    def numpy(self):
        def cv(x):
            return x.cpu().numpy() if hasattr(x, "cpu") else x
        return SweepResult(**{f.name: (cv(getattr(self, f.name)) if f.name != "scalars"
                                        else {k: cv(v) for k, v in self.scalars.items()})
                              for f in dataclasses.fields(self)})


def shard_indices(n: int, rank: int, world: int, sol_type: tp.Optional[np.ndarray] = None) -> np.ndarray:
    """Static partition of the points of a sweep over ranks.  Points are dealt cyclically WITHIN each solution-type
    bucket so that expensive hybrids/deflagrations and cheap detonations are balanced."""
    idx = np.arange(n)
    if sol_type is None:
        return idx[rank::world]
    order = np.argsort(sol_type, kind="stable")
    return np.sort(order[rank::world])


def bag_classify_host(v_wall: np.ndarray, alpha_n: np.ndarray, an_max: np.ndarray) -> np.ndarray:
    """identify_solution_type_bag on the host (closed form once alpha_n_max_deflagration is known;
    pttools/bubble/transition.py:19-44, alpha.py:192-204)."""
    ap_max_det = np.where(v_wall < CS0, 0.0, (1 - np.sqrt(3) * v_wall) ** 2 / (3 * (1 - v_wall ** 2)))
    out = np.full(v_wall.shape, -1, dtype=np.int32)
    defl = alpha_n < an_max
    out[defl & (v_wall <= CS0)] = 0
    out[defl & (v_wall > CS0)] = 1
    out[alpha_n < ap_max_det] = 2
    return out


def power_gw_scaled_bag_sweep(v_wall, alpha_n, z, npt=(engine.NXI_DEFAULT, engine.NT_DEFAULT, engine.NZ_LOOKUP_DEFAULT),
                              nuc_type="exponential", nuc_args=(1.0,), z_st_thresh=engine.Z_ST_THRESH,
                              chunk=2048, plan: tp.Optional[engine.SsmPlan] = None, want_pow_v=False,
                              an_max=None) -> SweepResult:
    """power_gw_scaled_bag(z, (v_wall, alpha_n, nuc_type, nuc_args), npt) for every point of a sweep
    (pttools/ssmtools/spectrum.py:243-294).  Inputs may be host arrays or device tensors; outputs are device tensors."""
    torch = engine._torch()
    vw = engine.dev_f64(v_wall).reshape(-1)
    an = engine.dev_f64(alpha_n).reshape(-1)
    n = vw.numel()
    z = np.asarray(z, dtype=np.float64)
    if plan is None:
        plan = engine.SsmPlan(z, cs=CS0, nt=npt[1], n_z_lookup=npt[2], nxi=npt[0], z_st_thresh=z_st_thresh,
                              nuc_type=nuc_type, a=nuc_args[0], mode="bag")
    if an_max is None:
        an_max, _ = engine.bag_alpha_n_max_unique(vw)
    dev = vw.device
    pow_gw = torch.empty((n, z.size), dtype=torch.float64, device=dev)
    ap = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    # The shell solve is one thread per point and latency bound: give it as many points as profile memory allows
    # (168-240 KB per shell), the spectrum stages (about 1 MB of scratch per point) run on sub-chunks of it.
    shell_chunk = max(chunk, min(n, shell_chunk_points(npt[0])))
    for s0 in range(0, n, shell_chunk):
        e0 = min(n, s0 + shell_chunk)
        shells = engine.sweep_shells_bag(vw[s0:e0], an[s0:e0], n_xi=npt[0], an_max=an_max[s0:e0])
        ap[s0:e0], st[s0:e0], status[s0:e0] = shells.alpha_plus, shells.sol_type, shells.status
        for s in range(s0, e0, chunk):
            e = min(e0, s + chunk)
            sub = shells.slice(s - s0, e - s0)
            pp = engine.point_params(vw[s:e], CS0, 0.75, 0.75, 0.75 * an[s:e], 0.0)   # quantities.py:38-54, w_n = 1
            res = engine.spectra_from_shells(plan, sub, pp)
            pow_gw[s:e] = res.pow_gw
    return SweepResult(vw, an, st, status, ap, z, pow_gw=pow_gw)


def shell_chunk_points(n_xi: int, budget_bytes: float = 24e9) -> int:
    """How many shells to solve per launch: bounded by the profile rows (3 arrays of row_capacity doubles each)."""
    cap = 5002 + n_xi + 2
    return int(max(1024, min(65536, budget_bytes // (3 * 8 * cap))))
