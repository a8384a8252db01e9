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


def sweep_spectra(model, v_wall, alpha_n, y=None, r_star=None, lifetime_multiplier=1.0, nuc_type="exponential",
                  nt=engine.NT_DEFAULT, n_z_lookup=engine.NZ_LOOKUP_DEFAULT, z_st_thresh=engine.Z_ST_THRESH,
                  n_xi=engine.N_XI_DEFAULT, Tn=None, g_star=None, gs_star=None, suppression="default",
                  chunk=4096, shell_chunk=None, want_pow_v=True, keep_bubbles=False, plan=None) -> SweepResult:
    """Spectrum(Bubble(model, v_wall, alpha_n), y, r_star=...) for every point of a sweep: the model-independent shell
    solve, the thermodynamic scalars, the SSM chain and Omega_gw,0 (pttools/analysis/parallel.py:157-187 ->
    bubble.py:232-352 -> ssmtools/spectrum.py:96-115 -> omgw0/omgw0_ssm.py:123-131), one model for all points."""
    from . import bubble as bub
    from . import omgw0 as om
    torch = engine._torch()
    to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a, dtype=float)
    vw = to_np(v_wall).astype(float).reshape(-1)
    an = to_np(alpha_n).astype(float).reshape(-1)
    n = vw.size
    # Points that share v_wall share the lifetime-convolution weights (K6'): process the sweep sorted by v_wall and
    # undo the permutation on the way out.
    perm = np.argsort(vw, kind="stable")
    sorted_already = bool(np.all(perm == np.arange(n)))
    if not sorted_already:
        vw, an = vw[perm], an[perm]
    y = engine.Y_DEFAULT if y is None else np.asarray(y, dtype=float)
    cs = float(np.sqrt(model.csb2))                                   # spectrum.py:99 (cs2 is constant per phase)
    if plan is None:
        plan = engine.SsmPlan(y, cs=cs, nt=nt, n_z_lookup=n_z_lookup, z_st_thresh=z_st_thresh, nuc_type=nuc_type,
                              a=1.0, mode="ssm", only_lookup_pass=not want_pow_v)
    want_pow_v = want_pow_v and plan.has_y_pass
    scale = None
    if r_star is not None:
        # omgw0_ssm.py:62-131: model temperatures are not physical for both in-scope models -> the overrides apply
        g_eff = om.G_STAR_DEFAULT if g_star is None else gs_star
        gs_eff = g_eff if gs_star is None else gs_star
        fgw0 = om.F_gw0(g_star=g_eff, gs_star=gs_eff)
        if suppression == "default":
            sup = om.DEFAULT_SUPPRESSION.points(vw, an)
        elif suppression is None:
            sup = np.ones_like(vw)
        else:
            sup = suppression.points(vw, an)
        scale = engine.dev_f64(r_star * fgw0 * sup)
    dev = "cuda"
    pow_gw = torch.empty((n, y.size), dtype=torch.float64, device=dev)
    pow_v = torch.empty((n, y.size), dtype=torch.float64, device=dev) if want_pow_v else None
    omgw0 = torch.empty((n, y.size), dtype=torch.float64, device=dev) if scale is not None else None
    status = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    scal: tp.Dict[str, tp.Any] = {}
    if shell_chunk is None:
        shell_chunk = max(chunk, min(n, int(24e9 // (3 * 8 * (2 * n_xi + 4)))))
    bubbles = []
    for s0 in range(0, n, shell_chunk):
        e0 = min(n, s0 + shell_chunk)
        # Points whose root search fails are retried the way the reference does (varied starts, backup scan); the
        # retries run on a side stream under the spectrum stages of the other points and are folded in below.
        bb = bub.sweep_bubbles(model, vw[s0:e0], an[s0:e0], n_xi=n_xi, retry="async")
        status[s0:e0], st[s0:e0] = bb.status, bb.sol_type
        for k, v in bb.scal.items():
            scal.setdefault(k, torch.empty(n, dtype=torch.float64, device=dev))[s0:e0] = v

        def lifetime_factor(b):
            # source lifetime factor, spectrum.py:123-136
            nu_ = b.scal["nu_gdh2024"]
            f = 1 / (1 + 2 * nu_)
            if r_star is not None and not np.isinf(lifetime_multiplier):
                f = f * (1 - (1 + lifetime_multiplier * 2 * r_star / torch.sqrt(b.scal["kinetic_energy_fraction"]))
                         ** (-1 - 2 * nu_))
            return f

        slf = lifetime_factor(bb)
        valid = bb.shells.npts.cpu().numpy() >= 3        # has a profile (points flagged solver_failed do, like the reference)

        def run_chunk(s):
            e = min(e0, s + chunk)
            sub = bb.shells.slice(s - s0, e - s0)
            res = engine.spectra_from_shells(plan, sub, bb.pp[s - s0:e - s0].contiguous(),
                                             slf=slf[s - s0:e - s0].contiguous(),
                                             scale=None if scale is None else scale[s:e].contiguous(),
                                             v_wall_host=vw[s:e], valid_host=valid[s - s0:e - s0])
            pow_gw[s:e] = res.pow_gw
            if want_pow_v:
                pow_v[s:e] = res.pow_v
            if omgw0 is not None:
                omgw0[s:e] = res.omgw0

        starts = list(range(s0, e0, chunk))
        n_free = 1
        if bb.retry_job is not None and len(starts) > 1:
            # Chunks without retried points go first (at least one chunk does): the retries finish under them and are
            # written into the batch before the chunks that hold such points start, so only the rescued points of chunks
            # that have already run need a pass of their own - normally none.
            where = bb.retry_job.pos + s0
            cnt = [np.count_nonzero((where >= s) & (where < s + chunk)) for s in starts]
            order = np.argsort(cnt, kind="stable")
            starts = [starts[i] for i in order]
            n_free = max(1, int(np.count_nonzero(np.asarray(cnt) == 0)))
        for s in starts[:n_free]:
            run_chunk(s)
        _, pos = bub.finish_retry(bb)                    # rescued profiles, scalars, status now sit in bb
        if pos.size:
            at = torch.as_tensor(pos + s0, device=dev, dtype=torch.int64)
            here = torch.as_tensor(pos, device=dev, dtype=torch.int64)
            status[at] = bb.status[here]
            for k, v in bb.scal.items():
                scal[k][at] = v[here]
            slf = lifetime_factor(bb)
            valid[pos] = True
        for s in starts[n_free:]:
            run_chunk(s)
        done_early = np.zeros(pos.shape, dtype=bool)
        for s in starts[:n_free]:
            done_early |= (pos + s0 >= s) & (pos + s0 < s + chunk)
        late = pos[done_early]
        if late.size:
            here = torch.as_tensor(late, device=dev, dtype=torch.int64)
            at = here + s0
            res = engine.spectra_from_shells(plan, bb.shells.take(here), bb.pp[here].contiguous(),
                                             slf=slf[here].contiguous(),
                                             scale=None if scale is None else scale[at].contiguous(),
                                             v_wall_host=vw[late + s0], valid_host=np.ones(late.size, dtype=bool))
            pow_gw[at] = res.pow_gw
            if want_pow_v:
                pow_v[at] = res.pow_v
            if omgw0 is not None:
                omgw0[at] = res.omgw0
        if keep_bubbles:
            bubbles.append(bb)
    if not sorted_already:
        if keep_bubbles:
            raise NotImplementedError("keep_bubbles needs points ordered by v_wall (grid sweeps are)")
        inv = torch.as_tensor(np.argsort(perm), device=dev)
        unperm = lambda t: None if t is None else t[inv]
        pow_gw, pow_v, omgw0, status, st = (unperm(t) for t in (pow_gw, pow_v, omgw0, status, st))
        scal = {k: v[inv] for k, v in scal.items()}
        vw, an = vw[np.argsort(perm)], an[np.argsort(perm)]
    out = SweepResult(engine.dev_f64(vw), engine.dev_f64(an), st, status, None, y, pow_gw=pow_gw, pow_v=pow_v,
                      omgw0=omgw0, scalars=scal)
    if keep_bubbles:
        out.bubbles = bubbles
    return out


def sweep_models(models, v_walls, alpha_ns, y=None, rank: int = 0, world: int = 1, **kw) -> tp.List[tp.Dict[str, tp.Any]]:
    """Spectra for several models on one (v_wall, alpha_n) grid - the multi-model sweeps of the reference's
    examples/const_cs/const_cs_gw.py:153-187 (create_spectra once per model).  Per model the grid is restricted to
    alpha_n >= model.alpha_n_min (Bubble raises below it, model.py:813-823); points are ordered v_wall-major so that
    the grouped spec_den_v path applies; plans are shared between models with the same broken-phase sound speed.
    With world > 1 the MODELS are dealt over the ranks (each model's sweep stays on one GPU).
    Returns one dict per model of this rank: {"model", "index", "mask" [n_alpha, n_vw], "result" SweepResult with rows
    in v_wall-major order of the masked grid, "i_alpha", "i_vw" row -> grid indices}."""
    v_walls = np.asarray(v_walls, dtype=float)
    alpha_ns = np.asarray(alpha_ns, dtype=float)
    y = engine.Y_DEFAULT if y is None else np.asarray(y, dtype=float)
    plans: tp.Dict[tp.Tuple, engine.SsmPlan] = {}
    out = []
    for im, model in enumerate(models):
        if im % world != rank:
            continue
        ok_alpha = alpha_ns >= model.alpha_n_min
        i_vw, i_al = np.meshgrid(np.arange(v_walls.size), np.nonzero(ok_alpha)[0], indexing="ij")     # v_wall-major
        i_vw, i_al = i_vw.ravel(), i_al.ravel()
        key = (round(float(model.csb2), 15), kw.get("nt", engine.NT_DEFAULT), kw.get("n_z_lookup", engine.NZ_LOOKUP_DEFAULT),
               kw.get("z_st_thresh", engine.Z_ST_THRESH), kw.get("nuc_type", "exponential"))
        if key not in plans:
            plans[key] = engine.SsmPlan(y, cs=float(np.sqrt(model.csb2)), nt=key[1], n_z_lookup=key[2], z_st_thresh=key[3],
                                        nuc_type=key[4], a=1.0, mode="ssm",
                                        only_lookup_pass=not kw.get("want_pow_v", True))
        res = sweep_spectra(model, v_walls[i_vw], alpha_ns[i_al], y=y, plan=plans[key], **kw) if i_vw.size else None
        mask = np.zeros((alpha_ns.size, v_walls.size), dtype=bool)
        mask[i_al, i_vw] = True
        out.append({"model": model, "index": im, "mask": mask, "result": res, "i_alpha": i_al, "i_vw": i_vw})
    return out


# ---------------------------------------------------------------------------------------------------------------
# Multi-GPU: static sharding, no collective in the solve, one gather of the result table at the end
# ---------------------------------------------------------------------------------------------------------------

def distributed_table(compute: tp.Callable[[np.ndarray], "np.ndarray | tp.Any"], n_points: int, n_cols: int,
                      sol_type: tp.Optional[np.ndarray] = None, device: str = "cuda", dst: int = 0):
    """Runs `compute(indices) -> [len(indices), n_cols] float64 tensor` on this rank's shard of range(n_points) and
    gathers the full [n_points, n_cols] table on rank `dst` (None elsewhere).

    Every (v_wall, alpha_n, model) point is independent (the reference maps one future per point,
    speedup/parallel.py:134-137), so the ranks never talk during the solve; the only exchange step is the final
    gather (NCCL over NVLink on a GPU box, gloo in the CPU tests).  Works without torch.distributed initialised
    (single process)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_indices(n_points, rank, world, sol_type)
    block = compute(mine)
    block = torch.as_tensor(block, dtype=torch.float64, device=device).reshape(len(mine), n_cols)
    if world == 1:
        out = torch.empty((n_points, n_cols), dtype=torch.float64, device=device)
        out[torch.as_tensor(mine, device=device)] = block
        return out
    # shards differ by at most one row: pad to a common size so that a plain gather can be used
    rows = (n_points + world - 1) // world
    padded = torch.zeros((rows, n_cols), dtype=torch.float64, device=device)
    padded[: len(mine)] = block
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n_points, n_cols), dtype=torch.float64, device=device)
    for r in range(world):
        idx = shard_indices(n_points, r, world, sol_type)
        out[torch.as_tensor(idx, device=device)] = bufs[r][: len(idx)]
    return out


def sweep_omgw0_distributed(model, v_wall, alpha_n, y=None, r_star=0.1, **kw):
    """10^6-point style sweep sharded over the ranks of torch.distributed: Omega_gw,0(f) table [n, n_y] on rank 0."""
    vw = np.asarray(v_wall, dtype=float).reshape(-1)
    an = np.asarray(alpha_n, dtype=float).reshape(-1)
    y = engine.Y_DEFAULT if y is None else np.asarray(y, dtype=float)
    codes = model.solution_type_codes(vw, an, np.asarray(model.wn(an), dtype=float))

    def compute(idx):
        return sweep_spectra(model, vw[idx], an[idx], y=y, r_star=r_star, want_pow_v=False, **kw).omgw0

    return distributed_table(compute, vw.size, y.size, sol_type=codes)
