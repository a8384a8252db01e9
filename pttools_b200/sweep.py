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

    invalid: tp.Any = None
    stats: tp.Any = None
    snr: tp.Any = None        # [n, 2]: signal_to_noise_ratio, signal_to_noise_ratio_instrument (sweep_spectra(snr=True))

    def take(self, rows):
        """The given rows (host int array) as a SweepResult of their own."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        on_device = {}

        def tk(x):
            if x is None:
                return None
            if hasattr(x, "cpu"):
                import torch
                if x.device not in on_device:
                    # ONE upload of the row list, from pinned memory and without waiting: a pageable copy would hold the
                    # host until the stream gets there, i.e. until the sweep that is filling these tables has finished
                    idx = torch.from_numpy(rows)
                    on_device[x.device] = idx.pin_memory().to(x.device, non_blocking=True) if x.is_cuda else idx
                return x[on_device[x.device]]
            return np.asarray(x)[rows]
        r = SweepResult(tk(self.v_wall), tk(self.alpha_n), tk(self.sol_type), tk(self.status), tk(self.alpha_plus), self.y,
                        pow_gw=tk(self.pow_gw), pow_v=tk(self.pow_v), omgw0=tk(self.omgw0),
                        scalars={k: tk(v) for k, v in self.scalars.items()})
        r.invalid = None if self.invalid is None else np.asarray(self.invalid)[rows]
        return r

    def numpy(self):
        def cv(x):
            return x.cpu().numpy() if hasattr(x, "cpu") else x
        return SweepResult(**{f.name: (cv(getattr(self, f.name)) if f.name not in ("scalars", "stats", "y")
                                        else ({k: cv(v) for k, v in self.scalars.items()} if f.name == "scalars"
                                              else getattr(self, f.name)))
                              for f in dataclasses.fields(self)})


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


def shell_chunk_points(n_xi: int, budget_bytes: tp.Optional[float] = None) -> int:
    """How many bag shells to solve per launch: bounded by the profile rows (3 arrays of row_capacity doubles each)
    against a share of the free device memory."""
    from . import _lib
    torch = engine._torch()
    cap = _lib.load().ptt_profile_row_capacity(n_xi)
    if budget_bytes is None:
        budget_bytes = 0.3 * torch.cuda.mem_get_info()[0]
    return int(max(1024, min(65536, budget_bytes // (3 * 8 * cap))))


def omgw0_scale(v_wall, alpha_n, r_star, g_star=None, gs_star=None, suppression="default", device=False):
    """Per-point factor r_star * F_gw0 * suppression(v_wall, alpha_n) of Spectrum.omgw0 (omgw0_ssm.py:62-131).  Both
    in-scope models have non-physical temperatures, so the overrides of the reference apply (g_star_override takes
    gs_star when g_star is given: omgw0_ssm.py:74, sic)."""
    from . import omgw0 as om
    g_eff = om.G_STAR_DEFAULT if g_star is None else gs_star
    gs_eff = g_eff if gs_star is None else gs_star
    fgw0 = om.F_gw0(g_star=g_eff, gs_star=gs_eff)
    table = om.DEFAULT_SUPPRESSION if isinstance(suppression, str) and suppression == "default" else suppression
    if table is None:
        sup = engine.dev_f64(np.ones_like(v_wall)) if device else np.ones_like(v_wall)
    elif device:
        sup = table.points_device(v_wall, alpha_n)           # ptt_suppression_batch: interpolated on the device
    else:
        sup = table.points(v_wall, alpha_n)
    return r_star * fgw0 * sup


def sweep_spectra(model, v_wall, alpha_n, y=None, r_star=None, lifetime_multiplier=1.0, nuc_type="exponential",
                  nt=engine.NT_DEFAULT, n_z_lookup=engine.NZ_LOOKUP_DEFAULT, z_st_thresh=engine.Z_ST_THRESH,
                  n_xi=engine.N_XI_DEFAULT, Tn=None, g_star=None, gs_star=None, suppression="default",
                  chunk=8000, shell_chunk=None, want_pow_v=True, want_pow_gw=True, keep_bubbles=False, plan=None,
                  mem="device",
                  allow_invalid=False, on_chunk=None, timing=False, out=None, sol_type=None, snr=False,
                  table_rows=None, out_rows=None) -> SweepResult:
    """Spectrum(Bubble(model, v_wall, alpha_n), y, r_star=...) for every point of a sweep: the model-independent shell
    solve, the thermodynamic scalars, the SSM chain and Omega_gw,0 (pttools/analysis/parallel.py:157-187 ->
    bubble.py:232-352 -> ssmtools/spectrum.py:96-115 -> omgw0/omgw0_ssm.py:123-131) in ONE call of the C ABI
    (ptt_sweep_spectra).  `model`: one Model, or a sequence with one Model per point that share csb2 (the sound speed of
    the GW integral, spectrum.py:99).  mem="device": result tables are torch tensors; "host": pinned numpy arrays,
    streamed out per chunk while the sweep runs.  Points the reference's Bubble constructor rejects come back with
    status PTT_ST_INVALID_INPUT and NaN rows (`invalid` marks them).
    snr=True (needs r_star): additionally the LISA signal-to-noise ratios of every point, Spectrum.signal_to_noise_ratio()
    and .signal_to_noise_ratio_instrument() (omgw0_ssm.py:143-147), as `snr[n, 2]`; the frequencies are
    f = y / r_star * f_star0(Tn, g_star) (omgw0_ssm.py:162-188)."""
    from . import bubble as bub
    from .models import Model
    torch = engine._torch()
    to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a, dtype=float)
    vw = to_np(v_wall).astype(float).reshape(-1)
    an = to_np(alpha_n).astype(float).reshape(-1)
    y = engine.Y_DEFAULT if y is None else np.asarray(y, dtype=float)
    if isinstance(model, Model):
        csb2 = float(model.csb2)
    else:
        csb2s = {float(m.csb2) for m in model}
        if len(csb2s) != 1:
            raise ValueError("the models of one sweep must share csb2 (it fixes the grids of the GW integral)")
        csb2 = csb2s.pop()
    cs = float(np.sqrt(csb2))                                         # spectrum.py:99 (cs2 is constant per phase)
    if plan is None:
        plan = engine.SsmPlan(y, cs=cs, nt=nt, n_z_lookup=n_z_lookup, z_st_thresh=z_st_thresh, nuc_type=nuc_type,
                              a=1.0, mode="ssm", only_lookup_pass=not want_pow_v)
    want_pow_v = want_pow_v and plan.has_y_pass
    if mem == "device":
        # the preparation runs on its own stream: its host reads must not wait for a previous sweep still in flight
        main, prep = torch.cuda.current_stream(), engine.prep_stream()
        with torch.cuda.stream(prep):
            pt = bub.point_table(model, vw, an, sol_type=sol_type, allow_invalid=allow_invalid, device=True)
            scale = None if r_star is None else omgw0_scale(vw, an, r_star, g_star, gs_star, suppression, device=True)
            vw_out, an_out = pt.pg[:, 0].contiguous(), pt.pg[:, 1].contiguous()
            codes = pt.pg[:, 2].to(torch.int32)
        main.wait_stream(prep)
        for t in (pt.pg, pt.pm, scale, vw_out, an_out, codes):
            if t is not None:
                t.record_stream(main)
    else:
        pt = bub.point_table(model, vw, an, sol_type=sol_type, allow_invalid=allow_invalid, device=False)
        scale = None if r_star is None else omgw0_scale(vw, an, r_star, g_star, gs_star, suppression, device=False)
        vw_out, an_out, codes = vw, an, pt.codes
    snr_w = None
    if snr:
        if r_star is None:
            raise ValueError("snr needs r_star (the frequencies today depend on it)")
        from . import noise as nz
        from . import omgw0 as om
        g_eff = om.G_STAR_DEFAULT if g_star is None else gs_star                 # sic: omgw0_ssm.py:74
        f = om.f(y, r_star, om.f_star0(om.T_default if Tn is None else Tn, g_eff))
        snr_w = nz.snr_weights(f, (nz.omega_noise(f), nz.omega_ins(f)))
    want = (("pow_v",) if want_pow_v else ()) + (("pow_gw",) if want_pow_gw or scale is None else ()) + \
        (("omgw0",) if scale is not None else ())
    pg_in, pm_in, sc_in = pt.pg, pt.pm, scale
    res = engine.sweep_c(plan, pg_in, pm_in, sc_in, n_xi=n_xi, r_star=r_star, lifetime_multiplier=lifetime_multiplier,
                         chunk=chunk, shell_chunk=shell_chunk or 0, want=want, rows=keep_bubbles, mem=mem,
                         on_chunk=on_chunk, timing=timing, out=out, snr_weights=snr_w, table_rows=table_rows,
                         out_rows=out_rows, host_cols=(vw, pt.codes) if mem == "device" else None)
    from . import _lib
    scal = {name: res.scalars[:, k] for k, name in enumerate(_lib.SWEEP_SCALARS)}
    out_ = SweepResult(vw_out, an_out, codes, res.status, None, y, pow_gw=res.pow_gw, pow_v=res.pow_v,
                       omgw0=res.omgw0, scalars=scal)
    out_.invalid = pt.invalid
    out_.stats = res.stats
    out_.snr = res.snr
    out_._keep = res
    if keep_bubbles:
        pg = engine.dev_f64(pt.pg)
        sol = torch.as_tensor(pt.codes, dtype=torch.int32, device="cuda")
        shells = engine.ShellBatch(pg[:, 0].contiguous(), pg[:, 1].contiguous(), None, sol, res.status, res.rows[0],
                                   res.rows[1], res.rows[2], res.off, res.npts, res.scalars, n_xi)
        out_.bubbles = bub.BubbleBatch(model, shells, shells.v_wall, shells.alpha_n, engine.dev_f64(pt.wn), sol,
                                       res.status, scal, None, n_xi, invalid=pt.invalid)
    return out_


def sweep_models(models, v_walls, alpha_ns, y=None, rank: int = 0, world: int = 1, **kw) -> tp.List[tp.Dict[str, tp.Any]]:
    """Spectra for several models on one (v_wall, alpha_n) grid - the multi-model sweeps of the reference's
    examples/const_cs/const_cs_gw.py:153-187 (create_spectra once per model).  Per model the grid is restricted to
    alpha_n >= model.alpha_n_min (Bubble raises below it, model.py:813-823).  Models that share the broken-phase sound
    speed share the grids of the spectrum chain AND the lifetime-convolution weights of K6' (they depend on v_wall and
    csb2 only), so each such family runs as ONE sweep ordered by v_wall: a group of K6' then holds the profiles of all
    its models.  With world > 1 the v_wall rows of every family are dealt over the ranks.
    Returns one dict per model: {"model", "index", "mask" [n_alpha, n_vw], "result" SweepResult with rows in v_wall-major
    order of the masked grid (None if the model has no point on this rank), "i_alpha", "i_vw" row -> grid indices}."""
    v_walls = np.asarray(v_walls, dtype=float)
    alpha_ns = np.asarray(alpha_ns, dtype=float)
    y = engine.Y_DEFAULT if y is None else np.asarray(y, dtype=float)
    my_vw = np.arange(v_walls.size)[rank::world]
    families: tp.Dict[float, tp.List[int]] = {}
    for im, m in enumerate(models):
        families.setdefault(round(float(m.csb2), 15), []).append(im)
    out = [None] * len(models)
    for _, members in families.items():
        blocks = []
        for im in members:
            ok_alpha = np.nonzero(alpha_ns >= models[im].alpha_n_min)[0]
            i_vw, i_al = np.meshgrid(my_vw, ok_alpha, indexing="ij")                  # v_wall-major
            blocks.append((im, i_vw.ravel(), i_al.ravel()))
        im_all = np.concatenate([np.full(b[1].size, b[0]) for b in blocks])
        ivw_all = np.concatenate([b[1] for b in blocks])
        ial_all = np.concatenate([b[2] for b in blocks])
        if im_all.size == 0:
            for im, i_vw, i_al in blocks:
                out[im] = {"model": models[im], "index": im, "mask": np.zeros((alpha_ns.size, v_walls.size), bool),
                           "result": None, "i_alpha": i_al, "i_vw": i_vw}
            continue
        order = np.argsort(v_walls[ivw_all], kind="stable")       # all models' profiles of one wall speed together
        per_point = [models[i] for i in im_all[order]]
        res = sweep_spectra(per_point, v_walls[ivw_all[order]], alpha_ns[ial_all[order]], y=y, **kw)
        inv = np.empty_like(order)
        inv[order] = np.arange(order.size)
        start = 0
        for im, i_vw, i_al in blocks:
            rows = inv[start:start + i_vw.size]
            start += i_vw.size
            mask = np.zeros((alpha_ns.size, v_walls.size), dtype=bool)
            mask[i_al, i_vw] = True
            out[im] = {"model": models[im], "index": im, "mask": mask, "result": res.take(rows) if rows.size else None,
                       "i_alpha": i_al, "i_vw": i_vw}
    return out


# ---------------------------------------------------------------------------------------------------------------
# Multi-GPU: whole v_wall rows per rank, no collective in the solve, result blocks sent to the root as they finish
# ---------------------------------------------------------------------------------------------------------------

def shard_rows(v_wall: np.ndarray, cost: tp.Optional[np.ndarray], world: int,
               only: tp.Optional[int] = None) -> tp.List[np.ndarray]:
    """Static partition of the points of a sweep over `world` ranks by WHOLE v_wall ROWS: all points that share a wall
    speed stay on one rank, because they share the lifetime-convolution weights of K6' (one banded matrix product per
    wall speed).  cost = None: the rows, ordered by v_wall, are dealt cyclically -- neighbouring wall speeds have nearly
    the same mix of solution types, so the shards are balanced without looking at the points; with cost[i] (estimated
    cost of point i) rows are dealt greedily, most expensive first, to the least loaded rank.  Returns per rank the
    point indices, ordered by (v_wall, original index).  only=k: only the shard of rank k is built (the others are None)."""
    v_wall = np.asarray(v_wall, dtype=float).reshape(-1)
    n = v_wall.size
    ranks = range(world) if only is None else (only,)
    out: tp.List[tp.Optional[np.ndarray]] = [None] * world

    def deal(n_rows, row_cost):
        if row_cost is None:
            return np.arange(n_rows) % world
        owner = np.empty(n_rows, dtype=np.int64)
        load = np.zeros(world)
        for r in np.argsort(-row_cost, kind="stable"):
            k = int(np.argmin(load))
            owner[r] = k
            load[k] += row_cost[r]
        return owner

    if n:
        # grid sweeps arrive as runs of equal wall speed: if every wall speed forms ONE run (in whatever order the runs
        # come), the rows are the runs, only the run values need sorting and a shard is a concatenation of index ranges
        starts = np.concatenate(([0], np.nonzero(v_wall[1:] != v_wall[:-1])[0] + 1))
        by_value = np.argsort(v_wall[starts], kind="stable")
        sorted_vals = v_wall[starts][by_value]
        if starts.size * 8 <= n and np.all(sorted_vals[1:] > sorted_vals[:-1]):
            lengths = np.diff(np.concatenate((starts, [n])))
            r_start, r_len = starts[by_value], lengths[by_value]          # rows in v_wall order
            row_cost = None if cost is None else \
                np.add.reduceat(np.asarray(cost, dtype=float).reshape(-1), starts)[by_value]
            owner = deal(starts.size, row_cost)
            for k in ranks:
                rs, rl = r_start[owner == k], r_len[owner == k]
                offs = np.concatenate(([0], np.cumsum(rl)[:-1]))
                out[k] = np.repeat(rs - offs, rl) + np.arange(int(rl.sum()))
            return out
    uniq, inv = np.unique(v_wall, return_inverse=True)
    order = np.lexsort((np.arange(n), v_wall))
    row_cost = None if cost is None else \
        np.bincount(inv, weights=np.asarray(cost, dtype=float).reshape(-1), minlength=uniq.size)
    owner = deal(uniq.size, row_cost)
    # split `order` by owner, keeping the order inside every shard: one stable radix sort on a 16-bit key instead of a
    # mask per rank
    key = owner[inv][order].astype(np.uint16 if world <= 65535 else np.int64)
    by_owner = order[np.argsort(key, kind="stable")]
    ends = np.cumsum(np.bincount(key, minlength=world))
    for k in ranks:
        out[k] = by_owner[(ends[k - 1] if k else 0):ends[k]]
    return out


def point_cost(codes: np.ndarray) -> np.ndarray:
    """Relative cost of a point from its solution type code: points without a solution (-1) skip the spectrum chain."""
    return np.where(np.asarray(codes) < 0, 0.05, 1.0)


def distributed_table(local_compute: tp.Callable, shards: tp.List[np.ndarray], n_cols: int, block: int, device="cuda",
                      dst: int = 0, dtype=None):
    """Runs `local_compute(idx, out, on_chunk)` on this rank's shard and assembles the full [n_points, n_cols] table on
    rank `dst` (None elsewhere).  `local_compute` fills out[r0:r1] for row ranges of its shard in any order and calls
    on_chunk(r0, r1, cuda_stream_ptr_or_None) when a range has been queued; ranges never straddle a multiple of
    `block`.  Completed blocks are sent to `dst` while the rest of the shard is still being computed (point-to-point,
    NCCL over NVLink on a GPU box, gloo in the CPU tests): the exchange hides under the compute except for the last
    block.  Every (v_wall, alpha_n, model) point is independent (the reference maps one future per point,
    speedup/parallel.py:134-137), so there is no other communication."""
    import torch
    import torch.distributed as dist
    dtype = dtype or torch.float64
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shards[rank]
    n_points = int(sum(len(s) for s in shards))
    local = torch.empty((len(mine), n_cols), dtype=dtype, device=device)
    blocks = lambda m: [(a, min(a + block, m)) for a in range(0, m, block)]
    works = []
    recv = {}
    use_cuda = torch.device(device).type == "cuda"
    if world > 1 and rank == dst:
        # post every receive up front, block by block round-robin over the senders
        recv = {r: torch.empty((len(shards[r]), n_cols), dtype=dtype, device=device) for r in range(world) if r != dst}
        todo = {r: blocks(len(shards[r])) for r in recv}
        k = 0
        while any(k < len(b) for b in todo.values()):
            for r, b in todo.items():
                if k < len(b):
                    works.append(dist.irecv(recv[r][b[k][0]:b[k][1]], src=r))
            k += 1
    my_blocks = blocks(len(mine))
    done = np.zeros(len(my_blocks), dtype=np.int64)      # rows delivered per block
    next_block = [0]

    def on_chunk(r0, r1, stream):
        if world == 1 or rank == dst:
            return
        done[r0 // block] += r1 - r0
        while next_block[0] < len(my_blocks):
            a, b = my_blocks[next_block[0]]
            if done[next_block[0]] < b - a:
                break
            if use_cuda and stream:
                with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                    works.append(dist.isend(local[a:b], dst=dst))
            else:
                works.append(dist.isend(local[a:b], dst=dst))
            next_block[0] += 1

    local_compute(mine, local, on_chunk)
    if world > 1 and rank != dst:
        # whatever the callbacks have not covered (a compute that never called back) goes now
        while next_block[0] < len(my_blocks):
            a, b = my_blocks[next_block[0]]
            works.append(dist.isend(local[a:b], dst=dst))
            next_block[0] += 1
    for w in works:
        w.wait()
    if rank != dst:
        return None
    if world == 1 and np.array_equal(mine, np.arange(n_points)):
        return local                      # one rank, points already in processing order: the shard IS the table
    out = torch.empty((n_points, n_cols), dtype=dtype, device=device)
    out[torch.as_tensor(mine, device=device)] = local
    for r, buf in recv.items():
        out[torch.as_tensor(shards[r], device=device)] = buf
    return out


LAST_LOCAL: tp.Dict[str, tp.Any] = {}      # the SweepResult of this rank's shard of the last distributed sweep (statistics)


class SharedTable:
    """A [n_rows, n_cols] float64 table in the memory of ONE GPU (rank `dst`) that every rank of the box can write:
    allocated by the library on `dst` (ptt_ipc_alloc), mapped by the other processes through CUDA IPC (ptt_ipc_open; peer
    access over NVLink / NVSwitch).  Collective over torch.distributed: every rank constructs it at the same point.
    `tensor`: torch view of the table (on rank dst its own memory, elsewhere the peer mapping)."""

    def __init__(self, n_rows: int, n_cols: int, dst: int = 0):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        torch = engine._torch()
        lib = _lib.load()
        self.shape, self.dst = (int(n_rows), int(n_cols)), dst
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self._ptr = C.c_void_p()
        self._owner = self.rank == dst
        nbytes = 8 * self.shape[0] * self.shape[1]
        handle = (C.c_ubyte * 64)()
        if self._owner:
            _lib.check(lib.ptt_ipc_alloc(nbytes, C.byref(self._ptr), handle), "ptt_ipc_alloc")
        if self.world > 1:
            box = [bytes(handle) if self._owner else None]
            dist.broadcast_object_list(box, src=dst)
            if not self._owner:
                buf = (C.c_ubyte * 64).from_buffer_copy(box[0])
                _lib.check(lib.ptt_ipc_open(buf, C.byref(self._ptr)), "ptt_ipc_open")

        class _Iface:
            pass
        holder = _Iface()
        holder.__cuda_array_interface__ = {"shape": self.shape, "typestr": "<f8", "data": (int(self._ptr.value), False),
                                           "version": 3, "strides": None}
        self._holder = holder
        self.tensor = torch.as_tensor(holder, device="cuda")

    def close(self):
        from . import _lib
        ptr, self._ptr = getattr(self, "_ptr", None), None
        if ptr is None or not ptr.value:
            return
        self.tensor = None
        lib = _lib.load()
        (lib.ptt_ipc_free if self._owner else lib.ptt_ipc_close)(ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass


def wait_distributed():
    """Completes sweep_omgw0_distributed(..., wait=False) calls: this rank's rows are in the table on `dst` and so are
    everybody else's (stream synchronisation + one barrier)."""
    import torch.distributed as dist
    engine._torch().cuda.current_stream().synchronize()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def sweep_omgw0_distributed(model, v_wall, alpha_n, y=None, r_star=0.1, chunk=8000, dst=0, table="omgw0", mem="device",
                            shared: tp.Optional[SharedTable] = None, transport: str = "auto", wait: bool = True, **kw):
    """The 10^6-point sweep of BASELINE.json config 5, sharded over the ranks of torch.distributed (one process per
    GPU): Omega_gw,0(f) table [n, n_y] on rank `dst` with rows in the order of the input points (None on the other
    ranks).  Whole v_wall rows per rank (shard_rows); every rank's sweep writes its rows STRAIGHT INTO the table on
    `dst`, chunk by chunk as they are produced: the table is shared through CUDA IPC (SharedTable) and the rows travel as
    peer stores over NVLink / NVSwitch under the compute of the next chunk -- no collective, no SM of the receiving GPU
    involved; one barrier at the end.  wait=False (needs `shared`): returns as soon as this rank's work is queued, like a
    single-GPU sweep does; the table is complete after wait_distributed() -- for back-to-back sweeps into different
    tables, whose host preparation then overlaps the previous sweep's tail.  shared: a SharedTable([n, n_y]) to write into (reuse it across sweeps of one
    shape; without it a table is allocated and mapped per call).  transport="sendrecv": point-to-point messages of
    torch.distributed instead (distributed_table; what the CPU tests run over gloo).
    mem="host": no exchange; every rank returns (indices of its shard, its [n_shard, n_y] block in pinned host memory),
    streamed out of the device chunk by chunk -- the host-side form of a sharded table."""
    from .models import Model
    import torch.distributed as dist
    vw = np.asarray(v_wall, dtype=float).reshape(-1)
    an = np.asarray(alpha_n, dtype=float).reshape(-1)
    y = engine.Y_DEFAULT if y is None else np.asarray(y, dtype=float)
    single = isinstance(model, Model)
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    # only the sendrecv transport needs the other ranks' shards (rank `dst` posts the receives)
    shards = shard_rows(vw, None, world, only=None if transport == "sendrecv" else rank)
    torch = engine._torch()
    # shell chunks that are multiples of the spectrum chunk keep every delivered range inside one block of `chunk` rows
    from . import _lib
    sc = kw.pop("shell_chunk", None) or \
        int(_lib.load().ptt_sweep_shell_chunk(kw.get("n_xi", engine.N_XI_DEFAULT), int(torch.cuda.mem_get_info()[0])))
    sc = max(chunk, sc // chunk * chunk)
    idx = shards[rank]
    m = model if single else [model[i] for i in idx]
    common = dict(y=y, r_star=r_star, want_pow_v=False, want_pow_gw=table == "pow_gw", chunk=chunk, shell_chunk=sc, **kw)
    if mem == "host":
        res = sweep_spectra(m, vw[idx], an[idx], mem="host", **common)
        LAST_LOCAL["result"] = res
        return idx, getattr(res, table)
    if world == 1:
        res = sweep_spectra(model, vw, an, out=None if shared is None else {table: shared.tensor}, **common)
        LAST_LOCAL["result"] = res
        return getattr(res, table)
    if transport == "sendrecv":
        def local_compute(ids, out, on_chunk):
            if len(ids):
                LAST_LOCAL["result"] = sweep_spectra(m, vw[ids], an[ids], on_chunk=on_chunk, out={table: out}, **common)
        return distributed_table(local_compute, shards, y.size, chunk, dst=dst)
    own = shared is None
    if own:
        shared = SharedTable(vw.size, y.size, dst=dst)
    if shared.shape != (vw.size, y.size):
        raise ValueError(f"shared table has shape {shared.shape}, the sweep needs {(vw.size, y.size)}")
    if len(idx):
        LAST_LOCAL["result"] = sweep_spectra(m, vw[idx], an[idx], out={table: shared.tensor}, table_rows=idx, **common)
    if not wait:
        if own:
            raise ValueError("wait=False needs a SharedTable that outlives the call")
        return shared.tensor if rank == dst else None
    torch.cuda.current_stream().synchronize()      # this rank's rows are in the table on `dst`
    dist.barrier()
    result = shared.tensor if rank == dst else None
    if own and rank != dst:
        shared.close()
    if own and rank == dst:
        result._shared = shared                       # keeps the allocation alive as long as the tensor
    return result
