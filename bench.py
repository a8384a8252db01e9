#!/usr/bin/env python
"""Benchmark of the hot path: bubble + GW spectra per second over the (v_wall, alpha_n) grid of BASELINE.json config 2.

A "step" is one pass of the hot path (fluid shell solve -> A^2 -> spec_den_v -> spec_den_gw -> pow_gw) over one
batch of points of the 1000 x 1000 BagModel grid v_wall = linspace(0.05, 0.95, 1000) x alpha_n = linspace(0.005, 0.5,
1000) at the reference's default resolutions.  A batch is a set of complete v_wall rows of the grid (32 rows x 1000
alpha_n by default, rows spread over the v_wall range with stride 619 mod 1000), i.e. the order in which a grid sweep
is processed; every step sees deflagrations, hybrids, detonations and no-solution points and no step re-uses another
step's data.  With N GPUs every rank works on its own batch (weak scaling) and the
result block is gathered on rank 0 over NCCL, as a sharded sweep would do.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the reference's own CPU implementation of the
same workload on the host cores: the unmodified package staged under oracle/_ref by build() (cpu_baseline.kind =
"reference"), or the oracle port when that does not import on the box ("port", with the reason in `sample`).
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_N = 1000
STRIDE = 999983                      # prime, coprime with 10^6
Y = np.logspace(-1, 3, 1000)         # ssmtools/const.py:20
NPT = (2000, 10000, 10000)           # NPTDEFAULT (ssmtools/const.py:10-19): n_xi, nt, n_z_lookup
N_XI = 5000                          # bubble/const.py:12
R_STAR = 0.1
WORKLOADS = {
    # the reference's new API: what create_spectra(model, v_walls, alpha_ns, spectrum_kwargs={"r_star": 0.1}) computes
    "generic": ("config2/5: BagModel(alpha_n_min=0.005) grid v_wall=linspace(0.05,0.95,1000) x alpha_n=linspace(0.005,0.5,"
                "1000); per point Bubble(model, v_wall, alpha_n) [model-independent solver, n_xi=5000, t_end=50] + "
                "Spectrum(bubble, r_star=0.1) [y=logspace(-1,3,1000), nt=10^4, n_z_lookup=10^4, nxi=2000] -> pow_v, pow_gw, "
                "omgw0"),
    # config 5: the same per-point work, the whole grid as ONE sweep per step
    "config5": ("config5: the full grid v_wall=linspace(0.05,0.95,1000) x alpha_n=linspace(0.005,0.5,1000) "
                "(BagModel(alpha_n_min=0.005)) in ONE call of sweep.sweep_omgw0_distributed per step -> Omega_gw,0 table "
                "[10^6, 1000] on rank 0; per point Bubble [n_xi=5000] + Spectrum(r_star=0.1) at the reference's default "
                "sizes; a step = one time-to-solution"),
    # the reference's old bag API
    "bag": ("config2: BagModel grid v_wall=linspace(0.05,0.95,1000) x alpha_n=linspace(0.005,0.5,1000); per point "
            "sound_shell_bag(n_xi=2000) + power_gw_scaled_bag(z=logspace(-1,3,1000), npt=(2000,10000,10000))"),
}


E2E_WARMUP = 2                       # untimed end-to-end calls before the timed ones
ROW_STRIDE = 619                     # coprime with 1000: consecutive batches take v_wall rows spread over the range


def batch_points(step: int, rank: int, world: int, batch: int):
    """Points of one batch.  A batch of a multiple of 1000 points is a set of complete v_wall rows of the grid (all
    1000 alpha_n each), rows spread over [0.05, 0.95] -- the order in which a grid sweep is processed; other batch
    sizes take a pseudo-random scatter of the grid (stride 999983 mod 10^6)."""
    v_walls = np.linspace(0.05, 0.95, GRID_N)
    alpha_ns = np.linspace(0.005, 0.5, GRID_N)
    if batch % GRID_N == 0:
        rows_per = batch // GRID_N
        rows = ((np.arange(rows_per, dtype=np.int64) + (step * world + rank) * rows_per) * ROW_STRIDE) % GRID_N
        rows = np.sort(rows)
        return np.repeat(v_walls[rows], GRID_N), np.tile(alpha_ns, rows_per)
    j = (np.arange(batch, dtype=np.int64) + (step * world + rank) * batch) * STRIDE % (GRID_N * GRID_N)
    return v_walls[j % GRID_N].copy(), alpha_ns[j // GRID_N].copy()


# ---------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference algorithm) -- also the cpu_baseline leg of the GPU arm
# ---------------------------------------------------------------------------------------------------------------

_CPU = {}


def _cpu_setup(path):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import models, shell_generic as sg
    _CPU["path"] = path
    if path in ("generic", "shells"):
        _CPU["model"] = models.bag_model(a_s=1.0 / (1 - 3 * 0.005 * 0.999), a_b=1.0, V_s=1.0)
        _CPU["ref"] = sg.FluidReference(dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_fluid_reference.npz"))))


def _cpu_point(args):
    vw, an = args
    from oracle import shell_bag as sb, shell_generic as sg, ssm
    if not _CPU.get("skip_bag_type") and sb.identify_solution_type(vw, an) == sb.ERROR:
        return 0
    if _CPU["path"] == "shells":
        try:
            sol = sg.sound_shell_generic(_CPU["model"], vw, an, _CPU["ref"])
            sg.bubble_scalars(_CPU["model"], sol, vw)
        except (ValueError, RuntimeError, IndexError):
            return 0
        return 1
    if _CPU["path"] == "bag":
        ssm.power_gw_scaled_bag(Y, (vw, an), NPT)
        return 1
    m = _CPU["model"]
    try:
        sol = sg.sound_shell_generic(m, vw, an, _CPU["ref"])
        sc = sg.bubble_scalars(m, sol, vw)
        sp = ssm.ssm_spectrum(m, sol, sc, vw, y=Y, r_star=R_STAR)
        ssm.omgw0_scale(R_STAR, 1.0) * sp["pow_gw"]
    except (ValueError, RuntimeError, IndexError):
        return 0
    return 1


def _ref_setup(path):
    """The REAL reference (oracle/_ref: the unmodified pttools package staged by oracle/build_ref.py, numba-jitted) when it
    imports here; returns None on success, else the reason (the caller then falls back to the oracle port and says so).
    The jitted helpers are compiled in the parent, before the fork, by solving one point of each solution type."""
    if "ref_reason" in _CPU:
        return _CPU["ref_reason"]
    from oracle import build_ref
    dst = build_ref.build()
    reason = None
    if dst is None:
        reason = "oracle/_ref not staged (build() did not see /root/reference)"
    else:
        os.environ.setdefault("NUMBA_THREADING_LAYER", "workqueue")
        os.environ.setdefault("NUMBA_NUM_THREADS", "1")      # one process per core, as the reference's own run_parallel
        added = [dst, os.path.join(ROOT, "tools", "refenv", "stubs")]
        sys.path[:0] = added
        try:
            import logging
            import warnings
            warnings.filterwarnings("ignore")
            logging.getLogger("pttools").setLevel(logging.CRITICAL)
            from pttools.models import BagModel
            from pttools.bubble import Bubble
            from pttools.omgw0 import Spectrum
            logging.getLogger("pttools").setLevel(logging.CRITICAL)
            _CPU["refmod"] = {"Bubble": Bubble, "Spectrum": Spectrum, "model": BagModel(alpha_n_min=0.005)}
            _CPU["path"] = path
            t0 = time.perf_counter()
            for pt in ((0.5, 0.1), (0.62, 0.1), (0.9, 0.05)):       # deflagration, hybrid, detonation
                _ref_point(pt)
            _CPU["ref_jit_s"] = time.perf_counter() - t0
        except Exception as e:      # noqa: BLE001 - numba / scipy missing or incompatible on this box
            for a in added:
                if a in sys.path:
                    sys.path.remove(a)
            reason = f"reference not importable here: {type(e).__name__}: {e}"
    _CPU["ref_reason"] = reason
    return reason


def _ref_point(args):
    """One grid point through the reference's own public API: Bubble(model, v_wall, alpha_n) [+ Spectrum(...).omgw0()]."""
    vw, an = args
    R = _CPU["refmod"]
    try:
        b = R["Bubble"](R["model"], float(vw), float(an))
    except (ValueError, RuntimeError):       # the constructor rejects the point (no solution type, alpha_n range)
        return 0
    if not b.solved:
        return 0
    if _CPU["path"] == "shells":
        return 1
    R["Spectrum"](b, r_star=R_STAR).omgw0()
    return 1


class CpuPool:
    """Fork pool over the host cores, created once and reused across steps.  kind = "reference" (the real reference from
    oracle/_ref) when it imports on this box and the path is one it covers, else "port" (the oracle restatement)."""

    def __init__(self, cores, path, prefer_reference=True):
        import multiprocessing as mp
        self.kind, self.note = "port", ""
        if prefer_reference and path in ("generic", "shells"):
            reason = _ref_setup(path)
            if reason is None:
                self.kind = "reference"
                self.note = f"; the unmodified reference from oracle/_ref (numba {_numba_version()}, jit warm-up " \
                            f"{_CPU['ref_jit_s']:.0f} s before the fork, not timed)"
            else:
                self.note = f"; oracle port, because: {reason}"
        if self.kind == "port":
            _cpu_setup(path)      # before the fork: workers inherit the model and the starting-guess table
        _CPU["path"] = path
        self.fn = _ref_point if self.kind == "reference" else _cpu_point
        self.pool = mp.get_context("fork").Pool(cores)

    def run(self, points):
        t0 = time.perf_counter()
        self.pool.map(self.fn, points, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def _numba_version():
    try:
        import numba
        return numba.__version__
    except ImportError:
        return "?"


def cpu_run(points, cores, path, prefer_reference=False):
    pool = CpuPool(cores, path, prefer_reference)
    try:
        return pool.run(points)
    finally:
        pool.close()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    per_step = 2 * max(cores, 8)
    times = []
    pool = CpuPool(cores, args.path, prefer_reference=not args.cpu_port)
    for s in range(args.warmup + args.steps):
        vw, an = batch_points(s, 0, 1, args.batch)
        pick = np.linspace(0, vw.size - 1, per_step).astype(int)      # spread over the rows and strengths of the batch
        vw, an = vw[pick], an[pick]
        dt = pool.run(list(zip(vw, an)))
        if s >= args.warmup:
            times.append(dt)
    pool.close()
    total = float(np.sum(times))
    value = per_step * args.steps / total
    sample = (f"{per_step} grid points per step spread evenly over the GPU arm's batch of the same step, "
              f"{args.steps} steps, one fork pool of {cores} processes" + pool.note)
    print(json.dumps({
        "impl": "reference", "metric": "bubble+GW spectra/sec", "value": value, "unit": "spectra/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.path], "batch_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "spectra/s", "cores": cores, "kind": pool.kind, "sample": sample},
        "e2e": {"value": value, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=RESULT, flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

class ClockSampler:
    # Started BEFORE the warm-up steps: the first nvidia-smi query on a fresh box initialises NVML and was seen to
    # stall a running step by ~250 ms; only samples whose timestamp falls inside the timed region are used.
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms",
                                       os.environ.get("PTT_BENCH_SAMPLER_MS", "100")], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def mark_start(self):
        self.t0 = datetime.datetime.now()

    def stop(self):
        t1 = datetime.datetime.now()
        t0 = getattr(self, "t0", None)
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                if t0 is not None and not (t0 <= ts <= t1):
                    continue
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def _alg_work(plan, stats, n_points_step, n_samples_step):
    """Algorithmic work per step of every stage (DESIGN.md section 3 states the per-unit figures)."""
    from pttools_b200 import engine
    n_sdv = plan.lookup_pass.z.size
    steps = max(1, stats["steps"])
    w = {}
    # K6': banded matrix product, 2 flop per (z_i, knot, profile) over the band ACTUALLY multiplied (trailing
    # no-solution profiles are trimmed: counted from the driver's own group list), + 10 flop per (z_i, T_j) for the
    # weights, once per wall speed
    w["spec_den_v_group"] = {"flop": (2.0 * stats["group_macs"] + 10.0 * n_sdv * NPT[1] * stats["group_count"]) / steps,
                             "bound": "tensor", "precision": "fp64 (DMMA m8n8k4)"}
    # K7': 2 mul + 4 fma per (wavenumber, cell, profile); cells counted from the plan's tables
    _, nseg, meta, _ = engine.gw_cells(plan)
    cells = float(int(nseg.sum().item()) + sum(int(meta[i, 2:2 + int(n)].sum().item()) for i, n in enumerate(nseg.tolist())))
    w["spec_den_gw_cells"] = {"flop": 10.0 * cells * stats["gw_profiles"] / steps, "bound": "fp64", "precision": "fp64 (FMA pipe)"}
    # K5: block-moment sine transform: 2^l blocks x 39 flop (16 moments: 14 fma Horner, 5 flop combine, 6 flop rotation)
    # x 2 functions per exact wavenumber (2^l >= z / 1.6, l <= 5), 16 moments x 2 flop x 2 functions per profile sample
    # for the finest level, ~60 flop per asymptotic wavenumber
    z, fr = plan.z_all, plan.frac_all
    zl = z[fr < 1.0]
    blocks = np.minimum(32.0, 2.0 ** np.ceil(np.log2(np.maximum(zl / 1.6, 1.0))))
    per_profile = 78.0 * blocks.sum() + 60.0 * np.count_nonzero(fr > 0.0)
    w["a2"] = {"flop": (per_profile * stats["a2_profiles"] + 64.0 * (n_samples_step + 2000.0 * stats["a2_profiles"])) / steps,
               "bound": "fp64", "precision": "fp64 (FMA pipe)"}
    # K3: one thread per point, latency bound; its roofline is the HBM write of the materialised profiles,
    # 3 arrays x 8 B per sample
    w["shells_generic"] = {"bytes": 24.0 * n_samples_step / steps, "bound": "hbm", "precision": "fp64"}
    return w


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pttools_b200 import engine, sweep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    full_grid = args.workload == "config5"
    if full_grid:
        # BASELINE.json config 5: the WHOLE 1000 x 1000 grid in one product call per step, sharded over the ranks
        # (strong scaling: the work of a step does not depend on N)
        if (GRID_N * GRID_N) % world:
            sys.exit("config5 needs a number of GPUs that divides 10^6")
        args.batch = GRID_N * GRID_N // world
    B = args.batch
    from pttools_b200.models import BagModel
    from pttools_b200 import fluid_reference
    model = BagModel(alpha_n_min=0.005)
    plan = engine.SsmPlan(Y, cs=float(np.sqrt(model.csb2)), nt=NPT[1], n_z_lookup=NPT[2], nxi=NPT[0], mode="ssm",
                          only_lookup_pass=not args.pow_v)
    fluid_reference.ref()         # starting-guess table: once per process, like the reference's pre-fork warm-up

    points_cache = {}

    def step_points(s):
        """The points of step s for ALL ranks (the product API shards them: whole v_wall rows per rank).  Synthetic
        input: generated before the timed regions (the inputs of a step are resident in host memory when it starts)."""
        if full_grid:
            s = 0                        # every step is the same sweep: the full grid, v_wall-major
        if s not in points_cache:
            if full_grid:
                v_walls, alpha_ns = np.linspace(0.05, 0.95, GRID_N), np.linspace(0.005, 0.5, GRID_N)
                points_cache[s] = (np.repeat(v_walls, GRID_N), np.tile(alpha_ns, GRID_N))
            else:
                parts = [batch_points(s, r, world, B) for r in range(world)]
                points_cache[s] = (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]))
        return points_cache[s]

    # the result table of a step: on rank 0, written by all ranks (CUDA IPC, peer stores over NVLink); allocated once
    # two of them, used alternately: a step's table stays untouched while the next step runs (a consumer on rank 0 has a
    # whole step to read it), so consecutive sweeps need no barrier between them
    shared = [sweep.SharedTable(B * world, Y.size) for _ in range(1 if full_grid else 2)] if world > 1 else None

    def device_step(s):
        """One pass of the hot path through the product's sweep API: rank 0 ends with the Omega_gw,0 table of all
        world x B points in HBM, in the order of the input points (config 5 of BASELINE.json)."""
        vw, an = step_points(s)
        return sweep.sweep_omgw0_distributed(model, vw, an, y=Y, r_star=R_STAR, n_xi=N_XI, chunk=args.chunk,
                                             shell_chunk=args.shell_chunk or None, plan=plan, timing=True, shared=None if shared is None else shared[s % len(shared)],
                                             wait=full_grid or world == 1)

    def barrier():
        torch.cuda.synchronize()      # this rank's rows are in the tables on rank 0 ...
        if world > 1:
            dist.barrier()            # ... and so are everybody else's
        torch.cuda.synchronize()

    total_steps = args.warmup + args.steps
    for s in list(range(total_steps)) + [1000 + k for k in range(args.steps + E2E_WARMUP)]:
        step_points(s)
    sampler = ClockSampler(local) if rank == 0 else None
    import gc
    gc.collect()
    gc.disable()          # no collector pauses inside the timed regions (re-enabled below)
    engine.STAGES.enable(True)          # the warm-up steps run exactly what the timed steps run
    for s in range(args.warmup):
        device_step(s)
    barrier()
    engine.STAGES.enable(True)          # resets the stage times of the warm-up
    if sampler:
        sampler.mark_start()
    launches0 = engine.STAGES.launches
    agg = {k: 0 for k in ("group_macs", "group_count", "group_profiles", "gw_profiles", "a2_profiles", "retried", "rescued",
                          "chunks_before_fold", "chunks_after_fold")}
    waits, step_stages = [], []
    from pttools_b200 import _lib
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    ev0.record()
    n_ok, n_samples = 0, 0.0
    step_ev = [ev0]
    per_step = []                   # results of the timed steps: reduced AFTER the timed region (no host sync inside it)
    for s in range(args.warmup, total_steps):
        table = device_step(s)
        st = sweep.LAST_LOCAL["result"]
        per_step.append((getattr(st._keep, "stats_box", st._keep), st.scalars["npts"], st.status))
        step_ev.append(torch.cuda.Event(enable_timing=True))
        step_ev[-1].record()
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    engine.finish_timing()          # the last step's stage times (read from its events)
    for keep, npts, status in per_step:
        stats = keep.stats
        for k in agg:
            agg[k] += int(getattr(stats, k))
        waits.append((round(stats.host_wait_shells_ms, 1), round(stats.host_wait_retry_ms, 1), int(stats.retried),
                      int(stats.chunks_before_fold)))
        step_stages.append({name: round(float(stats.stage_ms[i]), 1) for i, name in enumerate(_lib.STAGE_NAMES)
                            if stats.stage_calls[i]})
        n_samples += float(torch.nan_to_num(npts).sum())
        n_ok += int((status == 0).sum())
    del per_step
    dev_ms = ev0.elapsed_time(ev1)
    per_step_ms = [a.elapsed_time(b) for a, b in zip(step_ev[:-1], step_ev[1:])]
    stage_ms = engine.STAGES.summary()
    launches = engine.STAGES.launches - launches0
    engine.STAGES.enable(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([dev_ms, t_wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    grid_check = None
    if full_grid and rank == 0:
        # the gathered table against a single-GPU sweep of three of its v_wall rows
        vw_all, an_all = step_points(0)
        rows = [137, 577, 902]
        sel = np.concatenate([np.arange(r * GRID_N, (r + 1) * GRID_N) for r in rows])
        local = sweep.sweep_spectra(model, vw_all[sel], an_all[sel], y=Y, r_star=R_STAR, n_xi=N_XI, chunk=args.chunk, plan=plan,
                                    want_pow_v=False, want_pow_gw=False)
        got = table[torch.as_tensor(sel, device="cuda")]
        fin = torch.isfinite(local.omgw0)
        dev = float(((got[fin] / local.omgw0[fin]) - 1).abs().max()) if bool(fin.any()) else 0.0
        grid_check = {"v_wall_rows": rows, "max_rel_dev_vs_single_gpu_sweep": dev,
                      "nan_pattern_equal": bool(torch.equal(fin, torch.isfinite(got))),
                      "finite_rows_in_table": int(torch.isfinite(table[:, 0]).sum())}
        del local, got
    del table

    # ---- end-to-end: host buffers in, host result table out, through the reference-facing call
    # (analysis.create_spectra semantics at N = 1; at N > 1 every rank delivers its shard of the table to pinned host
    # memory, streamed out chunk by chunk under the compute)
    e2e_steps = min(args.steps, 2) if full_grid else args.steps

    def e2e_step(s):
        vw, an = step_points(1000 + s)
        shard, host_table = sweep.sweep_omgw0_distributed(model, vw, an, y=Y, r_star=R_STAR, n_xi=N_XI, chunk=args.chunk,
                                                          shell_chunk=args.shell_chunk or None, plan=plan, mem="host")
        return float(np.nanmax(host_table[:, 0])) if len(shard) else 0.0

    # warm-up: TWO calls -- the pinned result table of a call is released when the next call has returned, so the host
    # allocator settles on two cached 512 MB blocks; the second cudaHostAlloc (~0.2 s) would otherwise fall into the
    # timed region
    for s in range(E2E_WARMUP):
        e2e_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(E2E_WARMUP, E2E_WARMUP + e2e_steps):
        e2e_step(s)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t[0])
    gc.enable()

    if rank == 0:
        value = B * world * args.steps / (dev_ms * 1e-3)
        peak = engine.fp64_peak_flops()
        peaks_file = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks_file = json.load(f)
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks_file.get("hbm_gbs", 6554.2))
        agg["steps"] = args.steps
        alg = _alg_work(plan, agg, B, n_samples)
        traffic = {}
        for fname in ("r02_traffic.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fname)) as f:
                    tj = json.load(f)
                per = lambda k: tj[k]["read"] + tj[k]["write"]
                rows = max(1, min(args.chunk, B) // GRID_N)
                traffic = {"spec_den_v_group": rows * (per("sdv_gemm_kernel") + per("sdv_weights_kernel")),
                           "spec_den_gw_cells": (per("spec_den_gw_cells_kernel") + per("gw_table_kernel")) * min(args.chunk, B) / tj.get("gw_profiles_per_launch", 2000.0),
                           "_source": "profiles/" + fname}
                break
            except (OSError, KeyError, ValueError):
                continue
        stages = []
        for name, a in alg.items():
            if name not in stage_ms:
                continue
            ms_step = stage_ms[name]["ms"] / args.steps
            n_entries = max(1, stage_ms[name]["n"]) / args.steps
            if a["bound"] == "hbm":
                ach, pk, unit = a["bytes"] / (ms_step * 1e-3) / 1e9, hbm_peak, "GB/s"
            else:
                ach, pk, unit = a["flop"] / (ms_step * 1e-3) / 1e12, peak / 1e12, "TFLOP/s"
            stages.append({"kernel": name, "bound": a["bound"], "achieved": ach, "peak": pk, "unit": unit, "frac": ach / pk,
                           "traffic": traffic.get(name), "precision": a["precision"], "ms_per_step": ms_step,
                           "entries_per_step": n_entries, "share_of_step": stage_ms[name]["ms"] / dev_ms})
        stages.sort(key=lambda d: -d["share_of_step"])
        roofline = dict(stages[0]) if stages else None
        if roofline:
            roofline["peak_source"] = ("in-run FP64 FMA probe kernel (MEASURED_PEAKS.json has no FP64 entry); a DMMA-only probe "
                                       "measures the same peak (engine.fp64_pipe_flops)") if roofline["unit"] == "TFLOP/s" \
                else "MEASURED_PEAKS.json hbm_gbs"
            roofline["traffic_source"] = traffic.get("_source")
        cores = len(os.sched_getaffinity(0))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_leg(B, cores, args)
        out = {
            "metric": "bubble+GW spectra/sec", "value": value, "unit": "spectra/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if full_grid else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": (WORKLOADS["config5"] if full_grid else WORKLOADS["generic"]),
                       "batch_per_step": B * world, "batch_per_gpu_per_step": B,
                       "chunk": args.chunk, "api": "sweep.sweep_omgw0_distributed -> ptt_sweep_spectra (one C call per rank and "
                       "step); whole v_wall rows per rank; every rank writes its rows into the table on rank 0 (CUDA IPC, peer "
                       "stores over NVLink) chunk by chunk as they are produced" +
                       ("" if full_grid or world == 1 else "; consecutive steps write two tables alternately and are not separated "
                        "by a barrier (wait=False; the timed region ends with stream synchronisation + barrier)"),
                       "l2": "per-step working set (profiles + A2 + lookup tables, ~1 MB per point) exceeds the 126 MB L2; "
                             "every step uses fresh points",
                       **({"table_check": grid_check} if grid_check is not None else {}),
                       "valid_points_in_timed_region_rank0": n_ok, "per_step_ms": per_step_ms,
                       "wall_ms_per_step": wall_ms / args.steps, "stage_ms_by_step_rank0": step_stages,
                       "retried_points_rank0": agg["retried"],
                       "rescued_points_rank0": agg["rescued"],
                       "host_waits_rank0": {"per_step": waits, "columns": ["wait_shells_ms", "wait_retry_ms", "retried_points",
                                                                           "chunks_before_fold"]}},
            "clocks": clocks,
            "e2e": {"value": B * world * e2e_steps / e2e_s, "unit": "spectra/s", "h2d_bytes_per_step": (16 + 4 + 1) * 8 * B * world,
                    "d2h_bytes_per_step": (8 * int(Y.size) + 32 * 8 + 4) * B * world, "steps": e2e_steps,
                    "api": "sweep.sweep_omgw0_distributed(mem='host') -> ptt_sweep_spectra(PTT_MEM_HOST): numpy in, pinned numpy out"},
            "gpu_launches": launches,
            "stage_ms_per_step": {k: v["ms"] / args.steps for k, v in stage_ms.items()},
            "roofline": roofline,
            "roofline_stages": stages,
            "cpu_baseline": cpu,
        }
        print(json.dumps(out), file=RESULT, flush=True)
    if world > 1:
        dist.barrier()
        for sh in shared or ():
            sh.close()
        dist.destroy_process_group()


def cpu_baseline_leg(B, cores, args):
    n_cpu = 4 * max(cores, 8)
    vw, an = batch_points(0, 0, 1, B)
    pick = np.linspace(0, B - 1, n_cpu).astype(int)
    pool = CpuPool(cores, "generic", prefer_reference=not args.cpu_port)
    dt = pool.run(list(zip(vw[pick], an[pick])))
    pool.close()
    return {"value": n_cpu / dt, "unit": "spectra/s", "cores": cores, "kind": pool.kind,
            "sample": f"{n_cpu} points spread evenly over the batch of step 0 (same grid, same sizes), "
                      f"fork pool of {cores} processes, {dt:.1f} s" + pool.note}


# ---------------------------------------------------------------------------------------------------------------
# The other configurations of BASELINE.json (1 GPU): --workload config3 | config2-shells | config4
# ---------------------------------------------------------------------------------------------------------------

def _timed(step, steps, warmup, sync_each=True):
    """warmup untimed calls of step(s), then `steps` timed ones: device ms (CUDA events on the launching stream), wall ms and
    the time of every step.  sync_each: every step ends with a synchronisation (per-step wall times); without it the host
    preparation of a step overlaps the device work of the previous one and the per-step times are device times between
    events.  (The shells-only sweeps measured slower without it, 36.7 against 24.5 ms per step, gpurun_out/bench_final3_*.)"""
    import torch
    for s in range(warmup):
        step(s)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    per = []
    for k, s in enumerate(range(warmup, warmup + steps)):
        t1 = time.perf_counter()
        step(s)
        if sync_each:
            torch.cuda.synchronize()
            per.append(1e3 * (time.perf_counter() - t1))
        ev[k + 1].record()
    torch.cuda.synchronize()
    if not sync_each:
        per = [a.elapsed_time(b) for a, b in zip(ev[:-1], ev[1:])]
    return ev[0].elapsed_time(ev[-1]), 1e3 * (time.perf_counter() - t0), per


def _peaks(engine):
    hbm = 6554.2
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f).get("hbm_gbs", hbm))
    except (OSError, ValueError):
        pass
    return engine.fp64_peak_flops() / 1e12, hbm


def _emit(out, sampler):
    out["clocks"] = sampler.stop() if sampler else None
    print(json.dumps(out), file=RESULT, flush=True)


def config3_models():
    """Config 3: ConstCSModel(css2, csb2, a_s=5, a_b=1, V_s=1, alpha_n_min=0.05) for css2, csb2 in linspace(1/4, 1/3, 8)
    (the reference's example uses the corners {1/3, 1/4}^2, examples/const_cs/const_cs_gw.py:153-157), grouped by csb2.
    Combinations for which the reference's parameter search raises are left out (counted)."""
    import logging
    from pttools_b200.models import ConstCSModel
    logging.disable(logging.CRITICAL)
    vals = np.linspace(0.25, 1 / 3, 8)
    fams, skipped = [], 0
    for csb2 in vals:
        fam = []
        for css2 in vals:
            try:
                fam.append(ConstCSModel(css2=float(css2), csb2=float(csb2), a_s=5, a_b=1, V_s=1, alpha_n_min=0.05))
            except Exception:       # noqa: BLE001 - the reference's constructor rejects the combination
                skipped += 1
        fams.append(fam)
    logging.disable(logging.NOTSET)
    return fams, skipped


def workload_config3(args):
    import torch
    from pttools_b200 import engine, fluid_reference, sweep
    torch.cuda.set_device(0)
    fluid_reference.ref()
    fams, skipped = config3_models()
    v_walls, alpha_ns = np.linspace(0.05, 0.95, 128), np.linspace(0.05, 0.5, 128)
    kw = dict(y=Y, r_star=R_STAR, Tn=200, n_xi=N_XI, want_pow_v=False, chunk=args.chunk)
    counts = {}

    def step(s, mem="device"):
        fam = fams[s % len(fams)]
        res = sweep.sweep_models(fam, v_walls, alpha_ns, mem=mem, timing=mem == "device", **kw)
        counts[s % len(fams)] = sum(int(r["mask"].sum()) for r in res)
        return res

    sampler = ClockSampler(0)
    for s in range(len(fams)):       # every family once: plans, cell tables, point counts
        step(s)
    engine.STAGES.enable(True)          # the warm-up steps run exactly what the timed steps run
    for s in range(args.warmup):
        step(s)
    sampler.mark_start()
    engine.STAGES.enable(True)          # resets the stage times of the warm-up
    # sweeps queue behind one another (device tables): no synchronisation between the steps
    dev_ms, wall_ms, per = _timed(lambda s: step(s + args.warmup), args.steps, 0, sync_each=False)
    stage_ms = engine.STAGES.summary()
    macs = int(engine.STAGES.c_counts.get("group_macs", 0))
    engine.STAGES.enable(False)
    n_pts = sum(counts[s % len(fams)] for s in range(args.warmup, args.warmup + args.steps))
    e2e_ms, _, _ = _timed(lambda s: step(s, mem="host"), args.steps, 1)
    peak, hbm = _peaks(engine)
    roof = None
    if macs and "spec_den_v_group" in stage_ms:
        # the dominant stage as in the headline line: multiply-adds of the banded products the driver launched (after
        # trimming rows without a shell) over the device time of the stage, both summed over the same sweeps
        ach = 2.0 * macs / (stage_ms["spec_den_v_group"]["ms"] * 1e-3) / 1e12
        roof = {"kernel": "spec_den_v_group", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "precision": "fp64 (DMMA m8n8k4)",
                "share_of_step": stage_ms["spec_den_v_group"]["ms"] / max(sum(v["ms"] for v in stage_ms.values()), 1e-9)}
    cores = len(os.sched_getaffinity(0))
    cpu = None
    if not args.no_cpu_baseline:
        m = fams[-1][-1]
        n_cpu = 2 * max(cores, 8)
        iv = np.linspace(0, 127, n_cpu).astype(int)
        pts = [(float(v_walls[i]), float(alpha_ns[(37 * k) % 128])) for k, i in enumerate(iv)]
        dt = cpu_run_model(pts, cores, (m.css2, m.csb2, m.a_s, m.a_b, m.V_s))
        cpu = {"value": n_cpu / dt, "unit": "spectra/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} points of the 128x128 grid for the model (css2, csb2) = ({m.css2:.4f}, {m.csb2:.4f}), fork pool, {dt:.1f} s"}
    _emit({"metric": "bubble+GW spectra/sec", "value": n_pts / (dev_ms * 1e-3), "unit": "spectra/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "config3: 64 ConstCSModel(css2, csb2, a_s=5, a_b=1, V_s=1, alpha_n_min=0.05), css2, csb2 in "
                                  "linspace(1/4, 1/3, 8); v_wall=linspace(0.05,0.95,128) x alpha_n=linspace(0.05,0.5,128) restricted "
                                  "to alpha_n >= model.alpha_n_min; Spectrum(r_star=0.1, Tn=200) at the default sizes; one step = the "
                                  "8 models that share csb2 in ONE sweep ordered by v_wall",
                      "models_skipped_by_constructor": skipped, "points_per_family": [counts[k] for k in sorted(counts)],
                      "per_step_wall_ms": per, "l2": "working set per step far above the 126 MB L2"},
           "e2e": {"value": n_pts / (e2e_ms * 1e-3), "unit": "spectra/s", "steps": args.steps,
                   "h2d_bytes_per_step": 21 * 8 * n_pts // args.steps, "d2h_bytes_per_step": (8 * Y.size + 260) * n_pts // args.steps},
           "gpu_launches": int(engine.STAGES.launches), "stage_ms_per_step": {k: v["ms"] / args.steps for k, v in stage_ms.items()},
           "roofline": roof, "cpu_baseline": cpu}, sampler)


def workload_config2_shells(args):
    import torch
    from pttools_b200 import bubble, engine, fluid_reference
    from pttools_b200.models import BagModel
    torch.cuda.set_device(0)
    model = BagModel(alpha_n_min=0.005)
    fluid_reference.ref()
    B = args.batch
    sampler = ClockSampler(0)
    stats = {"samples": 0.0, "ok": 0, "n": 0}
    tables = {}

    def prep(s):
        if s not in tables:
            vw, an = batch_points(s, 0, 1, B)
            pt = bubble.point_table(model, vw, an)
            tables[s] = (vw, an, pt, engine.dev_f64(pt.pg), engine.dev_f64(pt.pm))
        return tables[s]

    def step_generic(s, count=True):
        vw, an, pt, pg, pm = prep(s)
        out = engine.sweep_c(None, pg, pm, n_xi=N_XI, rows=False, timing=True)
        if count:
            stats["samples"] += float(torch.nan_to_num(out.scalars[:, 31]).sum())
            stats["ok"] += int((out.status == 0).sum())
            stats["n"] += B
        return out

    def step_bag(s):
        vw, an, *_ = prep(s)
        return engine.sweep_shells_bag(vw, an, n_xi=N_XI)

    def step_host(s):
        vw, an, pt, *_ = prep(s)
        return engine.sweep_c(None, pt.pg, pt.pm, n_xi=N_XI, mem="host")

    total = args.warmup + args.steps
    for s in range(total):
        prep(s)
    sampler.mark_start()
    engine.STAGES.enable(True)
    dev_ms, wall_ms, per = _timed(step_generic, args.steps, args.warmup)
    stage_ms = engine.STAGES.summary()
    engine.STAGES.enable(False)
    bag_ms, _, _ = _timed(step_bag, args.steps, args.warmup)
    bag = step_bag(args.warmup)
    bag_samples = float(bag.npts[bag.status == 0].sum())
    e2e_ms, _, _ = _timed(step_host, args.steps, 1)
    # per solution type: pure batches of one type drawn from the grid
    per_type = {}
    v_walls, alpha_ns = np.linspace(0.05, 0.95, GRID_N), np.linspace(0.005, 0.5, GRID_N)
    vw_all, an_all = np.repeat(v_walls, 40), np.tile(alpha_ns[::25], GRID_N)
    codes = model.solution_type_codes(vw_all, an_all, np.asarray(model.wn(an_all)))
    for code, name in ((0, "sub_def"), (1, "hybrid"), (2, "detonation")):
        sel = np.nonzero(codes == code)[0][:16000]
        pt = bubble.point_table(model, vw_all[sel], an_all[sel])
        pg, pm = engine.dev_f64(pt.pg), engine.dev_f64(pt.pm)
        ms, _, _ = _timed(lambda s: engine.sweep_c(None, pg, pm, n_xi=N_XI), 3, 2)
        per_type[name] = {"points": int(sel.size), "shells_per_s": sel.size * 3 / (ms * 1e-3)}
    peak, hbm = _peaks(engine)
    k3 = stage_ms.get("shells_generic", {"ms": 0.0})["ms"] / args.steps
    bytes_step = 24.0 * stats["samples"] / args.steps
    roof = {"kernel": "shells_generic", "bound": "hbm", "achieved": bytes_step / (k3 * 1e-3) / 1e9 if k3 else None, "peak": hbm,
            "unit": "GB/s", "frac": (bytes_step / (k3 * 1e-3) / 1e9 / hbm) if k3 else None, "traffic": None,
            "ms_per_step": k3, "note": "one thread per point, latency bound: the profile rows (3 x 8 B per sample) are the only "
            "algorithmic HBM traffic"}
    cores = len(os.sched_getaffinity(0))
    cpu = None
    if not args.no_cpu_baseline:
        n_cpu = 16 * max(cores, 8)
        vw, an = batch_points(0, 0, 1, B)
        pick = np.linspace(0, B - 1, n_cpu).astype(int)
        dt = cpu_run(list(zip(vw[pick], an[pick])), cores, "shells")
        cpu = {"value": n_cpu / dt, "unit": "shells/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} points of the batch of step 0 through the port of sound_shell_generic, fork pool, {dt:.1f} s"}
    _emit({"metric": "fluid shells/sec", "value": B * args.steps / (dev_ms * 1e-3), "unit": "shells/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "config2-shells: BagModel(alpha_n_min=0.005) grid v_wall=linspace(0.05,0.95,1000) x "
                                  "alpha_n=linspace(0.005,0.5,1000), n_xi=5000; value = Bubble(model, v_wall, alpha_n).solve() "
                                  "(model-independent solver + the reference's retries + thermodynamic scalars, ptt_sweep_spectra "
                                  "without spectra), inputs resident in HBM",
                      "batch_per_step": B, "valid_fraction": stats["ok"] / max(1, stats["n"]),
                      "profile_samples_per_shell": stats["samples"] / max(1, stats["ok"]), "per_step_wall_ms": per,
                      "variants": {"sound_shell_bag (old API, ptt_sweep_shells_bag)": {
                          "shells_per_s": B * args.steps / (bag_ms * 1e-3), "ms_per_step": bag_ms / args.steps,
                          "hbm_write_GBps": 24.0 * bag_samples / (bag_ms / args.steps * 1e-3) / 1e9},
                          "per_solution_type (generic solver, pure batches)": per_type}},
           "e2e": {"value": B * args.steps / (e2e_ms * 1e-3), "unit": "shells/s", "steps": args.steps,
                   "h2d_bytes_per_step": 20 * 8 * B, "d2h_bytes_per_step": 260 * B},
           "gpu_launches": int(engine.STAGES.launches), "stage_ms_per_step": {k: v["ms"] / args.steps for k, v in stage_ms.items()},
           "roofline": roof, "cpu_baseline": cpu}, sampler)


def workload_config4(args):
    """Config 4: batched sin_transform, 10^4 profiles x 4096 wavenumbers x 8192 xi (f_p = max(0, cos(xi + phi_p)) on
    [xi_a, xi_b], numpy default_rng(0): generalises tests/test_performance.py:76-87 of the reference), then
    spec_den_gw_scaled with z_lookup[10^4] and y[4096] for the same number of profiles."""
    import torch
    from pttools_b200 import engine
    torch.cuda.set_device(0)
    P, n_z, n_xi = 10000, 4096, 8192
    rng = np.random.default_rng(0)
    phi, xa = rng.uniform(0, 1, P), rng.uniform(0.05, 0.5, P)
    xb = xa + rng.uniform(0.05, 0.45, P)
    xi = np.linspace(0, 1, n_xi)
    f = np.maximum(0, np.cos(xi[None, :] + phi[:, None])) * ((xi[None, :] > xa[:, None]) & (xi[None, :] < xb[:, None]))
    xi_h = np.broadcast_to(xi, (P, n_xi)).copy()
    xi_d, f_d = engine.dev_f64(xi_h), engine.dev_f64(f)
    off, npts = np.zeros(P, np.int32), np.full(P, n_xi, np.int32)
    z_mixed = np.logspace(-1, 3, n_z)
    z_exact = np.logspace(-1, np.log10(50), n_z)
    sampler = ClockSampler(0)
    sampler.mark_start()
    ms_mixed, _, per = _timed(lambda s: engine.sin_transform_batch(z_mixed, xi_d, f_d, off, npts, method="moments"), args.steps, args.warmup)
    ms_exact, _, _ = _timed(lambda s: engine.sin_transform_batch(z_exact, xi_d, f_d, off, npts, method="moments"), args.steps, args.warmup)
    pin_x, pin_f = torch.from_numpy(xi_h).pin_memory(), torch.from_numpy(f).pin_memory()
    pin_o = torch.empty((P, n_z), dtype=torch.float64).pin_memory()

    def e2e(s):
        r = engine.sin_transform_batch(z_mixed, pin_x.to("cuda", non_blocking=True), pin_f.to("cuda", non_blocking=True), off, npts,
                                       method="moments")
        pin_o.copy_(r, non_blocking=True)
    e2e_ms, _, _ = _timed(e2e, args.steps, 1)
    # GW integral on the same scale
    y4 = np.logspace(-1, 3, 4096)
    gplan = engine.GwOnlyPlan(10.0 ** np.linspace(np.log10(y4.min() * 0.5 * (1 - engine.CS0) / engine.CS0 - 1e-8),
                                                  np.log10(y4.max() * 0.5 * (1 + engine.CS0) / engine.CS0 + 1e-8), 10000), y4)
    zl = gplan.z_lookup
    pv = engine.dev_f64((zl ** 3 / (1 + (zl / 10) ** 8))[None, :] * rng.uniform(0.5, 2, (2000, 1)))
    ms_gw, _, _ = _timed(lambda s: engine.spec_den_gw_batch(gplan, pv, want_sd=False), 3, 2)
    peak, hbm = _peaks(engine)
    n_lo = int((z_mixed <= 50).sum())
    blocks = np.minimum(128.0, 2.0 ** np.ceil(np.log2(np.maximum(z_mixed[z_mixed <= 50] / 0.4, 1.0))))
    flop = P * (22.0 * blocks.sum() + 60.0 * (n_z - n_lo) + 24.0 * n_xi)
    byts = P * (2.0 * n_xi + n_z) * 8
    dur = ms_mixed / args.steps * 1e-3
    roof = {"kernel": "sin_transform (block moments)", "bound": "hbm", "achieved": byts / dur / 1e9, "peak": hbm, "unit": "GB/s",
            "frac": byts / dur / 1e9 / hbm, "traffic": None, "fp64_TFLOPs": flop / dur / 1e12, "fp64_frac": flop / dur / 1e12 / peak,
            "note": "algorithmic bytes = profiles in (xi, f) + transform out; flops = 22 per (exact wavenumber, block) + 24 per sample"}
    cores = len(os.sched_getaffinity(0))
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import ssm
        n_cpu = 8
        t0 = time.perf_counter()
        for p in range(n_cpu):
            ssm.sin_transform(z_mixed, xi, f[p])
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": "profiles/s", "cores": 1, "kind": "port",
               "sample": f"{n_cpu} profiles through the numpy port of sin_transform (one thread), {dt:.1f} s"}
    _emit({"metric": "sin_transform profiles/sec", "value": P * args.steps / (ms_mixed * 1e-3), "unit": "profiles/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_mixed / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "config4: sin_transform of 10^4 profiles x 4096 wavenumbers z=logspace(-1,3) (mixed exact / asymptotic) "
                                  "x 8192 xi samples, inputs resident in HBM (1.3 GB, above the L2)",
                      "all_exact_z_ms_per_step": ms_exact / args.steps, "per_step_wall_ms": per,
                      "sine_products_per_s_equivalent": P * n_lo * n_xi / dur,
                      "spec_den_gw_scaled": {"profiles": 2000, "n_y": 4096, "n_z_lookup": 10000, "ms": ms_gw / 3,
                                             "profiles_per_s": 2000 * 3 / (ms_gw * 1e-3)}},
           "e2e": {"value": P * args.steps / (e2e_ms * 1e-3), "unit": "profiles/s", "steps": args.steps,
                   "h2d_bytes_per_step": int(P * 2 * n_xi * 8), "d2h_bytes_per_step": int(P * n_z * 8)},
           "gpu_launches": 3 * (2 * args.steps + 2 * args.warmup), "roofline": roof, "cpu_baseline": cpu}, sampler)


def cpu_run_model(points, cores, model_params):
    import multiprocessing as mp
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import models, shell_generic as sg
    css2, csb2, a_s, a_b, V_s = model_params
    _CPU["path"] = "generic"
    _CPU["model"] = models.const_cs_model(css2, csb2, a_s, a_b, V_s)
    _CPU["ref"] = sg.FluidReference(dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_fluid_reference.npz"))))
    _CPU["skip_bag_type"] = True
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_point, points, chunksize=1)
    _CPU.pop("skip_bag_type", None)
    return time.perf_counter() - t0


RESULT = sys.stdout


def _claim_stdout():
    """The contract is ONE JSON line on stdout, but libraries write there too (NCCL prints its version banner to
    fd 1): point fd 1 at stderr for the duration of the run and keep the real stdout for the result line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64000)
    ap.add_argument("--chunk", type=int, default=8000)
    ap.add_argument("--shell-chunk", type=int, default=0, help="points per shell solve (0: from the free device memory)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="generic", choices=["generic"])
    ap.add_argument("--workload", default="headline", choices=["headline", "config5", "config3", "config2-shells", "config4"],
                    help="headline = configs 2/5 of BASELINE.json (the metric's configuration); the others are extra lines")
    ap.add_argument("--pow-v", action="store_true", help="also compute pow_v (the y pass of spec_den_v)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="CPU arm / cpu_baseline: the oracle port even when the real "
                    "reference (oracle/_ref) imports")
    a = ap.parse_args()
    RESULT = _claim_stdout()
    if a.impl == "reference":
        reference_arm(a)
    elif a.workload in ("headline", "config5"):
        gpu_arm(a)
    else:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            sys.exit("the extra workloads run on one GPU")
        {"config3": workload_config3, "config2-shells": workload_config2_shells, "config4": workload_config4}[a.workload](a)
