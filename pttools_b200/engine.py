"""Batched device operators: thin torch-tensor wrappers over the C ABI plus the sweep-level grid plans.

torch is plumbing here (device memory, streams); all arithmetic on the hot path runs in the hand-written CUDA
kernels of libpttools_b200.so.  The "plans" hold the grids that depend only on (y, cs, nt, ...): they are built
with the same numpy expressions the reference uses (pttools/ssmtools/spectrum.py:169-186, 503-515, 473) so that
the knots are bit-identical, and uploaded once per sweep.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import typing as tp

import numpy as np

from . import _lib
from ._lib import A2Plan, Eos, GwCells, GwPlan, SdvPlan

N_XI_DEFAULT = 5000
Y_DEFAULT = np.logspace(-1, 3, 1000)
Z_ST_THRESH = 50.0
DZ_ST_BLEND = np.pi
T_TILDE_MIN, T_TILDE_MAX = 0.01, 20.0
NXI_DEFAULT, NT_DEFAULT, NZ_LOOKUP_DEFAULT = 2000, 10000, 10000
GAMMA_DEFAULT = 4.0 / 3.0
CS0 = 1.0 / np.sqrt(3.0)

SMEM_WINDOW_BYTES = 200 * 1024     # dynamic shared memory budget for the lookup windows (B200: 227 KB per CTA)


def _torch():
    return _lib.require_cuda()


class _Stages:
    """Counts kernel launches and (when enabled) brackets every C-ABI call with CUDA events on the launching
    stream, so that bench.py can report per-kernel device time measured inside its timed region."""

    def __init__(self):
        self.enabled = False
        self.launches = 0
        self._ev = []
        self.c_stage_ms = {}        # stage times reported by the C sweep driver (ptt_sweep_last_stats)
        self.c_counts = {}          # what those stages processed (multiply-adds of the matrix products, profiles ...)

    def enable(self, on: bool):
        finish_timing()
        self.enabled = on
        self._ev = []
        self.c_stage_ms = {}
        self.c_counts = {}

    class _Ctx:
        def __init__(self, outer, name, launches):
            self.o, self.name, self.n = outer, name, launches

        def __enter__(self):
            self.o.launches += self.n
            if self.o.enabled:
                import torch
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record()
            return self

        def __exit__(self, *exc):
            if self.o.enabled:
                self.b.record()
                self.o._ev.append((self.name, self.a, self.b))
            return False

    def stage(self, name, launches=1):
        return self._Ctx(self, name, launches)

    def summary(self):
        import torch
        finish_timing()
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self._ev:
            d = out.setdefault(name, {"ms": 0.0, "n": 0})
            d["ms"] += a.elapsed_time(b)
            d["n"] += 1
        for name, v in self.c_stage_ms.items():
            d = out.setdefault(name, {"ms": 0.0, "n": 0})
            d["ms"] += v["ms"]
            d["n"] += v["n"]
        return out


STAGES = _Stages()


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_PREP_STREAMS = {}


def prep_stream():
    """High-priority side stream for the small device work that PREPARES a sweep (alpha_n_max per wall speed, starting
    guesses, suppression factors, parameter rows).  A device-table sweep returns with its last chunks in flight; on the
    main stream the next sweep's preparation -- which ends in host reads -- would queue behind them."""
    torch = _torch()
    dev = torch.cuda.current_device()
    if dev not in _PREP_STREAMS:
        _PREP_STREAMS[dev] = torch.cuda.Stream(device=dev, priority=-1)
    return _PREP_STREAMS[dev]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def dev_f64(a, device=None):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        return a.to(device=device or "cuda", dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64), device=device or "cuda")


# ---------------------------------------------------------------------------------------------------------------
# Shell operators
# ---------------------------------------------------------------------------------------------------------------

def integrate_batch(v0, w0, xi0, phase, t_end, n_xi, css2=1 / 3, csb2=1 / 3):
    """Batched fluid_integrate_param (pttools/bubble/integrate.py:81-135). Returns v, w, xi [n, n_xi], status[n]."""
    torch = _torch()
    lib = _lib.load()
    v0, w0, xi0, phase, t_end = (dev_f64(a) for a in (v0, w0, xi0, phase, t_end))
    n = v0.numel()
    out = [torch.empty((n, n_xi), dtype=torch.float64, device=v0.device) for _ in range(3)]
    status = torch.empty(n, dtype=torch.int32, device=v0.device)
    eos = Eos(kind=0, css2=css2, csb2=csb2, a_s=0, a_b=0, V_s=0, V_b=0, T_ref=1)
    _lib.check(lib.ptt_integrate_batch(n, _ptr(v0), _ptr(w0), _ptr(xi0), _ptr(phase), _ptr(t_end), n_xi, C.byref(eos),
                                       _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(status), _lib.MEM_DEVICE,
                                       _stream_ptr(torch)), "ptt_integrate_batch")
    return out[0], out[1], out[2], status


def bag_alpha_n_max(v_wall):
    """alpha_n_max_deflagration_bag for an array of wall speeds (pttools/bubble/alpha.py:44-50)."""
    torch = _torch()
    lib = _lib.load()
    v_wall = dev_f64(v_wall)
    out = torch.empty_like(v_wall)
    status = torch.empty(v_wall.numel(), dtype=torch.int32, device=v_wall.device)
    with STAGES.stage("bag_alpha_n_max", 1):
        _lib.check(lib.ptt_bag_alpha_n_max(v_wall.numel(), _ptr(v_wall), _ptr(out), _ptr(status), _lib.MEM_DEVICE,
                                           _stream_ptr(torch)), "ptt_bag_alpha_n_max")
    return out, status


def bag_alpha_n_max_unique(v_wall):
    """Same, but evaluated once per distinct wall speed and broadcast (the quantity depends on v_wall only; the
    reference recomputes it up to four times per point: transition.py:27, alpha.py:185,316, bubble.py:252)."""
    torch = _torch()
    v_wall = dev_f64(v_wall)
    uniq, inv = torch.unique(v_wall, return_inverse=True)
    am, st = bag_alpha_n_max(uniq)
    return am[inv], st[inv]


@dataclasses.dataclass
class ShellBatch:
    """Profiles of a batch of shells in device memory. Shell i is rows_*[i, off[i] : off[i] + npts[i]]."""
    v_wall: tp.Any
    alpha_n: tp.Any
    alpha_plus: tp.Any
    sol_type: tp.Any
    status: tp.Any
    rows_v: tp.Any
    rows_w: tp.Any
    rows_xi: tp.Any
    off: tp.Any
    npts: tp.Any
    scalars: tp.Any
    n_xi: int

    def __len__(self):
        return self.rows_v.shape[0]

    def slice(self, a, b):
        """Row-block view [a, b) of the batch (no copies)."""
        f = lambda t: None if t is None else t[a:b]
        return ShellBatch(f(self.v_wall), f(self.alpha_n), f(self.alpha_plus), f(self.sol_type), f(self.status),
                          self.rows_v[a:b], self.rows_w[a:b], self.rows_xi[a:b], self.off[a:b], self.npts[a:b],
                          f(self.scalars), self.n_xi)

    def take(self, idx):
        """The shells at positions idx (int64 tensor) as a batch of their own (copies)."""
        f = lambda t: None if t is None else t[idx].contiguous()
        return ShellBatch(f(self.v_wall), f(self.alpha_n), f(self.alpha_plus), f(self.sol_type), f(self.status),
                          self.rows_v[idx], self.rows_w[idx], self.rows_xi[idx], f(self.off), f(self.npts),
                          f(self.scalars), self.n_xi)

    def profile(self, i):
        o, n = int(self.off[i]), int(self.npts[i])
        return (self.rows_v[i, o:o + n].cpu().numpy(), self.rows_w[i, o:o + n].cpu().numpy(),
                self.rows_xi[i, o:o + n].cpu().numpy())


def sweep_shells_bag(v_wall, alpha_n, n_xi=N_XI_DEFAULT, an_max=None) -> ShellBatch:
    """sound_shell_bag for a batch of points (pttools/bubble/fluid_bag.py:33-79)."""
    torch = _torch()
    lib = _lib.load()
    v_wall, alpha_n = dev_f64(v_wall).reshape(-1), dev_f64(alpha_n).reshape(-1)
    n = v_wall.numel()
    if an_max is None:
        an_max, _ = bag_alpha_n_max_unique(v_wall)
    cap = lib.ptt_profile_row_capacity(n_xi)
    dev = v_wall.device
    rows = [torch.empty((n, cap), dtype=torch.float64, device=dev) for _ in range(3)]
    off = torch.empty(n, dtype=torch.int32, device=dev)
    npts = torch.empty(n, dtype=torch.int32, device=dev)
    ap = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    scalars = torch.empty((n, 4), dtype=torch.float64, device=dev)
    with STAGES.stage("shells_bag", 2):
        _lib.check(lib.ptt_sweep_shells_bag(n, _ptr(v_wall), _ptr(alpha_n), n_xi, _ptr(an_max), _ptr(rows[0]),
                                            _ptr(rows[1]), _ptr(rows[2]), cap, _ptr(off), _ptr(npts), _ptr(ap), _ptr(st),
                                            _ptr(scalars), _ptr(status), _lib.MEM_DEVICE, _stream_ptr(torch)),
                   "ptt_sweep_shells_bag")
    return ShellBatch(v_wall, alpha_n, ap, st, status, rows[0], rows[1], rows[2], off, npts, scalars, n_xi)


def bag_find_alpha_plus(v_wall, alpha_n, n_xi=N_XI_DEFAULT, an_max=None):
    torch = _torch()
    lib = _lib.load()
    v_wall, alpha_n = dev_f64(v_wall).reshape(-1), dev_f64(alpha_n).reshape(-1)
    n = v_wall.numel()
    if an_max is None:
        an_max, _ = bag_alpha_n_max_unique(v_wall)
    ap = torch.empty(n, dtype=torch.float64, device=v_wall.device)
    st = torch.empty(n, dtype=torch.int32, device=v_wall.device)
    status = torch.empty(n, dtype=torch.int32, device=v_wall.device)
    _lib.check(lib.ptt_bag_find_alpha_plus(n, _ptr(v_wall), _ptr(alpha_n), n_xi, _ptr(an_max), _ptr(ap), _ptr(st),
                                           _ptr(status), _lib.MEM_DEVICE, _stream_ptr(torch)), "ptt_bag_find_alpha_plus")
    return ap, st, status


def bag_profiles(v_wall, alpha_plus, sol_type, n_xi=N_XI_DEFAULT, w_n=None) -> ShellBatch:
    """sound_shell_alpha_plus for a batch (pttools/bubble/fluid_bag.py:83-222)."""
    torch = _torch()
    lib = _lib.load()
    v_wall, alpha_plus = dev_f64(v_wall).reshape(-1), dev_f64(alpha_plus).reshape(-1)
    n = v_wall.numel()
    dev = v_wall.device
    sol_type = torch.as_tensor(sol_type, dtype=torch.int32, device=dev).reshape(-1).contiguous()
    w_n_t = dev_f64(w_n).reshape(-1) if w_n is not None else None
    cap = lib.ptt_profile_row_capacity(n_xi)
    rows = [torch.empty((n, cap), dtype=torch.float64, device=dev) for _ in range(3)]
    off = torch.empty(n, dtype=torch.int32, device=dev)
    npts = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    scalars = torch.empty((n, 4), dtype=torch.float64, device=dev)
    _lib.check(lib.ptt_bag_profiles(n, _ptr(v_wall), _ptr(alpha_plus), _ptr(sol_type), n_xi, _ptr(w_n_t), _ptr(rows[0]),
                                    _ptr(rows[1]), _ptr(rows[2]), cap, _ptr(off), _ptr(npts), _ptr(scalars), _ptr(status),
                                    _lib.MEM_DEVICE, _stream_ptr(torch)), "ptt_bag_profiles")
    return ShellBatch(v_wall, None, alpha_plus, sol_type, status, rows[0], rows[1], rows[2], off, npts, scalars, n_xi)


GEN_SCALARS = ("vp", "vm", "vp_tilde", "vm_tilde", "v_sh", "vm_sh", "vm_tilde_sh", "wp", "wm", "wm_sh", "v_cj", "wn",
               "wn_estimate", "solver_failed")


def sweep_shells_generic(pg, n_xi=N_XI_DEFAULT, t_end=50.0, thin_shell_limit=100, wn_rtol=1e-4) -> ShellBatch:
    """sound_shell_generic for a batch of points (pttools/bubble/fluid.py:848-1052); pg[n,16] as in the C header."""
    torch = _torch()
    lib = _lib.load()
    pg = dev_f64(pg)
    n = pg.shape[0]
    cap = lib.ptt_generic_row_capacity(n_xi)
    dev = pg.device
    rows = [torch.empty((n, cap), dtype=torch.float64, device=dev) for _ in range(3)]
    off = torch.empty(n, dtype=torch.int32, device=dev)
    npts = torch.empty(n, dtype=torch.int32, device=dev)
    status = torch.empty(n, dtype=torch.int32, device=dev)
    scalars = torch.empty((n, 16), dtype=torch.float64, device=dev)
    with STAGES.stage("shells_generic", 1):
        _lib.check(lib.ptt_sweep_shells_generic(n, _ptr(pg), n_xi, float(t_end), int(thin_shell_limit), float(wn_rtol),
                                                _ptr(rows[0]), _ptr(rows[1]), _ptr(rows[2]), cap, _ptr(off), _ptr(npts),
                                                _ptr(scalars), _ptr(status), _lib.MEM_DEVICE, _stream_ptr(torch)),
                   "ptt_sweep_shells_generic")
    return ShellBatch(pg[:, 0].contiguous(), pg[:, 1].contiguous(), None, pg[:, 2].to(torch.int32), status, rows[0],
                      rows[1], rows[2], off, npts, scalars, n_xi)


def generic_retry(pg, idx_host, n_xi=N_XI_DEFAULT, t_end=50.0, thin_shell_limit=100, wn_rtol=1e-4, stream=None):
    """The retries the reference makes when its first root search fails (fsolve_vary, speedup/solvers.py:12-55; backup
    scan of the hybrid solver, fluid.py:731-795) for rows idx_host of pg (ptt_generic_rescue).  Returns a compact
    ShellBatch (entry f = point idx_host[f]) and the int32 tensor `rescued`.  With `stream` the kernel is queued there
    (after everything already queued on the current stream) and the caller synchronises before reading."""
    torch = _torch()
    lib = _lib.load()
    idx_host = np.asarray(idx_host, dtype=np.int32)
    n_f = idx_host.size
    cap = lib.ptt_generic_row_capacity(n_xi)
    dev = pg.device
    idx = torch.as_tensor(idx_host, device=dev)
    rows = [torch.full((n_f, cap), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    off = torch.zeros(n_f, dtype=torch.int32, device=dev)
    npts = torch.ones(n_f, dtype=torch.int32, device=dev)
    status = torch.full((n_f,), _lib.ST_SOLVER_FAILED, dtype=torch.int32, device=dev)
    scalars = torch.full((n_f, 16), float("nan"), dtype=torch.float64, device=dev)
    rescued = torch.zeros(n_f, dtype=torch.int32, device=dev)

    def launch():
        _lib.check(lib.ptt_generic_rescue(n_f, _ptr(idx), _ptr(pg), n_xi, float(t_end), int(thin_shell_limit),
                                          float(wn_rtol), _ptr(rows[0]), _ptr(rows[1]), _ptr(rows[2]), cap, _ptr(off),
                                          _ptr(npts), _ptr(scalars), _ptr(status), _ptr(rescued), _stream_ptr(torch)),
                   "ptt_generic_rescue")

    if stream is None:
        with STAGES.stage("shells_retry", 1):
            launch()
    else:
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            launch()
        STAGES.launches += 1
    idx_l = idx.long()
    sub = ShellBatch(pg[idx_l, 0].contiguous(), pg[idx_l, 1].contiguous(), None, pg[idx_l, 2].to(torch.int32), status,
                     rows[0], rows[1], rows[2], off, npts, scalars, n_xi)
    return sub, rescued


class SweepOutput:
    """Result tables of ptt_sweep_spectra: torch tensors (mem="device") or pinned numpy arrays (mem="host")."""

    def __init__(self):
        self.pow_v = self.pow_gw = self.omgw0 = self.scalars = self.status = None
        self.rows = self.off = self.npts = self.snr = None
        self.stats = None
        self._keep = []

    def scal(self, name):
        return self.scalars[:, _lib.SWEEP_SCALARS.index(name)]


_LIVE_CALLBACKS = []


def sweep_c(plan, pg, pm, scale=None, *, n_xi=N_XI_DEFAULT, t_end=50.0, thin_shell_limit=100, wn_rtol=1e-4, r_star=None,
            lifetime_multiplier=1.0, chunk=8000, shell_chunk=0, want=("pow_v", "pow_gw", "omgw0"), scalars=True,
            rows=False, mem="device", retry=True, timing=False, on_chunk=None, out=None, snr_weights=None,
            table_rows=None, out_rows=None, host_cols=None) -> SweepOutput:
    """The whole sweep in one C call (ptt_sweep_spectra, csrc/ptt_sweep.cu): shells, the reference's retries,
    thermodynamic scalars, A^2, spec_den_v, spec_den_gw, pow_gw, omgw0.  pg[n,16], pm[n,4] as in the C header (host
    numpy arrays or device tensors according to `mem`); plan: SsmPlan / CPlan or None (shells only).
    on_chunk(row0, row1, cuda_stream_ptr) is called as chunks are queued for delivery; out: dict of preallocated
    destination tables (same kind as `mem`) to write into instead of fresh ones.  host_cols = (v_wall, sol_type codes) on
    the host, in the order of pg (device mode): saves the driver the device->host copy of pg and the wait for the stream
    that goes with it, so that this sweep is queued while the previous one still runs (ptt_sweep_opts_t.host_v_wall)."""
    torch = _torch()
    lib = _lib.load()
    host = mem == "host"
    if mem not in ("host", "device"):
        raise ValueError(mem)
    res = SweepOutput()
    if host:
        pg_a = np.ascontiguousarray(pg, dtype=np.float64)
        pm_a = np.ascontiguousarray(pm, dtype=np.float64)
        sc_a = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        n = pg_a.shape[0]
        ptr = lambda a: None if a is None else C.c_void_p(a.ctypes.data)

        def alloc(shape, dtype=torch.float64):
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            res._keep.append(t)
            return t.numpy()
    else:
        pg_a, pm_a = dev_f64(pg), dev_f64(pm)
        sc_a = None if scale is None else dev_f64(scale).reshape(-1)
        n = pg_a.shape[0]
        ptr = lambda a: None if a is None else C.c_void_p(a.data_ptr())

        def alloc(shape, dtype=torch.float64):
            return torch.empty(shape, dtype=dtype, device="cuda")
    view = None if plan is None else plan.sweep_view()
    want = tuple(want) if plan is not None else ()
    if "pow_v" in want and not (view.n_pass == 2 and view.has_y_pass):
        want = tuple(w for w in want if w != "pow_v")
    if "omgw0" in want and sc_a is None:
        want = tuple(w for w in want if w != "omgw0")
    n_y = int(view.gw.n_y) if view is not None else 0
    out = out or {}
    o = _lib.SweepOut()
    for name in ("pow_v", "pow_gw", "omgw0"):
        if name in want:
            t = out.get(name)
            if t is None:
                t = alloc((n if table_rows is None else int(out_rows or n), n_y))
            setattr(res, name, t)
            setattr(o, name, ptr(t))
    res.status = out.get("status")
    if res.status is None:
        res.status = alloc((n,), torch.int32)
    o.status = ptr(res.status)
    if scalars:
        res.scalars = out.get("scalars")
        if res.scalars is None:
            res.scalars = alloc((n, _lib.SWEEP_NSCAL))
        o.scalars = ptr(res.scalars)
    if table_rows is not None:
        if host:
            raise ValueError("table_rows needs device tables")
        tr = np.ascontiguousarray(table_rows, dtype=np.int64)
        if tr.shape != (n,):
            raise ValueError("table_rows: one destination row per point")
        res._keep.append(tr)
        o.table_rows = C.c_void_p(tr.ctypes.data)
    if snr_weights is not None and "omgw0" in want:
        w = dev_f64(np.ascontiguousarray(snr_weights, dtype=np.float64))
        res._keep.append(w)
        res.snr = alloc((n, w.shape[0]))
        o.snr, o.snr_weight, o.n_snr = ptr(res.snr), C.c_void_p(w.data_ptr()), int(w.shape[0])
    if rows:
        if host:
            raise ValueError("profile rows are returned in device memory only")
        cap = lib.ptt_generic_row_capacity(n_xi)
        res.rows = [torch.empty((n, cap), dtype=torch.float64, device="cuda") for _ in range(3)]
        res.off = torch.empty(n, dtype=torch.int32, device="cuda")
        res.npts = torch.empty(n, dtype=torch.int32, device="cuda")
        o.rows_v, o.rows_w, o.rows_xi = (ptr(r) for r in res.rows)
        o.row_stride = cap
        o.off, o.npts = ptr(res.off), ptr(res.npts)
    opts = _lib.SweepOpts(n_xi=int(n_xi), thin_shell_limit=int(thin_shell_limit), chunk=int(chunk),
                          shell_chunk=int(shell_chunk or 0),
                          flags=(_lib.SWEEP_TIMING if timing else 0) | (0 if retry else _lib.SWEEP_NO_RETRY),
                          group_min=int(GROUP_MIN) if USE_GROUPED_SDV else 1 << 30,
                          t_end=float(t_end), wn_rtol=float(wn_rtol), r_star=-1.0 if r_star is None else float(r_star),
                          lifetime_multiplier=float(lifetime_multiplier))
    if host_cols is not None and not host:
        hv = np.ascontiguousarray(host_cols[0], dtype=np.float64).reshape(-1)
        hc = np.ascontiguousarray(host_cols[1], dtype=np.int32).reshape(-1)
        if hv.size != n or hc.size != n:
            raise ValueError("host_cols: one v_wall and one solution type per point")
        res._keep += [hv, hc]
        opts.host_v_wall, opts.host_sol_type = C.c_void_p(hv.ctypes.data), C.c_void_p(hc.ctypes.data)
    cb = None
    if on_chunk is not None:
        cb = _lib.CHUNK_CB(lambda user, r0, r1, stream: on_chunk(int(r0), int(r1), stream))
    finish_timing(keep_last=True)
    rc = lib.ptt_sweep_spectra(C.byref(view) if view is not None else None, C.byref(opts), n, ptr(pg_a), ptr(pm_a),
                               ptr(sc_a), C.byref(o), C.cast(cb, C.c_void_p) if cb is not None else None, None,
                               _lib.MEM_HOST if host else _lib.MEM_DEVICE, _stream_ptr(torch))
    _lib.check(rc, "ptt_sweep_spectra")
    st = _lib.SweepStats()
    lib.ptt_sweep_last_stats(C.byref(st))
    res.stats = st
    STAGES.launches += int(st.launches)
    if timing:
        if host:
            _add_stage_times(st)
        else:
            # device tables: the call has returned with its last chunks in flight; the stage times are read from the
            # events when somebody asks for them (finish_timing) or when the sweep after the next one starts
            res.stats_box = _StatsBox(res)
            _PENDING_TIMING.append(res.stats_box)
    res._keep += [pg_a, pm_a, sc_a]
    return res


class _StatsBox:
    """Where the completed stats of a timed device-table sweep arrive (does not keep the result tables alive)."""

    def __init__(self, res):
        import weakref
        self.stats, self._res = res.stats, weakref.ref(res)


_PENDING_TIMING = []


def _add_stage_times(st):
    for i, name in enumerate(_lib.STAGE_NAMES):
        if st.stage_calls[i]:
            d = STAGES.c_stage_ms.setdefault(name, {"ms": 0.0, "n": 0})
            d["ms"] += st.stage_ms[i]
            d["n"] += int(st.stage_calls[i])
    for k in ("group_macs", "group_profiles", "group_count", "a2_profiles", "gw_profiles"):
        STAGES.c_counts[k] = STAGES.c_counts.get(k, 0) + int(getattr(st, k))


def finish_timing(keep_last=False):
    """Completes the stage times of the timed device-table sweeps queued so far (waits for them to finish).
    keep_last: all but the most recent one - what a new sweep does before it is queued behind that one."""
    while len(_PENDING_TIMING) > (1 if keep_last else 0):
        box = _PENDING_TIMING.pop(0)
        lib = _lib.load()
        st = _lib.SweepStats()
        _lib.check(lib.ptt_sweep_timing_collect(int(box.stats.timing_serial), C.byref(st)), "ptt_sweep_timing_collect")
        box.stats = st
        res = box._res()
        if res is not None:
            res.stats = st
        if STAGES.enabled:
            _add_stage_times(st)


def shell_wall_values(shells: ShellBatch):
    """props.v_and_w_from_solution for a batch of bag shells -> [n,8] = vp, vm, vp_tilde, vm_tilde, wp, wm, wn, 0."""
    torch = _torch()
    lib = _lib.load()
    n = len(shells)
    out = torch.empty((n, 8), dtype=torch.float64, device=shells.rows_v.device)
    with STAGES.stage("wall_values", 1):
        _lib.check(lib.ptt_shell_wall_values(n, _ptr(shells.rows_v), _ptr(shells.rows_w), _ptr(shells.rows_xi),
                                             shells.rows_v.shape[1], _ptr(shells.off), _ptr(shells.npts),
                                             _ptr(shells.v_wall), _ptr(shells.sol_type), _ptr(out), _stream_ptr(torch)),
                   "ptt_shell_wall_values")
    return out


# ---------------------------------------------------------------------------------------------------------------
# Spectrum plans (host: numpy, identical expressions to the reference; device: uploaded once)
# ---------------------------------------------------------------------------------------------------------------

def blend_fraction(z: np.ndarray, z_st_thresh: float = Z_ST_THRESH) -> np.ndarray:
    """Per-wavenumber weight of the asymptotic formula in sin_transform (calculators.py:119-158):
    0 = exact trapezoid only, 1 = asymptotic only, in between = cubic blend on (thresh - pi, thresh]."""
    z = np.asarray(z, dtype=np.float64)
    frac = np.zeros_like(z)
    if z.size <= 1:
        # scalar branch (calculators.py:104-116)
        frac[z > z_st_thresh] = 1.0
        return frac
    frac[z > z_st_thresh] = 1.0
    blend = (z > z_st_thresh - DZ_ST_BLEND) & (z <= z_st_thresh)
    if np.any(blend):
        zb = z[blend]
        zmax, zmin = zb.max(), zb.min()
        s = (zb - zmin) / (zmax - zmin) if zmax > zmin else 0.5 * np.ones_like(zb)
        frac[blend] = 3 * s ** 2 - 2 * s ** 3
    return frac


FORCE_DIRECT_ST = False     # validation switch: evaluate the exact sine transform with one sine per (z, xi) pair


def z_exact_max(z: np.ndarray, frac: np.ndarray) -> float:
    sel = frac < 1.0
    return float(np.max(z[sel])) if np.any(sel) else 0.0


def qt_lookup_grid(z: np.ndarray):
    """qT_lookup of spec_den_v (spectrum.py:503-515) and its log-grid parameters (l0, dlog10)."""
    lmin, lmax = np.log10(np.min(z)), np.log10(np.max(z))
    dl = (lmax - lmin) / z.size
    l0 = lmin + np.log10(T_TILDE_MIN)
    return 10 ** np.arange(l0, lmax + np.log10(T_TILDE_MAX), dl), l0, dl


def nu(t, nuc_type="exponential", a=1.0):
    """Bubble lifetime distribution (spectrum.py:189-205)."""
    if nuc_type == "simultaneous":
        return 0.5 * a * (a * t) ** 2 * np.exp(-(a * t) ** 3 / 6)
    if nuc_type == "exponential":
        return a * np.exp(-a * t)
    raise ValueError("Nucleation type not recognized")


def _trapz_weights(x: np.ndarray) -> np.ndarray:
    w = np.empty_like(x)
    w[1:-1] = 0.5 * (x[2:] - x[:-2])
    w[0] = 0.5 * (x[1] - x[0])
    w[-1] = 0.5 * (x[-1] - x[-2])
    return w


class SdvPass:
    """One spec_den_v evaluation: A^2 tabulated on qT_lookup(z), convolved with the lifetime distribution at z."""

    def __init__(self, z: np.ndarray, nt: int, nuc_type: str, a: float):
        z = np.asarray(z, dtype=np.float64)
        if np.any(np.diff(z) < 0):
            raise ValueError("z must be ascending")
        self.z = z
        self.qT, self.l0, self.dl = qt_lookup_grid(z)
        self.nt = nt
        lt = np.linspace(np.log10(T_TILDE_MIN), np.log10(T_TILDE_MAX), nt)
        self.T = 10.0 ** lt
        self.LT = lt / self.dl
        self.W = _trapz_weights(self.T) * self.T ** 6 * nu(self.T, nuc_type, a)
        self.zterm = (np.log10(z) - self.l0) / self.dl
        # tile size: the lookup window of a tile must fit in shared memory
        cap_entries = SMEM_WINDOW_BYTES // 16
        t_span = self.LT[-1] - self.LT[0]
        tile = 512
        best = None
        while True:
            spans = self.zterm[np.minimum(np.arange(0, z.size, tile) + tile - 1, z.size - 1)] - self.zterm[::tile]
            need = int(np.ceil(spans.max() + t_span)) + 6
            if need <= cap_entries:
                best = (tile, need)
                if tile >= z.size:
                    break
                tile += 512
            else:
                break
        if best is None:
            raise ValueError("spec_den_v lookup window does not fit in shared memory; reduce nt range or z density")
        self.z_tile, self.win_entries = best
        self.smem_bytes = self.win_entries * 16
        self.omega_ab = _omega_pieces(self.LT, self.T, self.W, self.dl)


def _omega_pieces(LT, T, W, dl):
    """(alpha, beta) of the knot weight of K6' on every piece between consecutive kinks of {LT_j - 1, LT_j, LT_j + 1}.
    With x = e - (zterm_i + shift) the samples whose lookup lands in the segment right of knot e are those with
    x <= LT_j < x + 1 (weight 1 - t), left of it x - 1 <= LT_j < x (weight t), t = (X / x0 - 1) / D on the log grid
    (D = 10^dl - 1), so that  omega = alpha + beta * X_i / qT[e]  with
        alpha = (1 + 1/D) sum_A W_j - (1/D) sum_B W_j,   beta = -(1/D) sum_A W_j T_j + (10^dl / D) sum_B W_j T_j."""
    nt = LT.size
    D = 10.0 ** dl - 1.0
    kinks = np.concatenate([LT - 1.0, LT, LT + 1.0])
    label = np.repeat([0, 1, 2], nt)                       # 0: LT - 1, 1: LT, 2: LT + 1
    order = np.argsort(kinks, kind="stable")
    lab = label[order]
    cnt = np.zeros((3, 3 * nt + 1), dtype=np.int64)        # kinks of each label among the first p sorted kinks
    for k in range(3):
        cnt[k, 1:] = np.cumsum(lab == k)
    n_m1, n_0, n_p1 = cnt[0], cnt[1], cnt[2]               # A = [n_0, n_m1), B = [n_p1, n_0)
    WT = W * T
    sums = np.zeros((4, 3 * nt + 1))
    for lo, hi, iw, iwt in ((n_0, n_m1, 0, 1), (n_p1, n_0, 2, 3)):
        for k in range(int((hi - lo).max())):
            idx = lo + k
            ok = idx < hi
            idx = np.minimum(idx, nt - 1)
            sums[iw] += np.where(ok, W[idx], 0.0)
            sums[iwt] += np.where(ok, WT[idx], 0.0)
    alpha = (1.0 + 1.0 / D) * sums[0] - sums[2] / D
    beta = -sums[1] / D + (10.0 ** dl / D) * sums[3]
    return np.ascontiguousarray(np.stack([alpha, beta], axis=1))


class SsmPlan:
    """Grids for a batch of SSM spectra that share (y, cs, sizes).

    mode="ssm": the SSMSpectrum.compute chain (spectrum.py:96-115): pass A on z=y (gives pow_v and a2), z_lookup =
    gen_lookup(y, cs, n_z_lookup, eps=1e-8), pass B on z_lookup, GW integral on (z_lookup, y).
    mode="bag": power_gw_scaled_bag (spectrum.py:243-294): x = np.logspace(log10 xmin, log10 xmax, npt[2]),
    one pass on x, GW integral on (x, z).
    """

    def __init__(self, y, cs=CS0, nt=NT_DEFAULT, n_z_lookup=NZ_LOOKUP_DEFAULT, nxi=NXI_DEFAULT,
                 z_st_thresh=Z_ST_THRESH, nuc_type="exponential", a=1.0, gamma=GAMMA_DEFAULT, mode="ssm",
                 phase_rule=None, nz_int=None, only_first_pass=False, only_lookup_pass=False):
        torch = _torch()
        y = np.asarray(y, dtype=np.float64)
        if np.any(y <= 0.0):
            raise ValueError("z values must all be positive.")
        self.y, self.cs, self.nt, self.nxi, self.mode = y, float(cs), int(nt), int(nxi), mode
        self.z_st_thresh, self.nuc_type, self.a, self.gamma = z_st_thresh, str(nuc_type), float(a), gamma
        eps = 1e-8
        lo = y.min() * 0.5 * (1.0 - cs) / cs - eps      # lookup_limits, spectrum.py:183-186
        hi = y.max() * 0.5 * (1.0 + cs) / cs + eps
        self.z_lookup = 10.0 ** np.linspace(np.log10(lo), np.log10(hi), n_z_lookup)
        self.passes: tp.List[SdvPass] = []
        self.has_y_pass = (mode == "ssm") and not only_lookup_pass
        if mode == "ssm":
            if self.has_y_pass:
                self.passes.append(SdvPass(y, nt, self.nuc_type, self.a))
            self.phase_rule = 1 if phase_rule is None else phase_rule
        elif mode == "bag":
            self.phase_rule = 0 if phase_rule is None else phase_rule
        else:
            raise ValueError(mode)
        if not only_first_pass:
            self.passes.append(SdvPass(self.z_lookup, nt, self.nuc_type, self.a))
        self.lookup_pass = self.passes[-1]

        # A^2 wavenumbers: all lookup grids concatenated
        self.seg = np.cumsum([0] + [p.qT.size for p in self.passes]).astype(np.int32)
        self.z_all = np.concatenate([p.qT for p in self.passes])
        self.frac_all = np.concatenate([blend_fraction(p.qT, z_st_thresh) for p in self.passes])
        self.xi_re = np.linspace(0, 1 - 1 / nxi, nxi)     # calculators.py:100

        self.gw = None if only_first_pass else build_gw_tables(self.z_lookup, y, cs, gamma, nz_int)

        # upload
        self._t = {}
        up = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
        self._t["seg"] = up(self.seg, torch.int32)
        for k in ("z_all", "frac_all", "xi_re", "z_lookup", "y"):
            self._t[k] = up(getattr(self, k))
        for i, p in enumerate(self.passes):
            for k in ("qT", "z", "zterm", "T", "LT", "W"):
                self._t[f"p{i}_{k}"] = up(getattr(p, k))
            self._t[f"p{i}_omega_ab"] = up(p.omega_ab)
        if self.gw is not None:
            for k in ("L1", "L2", "Z1", "Z2", "G", "yterm"):
                self._t["gw_" + k] = up(self.gw[k])

        self.a2_plan = A2Plan(n_z=self.z_all.size, n_seg=len(self.passes), seg=self._t["seg"].data_ptr(),
                              z=self._t["z_all"].data_ptr(), frac=self._t["frac_all"].data_ptr(), nxi=self.nxi,
                              phase_rule=self.phase_rule, xi_re=self._t["xi_re"].data_ptr(),
                              z_exact_max=z_exact_max(self.z_all, self.frac_all), force_direct=int(FORCE_DIRECT_ST))
        self.sdv_plans = []
        for i, p in enumerate(self.passes):
            self.sdv_plans.append(SdvPlan(
                n_q=p.qT.size, n_z=p.z.size, nt=p.nt, z_tile=p.z_tile, qT=self._t[f"p{i}_qT"].data_ptr(),
                z=self._t[f"p{i}_z"].data_ptr(), zterm=self._t[f"p{i}_zterm"].data_ptr(),
                T=self._t[f"p{i}_T"].data_ptr(), LT=self._t[f"p{i}_LT"].data_ptr(), W=self._t[f"p{i}_W"].data_ptr(),
                inv_dlog=1.0 / p.dl, smem_bytes=p.smem_bytes,
                omega_ab=self._t[f"p{i}_omega_ab"].data_ptr() if USE_OMEGA_PIECES else None))
        self.gw_plan = None if self.gw is None else make_gw_plan(self.gw, self._t, self.z_lookup.size, y.size)

    @property
    def n_a2(self):
        return int(self.z_all.size)

    def sweep_view(self) -> "_lib.SweepPlan":
        """The POD view of this plan that ptt_sweep_spectra takes (include/pttools_b200.h, ptt_sweep_plan_t)."""
        if self.gw is None:
            raise ValueError("a sweep needs a plan with the z_lookup pass and the GW tables")
        v = getattr(self, "_sweep_view", None)
        if v is not None:
            return v
        v = _lib.SweepPlan()
        v.a2 = self.a2_plan
        v.n_pass = len(self.passes)
        v.has_y_pass = int(self.has_y_pass)
        self._h_zterm = [np.ascontiguousarray(p.zterm, dtype=np.float64) for p in self.passes]
        for i, p in enumerate(self.passes):
            v.sdv[i] = self.sdv_plans[i]
            v.h_zterm[i] = self._h_zterm[i].ctypes.data
            v.lt_first[i], v.lt_last[i], v.dl[i] = float(p.LT[0]), float(p.LT[-1]), float(p.dl)
        for i, sg in enumerate(self.seg):
            v.seg[i] = int(sg)
        v.gw = self.gw_plan
        tabs = gw_cells(self) if USE_GW_CELLS else False
        if tabs:
            v.cells = tabs[0]
        self._sweep_view = v
        return v


class CPlan:
    """A plan built by the library itself (ptt_ssm_plan_create, csrc/ptt_plan.cu): what a C caller of ptt_sweep_spectra
    uses.  Same grids as SsmPlan to rounding (tests/test_plan_cabi.py)."""

    def __init__(self, y, cs=CS0, nt=NT_DEFAULT, n_z_lookup=NZ_LOOKUP_DEFAULT, nxi=NXI_DEFAULT, z_st_thresh=Z_ST_THRESH,
                 nuc_type="exponential", a=1.0, gamma=GAMMA_DEFAULT, mode="ssm", phase_rule=-1, nz_int=0, flags=0,
                 upload=True):
        lib = _lib.load()
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        if nuc_type not in ("exponential", "simultaneous"):
            raise ValueError("Nucleation type not recognized")
        d = _lib.PlanDesc(n_y=self.y.size, nt=int(nt), n_z_lookup=int(n_z_lookup), nxi=int(nxi),
                          nuc_type=int(nuc_type == "simultaneous"), mode={"ssm": 0, "bag": 1}[mode], phase_rule=int(phase_rule),
                          nz_int=int(nz_int or 0), flags=int(flags), cs=float(cs), z_st_thresh=float(z_st_thresh),
                          nuc_a=float(a), gamma=float(gamma))
        self._h = C.c_void_p()
        if upload:
            _torch()
            _lib.check(lib.ptt_ssm_plan_create(C.byref(d), self.y.ctypes.data, C.byref(self._h)), "ptt_ssm_plan_create")
        else:
            _lib.check(lib.ptt_ssm_plan_create_host(C.byref(d), self.y.ctypes.data, C.byref(self._h)),
                       "ptt_ssm_plan_create_host")
        self.uploaded = upload

    def array(self, name: str) -> np.ndarray:
        lib = _lib.load()
        n = lib.ptt_ssm_plan_array(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        lib.ptt_ssm_plan_array(self._h, name.encode(), out.ctypes.data, n)
        return out

    def info(self):
        lib = _lib.load()
        n_pass = C.c_int32()
        seg, zt, win, sc = (C.c_int32 * 4)(), (C.c_int32 * 2)(), (C.c_int32 * 2)(), (C.c_double * 8)()
        _lib.check(lib.ptt_ssm_plan_info(self._h, C.byref(n_pass), seg, zt, win, sc), "ptt_ssm_plan_info")
        return {"n_pass": n_pass.value, "seg": list(seg)[: n_pass.value + 1], "z_tile": list(zt)[: n_pass.value],
                "win_entries": list(win)[: n_pass.value], "l0_dl": list(sc)[:4], "z_exact_max": sc[4],
                "gw_prefactor": sc[5], "gw_lrange": sc[6], "gw_win_entries": int(sc[7])}

    def sweep_view(self) -> "_lib.SweepPlan":
        ptr = _lib.load().ptt_ssm_plan_view(self._h)
        if not ptr:
            raise _lib.PttError("the plan has not been uploaded")
        return ptr.contents

    @property
    def has_y_pass(self):
        return bool(self.sweep_view().has_y_pass)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().ptt_ssm_plan_destroy(h)
            except Exception:       # interpreter shutdown
                pass


def build_gw_tables(z_lookup, y, cs, gamma=GAMMA_DEFAULT, nz_int=None):
    """Tables of the GW integral (spectrum.py:327-368) with the substitution z = y * zeta; z_lookup is log-uniform."""
    n_zl = z_lookup.size
    n_int = n_zl if nz_int is None else int(nz_int)
    zpf, zmf = (1 + cs) / (2 * cs), (1 - cs) / (2 * cs)
    lz = np.linspace(np.log10(zmf), np.log10(zpf), n_int)
    zeta = 10.0 ** lz
    zeta2 = zpf + zmf - zeta
    lz0 = np.log10(z_lookup[0])
    dlz = (np.log10(z_lookup[-1]) - lz0) / (n_zl - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        kern = ((zeta - zpf) ** 2 * (zeta - zmf) ** 2) / (zeta * zeta2)
    gw = dict(n_int=n_int, L1=lz / dlz, L2=np.log10(zeta2) / dlz, Z1=zeta, Z2=zeta2, G=_trapz_weights(zeta) * kern,
              yterm=(np.log10(y) - lz0) / dlz,
              prefactor=0.75 * gamma ** 2 * ((1 - cs ** 2) / cs ** 2) ** 2 / (4 * np.pi * cs))
    y_tile = 32
    span = (gw["yterm"][np.minimum(np.arange(0, y.size, y_tile) + y_tile - 1, y.size - 1)] - gw["yterm"][::y_tile]).max()
    lrange = max(gw["L1"].max(), gw["L2"].max()) - min(gw["L1"].min(), gw["L2"].min())
    gw["y_tile"] = y_tile
    gw["win_entries"] = int(np.ceil(abs(span) + lrange)) + 6
    if gw["win_entries"] * 16 > SMEM_WINDOW_BYTES:
        raise ValueError("spec_den_gw lookup window does not fit in shared memory")
    return gw


def make_gw_plan(g, t, n_zl, n_y):
    return GwPlan(n_zl=n_zl, n_y=n_y, n_int=g["n_int"], y_tile=g["y_tile"], z_lookup=t["z_lookup"].data_ptr(),
                  y=t["y"].data_ptr(), yterm=t["gw_yterm"].data_ptr(), L1=t["gw_L1"].data_ptr(), L2=t["gw_L2"].data_ptr(),
                  Z1=t["gw_Z1"].data_ptr(), Z2=t["gw_Z2"].data_ptr(), G=t["gw_G"].data_ptr(), prefactor=g["prefactor"],
                  smem_bytes=g["win_entries"] * 16)


class GwOnlyPlan:
    """Plan for spec_den_gw_scaled on a user-supplied z_lookup (spectrum.py:402-430).  Log-uniform z_lookup with ascending
    y (what gen_lookup / np.logspace give, and what the reference itself always passes) runs on the arithmetic-index
    kernels; any other ascending z_lookup, or unordered y, on the binary-search kernel (`general`)."""

    def __init__(self, z_lookup, y, cs=CS0, gamma=GAMMA_DEFAULT, nz_int=None):
        torch = _torch()
        z_lookup = np.asarray(z_lookup, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if z_lookup.size < 2 or np.any(np.diff(z_lookup) <= 0):
            raise ValueError("z_lookup must be strictly ascending (np.interp needs it)")
        lz = np.log10(z_lookup)
        uniform = np.all(z_lookup > 0) and np.allclose(lz, np.linspace(lz[0], lz[-1], z_lookup.size), rtol=0, atol=1e-9)
        self.general = not (uniform and np.all(np.diff(y) >= 0))
        self.y, self.z_lookup = y, z_lookup
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")
        self._t = {"z_lookup": up(z_lookup), "y": up(y)}
        if self.general:
            # the same zeta tables; the log-index tables are not used by the search kernel
            grid = np.logspace(0, 12, 16)
            self.gw = build_gw_tables(grid, np.array([1.0, 1.0]), cs, gamma, z_lookup.size if nz_int is None else nz_int)
        else:
            self.gw = build_gw_tables(z_lookup, y, cs, gamma, nz_int)
        for k in ("L1", "L2", "Z1", "Z2", "G", "yterm"):
            self._t["gw_" + k] = up(self.gw[k])
        self.gw_plan = None if self.general else make_gw_plan(self.gw, self._t, z_lookup.size, y.size)


class A2OnlyPlan:
    """Plan for a2_e_conserving on one arbitrary ascending z grid (ssm.py:35-140)."""

    def __init__(self, z, nxi=NXI_DEFAULT, z_st_thresh=Z_ST_THRESH, phase_rule=1, force_direct=False):
        torch = _torch()
        z = np.asarray(z, dtype=np.float64)
        self.nxi = int(nxi)
        self.z_all = z
        self.seg = np.array([0, z.size], dtype=np.int32)
        self.xi_re = np.linspace(0, 1 - 1 / nxi, nxi)
        up = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device="cuda")
        self._t = {"seg": up(self.seg, torch.int32), "z": up(z), "frac": up(blend_fraction(z, z_st_thresh)),
                   "xi_re": up(self.xi_re)}
        frac = blend_fraction(z, z_st_thresh)
        self.a2_plan = A2Plan(n_z=z.size, n_seg=1, seg=self._t["seg"].data_ptr(), z=self._t["z"].data_ptr(),
                              frac=self._t["frac"].data_ptr(), nxi=self.nxi, phase_rule=phase_rule,
                              xi_re=self._t["xi_re"].data_ptr(), z_exact_max=z_exact_max(z, frac),
                              force_direct=int(force_direct or FORCE_DIRECT_ST))

    @property
    def n_a2(self):
        return int(self.z_all.size)


@dataclasses.dataclass
class SpectrumBatch:
    y: np.ndarray
    pow_gw: tp.Any
    sd_gw: tp.Any
    omgw0: tp.Any = None
    pow_v: tp.Any = None
    sd_v: tp.Any = None
    a2: tp.Any = None          # A^2 on the first pass' qT_lookup (SSMSpectrum.a2)
    sd_v_lookup: tp.Any = None


def point_params(v_wall, cs, e_coef_s, e_coef_b, V_s, V_b):
    """pp[P,8] rows of ptt_a2_batch."""
    torch = _torch()
    v_wall = dev_f64(v_wall).reshape(-1)
    n = v_wall.numel()
    pp = torch.zeros((n, 8), dtype=torch.float64, device=v_wall.device)
    pp[:, 0] = v_wall
    for j, val in enumerate((cs, e_coef_s, e_coef_b, V_s, V_b), start=1):
        pp[:, j] = dev_f64(val) if not np.isscalar(val) else float(val)
    return pp


def a2_batch(plan: SsmPlan, shells: ShellBatch, pp, want_parts=False):
    """A^2 on all lookup grids of the plan for a batch of shells (ssm.py:35-140). Returns [P, plan.n_a2]."""
    torch = _torch()
    lib = _lib.load()
    P = len(shells)
    cap = shells.rows_v.shape[1]
    ws = lib.ptt_a2_work_stride(cap, plan.nxi, plan.n_a2)
    dev = shells.rows_v.device
    work = torch.empty((P, ws), dtype=torch.float64, device=dev)
    a2 = torch.empty((P, plan.n_a2), dtype=torch.float64, device=dev)
    fp = torch.empty_like(a2) if want_parts else None
    l2 = torch.empty_like(a2) if want_parts else None
    with STAGES.stage("a2", 3):
        _lib.check(lib.ptt_a2_batch(C.byref(plan.a2_plan), P, _ptr(shells.rows_v), _ptr(shells.rows_w),
                                    _ptr(shells.rows_xi), cap, _ptr(shells.off), _ptr(shells.npts), _ptr(pp), _ptr(work),
                                    ws, _ptr(a2), _ptr(fp), _ptr(l2), _stream_ptr(torch)), "ptt_a2_batch")
    if want_parts:
        return a2, fp, l2
    return a2


def spec_den_v_batch(plan: SsmPlan, ipass: int, a2, v_wall):
    """_spec_den_v_core for pass `ipass` (spectrum.py:451-486); a2 is the full [P, n_a2] table."""
    torch = _torch()
    lib = _lib.load()
    P = a2.shape[0]
    sp = plan.sdv_plans[ipass]
    out = torch.empty((P, sp.n_z), dtype=torch.float64, device=a2.device)
    a2_view_ptr = C.c_void_p(a2.data_ptr() + 8 * int(plan.seg[ipass]))
    with STAGES.stage("spec_den_v" if ipass == len(plan.passes) - 1 else "spec_den_v_y", 1):
        _lib.check(lib.ptt_spec_den_v_batch(C.byref(sp), P, a2_view_ptr, a2.stride(0), _ptr(v_wall), _ptr(out), sp.n_z,
                                            _stream_ptr(torch)), "ptt_spec_den_v_batch")
    return out


USE_OMEGA_PIECES = True      # K6' weights from the per-piece closed form (False: sample loop for every entry)
GROUP_MIN = 24          # smallest run of equal v_wall worth the banded matrix product (break-even ~15 profiles)
USE_GROUPED_SDV = True


def _group_band(p: SdvPass, v_wall: float, parity=None):
    """First knot of every 128-row tile and the common band width for one wall speed (host side of K6').  parity (0 or
    1): every tile's first knot gets this parity (one knot earlier where needed; -1 = a knot before the table, weight
    0), so that its A^2 entries start on a 16-byte boundary (PTT_SDV_ALIGNED_BAND)."""
    b_r = (8.0 * np.pi) ** (1.0 / 3.0)
    shift = -np.log10(b_r * v_wall) / p.dl
    n_z, n_q = p.z.size, p.qT.size
    i0 = np.arange(0, n_z, 128)
    i1 = np.minimum(i0 + 128, n_z) - 1
    e_first = np.floor(p.zterm[i0] + p.LT[0] + shift).astype(np.int64) - 1
    e_base = np.clip(e_first, 0, max(n_q - 2, 0))
    if parity is not None:
        e_base = e_base - ((e_base + int(parity)) & 1)
    e_last = np.clip(np.floor(p.zterm[i1] + p.LT[-1] + shift).astype(np.int64) + 1, 0, n_q - 2) + 1
    kw = int((e_last - e_base).max()) + 3
    kw = ((kw + 15) // 16) * 16
    return e_base.astype(np.int32), kw


def spec_den_v_group(plan: SsmPlan, ipass: int, a2, v_wall: float, out, slot: int = 0, timed: bool = True):
    """_spec_den_v_core for profiles that share v_wall, as a banded FP64 GEMM (ptt_spec_den_v_group).
    a2: [n_p, n_a2] rows of the group; out: [n_p, n_z] destination view; slot: which Omega scratch buffer to use
    (one per concurrently running group)."""
    torch = _torch()
    lib = _lib.load()
    sp, p = plan.sdv_plans[ipass], plan.passes[ipass]
    # rows of A^2 on 16-byte boundaries: the band of every tile can be aligned too, and the operands go by TMA
    a2_addr = a2.data_ptr() + 8 * int(plan.seg[ipass])
    aligned = a2.stride(0) % 2 == 0 and a2.data_ptr() % 16 == 0
    e_base, kw = _group_band(p, v_wall, parity=(a2_addr >> 3) & 1 if aligned else None)
    n_tiles = e_base.size
    key = ("omega", ipass, slot)
    need = n_tiles * 128 * kw + 2 * p.qT.size
    buf = plan._t.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty(int(need * 1.02) + 1024, dtype=torch.float64, device=a2.device)
        plan._t[key] = buf
    eb = torch.as_tensor(e_base, device=a2.device)
    a2_ptr = C.c_void_p(a2_addr)
    call = lambda: _lib.check(lib.ptt_spec_den_v_group_ex(C.byref(sp), a2.shape[0], a2_ptr, a2.stride(0), float(v_wall), kw,
                                                          _ptr(eb), _ptr(buf), _ptr(out), out.stride(0),
                                                          _lib.SDV_ALIGNED_BAND if aligned else 0, _stream_ptr(torch)),
                              "ptt_spec_den_v_group_ex")
    if timed:
        with STAGES.stage(_group_stage_name(plan, ipass), 2):
            call()
    else:
        call()
    return out


def _group_stage_name(plan, ipass):
    return "spec_den_v_group" if ipass == len(plan.passes) - 1 else "spec_den_v_group_y"


SDV_STREAMS = int(__import__("os").environ.get("PTT_SDV_STREAMS", "2"))     # groups of one call run on this many side
                                                                      # streams (fills the last partial wave of each GEMM)
_side_streams = []


def _streams(torch, n):
    while len(_side_streams) < n:
        _side_streams.append(torch.cuda.Stream())
    return _side_streams[:n]


def spec_den_v_auto(plan: SsmPlan, ipass: int, a2, v_wall, v_wall_host=None, valid_host=None):
    """spec_den_v for a batch: runs of >= GROUP_MIN consecutive profiles with the same v_wall go through the banded
    matrix product, everything else through the gather kernel.  v_wall_host: the same wall speeds on the host (saves
    a device read-back); valid_host: bool per profile - trailing profiles of a run that have no shell (their A^2 is
    NaN) are left out of the matrix product and get NaN directly."""
    torch = _torch()
    P = a2.shape[0]
    n_z = plan.sdv_plans[ipass].n_z
    if not USE_GROUPED_SDV or P < GROUP_MIN:
        return spec_den_v_batch(plan, ipass, a2, v_wall)
    vw = v_wall.detach().cpu().numpy() if v_wall_host is None else np.asarray(v_wall_host, dtype=float)
    cuts = np.concatenate(([0], np.nonzero(np.diff(vw) != 0)[0] + 1, [P]))
    runs = [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:])]
    if not any(b - a >= GROUP_MIN for a, b in runs):
        return spec_den_v_batch(plan, ipass, a2, v_wall)
    out = torch.empty((P, n_z), dtype=torch.float64, device=a2.device)
    small_start = None
    big_runs = []
    for a, b in runs + [(P, P)]:
        big = (b - a) >= GROUP_MIN
        if (big or a == P) and small_start is not None:          # flush the pending run of small groups
            out[small_start:a] = spec_den_v_batch(plan, ipass, a2[small_start:a], v_wall[small_start:a].contiguous())
            small_start = None
        if a == P:
            break
        if big:
            b_eff = b
            if valid_host is not None:
                ok = np.nonzero(valid_host[a:b])[0]
                b_eff = a + (int(ok[-1]) + 1 if ok.size else 0)
                if b_eff < b:
                    out[b_eff:b] = float("nan")
            if b_eff > a:
                big_runs.append((a, b_eff))
        elif small_start is None:
            small_start = a
    if len(big_runs) < 2 or SDV_STREAMS < 2:
        for a, b in big_runs:
            spec_den_v_group(plan, ipass, a2[a:b], float(vw[a]), out[a:b])
        return out
    # Independent groups on side streams: the tail wave of one matrix product overlaps the next group's work.
    # One stage entry spans fork -> join on the launching stream.
    main = torch.cuda.current_stream()
    side = _streams(torch, min(SDV_STREAMS, len(big_runs)))
    with STAGES.stage(_group_stage_name(plan, ipass), 2 * len(big_runs)):
        fork = torch.cuda.Event()
        fork.record(main)
        for k, (a, b) in enumerate(big_runs):
            st = side[k % len(side)]
            if k < len(side):
                st.wait_event(fork)
            with torch.cuda.stream(st):
                spec_den_v_group(plan, ipass, a2[a:b], float(vw[a]), out[a:b], slot=k % len(side), timed=False)
        for st in side:
            join = torch.cuda.Event()
            join.record(st)
            main.wait_event(join)
    return out


GW_CELLS_MIN = 128      # smallest batch for which building/using the per-plan cell tables pays
USE_GW_CELLS = True


def gw_cells(plan):
    """Cell tables of the GW integral for this plan (csrc/ptt_gw_cells.cu), built on the device on first use."""
    cached = getattr(plan, "_gw_cells", None)
    if cached is not None:
        return cached
    torch = _torch()
    lib = _lib.load()
    g = plan.gw
    lrange = max(g["L1"].max(), g["L2"].max()) - min(g["L1"].min(), g["L2"].min())
    n_zl, n_y = int(plan.z_lookup.size), int(plan.y.size)
    cap = lib.ptt_gw_cells_capacity(n_zl, int(g["n_int"]), float(lrange))
    if cap <= 0:
        raise _lib.PttError("ptt_gw_cells_capacity: sizes out of range")
    cap_c = lib.ptt_gw_cells_coef_capacity(cap)
    ncell = torch.empty(n_y, dtype=torch.int32, device="cuda")
    meta = torch.zeros((n_y, cap + 2), dtype=torch.int32, device="cuda")
    coef = torch.zeros((n_y, cap_c, 2), dtype=torch.float64, device="cuda")
    with STAGES.stage("gw_cells_build", 1):
        _lib.check(lib.ptt_gw_cells_build(C.byref(plan.gw_plan), cap, _ptr(ncell), _ptr(meta), _ptr(coef),
                                          _stream_ptr(torch)), "ptt_gw_cells_build")
    if int(ncell.min().item()) < 0:
        # capacity exceeded or a lookup that does not move monotonically: this plan stays on the per-sample kernel
        plan._gw_cells = False
        return False
    cells = GwCells(cap=cap, reserved=0, ncell=ncell.data_ptr(), meta=meta.data_ptr(), coef=coef.data_ptr())
    plan._gw_cells = (cells, ncell, meta, coef)
    return plan._gw_cells


def spec_den_gw_batch(plan: SsmPlan, pv, slf=None, scale=None, want_sd=True, cells=None):
    """spec_den_gw_scaled + pow_spec (+ omgw0 scale) for a batch.  cells=None: the cell formulation for batches of
    GW_CELLS_MIN profiles or more, the per-sample kernel below that; True / False force one of them."""
    torch = _torch()
    lib = _lib.load()
    P = pv.shape[0]
    dev = pv.device
    n_y = plan.y.size
    slf_t = torch.ones(P, dtype=torch.float64, device=dev) if slf is None else dev_f64(slf).reshape(-1)
    scale_t = None if scale is None else dev_f64(scale).reshape(-1)
    sd = torch.empty((P, n_y), dtype=torch.float64, device=dev) if want_sd else None
    pw = torch.empty((P, n_y), dtype=torch.float64, device=dev)
    om = torch.empty((P, n_y), dtype=torch.float64, device=dev) if scale_t is not None else None
    if getattr(plan, "general", False):
        with STAGES.stage("spec_den_gw_general", 1):
            t = plan._t
            _lib.check(lib.ptt_spec_den_gw_general_batch(int(plan.z_lookup.size), _ptr(t["z_lookup"]), n_y, _ptr(t["y"]),
                                                         int(plan.gw["n_int"]), _ptr(t["gw_Z1"]), _ptr(t["gw_Z2"]),
                                                         _ptr(t["gw_G"]), float(plan.gw["prefactor"]), P, _ptr(pv),
                                                         pv.shape[1], _ptr(slf_t), _ptr(scale_t), _ptr(sd), _ptr(pw),
                                                         _ptr(om), _stream_ptr(torch)), "ptt_spec_den_gw_general_batch")
        return sd, pw, om
    if cells is None:
        cells = USE_GW_CELLS and P >= GW_CELLS_MIN
    tabs = gw_cells(plan) if cells else False
    if tabs:
        cs = tabs[0]
        table = torch.empty(lib.ptt_gw_cells_table_size(int(plan.z_lookup.size), P), dtype=torch.float64, device=dev)
        with STAGES.stage("spec_den_gw_cells", 2):
            _lib.check(lib.ptt_spec_den_gw_cells_batch(C.byref(plan.gw_plan), C.byref(cs), P, _ptr(pv), pv.shape[1],
                                                       _ptr(slf_t), _ptr(scale_t), _ptr(table), _ptr(sd), _ptr(pw),
                                                       _ptr(om), _stream_ptr(torch)), "ptt_spec_den_gw_cells_batch")
        return sd, pw, om
    with STAGES.stage("spec_den_gw", 1):
        _lib.check(lib.ptt_spec_den_gw_batch(C.byref(plan.gw_plan), P, _ptr(pv), pv.shape[1], _ptr(slf_t), _ptr(scale_t),
                                             _ptr(sd), _ptr(pw), _ptr(om), _stream_ptr(torch)), "ptt_spec_den_gw_batch")
    return sd, pw, om


def pow_spec_batch(z_dev, sd):
    torch = _torch()
    lib = _lib.load()
    out = torch.empty_like(sd)
    with STAGES.stage("pow_spec", 1):
        _lib.check(lib.ptt_pow_spec_batch(sd.shape[0], sd.shape[1], _ptr(z_dev), _ptr(sd), sd.shape[1], _ptr(out),
                                          _stream_ptr(torch)), "ptt_pow_spec_batch")
    return out


def spectra_from_shells(plan: SsmPlan, shells: ShellBatch, pp, slf=None, scale=None, keep=False,
                        v_wall_host=None, valid_host=None) -> SpectrumBatch:
    """Shell profiles -> A^2 -> spec_den_v -> spec_den_gw -> pow_gw (-> omgw0), all on the device.
    v_wall_host / valid_host: optional host copies of the wall speeds and of "this point has a shell" (see
    spec_den_v_auto); results are the same with or without them."""
    a2 = a2_batch(plan, shells, pp)
    res = SpectrumBatch(y=plan.y, pow_gw=None, sd_gw=None)
    v_wall = pp[:, 0].contiguous()
    if plan.mode == "ssm" and plan.has_y_pass:
        sdv_y = spec_den_v_auto(plan, 0, a2, v_wall, v_wall_host, valid_host)
        res.sd_v = sdv_y
        res.pow_v = pow_spec_batch(plan._t["y"], sdv_y)
        if keep:
            res.a2 = a2[:, : int(plan.seg[1])]
    sdv_l = spec_den_v_auto(plan, len(plan.passes) - 1, a2, v_wall, v_wall_host, valid_host)
    if keep:
        res.sd_v_lookup = sdv_l
    res.sd_gw, res.pow_gw, res.omgw0 = spec_den_gw_batch(plan, sdv_l, slf, scale)
    return res


def sin_transform_batch(z, xi, f, off, npts, z_st_thresh=Z_ST_THRESH, method="auto"):
    """Batched sin_transform over ragged profiles stored as rows (calculators.py:161-190).
    method: "direct" = one sine per (z, xi) pair; "moments" = block-moment evaluation of the same sum (needs
    0 <= xi <= 1 and an exact branch that ends at z <= 51.2); "auto" picks moments when both hold."""
    torch = _torch()
    lib = _lib.load()
    zz = np.asarray(z, dtype=np.float64)
    frac = dev_f64(blend_fraction(zz, z_st_thresh))
    z_d = dev_f64(zz)
    xi, f = dev_f64(xi), dev_f64(f)
    if xi.dim() == 1:
        xi, f = xi[None, :], f[None, :]
    P, stride = xi.shape
    off_t = torch.as_tensor(off, dtype=torch.int32, device=xi.device).reshape(-1).contiguous()
    npts_t = torch.as_tensor(npts, dtype=torch.int32, device=xi.device).reshape(-1).contiguous()
    out = torch.empty((P, zz.size), dtype=torch.float64, device=xi.device)
    if method == "auto":
        fr = blend_fraction(zz, z_st_thresh)
        col = torch.arange(stride, device=xi.device)[None, :]
        valid = (col >= off_t[:, None]) & (col < (off_t + npts_t)[:, None])
        in_unit = bool(((xi >= 0.0) & (xi <= 1.0) | ~valid).all().item())
        method = "moments" if (in_unit and z_exact_max(zz, fr) <= 51.2) else "direct"
    if method == "moments":
        _lib.check(lib.ptt_sin_transform_unit_batch(P, _ptr(xi), _ptr(f), stride, _ptr(off_t), _ptr(npts_t), zz.size,
                                                    _ptr(z_d), _ptr(frac), _ptr(out), _stream_ptr(torch)),
                   "ptt_sin_transform_unit_batch")
        return out
    _lib.check(lib.ptt_sin_transform_batch(P, _ptr(xi), _ptr(f), stride, _ptr(off_t), _ptr(npts_t), zz.size, _ptr(z_d),
                                           _ptr(frac), _ptr(out), _stream_ptr(torch)), "ptt_sin_transform_batch")
    return out


def profile_integrals(shells: ShellBatch, pp2):
    """Raw trapezoid integrals over xi^3 per shell (pttools/bubble/thermo.py:139-221); see ptt_profile_integrals_batch."""
    torch = _torch()
    lib = _lib.load()
    P = len(shells)
    out = torch.empty((P, 8), dtype=torch.float64, device=shells.rows_v.device)
    with STAGES.stage("profile_integrals", 1):
        _lib.check(lib.ptt_profile_integrals_batch(P, _ptr(shells.rows_v), _ptr(shells.rows_w), _ptr(shells.rows_xi),
                                                   shells.rows_v.shape[1], _ptr(shells.off), _ptr(shells.npts), _ptr(pp2),
                                                   _ptr(out), _stream_ptr(torch)), "ptt_profile_integrals_batch")
    return out


def fp64_pipe_flops(mode, iters=20000):
    """Measured FP64 throughput of the FMA pipe (mode 0), of DMMA m8n8k4 (1), and of both interleaved (2)."""
    torch = _torch()
    lib = _lib.load()
    out = C.c_double(0.0)
    _lib.check(lib.ptt_fp64_pipe_probe(int(mode), iters, C.byref(out), _stream_ptr(torch)), "ptt_fp64_pipe_probe")
    return out.value


def fp64_peak_flops(iters=20000):
    torch = _torch()
    lib = _lib.load()
    out = C.c_double(0.0)
    _lib.check(lib.ptt_fp64_peak_probe(iters, C.byref(out), _stream_ptr(torch)), "ptt_fp64_peak_probe")
    return out.value
