"""Starting guesses for the model-independent solver: the 100 x 100 bag-model table of the reference
(pttools/bubble/fluid_reference.py:24-210), built on the device with the bag kernels instead of a process pool
(10^4 bag solves: one kernel launch here, ~27 s on 8 cores there), kept in memory instead of an HDF5 file next to
the module, and queried for whole batches at once (nearest valid point of the same solution type)."""
from __future__ import annotations

import functools

import numpy as np

from . import engine


class FluidReference:
    def __init__(self, v_wall_min=0.05, v_wall_max=0.95, alpha_n_min=0.01, alpha_n_max=0.99, n_v_wall=100, n_alpha_n=100):
        torch = engine._torch()
        self.v_wall = np.linspace(v_wall_min, v_wall_max, n_v_wall, endpoint=True)
        self.alpha_n = np.linspace(alpha_n_min, alpha_n_max, n_alpha_n, endpoint=True)
        vw, an = np.meshgrid(self.v_wall, self.alpha_n)           # [n_alpha, n_vw], fluid_reference.py:96-97
        vw_f, an_f = vw.ravel(), an.ravel()
        an_max, _ = engine.bag_alpha_n_max(self.v_wall)
        an_max_f = an_max.cpu().numpy()[None, :].repeat(n_alpha_n, 0).ravel()
        data = np.full((vw_f.size, 6), np.nan)
        sol_type = np.full(vw_f.size, -1, dtype=np.int64)
        chunk = 2000
        for s in range(0, vw_f.size, chunk):
            e = min(vw_f.size, s + chunk)
            sh = engine.sweep_shells_bag(vw_f[s:e], an_f[s:e], an_max=engine.dev_f64(an_max_f[s:e]))
            wv = engine.shell_wall_values(sh).cpu().numpy()
            st = sh.sol_type.cpu().numpy()
            ok = (sh.status.cpu().numpy() == 0) & ~(an_f[s:e] > an_max_f[s:e]) & np.all(np.isfinite(wv[:, :7]), axis=1)
            # junction_condition_deviation1 check of fluid_reference.py:183-187
            g2 = lambda v: 1.0 / (1.0 - v ** 2)
            dev = wv[:, 4] * g2(wv[:, 2]) * wv[:, 2] - wv[:, 5] * g2(wv[:, 3]) * wv[:, 3]
            ok &= np.abs(dev) <= 0.025
            ok &= ~np.any(wv[:, :6] < 0, axis=1)
            data[s:e][ok] = wv[ok, :6]
            sol_type[s:e][ok] = st[ok]
        self.data = data.reshape(n_alpha_n, n_v_wall, 6)
        self.sol_type = sol_type.reshape(n_alpha_n, n_v_wall)
        self.vp, self.vm, self.vp_tilde, self.vm_tilde, self.wp, self.wm = (self.data[:, :, k] for k in range(6))
        # device copies for batched nearest-neighbour queries: the valid entries of the three solution types back to back
        sets = [np.nonzero(sol_type == code)[0] for code in (0, 1, 2)]
        self._set_off = torch.as_tensor(np.cumsum([0] + [s.size for s in sets]).astype(np.int32), device="cuda")
        order = np.concatenate(sets)
        self._cx = torch.as_tensor(np.ascontiguousarray(vw_f[order]), device="cuda")
        self._cy = torch.as_tensor(np.ascontiguousarray(an_f[order]), device="cuda")
        self._vals = torch.as_tensor(np.ascontiguousarray(data[order]), device="cuda")

    def get_batch(self, v_wall, alpha_n, sol_type):
        """Nearest valid table entry of the same solution type for every point: tensor [n, 6]
        (vp, vm, vp_tilde, vm_tilde, wp, wm), like NearestNDInterpolator in fluid_reference.py:60-62,152-163; NaN rows
        for points without a solution type."""
        import ctypes as C
        from . import _lib
        torch = engine._torch()
        lib = _lib.load()
        v_wall, alpha_n = engine.dev_f64(v_wall).reshape(-1), engine.dev_f64(alpha_n).reshape(-1)
        st = torch.as_tensor(np.asarray(sol_type), device=v_wall.device).reshape(-1).to(torch.int32).contiguous()
        n = v_wall.numel()
        idx = torch.empty(n, dtype=torch.int32, device=v_wall.device)
        p = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.ptt_nearest_batch(n, p(v_wall), p(alpha_n), p(st), 3, p(self._set_off), p(self._cx), p(self._cy),
                                         p(idx), engine._stream_ptr(torch)), "ptt_nearest_batch")
        idx = idx.long()
        out = self._vals[idx.clamp(min=0)]
        out[idx < 0] = float("nan")
        return out

    def get(self, v_wall, alpha_n, sol_type):
        from .models import SOL_CODE
        code = SOL_CODE[sol_type] if not isinstance(sol_type, (int, np.integer)) else int(sol_type)
        return self.get_batch([v_wall], [alpha_n], [code])[0].cpu().numpy()


@functools.lru_cache(maxsize=1)
def ref() -> FluidReference:
    return FluidReference()
