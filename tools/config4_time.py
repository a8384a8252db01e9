"""BASELINE.json config 4: batched sin_transform, 10^4 profiles x 4096 wavenumbers x 8192 xi samples
(f_p(xi) = max(0, cos(xi + phi_p)) on [xi_a, xi_b], numpy default_rng(0))."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pttools_b200 import engine

P, n_z, n_xi = 10000, 4096, 8192
rng = np.random.default_rng(0)
phi, xa = rng.uniform(0, 1, P), rng.uniform(0.05, 0.5, P)
xb = xa + rng.uniform(0.05, 0.45, P)
xi = np.linspace(0, 1, n_xi)
f = np.maximum(0, np.cos(xi[None, :] + phi[:, None])) * ((xi[None, :] > xa[:, None]) & (xi[None, :] < xb[:, None]))
xi_d = engine.dev_f64(np.broadcast_to(xi, (P, n_xi)).copy())
f_d = engine.dev_f64(f)
off, npts = np.zeros(P, np.int32), np.full(P, n_xi, np.int32)
out = {}
for name, z in (("all_exact", np.logspace(-1, np.log10(50), n_z)), ("mixed", np.logspace(-1, 3, n_z))):
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = engine.sin_transform_batch(z, xi_d, f_d, off, npts, method="moments"); b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    sub = 100
    a.record(); rd = engine.sin_transform_batch(z, xi_d[:sub], f_d[:sub], off[:sub], npts[:sub], method="direct"); b.record()
    torch.cuda.synchronize()
    ms_direct = a.elapsed_time(b) * P / sub
    err = float((r[:sub] - rd).abs().max() / rd.abs().max())
    n_exact = int((z <= 50).sum())
    out[name] = {"moments_ms": ms, "direct_ms_extrapolated_from_100_profiles": ms_direct,
                 "sine_products_per_s_equivalent": P * n_exact * n_xi / (ms * 1e-3), "max_abs_dev_over_row_max": err}
print(json.dumps({"config": "4: 10^4 profiles x 4096 z x 8192 xi", **out}))
