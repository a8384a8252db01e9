"""Small end-to-end run of every kernel, meant to be run under compute-sanitizer on the GPU box."""
import sys
import numpy as np
sys.path.insert(0, ".")
from pttools_b200 import engine, sweep, bubble
from pttools_b200.models import BagModel, ConstCSModel

y = np.logspace(-1, 3, 40)
vw = np.repeat([0.3, 0.62, 0.9], 100)
an = np.tile(np.linspace(0.05, 0.3, 100), 3)
engine.GROUP_MIN = 50
r = sweep.sweep_spectra(BagModel(alpha_n_min=0.005), vw, an, y=y, r_star=0.1, nt=300, n_z_lookup=400, n_xi=600, chunk=128)
print("generic", int((r.status == 0).sum()), float(np.nanmax(r.omgw0.cpu().numpy())))
m = ConstCSModel(1 / 4, 1 / 3, a_s=1.5, a_b=1, V_s=1)
r = sweep.sweep_spectra(m, [0.5, 0.7, 0.9], [0.2, 0.2, 0.2], y=y, r_star=0.1, nt=300, n_z_lookup=400, n_xi=600)
print("const_cs", r.status.cpu().numpy())
r = sweep.power_gw_scaled_bag_sweep([0.5, 0.7, 0.9, 0.2], [0.1, 0.1, 0.1, 0.45], y, (600, 300, 400))
print("bag", r.status.cpu().numpy())
v, w, xi, st = engine.integrate_batch([0.25, 0.4], [1.0, 1.0], [0.5, 0.7], [0.0, 0.0], [-50.0, 20.0], 300)
print("integrate", st.cpu().numpy())
xi_ = np.linspace(0, 1, 257)
f = np.maximum(0, np.cos(xi_ + 0.3)) * ((xi_ > 0.2) & (xi_ < 0.8))
out = engine.sin_transform_batch(np.logspace(-1, 3, 100), xi_, f, [0], [257])
print("sin_transform", float(out.abs().max()))
