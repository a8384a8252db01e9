"""Where the host time of a config-3 step goes (bench.py --workload config3: 8 ConstCS models that share csb2 in one
sweep): cProfile over three steps after warm-up, by cumulative time, plus the device-side timeline of the sweeps
(PTT_SWEEP_TRACE=1 python tools/config3_phases.py 2> trace.txt)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from pttools_b200 import fluid_reference, sweep

fluid_reference.ref()
fams, _ = bench.config3_models()
v_walls, alpha_ns = np.linspace(0.05, 0.95, 128), np.linspace(0.05, 0.5, 128)
kw = dict(y=bench.Y, r_star=bench.R_STAR, Tn=200, n_xi=bench.N_XI, want_pow_v=False, chunk=8000)
for s in range(3):
    sweep.sweep_models(fams[s], v_walls, alpha_ns, timing=True, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for s in range(3, 6):
    sweep.sweep_models(fams[s], v_walls, alpha_ns, timing=True, **kw)
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"3 steps: host returned after {1e3 * (t1 - t0):.0f} ms, device done after {1e3 * (t2 - t0):.0f} ms")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
