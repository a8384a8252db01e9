"""Where does one bench step spend its time?  torch.profiler (kernels incl. torch's own) + cProfile (host)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from pttools_b200 import engine, sweep, fluid_reference
from pttools_b200.models import BagModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
CHUNK = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
model = BagModel(alpha_n_min=0.005)
plan = engine.SsmPlan(bench.Y, cs=float(np.sqrt(model.csb2)), nt=bench.NPT[1], n_z_lookup=bench.NPT[2], nxi=bench.NPT[0],
                      mode="ssm")
fluid_reference.ref()


def step(s):
    vw, an = bench.batch_points(s, 0, 1, B)
    res = sweep.sweep_spectra(model, engine.dev_f64(vw), engine.dev_f64(an), y=bench.Y, r_star=bench.R_STAR,
                              n_xi=bench.N_XI, chunk=CHUNK, plan=plan)
    torch.cuda.synchronize()
    return res


for s in range(2):
    step(s)
t0 = time.perf_counter(); step(2); print("step wall ms", (time.perf_counter() - t0) * 1e3)
pr = cProfile.Profile(); pr.enable(); step(3); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    step(4)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
