"""Where the host time of an end-to-end step goes (bench.py's e2e leg: sweep.sweep_omgw0_distributed(mem="host"), numpy in,
pinned numpy out): cProfile over three steps after warm-up, by cumulative time."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from pttools_b200 import engine, fluid_reference, sweep
from pttools_b200.models import BagModel

model = BagModel(alpha_n_min=0.005)
plan = engine.SsmPlan(bench.Y, cs=float(np.sqrt(model.csb2)), nt=10000, n_z_lookup=10000, nxi=2000, mode="ssm", only_lookup_pass=True)
fluid_reference.ref()
pts = [bench.batch_points(1000 + s, 0, 1, 64000) for s in range(6)]


def step(s):
    vw, an = pts[s]
    shard, table = sweep.sweep_omgw0_distributed(model, vw, an, y=bench.Y, r_star=bench.R_STAR, n_xi=bench.N_XI, chunk=8000, plan=plan, mem="host")
    return float(np.nanmax(table[:, 0]))


for s in range(3):
    step(s)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for s in range(3, 6):
    step(s)
pr.disable()
print(f"3 steps: {1e3 * (time.perf_counter() - t0) / 3:.1f} ms per step")
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
