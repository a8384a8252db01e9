"""Per-step times of the bench loop with and without the clock sampler / stage timers, plus allocator activity."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from pttools_b200 import engine, sweep, fluid_reference
from pttools_b200.models import BagModel
B = 32000
model = BagModel(alpha_n_min=0.005)
plan = engine.SsmPlan(bench.Y, cs=float(np.sqrt(model.csb2)), nt=bench.NPT[1], n_z_lookup=bench.NPT[2], nxi=bench.NPT[0], mode="ssm")
fluid_reference.ref()
inputs = [tuple(engine.dev_f64(a) for a in bench.batch_points(s, 0, 1, B)) for s in range(14)]
def step(s):
    return sweep.sweep_spectra(model, *inputs[s], y=bench.Y, r_star=bench.R_STAR, n_xi=bench.N_XI, chunk=2000, plan=plan)
def run(tag, s0, n=4):
    torch.cuda.synchronize()
    ts = []
    for s in range(s0, s0 + n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0 = torch.cuda.memory_stats()["num_device_alloc"]
        t0 = time.perf_counter()
        a.record(); r = step(s); int((r.status == 0).sum()); b.record(); torch.cuda.synchronize()
        ts.append((round(a.elapsed_time(b)), round((time.perf_counter() - t0) * 1e3), torch.cuda.memory_stats()["num_device_alloc"] - m0))
    print(tag, ts, flush=True)
step(0); step(1); step(2)
run("plain", 3)
smp = bench.ClockSampler(0); run("sampler", 3, 6); print(smp.stop())
engine.STAGES.enable(True); run("stages", 3, 6); engine.STAGES.enable(False)
run("plain again", 9)
