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
inputs = [tuple(engine.dev_f64(a) for a in bench.batch_points(s, 0, 1, B)) for s in range(12)]
def step(s):
    return sweep.sweep_spectra(model, *inputs[s], y=bench.Y, r_star=bench.R_STAR, n_xi=bench.N_XI, chunk=2000, plan=plan)
def run(tag, s0):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(s0, s0 + 3):
        step(s)
    b.record(); torch.cuda.synchronize()
    print(tag, a.elapsed_time(b) / 3, "ms/step")
step(0); step(1)
run("plain", 2)
engine.STAGES.enable(True); run("stages", 5); engine.STAGES.enable(False)
smp = bench.ClockSampler(0); run("sampler", 8); print(smp.stop())
run("plain again", 2)
