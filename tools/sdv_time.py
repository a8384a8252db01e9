import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pttools_b200 import engine
y = np.logspace(-1, 3, 1000)
plan = engine.SsmPlan(y, cs=engine.CS0, nt=10000, n_z_lookup=10000, nxi=2000, mode="ssm")
n_p = 1000
rng = np.random.default_rng(0)
a2 = torch.as_tensor(np.abs(rng.normal(size=(n_p, plan.n_a2))) * np.logspace(0, -12, plan.n_a2)[None, :], device="cuda")
for ipass in (0, 1):
    out = torch.empty((n_p, plan.sdv_plans[ipass].n_z), dtype=torch.float64, device="cuda")
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); engine.spec_den_v_group(plan, ipass, a2, 0.5, out); b.record(); torch.cuda.synchronize()
        print("pass", ipass, "ms", a.elapsed_time(b))
    from pttools_b200.engine import _group_band
    eb, kw = _group_band(plan.passes[ipass], 0.5)
    print("kw", kw, "tiles", eb.size, "flops", 2.0 * plan.sdv_plans[ipass].n_z * kw * n_p)
