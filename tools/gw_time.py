"""Time of the K7' cell kernel (+ table) on synthetic P_v tables; agreement with the per-sample kernel on a small batch.
Usage: python tools/gw_time.py [P]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from pttools_b200 import engine
y = np.logspace(-1, 3, 1000)
plan = engine.SsmPlan(y, cs=engine.CS0, nt=10000, n_z_lookup=10000, nxi=2000, mode="ssm", only_lookup_pass=True)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
zl = plan.z_lookup
rng = np.random.default_rng(0)
small = torch.as_tensor(zl[None, :] ** 3 / (1 + (zl[None, :] / rng.uniform(1, 30, (200, 1))) ** 8), device="cuda")
o0 = engine.spec_den_gw_batch(plan, small, None, None, cells=False)[1].cpu().numpy()
o1 = engine.spec_den_gw_batch(plan, small, None, None, cells=True)[1].cpu().numpy()
pv = torch.rand((P, zl.size), dtype=torch.float64, device="cuda") + 0.5
ms = []
for rep in range(5):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); engine.spec_den_gw_batch(plan, pv, None, None, cells=True, want_sd=False); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
print(f"cells kernel + table: {np.median(ms[1:]):.2f} ms per {P} profiles; max rel diff vs per-sample kernel {np.nanmax(np.abs(o1 / o0 - 1)):.1e}")
