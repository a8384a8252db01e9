import numpy as np, torch, time
from pttools_b200 import engine
y = np.logspace(-1, 3, 1000)
plan = engine.SsmPlan(y, cs=engine.CS0, nt=10000, n_z_lookup=10000, nxi=2000, mode="ssm")
P = 2000
zl = plan.z_lookup
rng = np.random.default_rng(0)
pv = torch.as_tensor(zl[None, :] ** 3 / (1 + (zl[None, :] / rng.uniform(1, 30, (P, 1))) ** 8), device="cuda")
t0 = time.time(); engine.gw_cells(plan); torch.cuda.synchronize(); print("cells build s", time.time() - t0, "cap", engine.gw_cells(plan)[0].cap)
nc = engine.gw_cells(plan)[1].cpu().numpy(); print("ncell min/mean/max", nc.min(), nc.mean(), nc.max())
for cells in (False, True):
    for rep in range(6):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); out = engine.spec_den_gw_batch(plan, pv, None, None, cells=cells); b.record(); torch.cuda.synchronize()
        print("cells", cells, "ms", a.elapsed_time(b))
    if cells: o1 = out[1].cpu().numpy()
    else: o0 = out[1].cpu().numpy()
print("max rel diff", np.nanmax(np.abs(o1 / o0 - 1)))
