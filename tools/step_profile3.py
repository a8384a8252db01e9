"""Where do the slow steps come from?  Per-step wall time of the host preparation and of the C call, 14 steps."""
import gc, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pttools_b200 import engine, sweep, bubble, fluid_reference
from pttools_b200.models import BagModel

torch.cuda.set_device(0)
model = BagModel(alpha_n_min=0.005)
plan = engine.SsmPlan(bench.Y, cs=float(np.sqrt(model.csb2)), nt=bench.NPT[1], n_z_lookup=bench.NPT[2], nxi=bench.NPT[0], mode="ssm", only_lookup_pass=True)
fluid_reference.ref()
B = 64000
if "--nogc" in sys.argv:
    gc.disable()
rows = []
for s in range(14):
    vw, an = bench.batch_points(s, 0, 1, B)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    codes = np.array(model.solution_type_codes(vw, an, np.asarray(model.wn(an))), dtype=np.int64)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    pt = bubble.point_table(model, vw, an, sol_type=codes)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    scale = sweep.omgw0_scale(vw, an, 0.1)
    pg, pm, sc = engine.dev_f64(pt.pg), engine.dev_f64(pt.pm), engine.dev_f64(scale)
    out = torch.empty((B, bench.Y.size), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    res = engine.sweep_c(plan, pg, pm, sc, n_xi=5000, r_star=0.1, chunk=8000, want=("pow_gw", "omgw0"), out={"omgw0": out}, timing=True)
    torch.cuda.synchronize(); t4 = time.perf_counter()
    st = res.stats
    stage = sum(st.stage_ms[i] for i in (0, 2, 3, 5, 6))
    rows.append([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), stage, st.stage_ms[0], st.stage_ms[3], st.stage_ms[5], st.stage_ms[6], st.stage_ms[1], st.host_wait_shells_ms, st.host_wait_retry_ms])
print("codes  table  scale+up  sweep_c | sum(K3,K4,K5,K6,K7) K3 K5 K6 K7 retry waitS waitR")
for r in rows:
    print(" ".join(f"{x:8.1f}" for x in r))
