"""Wall-clock breakdown of one bench step on the GPU box (host pieces vs the C call)."""
import os, sys, time
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


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:40s} {1e3 * (t1 - t0):8.2f} ms", flush=True)
    return t1


for s in range(5):
    print("step", s)
    vw, an = bench.batch_points(s, 0, 1, B)
    torch.cuda.synchronize()
    t = t00 = time.perf_counter()
    codes = np.array(model.solution_type_codes(vw, an, np.asarray(model.wn(an))), dtype=np.int64); t = tick("solution_type_codes", t)
    shards = sweep.shard_rows(vw, sweep.point_cost(codes), 1); t = tick("shard_rows", t)
    idx = shards[0]
    pt = bubble.point_table(model, vw[idx], an[idx]); t = tick("point_table", t)
    scale = sweep.omgw0_scale(vw[idx], an[idx], 0.1); t = tick("omgw0_scale", t)
    pg, pm, sc = engine.dev_f64(pt.pg), engine.dev_f64(pt.pm), engine.dev_f64(scale); t = tick("uploads", t)
    out = torch.empty((B, bench.Y.size), dtype=torch.float64, device="cuda"); t = tick("alloc out", t)
    for timing in (False, True):
        res = engine.sweep_c(plan, pg, pm, sc, n_xi=5000, r_star=0.1, chunk=8000, want=("pow_gw", "omgw0"), out={"omgw0": out}, timing=timing)
        t = tick(f"sweep_c timing={timing}", t)
    full = torch.empty((B, bench.Y.size), dtype=torch.float64, device="cuda")
    full[torch.as_tensor(idx, device="cuda")] = out; t = tick("reassemble", t)
    t = time.perf_counter()
    r = sweep.sweep_omgw0_distributed(model, vw, an, y=bench.Y, r_star=0.1, n_xi=5000, chunk=8000, plan=plan); t = tick("whole sweep_omgw0_distributed", t)
    r = sweep.sweep_omgw0_distributed(model, vw, an, y=bench.Y, r_star=0.1, n_xi=5000, chunk=8000, plan=plan, mem="host"); t = tick("same, mem=host", t)
    st = engine._lib.SweepStats(); 
    print("  stage ms:", {n: round(sweep.LAST_LOCAL["result"].stats.stage_ms[i], 1) for i, n in enumerate(engine._lib.STAGE_NAMES)})
