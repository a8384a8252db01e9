"""Host-side time of the phases of one bench step (1 GPU): where the device can idle between two sweeps."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from pttools_b200 import bubble, engine, fluid_reference, sweep
from pttools_b200.models import BagModel

model = BagModel(alpha_n_min=0.005)
plan = engine.SsmPlan(bench.Y, cs=float(np.sqrt(model.csb2)), nt=10000, n_z_lookup=10000, nxi=2000, mode="ssm", only_lookup_pass=True)
fluid_reference.ref()
for s in range(6):
    vw, an = bench.batch_points(s, 0, 1, 64000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh = sweep.shard_rows(vw, None, 1, only=0)
    t1 = time.perf_counter()
    free = torch.cuda.mem_get_info()[0]
    t2 = time.perf_counter()
    pt = bubble.point_table(model, vw, an, device=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    sc = sweep.omgw0_scale(vw, an, 0.1, None, None, "default", device=True)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    res = sweep.sweep_spectra(model, vw, an, y=bench.Y, r_star=0.1, want_pow_v=False, want_pow_gw=False, chunk=8000, plan=plan)
    t5 = time.perf_counter()
    torch.cuda.synchronize()
    t6 = time.perf_counter()
    print(f"step {s}: shard {1e3*(t1-t0):.1f}  mem_get_info {1e3*(t2-t1):.1f}  point_table {1e3*(t3-t2):.1f}  scale {1e3*(t4-t3):.1f}  "
          f"sweep_spectra call {1e3*(t5-t4):.1f}  tail sync {1e3*(t6-t5):.1f} ms")

print("--- mem='host' (the e2e path)")
for s in range(5):
    vw, an = bench.batch_points(100 + s, 0, 1, 64000)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pt = bubble.point_table(model, vw, an, device=False)
    t1 = time.perf_counter()
    sc = sweep.omgw0_scale(vw, an, 0.1, None, None, "default", device=False)
    t2 = time.perf_counter()
    res = sweep.sweep_spectra(model, vw, an, y=bench.Y, r_star=0.1, want_pow_v=False, want_pow_gw=False, chunk=8000, plan=plan, mem="host")
    t3 = time.perf_counter()
    st = res.stats
    print(f"step {s}: point_table {1e3*(t1-t0):.1f}  scale {1e3*(t2-t1):.1f}  sweep_spectra(host) call {1e3*(t3-t2):.1f} ms "
          f"(includes its own point_table + scale); wait_shells {st.host_wait_shells_ms:.1f}")
