"""Where the shell solver (generic_solve_kernel) spends its time on a bench step: per-point cycle counts (build with
-DPTT_K3_CLOCKS), by solution type, and a simulation of the warp schedule in the current order against slow-first orders.
Usage (GPU box): PTT_NVCC_EXTRA=-DPTT_K3_CLOCKS python pttools_b200/build.py --force && python tools/k3_clocks.py
(or PTT_B200_LIB=<an instrumented build of the library> python tools/k3_clocks.py); per-point arrays go to
gpurun_out/k3_clocks.npz."""
import sys, heapq
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from pttools_b200 import bubble, engine, fluid_reference
from pttools_b200.models import BagModel

model = BagModel(alpha_n_min=0.005)
fluid_reference.ref()
vw, an = bench.batch_points(3, 0, 1, 64000)
pt = bubble.point_table(model, vw, an, device=True)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sh = engine.sweep_shells_generic(pt.pg)
ev0.record(); sh = engine.sweep_shells_generic(pt.pg); ev1.record(); torch.cuda.synchronize()
print("kernel ms", ev0.elapsed_time(ev1))
sc = sh.scalars.cpu().numpy()
clk, nres = sc[:, 15], sc[:, 14]
codes = pt.codes
import os
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/k3_clocks.npz", vw=vw, an=an, codes=codes, clk=clk, n_resid=nres)
ok = np.isfinite(clk) & (clk > 0)
print("points with a profile", ok.sum(), "of", clk.size)
for c, name in ((0, "sub_def"), (1, "hybrid"), (2, "deton")):
    m = ok & (codes == c)
    if m.any():
        print(f"{name:8s} n={m.sum():6d} clocks mean {clk[m].mean()/1e6:6.2f} M  p50 {np.median(clk[m])/1e6:6.2f}  p99 {np.percentile(clk[m],99)/1e6:6.2f}  "
              f"max {clk[m].max()/1e6:6.2f} M   n_resid mean {nres[m].mean():.1f} max {nres[m].max():.0f}")
top = np.argsort(-np.nan_to_num(clk))[:12]
for i in top:
    print(f"  slow: vw={vw[i]:.4f} an={an[i]:.4f} code={codes[i]} clocks={clk[i]/1e6:.2f} M n_resid={nres[i]:.0f}")
# warp durations = max over the 32 lanes (lanes of one warp run in lockstep through the union of their paths)
c = np.nan_to_num(clk, nan=0.0)
n = c.size
pad = (-n) % 32
wd = np.concatenate([c, np.zeros(pad)]).reshape(-1, 32).max(axis=1)


def makespan(durs, slots=148 * 8):
    """CTAs of 2 warps, 4 CTAs per SM: approximate with `slots` warp slots filled greedily in launch order."""
    h = [0.0] * slots
    heapq.heapify(h)
    end = 0.0
    for d in durs:
        t = heapq.heappop(h) + d
        end = max(end, t)
        heapq.heappush(h, t)
    return end

print(f"warps {wd.size}; sum of warp durations / slots = {wd.sum()/1184/1e6:.2f} M clocks; longest warp {wd.max()/1e6:.2f} M")
print(f"makespan, launch order      : {makespan(wd)/1e6:.2f} M clocks")
print(f"makespan, slowest warp first: {makespan(np.sort(wd)[::-1])/1e6:.2f} M clocks")
# a-priori key: hybrids first, then deflagrations, then detonations, rows without solution last
key = np.where(codes == 1, 0, np.where(codes == 0, 1, np.where(codes == 2, 2, 3)))
order = np.argsort(key, kind="stable")
co = np.concatenate([c[order], np.zeros(pad)]).reshape(-1, 32).max(axis=1)
print(f"makespan, points ordered hybrid/defl/deton: {makespan(co)/1e6:.2f} M clocks (warps regrouped)")
