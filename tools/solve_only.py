import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from pttools_b200 import engine, fluid_reference, bubble
from pttools_b200.models import BagModel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
model = BagModel(alpha_n_min=0.005)
fluid_reference.ref()
vw, an = bench.batch_points(3, 0, 1, B)
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); bb = bubble.sweep_bubbles(model, vw, an, n_xi=bench.N_XI); b.record(); torch.cuda.synchronize()
    print("sweep_bubbles ms", a.elapsed_time(b))
print("status", np.bincount(bb.status.cpu().numpy()))
