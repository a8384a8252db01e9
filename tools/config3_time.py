"""BASELINE.json config 3 (one model-quadrant of it): ConstCSModel(css2, csb2, a_s=5, a_b=1, V_s=1) for
css2, csb2 in {1/3, 1/4}, v_wall = linspace(0.05, 0.95, 128) x alpha_n = linspace(0.05, 0.5, 128) restricted to
alpha_n >= model.alpha_n_min, r_star = 0.1, Tn = 200, default resolutions."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pttools_b200 import sweep, fluid_reference
from pttools_b200.models import ConstCSModel

fluid_reference.ref()
v_walls, alpha_ns = np.linspace(0.05, 0.95, 128), np.linspace(0.05, 0.5, 128)
models = [ConstCSModel(css2=a, csb2=b, a_s=5, a_b=1, V_s=1) for a in (1 / 3, 1 / 4) for b in (1 / 3, 1 / 4)]
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = sweep.sweep_models(models, v_walls, alpha_ns, r_star=0.1, Tn=200)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
n = sum(int(r["mask"].sum()) for r in res)
ok = sum(int((r["result"].status == 0).sum()) for r in res)
types = np.concatenate([r["result"].sol_type.cpu().numpy() for r in res])
print(json.dumps({"config": "3 (4 of 64 models)", "points": n, "valid": ok, "seconds": dt, "spectra_per_s": n / dt,
                  "solution_types": np.bincount(types[types >= 0], minlength=3).tolist(),
                  "alpha_n_min": [m.alpha_n_min for m in models]}))
