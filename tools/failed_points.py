"""Lists the points of the bench grid whose shell solve does not end with status 0/1 (solved / no solution):
candidates for a comparison with the reference's validity flags.  usage: failed_points.py [n_batches] [sync|off]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
from pttools_b200 import bubble, fluid_reference
from pttools_b200.models import BagModel

n_batches = int(sys.argv[1]) if len(sys.argv) > 1 else 8
RETRY = sys.argv[2] if len(sys.argv) > 2 else "sync"
model = BagModel(alpha_n_min=0.005)
fluid_reference.ref()
out, total = [], 0
for s in range(n_batches):
    vw, an = bench.batch_points(s, 0, 1, 64000)
    bb = bubble.sweep_bubbles(model, vw, an, thermo=False, retry=RETRY)
    st = bb.shells.status.cpu().numpy()
    sol = bb.sol_type.cpu().numpy() if hasattr(bb, "sol_type") else np.full(st.shape, -9)
    sc = bb.shells.scalars.cpu().numpy() if hasattr(bb.shells, "scalars") else None
    total += st.size
    for i in np.nonzero((st != 0) & (st != 1))[0]:
        out.append({"v_wall": float(vw[i]), "alpha_n": float(an[i]), "status": int(st[i]), "sol_type": int(sol[i]),
                    "wn": float(sc[i, 11]) if sc is not None else None,
                    "wn_estimate": float(sc[i, 12]) if sc is not None else None})
print(json.dumps({"points": total, "failed": out}))
