"""One-off validation: the CUDA generic solver against the oracle (the reference's algorithm with LSODA) on random
points of the bench grid: root w_-, shock position, kappa, and the profile as a curve."""
import sys, warnings
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
warnings.filterwarnings("ignore")
from oracle import models as om, shell_generic as sg
from util_compare import curve_errors
from pttools_b200 import bubble
from pttools_b200.models import BagModel, ConstCSModel

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ref = sg.FluidReference(dict(np.load("tests/golden/ref_fluid_reference.npz")))
for name, pm, mo in (("bag", BagModel(alpha_n_min=0.005), None), ("cs_third_quarter", ConstCSModel(css2=1 / 3, csb2=1 / 4, a_s=5, a_b=1, V_s=1), None)):
    mo = om.bag_model(pm.a_s, pm.a_b, pm.V_s, pm.V_b) if name == "bag" else om.const_cs_model(pm.css2, pm.csb2, pm.a_s, pm.a_b, pm.V_s, pm.V_b)
    vw = np.sort(rng.uniform(0.06, 0.94, N))
    an = rng.uniform(max(0.006, 1.02 * pm.alpha_n_min), 0.5, N)
    bb = bubble.sweep_bubbles(pm, vw, an)
    st = bb.status.cpu().numpy()
    worst = dict(wm=0.0, v_sh=0.0, kappa=0.0, curve=0.0)
    n_cmp = n_disagree = 0
    for i in range(N):
        try:
            sol = sg.sound_shell_generic(mo, vw[i], an[i], ref)
            ok_ref = sol["v"].size > 3 and not sol.get("solver_failed", False)
        except Exception:
            ok_ref = False
        ok_gpu = st[i] == 0
        if ok_ref != ok_gpu:
            n_disagree += 1
            print(f"  {name}: validity differs at v_wall={vw[i]:.5f} alpha_n={an[i]:.5f}: gpu status {st[i]}, reference ok={ok_ref}")
            continue
        if not ok_ref:
            continue
        n_cmp += 1
        v, w, xi = bb.shells.profile(i)
        ev, ew = curve_errors(xi, v, w, sol["xi"], sol["v"], sol["w"], vw[i])
        worst["curve"] = max(worst["curve"], ev, ew)
        worst["wm"] = max(worst["wm"], abs(float(bb.scal["wm"][i]) / sol["wm"] - 1))
        worst["v_sh"] = max(worst["v_sh"], abs(float(bb.scal["v_sh"][i]) / sol["v_sh"] - 1))
        bs = sg.bubble_scalars(mo, sol, vw[i])
        worst["kappa"] = max(worst["kappa"], abs(float(bb.scal["kappa"][i]) / bs["kappa"] - 1))
    print(name, "compared", n_cmp, "validity disagreements", n_disagree, "worst relative deviations", worst)
