import numpy as np, torch, sys
sys.path.insert(0, '.')
from pttools_b200 import engine
from oracle import shell_bag as sb, ssm
vw, an = np.array([0.5, 0.9]), np.array([0.1, 0.1])
n_xi = 1000
z = np.logspace(-1, 3, 60)
npt = (n_xi, 300, 400)
shells = engine.sweep_shells_bag(vw, an, n_xi=n_xi)
plan = engine.SsmPlan(z, nt=npt[1], n_z_lookup=npt[2], nxi=n_xi, mode="bag")
pp = engine.point_params(vw, sb.CS0, 0.75, 0.75, 0.75 * an, 0.0)
a2 = engine.a2_batch(plan, shells, pp).cpu().numpy()
for i in range(2):
    prof = shells.profile(i)
    qt = plan.passes[-1].qT
    a2o = ssm.a2_e_conserving_bag(qt, vw[i], an[i], npt, profile=prof)[0]
    print('a2 rel err', np.max(np.abs(a2[i]/a2o-1)), a2[i][:3], a2o[:3])
    sdv = engine.spec_den_v_batch(plan, 0, torch.as_tensor(a2o[None,:], device='cuda'), engine.dev_f64(vw[i:i+1])).cpu().numpy()[0]
    sdvo = ssm.spec_den_v_core(plan.z_lookup, a2o, qt, vw[i], nt=npt[1])
    print('sdv rel err', np.max(np.abs(sdv/sdvo-1)), sdv[:3], sdvo[:3], sdv[-3:], sdvo[-3:])
    sd, pw, om = engine.spec_den_gw_batch(plan, torch.as_tensor(sdvo[None,:], device='cuda'))
    sdo, _ = ssm.spec_den_gw_scaled(plan.z_lookup, sdvo, z)
    sd = sd.cpu().numpy()[0]
    print('gw rel err', np.max(np.abs(sd/sdo-1)), sd[:3], sdo[:3])
