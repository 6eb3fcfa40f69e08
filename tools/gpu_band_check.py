import sys; sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
t = TP.wind_battery(24)
band = S.BatchLPSolver(t, kernel=S.KERNEL_BAND)
lmp, cf, W, P = SC.c2(64); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
r = band.solve_host(lmp, rp, want_x=True, want_y=True)
print("band", S.last_launch(), np.bincount(r.status), r.iters.mean())
tn = TP.nuclear(48); sn = S.BatchLPSolver(tn); rn = sn.solve_host(SC.c3(32), None); print("nuclear", np.bincount(rn.status), rn.iters.mean())
tp = TP.wind_battery_pem(24); sp = S.BatchLPSolver(tp)
rpp = TP.wind_battery_rparams(24, cf, W, 150.0, pem_mw=200.0)[0]
rr = sp.solve_host(np.concatenate([lmp, np.full((64, 1), 2.5)], axis=1), rpp); print("pem w=8", np.bincount(rr.status), rr.iters.mean())
