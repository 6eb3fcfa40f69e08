import sys; sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
t = TP.wind_battery(24)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
sol = S.BatchLPSolver(t, tol=1e-8, feas_tol=1e-9)
r = sol.solve_host(lmp[:2000], rp, want_x=True, want_y=True)
np.savez_compressed("gpurun_out/dump_wb24.npz", obj=r.obj, iters=r.iters, status=r.status, x=r.x[:400], y=r.y[:400])
print("ok", r.iters.mean())
