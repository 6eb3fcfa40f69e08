import sys; sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
from oracle import highs as H
t = TP.wind_battery(24)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
nref = 1500
ref, _, _ = H.solve_batch("wind_battery", lmp[:nref], kwargs=dict(cf=cf, wind_mw=W, batt_mw=P))
c0, b, u, k = t.instantiate(lmp[0], rp); scale = np.abs(b).max(); C = (t.Cmap @ lmp.T).T + t.c0
for tol, ft in ((1e-8, 1e-8), (1e-8, 1e-9), (1e-8, 1e-10), (1e-9, 1e-10), (1e-9, 1e-11)):
    sol = S.BatchLPSolver(t, tol=tol, feas_tol=ft)
    r = sol.solve_host(lmp, rp, want_x=True, want_y=True)
    err = np.abs(r.obj[:nref] - ref) / np.maximum(1, np.abs(ref))
    rc = C - r.y @ t.A; ueff = np.where(np.isfinite(u), u, 10 * scale)
    lower = r.y @ b + (np.minimum(rc, 0) * ueff).sum(1) + k
    gap = (r.obj - lower) / np.maximum(1, np.abs(r.obj))
    print(tol, ft, "status", np.bincount(r.status), "iters %.2f max %d" % (r.iters.mean(), r.iters.max()),
          "err max %.2e" % err.max(), "cert gap max %.2e" % gap.max(), "pfeas %.1e" % (np.abs(r.x @ t.A.T - b).max() / scale), flush=True)
