import sys, time; sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
from oracle import highs as H, lp_models as L
p = SC.pool(); lam, cfs = p["dalmp_303"], p["dacf_303"]
for T in (672, 2184, 8736):
    t = TP.wind_battery(T)
    sol = S.BatchLPSolver(t, max_iter=120)
    rp = TP.wind_battery_rparams(T, cfs[:T], 847.0, 211.75)[0]
    t0 = time.perf_counter(); r = sol.solve_host(lam[None, :T], rp); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); ref = H.solve(L.wind_battery_raw(lam[:T], cfs[:T], 847.0, 211.75))[0]; dto = time.perf_counter() - t0
    print("T=%d: gpu %.2f s status %s iters %s obj %.8f | oracle %.1f s obj %.8f rel %.2e" % (T, dt, r.status, r.iters, r.obj[0], dto, ref, abs(r.obj[0] - ref) / max(1, abs(ref))), flush=True)
