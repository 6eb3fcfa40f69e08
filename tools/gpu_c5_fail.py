import sys; sys.path.insert(0, ".")
import numpy as np, time
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
t = TP.wind_battery(24)
lmp, cf, w, b = SC.c5()
rp = TP.wind_battery_rparams(24, cf, w, b)
sol = S.BatchLPSolver(t)
t0 = time.perf_counter(); r = sol.solve_host(lmp, rp); dt = time.perf_counter() - t0
t0 = time.perf_counter(); r = sol.solve_host(lmp, rp); dt = time.perf_counter() - t0
bad = np.flatnonzero(r.status != 0)
print("C5 host e2e %.1f ms -> %.2fM LP/s; bad" % (dt * 1e3, len(lmp) / dt / 1e6), bad, r.status[bad], r.iters[bad])
np.savez("gpurun_out/c5_bad.npz", idx=bad, lmp=lmp[bad], cf=cf[bad], w=w[bad], b=b[bad])
for name, kw in (("band", dict(kernel=S.KERNEL_BAND)), ("step0.99", dict(step_frac=0.99)), ("reg1e-7", dict(reg_primal=1e-7)), ("reg1e-6", dict(reg_primal=1e-6)),
                 ("tol1e-8", dict(tol=1e-8, feas_tol=1e-8)), ("step0.9", dict(step_frac=0.9))):
    s2 = S.BatchLPSolver(t, **kw)
    r2 = s2.solve_host(lmp[bad], rp[bad])
    print(name, r2.status, r2.iters, r2.obj)
l2, cf2, W2, P2 = SC.c2(10000); rp2 = TP.wind_battery_rparams(24, cf2, W2, P2)[0]
for _ in range(3):
    t0 = time.perf_counter(); r = sol.solve_host(l2, rp2); dt = time.perf_counter() - t0
print("C2 host e2e %.3f ms -> %.2fM LP/s (chunked pipeline)" % (dt * 1e3, 10000 / dt / 1e6), S.last_launch(), np.bincount(r.status))
