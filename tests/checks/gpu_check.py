"""First-contact GPU check: small parity run vs HiGHS + a timing of the C2 batch.  (dev tool, not a test)"""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S

n_par = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_big = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
T = 24
t = TP.wind_battery(T)
sol = S.BatchLPSolver(t)
lmp, cf, wind_mw, batt_mw = SC.c2(max(n_big, n_par))
rp = TP.wind_battery_rparams(T, cf, wind_mw, batt_mw)[0]
r = sol.solve_host(lmp[:n_par], rp, want_x=True, want_y=True)
print("status", np.bincount(r.status), "iters mean %.2f max %d" % (r.iters.mean(), r.iters.max()), S.last_launch())
from oracle import lp_models as L, highs as H
ref = np.array([H.solve(L.wind_battery_raw(lmp[k], cf, wind_mw, batt_mw))[0] for k in range(n_par)])
err = np.abs(r.obj - ref) / np.maximum(1.0, np.abs(ref))
print("max rel err vs HiGHS(raw LP)", err.max(), "median", np.median(err))
if n_big:
    dev = torch.device("cuda:0")
    cp = torch.tensor(lmp[:n_big], device=dev); rpt = torch.tensor(rp, device=dev)
    out = sol.solve(cp, rpt)
    torch.cuda.synchronize()
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = sol.solve(cp, rpt, out=out); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("N=%d kernel %.3f ms  -> %.3e LP/s" % (n_big, ms, n_big / ms * 1e3), S.last_launch())
    st = out.status.cpu().numpy(); it = out.iters.cpu().numpy()
    print("status", np.bincount(st), "iters mean %.2f max %d" % (it.mean(), it.max()))
    t0 = time.perf_counter(); r2 = sol.solve_host(lmp[:n_big], rp); dt = time.perf_counter() - t0
    t0 = time.perf_counter(); r2 = sol.solve_host(lmp[:n_big], rp); dt = time.perf_counter() - t0
    print("host e2e %.3f ms -> %.3e LP/s" % (dt * 1e3, n_big / dt))
    print("obj agreement device vs host path", np.abs(out.obj.cpu().numpy() - r2.obj).max())
