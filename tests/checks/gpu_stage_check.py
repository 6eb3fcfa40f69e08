"""Stage kernel vs band kernel vs HiGHS; timing of both kernels on C2 (dev tool)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S

small = len(sys.argv) > 1 and sys.argv[1] == "small"
T = 24
t = TP.wind_battery(T)
stage = S.BatchLPSolver(t, kernel=S.KERNEL_STAGE)
band = S.BatchLPSolver(t, kernel=S.KERNEL_BAND)
N = 64 if small else 10000
lmp, cf, W, P = SC.c2(N)
rp = TP.wind_battery_rparams(T, cf, W, P)[0]
a = stage.solve_host(lmp, rp, want_x=True, want_y=True)
print("stage launch", S.last_launch(), "status", np.bincount(a.status), "iters %.2f max %d" % (a.iters.mean(), a.iters.max()))
b = band.solve_host(lmp, rp, want_x=True, want_y=True)
print("band  launch", S.last_launch(), "status", np.bincount(b.status), "iters %.2f max %d" % (b.iters.mean(), b.iters.max()))
print("stage vs band: obj rel diff max %.2e, iters equal %.4f" % ((np.abs(a.obj - b.obj) / np.maximum(1, np.abs(b.obj))).max(), (a.iters == b.iters).mean()))
c0, bb, u, k = t.instantiate(lmp[0], rp); scale = np.abs(bb).max()
print("stage primal feas %.2e  x vs band x max diff (scaled) %.2e" % (np.abs(a.x @ t.A.T - bb).max() / scale, np.abs(a.x - b.x).max() / scale))
if small:
    sys.exit(0)
from oracle import highs as H
nref = 1500
ref, _, _ = H.solve_batch("wind_battery", lmp[:nref], kwargs=dict(cf=cf, wind_mw=W, batt_mw=P))
print("err vs HiGHS: stage %.2e band %.2e" % ((np.abs(a.obj[:nref] - ref) / np.maximum(1, np.abs(ref))).max(), (np.abs(b.obj[:nref] - ref) / np.maximum(1, np.abs(ref))).max()))
dev = torch.device("cuda:0")
cp = torch.tensor(lmp, device=dev); rpt = torch.tensor(rp, device=dev)
for name, sol in (("stage", stage), ("band", band)):
    out = sol.solve(cp, rpt); torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rpt, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%s kernel N=%d: %.3f ms (min %.3f) -> %.3e LP/s" % (name, N, np.median(ts), min(ts), N / np.median(ts) * 1e3), S.last_launch())
    t0 = time.perf_counter(); sol.solve_host(lmp, rp); dt = time.perf_counter() - t0
    print("%s host e2e %.3f ms -> %.3e LP/s" % (name, dt * 1e3, N / dt))
# bigger batch for throughput mode
lmp5, cf5, w5, b5 = SC.c5(8, 8, 1000)
rp5 = TP.wind_battery_rparams(T, cf5, w5, b5)
cp5 = torch.tensor(lmp5, device=dev); rp5t = torch.tensor(rp5, device=dev)
out = stage.solve(cp5, rp5t); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); stage.solve(cp5, rp5t, out=out); e1.record(); torch.cuda.synchronize()
st = out.status.cpu().numpy(); it = out.iters.cpu().numpy()
print("stage C5-like N=%d: %.3f ms -> %.3e LP/s; status %s iters %.2f max %d" % (len(lmp5), e0.elapsed_time(e1), len(lmp5) / e0.elapsed_time(e1) * 1e3, np.bincount(st), it.mean(), it.max()))
