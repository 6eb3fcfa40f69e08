import sys, time; sys.path.insert(0, ".")
import numpy as np, ctypes as C, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
t = TP.wind_battery(24); sol = S.BatchLPSolver(t)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
lmp = np.ascontiguousarray(lmp); rp = np.ascontiguousarray(rp)
for _ in range(5): sol.solve_host(lmp, rp)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); sol.solve_host(lmp, rp); ts.append(time.perf_counter() - t0)
print("python solve_host: median %.3f ms min %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3))
obj = np.empty(10000); st = np.empty(10000, np.int32); it = np.empty(10000, np.int32)
vp = lambda a: a.ctypes.data_as(C.c_void_p)
args = (sol.handle, 10000, vp(lmp), vp(rp), 0, C.byref(sol.opts), vp(obj), vp(st), vp(it), None, None)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); sol.lib.dsp_lp_solve_batch_host(*args); ts.append(time.perf_counter() - t0)
print("raw C call:        median %.3f ms min %.3f" % (np.median(ts) * 1e3, min(ts) * 1e3))
dev = torch.device("cuda:0"); cp = torch.tensor(lmp, device=dev); rpt = torch.tensor(rp, device=dev)
out = sol.solve(cp, rpt); torch.cuda.synchronize()
ts = []
for _ in range(20):
    t0 = time.perf_counter(); sol.solve(cp, rpt, out=out); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("device path + sync (wall): median %.3f ms" % (np.median(ts) * 1e3))
pin = torch.empty(lmp.shape, dtype=torch.float64).pin_memory(); pin.numpy()[:] = lmp
ts = []
for _ in range(20):
    t0 = time.perf_counter(); pin.numpy()[:] = lmp; ts.append(time.perf_counter() - t0)
print("host memcpy 1.92 MB into pinned: median %.3f ms" % (np.median(ts) * 1e3))
ts = []
for _ in range(20):
    t0 = time.perf_counter(); cp.copy_(pin, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("H2D 1.92 MB pinned + sync: median %.3f ms" % (np.median(ts) * 1e3))
