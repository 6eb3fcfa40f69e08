"""C2 + C5-slice timing of the stage kernels of the library named by DSP_LP_LIB, parity of generation 2 against generation 1"""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
t = TP.wind_battery(24)
v2 = S.BatchLPSolver(t, kernel=S.KERNEL_STAGE); v1 = S.BatchLPSolver(t, kernel=S.KERNEL_STAGE_V1)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
cp = torch.tensor(lmp, device=dev); rpt = torch.tensor(rp, device=dev)
l5, c5, w5, b5 = SC.c5(8, 8, 2000)
cp5 = torch.tensor(l5, device=dev); rp5 = torch.tensor(TP.wind_battery_rparams(24, c5, w5, b5), device=dev)
def timed(sol, cp, rp, reps=9):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return o, float(np.median(ts))
a, ms = timed(v2, cp, rpt); la = S.last_launch(); a5, ms5 = timed(v2, cp5, rp5, 3)
b = v1.solve(cp, rpt); torch.cuda.synchronize()
rel = float(((a.obj - b.obj).abs() / b.obj.abs().clamp(min=1)).max())
print("%s: C2 %.3f ms (nonopt %d, iters %.2f, rel vs v1 %.1e)  C5/128k %.3f ms (nonopt %d) %s" % (
    os.path.basename(os.environ.get("DSP_LP_LIB", "default")), ms, int((a.status != 0).sum()), float(a.iters.float().mean()), rel, ms5,
    int((a5.status != 0).sum()), la), flush=True)
