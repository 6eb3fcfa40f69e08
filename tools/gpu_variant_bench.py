"""times the stage kernel of the library named by DSP_LP_LIB on C2 and a C5 slice; checks against the band kernel"""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
T = 24
t = TP.wind_battery(T)
stage = S.BatchLPSolver(t, kernel=S.KERNEL_STAGE); band = S.BatchLPSolver(t, kernel=S.KERNEL_BAND)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(T, cf, W, P)[0]
dev = torch.device("cuda:0")
cp = torch.tensor(lmp, device=dev); rpt = torch.tensor(rp, device=dev)
a = stage.solve(cp, rpt); b = band.solve(cp, rpt); torch.cuda.synchronize()
ao, bo = a.obj.cpu().numpy(), b.obj.cpu().numpy()
ok = (a.status.cpu().numpy() == 0).all()
ts = []
for rep in range(7):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); stage.solve(cp, rpt, out=a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
lmp5, cf5, w5, b5 = SC.c5(8, 8, 2000)
cp5 = torch.tensor(lmp5, device=dev); rp5 = torch.tensor(TP.wind_battery_rparams(T, cf5, w5, b5), device=dev)
o5 = stage.solve(cp5, rp5); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); stage.solve(cp5, rp5, out=o5); e1.record(); torch.cuda.synchronize()
ms5 = e0.elapsed_time(e1)
print(os.environ.get("DSP_LP_LIB", "default"), S.last_launch(), "C2 10k: %.3f ms (%.2fM LP/s) | C5 128k: %.3f ms (%.2fM LP/s) | all optimal %s, c5 nonopt %d, max rel diff vs band %.1e"
      % (np.median(ts), 10 / np.median(ts), ms5, len(lmp5) / ms5 / 1e3, ok, int((o5.status != 0).sum()), (np.abs(ao - bo) / np.maximum(1, np.abs(bo))).max()))
