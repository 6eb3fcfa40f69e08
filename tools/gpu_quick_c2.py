"""minimal C2 timing of the stage kernel of the library named by DSP_LP_LIB (round-2 variants), with a parity check vs the default build's band kernel"""
import sys, os
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
t = TP.wind_battery(24)
sf = float(os.environ.get("DSP_STEP_FRAC", "0.9995"))      # e.g. 0.99995 with the -DDSP_STAGE_START=1 build
stage = S.BatchLPSolver(t, kernel=S.KERNEL_STAGE, step_frac=sf); band = S.BatchLPSolver(t, kernel=S.KERNEL_BAND)
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
dev = torch.device("cuda:0")
cp = torch.tensor(lmp, device=dev); rpt = torch.tensor(rp, device=dev)
a = stage.solve(cp, rpt); b = band.solve(cp[:512], rpt); torch.cuda.synchronize()
ts = []
for rep in range(7):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); stage.solve(cp, rpt, out=a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ao, bo = a.obj.cpu().numpy()[:512], b.obj.cpu().numpy()
print(os.environ.get("DSP_LP_LIB", "default"), "step_frac", sf, S.last_launch(), "C2 10k: %.3f ms | iters %.2f | all optimal %s | max rel diff vs band %.1e"
      % (np.median(ts), float(a.iters.float().mean()), bool((a.status == 0).all()), (np.abs(ao - bo) / np.maximum(1, np.abs(bo))).max()))
