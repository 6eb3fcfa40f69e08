"""a single C2 launch of the stage kernel (generation 2 by default; DSP_KERNEL=3 for generation 1) for `ncu --set full`"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
t = TP.wind_battery(24)
sol = S.BatchLPSolver(t, kernel=int(os.environ.get("DSP_KERNEL", S.KERNEL_STAGE)))
lmp, cf, W, P = SC.c2(10000); rp = TP.wind_battery_rparams(24, cf, W, P)[0]
dev = torch.device("cuda:0")
cp = torch.tensor(lmp, device=dev); rpt = torch.tensor(rp, device=dev)
o = sol.solve(cp, rpt); torch.cuda.synchronize()
o = sol.solve(cp, rpt, out=o); torch.cuda.synchronize()
print(S.last_launch(), int((o.status != 0).sum()), float(o.iters.float().mean()))
