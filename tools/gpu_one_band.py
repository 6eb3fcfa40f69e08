"""one band-kernel workload for an ncu capture: day-ahead bidder T=48 (hybrid placement), 2368 LPs = 16 per SM"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0"); rng = np.random.default_rng(3)
T = 48; N = 2368; cf = np.tile(SC.c2(1)[1], 2)
da = rng.uniform(5, 80, (N, T)); rt = da + rng.normal(0, 10, (N, T))
sol = S.BatchLPSolver(TP.wind_battery_operation(T, "bidder_da"))
cp = torch.tensor(np.concatenate([da, rt, np.full((N, 1), 1e3)], 1), device=dev)
rp = torch.tensor(np.repeat(TP.wind_battery_operation_rparams(T, cf, 200, 25, 100), N, 0), device=dev)
for _ in range(2):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
print(S.last_launch(), int((o.status != 0).sum()), float(o.iters.float().mean()))
