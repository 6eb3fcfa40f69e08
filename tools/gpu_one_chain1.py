"""a single C3 launch (nuclear T = 48, 5 000 LPs) of the chain stage kernel for `ncu --set full`"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
sol = S.BatchLPSolver(TP.nuclear(48))
cp = torch.tensor(SC.c3(5000), device="cuda")
o = sol.solve(cp, None); torch.cuda.synchronize()
o = sol.solve(cp, None, out=o); torch.cuda.synchronize()
print(S.last_launch(), int((o.status != 0).sum()), float(o.iters.float().mean()))
