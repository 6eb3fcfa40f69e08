"""C2 timing of the generation-2 stage kernel over warps-per-CTA and CTA-synchronised on/off (+ generation 1 for reference)"""
import sys, os, json
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

def timed(sol, cp, rp, reps=7):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return o, float(np.median(ts))

o, ms = timed(v1, cp, rpt); o5, ms5 = timed(v1, cp5, rp5, 3)
print("v1: C2 %.3f ms   C5/128k %.3f ms" % (ms, ms5), flush=True)
for sync in (1, 0):
    for w in (7, 6, 5, 4, 3, 2):
        os.environ["DSP_STAGE2_SYNC"] = str(sync); os.environ["DSP_STAGE2_WARPS"] = str(w)
        o, ms = timed(v2, cp, rpt); o5, ms5 = timed(v2, cp5, rp5, 3)
        print("v2 sync=%d warps=%d: C2 %.3f ms (nonopt %d)  C5/128k %.3f ms (nonopt %d) %s" % (sync, w, ms, int((o.status != 0).sum()), ms5, int((o5.status != 0).sum()), S.last_launch()), flush=True)
