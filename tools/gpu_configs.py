"""Throughput + status of the other BASELINE configs on one GPU (C3 nuclear, C4 fossil surrogate, C5 design sweep)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
out = {}

def timed(sol, cp, rp, reps=3):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    st = o.status.cpu().numpy(); it = o.iters.cpu().numpy()
    return dict(ms=float(np.median(ts)), lps=float(cp.shape[0] / np.median(ts) * 1e3), status=np.bincount(st, minlength=3).tolist(),
                iters_mean=float(it.mean()), iters_max=int(it.max()), launch=S.last_launch())

t = TP.nuclear(48); sol = S.BatchLPSolver(t)
out["C3_nuclear_T48_N5000"] = timed(sol, torch.tensor(SC.c3(5000), device=dev), None)
t = TP.fossil_surrogate(168); sol = S.BatchLPSolver(t)
out["C4_fossil_surrogate_T168_N2000"] = timed(sol, torch.tensor(SC.c4(2000), device=dev), None, reps=2)
t = TP.wind_battery(24); sol = S.BatchLPSolver(t)
lmp, cf, w, b = SC.c5()
cp = torch.tensor(lmp, device=dev); rp = torch.tensor(TP.wind_battery_rparams(24, cf, w, b), device=dev)
out["C5_design_sweep_N560640"] = timed(sol, cp, rp)
t = TP.wind_battery_pem(24); sol = S.BatchLPSolver(t)
l2, cf2, W2, P2 = SC.c2(10000)
cpp = torch.tensor(np.concatenate([l2, np.full((10000, 1), 2.5)], axis=1), device=dev)
rpp = torch.tensor(TP.wind_battery_rparams(24, cf2, W2, 150.0, pem_mw=200.0)[0], device=dev)
out["wind_battery_pem_T24_N10000_band"] = timed(sol, cpp, rpp)
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/configs_r1.json", "w"), indent=1)
