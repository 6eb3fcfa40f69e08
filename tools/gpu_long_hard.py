"""the full-year design sweep of tools/gpu_long_check.py (64 LPs, T = 8736): long stage kernel vs band kernel, per-LP status / iterations"""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
p = SC.pool()
T, N = 8736, 64
t = TP.wind_battery(T)
lam, cf = p["dalmp_303"][:T], p["dacf_303"][:T]
rng = np.random.default_rng(T)
lmp = np.tile(lam, (N, 1)) * rng.lognormal(0, 0.1, (N, T))
wind = np.repeat(np.linspace(200, 1600, 8), 8); batt = np.tile(np.linspace(0.05, 1.0, 8), 8) * wind
rp = TP.wind_battery_rparams(T, np.tile(cf, (N, 1)), wind, batt)
cpd = torch.tensor(lmp, device=dev); rpd = torch.tensor(rp, device=dev)
res = {}
for name, kern in (("stage_long", S.KERNEL_STAGE), ("band", S.KERNEL_BAND)):
    sol = S.BatchLPSolver(t, kernel=kern)
    o = sol.solve(cpd, rpd); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); o = sol.solve(cpd, rpd, out=o); e1.record(); torch.cuda.synchronize()
    res[name] = dict(ms=e0.elapsed_time(e1), status=o.status.cpu().numpy().tolist(), iters=o.iters.cpu().numpy().tolist(), obj=o.obj.cpu().numpy().tolist())
    print(name, "%.1f ms" % res[name]["ms"], "non-optimal", [k for k, s in enumerate(res[name]["status"]) if s != 0], "iters", res[name]["iters"], flush=True)
a, b = np.array(res["stage_long"]["obj"]), np.array(res["band"]["obj"])
print("max rel diff", float(np.max(np.abs(a - b) / np.maximum(1, np.abs(b)))))
json.dump(res, open("gpurun_out/long_hard.json", "w"))
