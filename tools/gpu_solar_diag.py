"""diagnostic: the PV + battery + hydrogen batch of tests/test_solar_battery_hydrogen.py -- which LPs end non-optimal, under which options"""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
g = json.load(open("tests/golden/solar_golden.json"))
LMP = np.array(g["lmp_24"])
cf0 = np.sin(np.deg2rad(np.linspace(0, 180, 24))); ld0 = np.full(24, 100.0)
rng = np.random.default_rng(17)
N = 256
lmp = np.vstack([LMP[None], SC.c2(N - 1)[0] * 0.5])
load = np.vstack([ld0[None], ld0[None] * rng.uniform(0.7, 1.2, (N - 1, 24))])
cfs = np.vstack([cf0[None], cf0[None] * rng.uniform(0.5, 1.0, (N - 1, 24))])
rp = TP.solar_rparams(24, cfs, 200.0, load)
for kw in (dict(), dict(batt_mw=50.0, batt_mwh=200.0), dict(batt_mw=50.0, batt_mwh=200.0, pem_mw=20.0)):
    t = TP.solar_battery_hydrogen(24, **kw)
    for opts in (dict(), dict(reg_primal=1e-7), dict(step_frac=0.99), dict(max_iter=200)):
        sol = S.BatchLPSolver(t, **opts)
        out = sol.solve(torch.tensor(lmp, device="cuda"), torch.tensor(rp, device="cuda"))
        st = out.status.cpu().numpy(); it = out.iters.cpu().numpy()
        bad = np.nonzero(st)[0]
        print(kw, opts, "w", t.w, "launch", S.last_launch(), "non-optimal", bad.tolist(), st[bad].tolist(), it[bad].tolist(), "iters mean %.1f max %d" % (it.mean(), it.max()),
              "obj", out.obj.cpu().numpy()[bad].tolist(), flush=True)
