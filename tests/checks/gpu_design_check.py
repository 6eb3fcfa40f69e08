import sys; sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
from oracle import highs as H, lp_models as L
t = TP.wind_battery_design(24)
lmp, cf, W, P = SC.c2(48); lmp[::3] *= 40.0
rp = TP.wind_battery_rparams(24, cf, W, 0.0)[0]
sol = S.BatchLPSolver(t)
r = sol.solve_host(lmp, rp, want_x=True)
print("launch", S.last_launch(), "status", r.status, "iters", r.iters)
sols = [H.solve(L.wind_battery_raw(l, cf, W, 0.0, design_opt=True, extant_wind=True)) for l in lmp]
ref = np.array([s[0] for s in sols])
err = np.abs(r.obj - ref) / np.maximum(1, np.abs(ref)); print("err", err.max(), err)
lp0 = L.wind_battery_raw(lmp[0], cf, W, 0.0, design_opt=True, extant_wind=True)
p_ref = np.array([s[1][lp0.meta["Bc"]] for s in sols]); print("P ref", p_ref[:9]); print("P gpu", r.x[:9, t.col_names.index("blk[0].fs.battery.nameplate_power")])
