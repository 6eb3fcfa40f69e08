"""Long horizons on the band kernel (global-workspace mode): full-year LPs of the reference's sweeps."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S, pricetaker as PT
from oracle import highs as H, lp_models as L
p = SC.pool()
gold = json.load(open("tests/golden/wind_pem_golden.json"))["wind_PEM_RT_1000"]
# 1. wind + PEM, no battery, 8784 periods: the reference's committed results table
lmp, cf = p["pq1000_rt_lmp"], p["pq1000_rt_cf"]
rows = [1, 4, 8]
params = {"wind_mw": 847.0, "batt_mw": 0.0, "pem_mw": np.array([gold["pem_mw"][r] for r in rows]),
          "h2_price_per_kg": np.array([gold["h2_price_per_kg"][r] for r in rows]), "design_opt": False, "extant_wind": True,
          "wind_resource": np.tile(cf, (3, 1)), "DA_LMPs": np.tile(lmp, (3, 1))}
t0 = time.perf_counter(); res = PT.wind_battery_pem_optimize(8784, params, want_solution=True); dt = time.perf_counter() - t0
print("wind+PEM T=8784 x3: %.2f s, status %s iters %s launch %s" % (dt, res.status, res.iters, S.last_launch()))
for k, r in enumerate(rows):
    print("   row %d NPV gpu %.2f  reference csv %.2f  rel %.2e   annual_rev_h2 rel %.2e" % (r, res.NPV[k], gold["NPV"][r], abs(res.NPV[k] - gold["NPV"][r]) / abs(gold["NPV"][r]),
          abs(res.annual_rev_h2[k] - gold["annual_rev_h2"][r]) / abs(gold["annual_rev_h2"][r])))
# 2. wind + battery, weekly horizon
lam, cfs = p["dalmp_303"], p["dacf_303"]
N = 16
w = np.stack([lam[k * 168:(k + 1) * 168] for k in range(N)]); c = np.stack([cfs[k * 168:(k + 1) * 168] for k in range(N)])
par = {"wind_mw": 847.0, "batt_mw": 211.75, "design_opt": False, "extant_wind": True, "wind_resource": c, "DA_LMPs": w}
t0 = time.perf_counter(); r168 = PT.wind_battery_optimize(168, par, want_solution=False); dt = time.perf_counter() - t0
ref = np.array([H.solve(L.wind_battery_raw(w[k], c[k], 847.0, 211.75))[0] for k in range(N)])
print("wind+battery T=168 x16: %.2f s status %s iters max %d err %.2e" % (dt, np.bincount(r168.status), r168.iters.max(), (np.abs(r168.obj - ref) / np.maximum(1, np.abs(ref))).max()), S.last_launch())
# 3. wind + battery, full year (run_pricetaker_wind_battery.run_design with n_time_points = 8736)
par = {"wind_mw": 847.0, "batt_mw": np.array([84.7, 211.75]), "design_opt": False, "extant_wind": True,
       "wind_resource": np.tile(cfs, (2, 1)), "DA_LMPs": np.tile(lam, (2, 1))}
t0 = time.perf_counter(); ry = PT.wind_battery_optimize(8736, par, want_solution=False); dt = time.perf_counter() - t0
print("wind+battery T=8736 x2: %.2f s status %s iters %s" % (dt, ry.status, ry.iters), S.last_launch())
t0 = time.perf_counter(); refy = H.solve(L.wind_battery_raw(lam, cfs, 847.0, 211.75))[0]; dto = time.perf_counter() - t0
print("   oracle (HiGHS, 1 LP) %.1f s; obj gpu %.6f oracle %.6f rel %.2e" % (dto, ry.obj[1], refy, abs(ry.obj[1] - refy) / max(1, abs(refy))))
