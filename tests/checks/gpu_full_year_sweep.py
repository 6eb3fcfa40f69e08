"""The reference's wind+battery design sweep as it is actually run (run_pricetaker_wind_battery.py:37-58:
n_time_points = 8736, design_opt False, one LP per (wind size, battery ratio)) -- all design points in one GPU batch."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import run_pricetaker as RP, scenarios as SC, solver as S
from oracle import highs as H, lp_models as L
p = SC.pool(); lam, cf = p["dalmp_303"], p["dacf_303"]
wind = [float(w) for w in np.linspace(200, 1600, 8)]; ratio = [float(r) for r in np.linspace(0.05, 1.0, 8)]
t0 = time.perf_counter(); out = RP.run_wind_battery_sweep(wind, ratio, lam[None, :], cf, n_time_points=8736); dt = time.perf_counter() - t0
t0 = time.perf_counter(); out = RP.run_wind_battery_sweep(wind, ratio, lam[None, :], cf, n_time_points=8736); dt = time.perf_counter() - t0
bad = [d for d in out if d["termination_condition"] != "optimal"]
print("64 design points x 8736 periods: %.2f s end to end (2nd call), non-optimal %d, launch %s" % (dt, len(bad), S.last_launch()))
chk = [(0, 0), (3, 4), (7, 7)]
for (a, b) in chk:
    d = out[a * 8 + b]
    t0 = time.perf_counter(); ref = -H.solve(L.wind_battery_raw(lam, cf, wind[a], wind[a] * ratio[b]))[0] * 1e5; dto = time.perf_counter() - t0
    print("   wind %.0f MW ratio %.2f: NPV gpu %.2f oracle %.2f rel %.2e (HiGHS %.1f s)" % (wind[a], ratio[b], d["NPV"], ref, abs(d["NPV"] - ref) / abs(ref), dto))
json.dump({"seconds": dt, "design_points": 64, "T": 8736, "non_optimal": len(bad)}, open("gpurun_out/full_year_sweep.json", "w"))
