"""compute-sanitizer target: small batches through every kernel family (stage kernel with x/y write-back, band kernel in
shared-memory / hybrid / workspace placement).  Run as
    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/gpu_sanitize.py > profiles/sanitizer_r2.log
"""
import sys
sys.path.insert(0, ".")
import numpy as np
from dispatches_b200 import templates as TP, scenarios as SC, solver as S

lmp, cf, W, P = SC.c2(192)
rp = TP.wind_battery_rparams(24, cf, W, P)[0]
t = TP.wind_battery(24)
for kern, name in ((S.KERNEL_STAGE, "stage2<8,3>"), (S.KERNEL_STAGE_V1, "stage v1"), (S.KERNEL_BAND, "band<4> smem")):
    r = S.BatchLPSolver(t, kernel=kern).solve_host(lmp, rp, want_x=True, want_y=True)
    print(name, "optimal", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
l5, c5, w5, b5 = SC.c5(2, 2, 40)
r = S.BatchLPSolver(t).solve_host(l5, TP.wind_battery_rparams(24, c5, w5, b5))
print("stage, rhs batched", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
r = S.BatchLPSolver(TP.nuclear(48)).solve_host(SC.c3(64), None, want_x=True, want_y=True)
print("chain1<16,3,2> nuclear", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
pp = SC.pool()
for T in (12, 48, 96):
    r = S.BatchLPSolver(TP.wind_battery(T)).solve_host(np.tile(pp["dalmp_303"][:T], (40, 1)) * np.random.default_rng(T).lognormal(0, .2, (40, T)),
                                                       TP.wind_battery_rparams(T, pp["dacf_303"][:T], 847.0, 211.75)[0], want_x=True)
    print(f"stage2 T={T}", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
r = S.BatchLPSolver(TP.nuclear(48), kernel=S.KERNEL_BAND).solve_host(SC.c3(64), None, want_x=True)
print("band<1> nuclear", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
r = S.BatchLPSolver(TP.fossil_surrogate(168)).solve_host(SC.c4(24), None)
print("band fossil T=168", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
p = SC.pool()
T = 672
r = S.BatchLPSolver(TP.wind_battery(T)).solve_host(np.tile(p["dalmp_303"][:T], (2, 1)), TP.wind_battery_rparams(T, p["dacf_303"][:T], 847.0, 211.75)[0])
print("long-horizon stage kernel T=672 (+ retry pass of the band kernel)", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
r = S.BatchLPSolver(TP.wind_battery(T), kernel=S.KERNEL_BAND).solve_host(np.tile(p["dalmp_303"][:T], (2, 1)), TP.wind_battery_rparams(T, p["dacf_303"][:T], 847.0, 211.75)[0])
print("band ws T=672", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
ts = TP.solar_battery_hydrogen(24, batt_mw=50.0, batt_mwh=200.0, pem_mw=20.0)
cfs = np.sin(np.deg2rad(np.linspace(0, 180, 24)))
r = S.BatchLPSolver(ts).solve_host(np.abs(lmp[:16]) * 0.5, TP.solar_rparams(24, cfs, 200.0, np.full(24, 100.0))[0], want_x=True, want_y=True)
print("band<16> solar (cyclic)", int((r.status == 0).sum()), "/", len(r.status), S.last_launch())
