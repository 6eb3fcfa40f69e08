"""Workspace placement of the band kernel with the register-window sweeps (dsp_band.cuh) vs the in-place sweeps (DSP_BAND_NO_RW=1):
C4 (fossil surrogate T = 168), the wind+battery band path at T = 168 / 672, the cyclic wind+PEM template at T = 168 / 2184 --
time, statuses, and bit-level agreement of the objectives of the two variants"""
import sys, os, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
p = SC.pool()
def timed(sol, cp, rp, reps=2):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return o, float(np.median(ts))
cases = []
t = TP.fossil_surrogate(168)
cases.append(("C4_fossil_T168_N2000", t, SC.c4(2000), None, S.KERNEL_AUTO))
for T, N in ((168, 2000), (672, 256)):
    t = TP.wind_battery(T)
    rng = np.random.default_rng(T)
    st = rng.integers(0, 8736 - T, N)
    lmp = np.stack([p["dalmp_303"][s:s + T] for s in st]) * rng.lognormal(0, 0.25, (N, T))
    cfs = np.stack([p["dacf_303"][s:s + T] for s in st])
    cases.append((f"wind_battery_band_T{T}_N{N}", t, lmp, TP.wind_battery_rparams(T, cfs, rng.uniform(200, 1600, N), rng.uniform(10, 800, N)), S.KERNEL_BAND))
for T, N in ((168, 1000), (2184, 32)):
    t = TP.wind_battery_pem(T)
    rng = np.random.default_rng(T + 1)
    st = rng.integers(0, 8736 - T, N)
    lmp = np.stack([p["dalmp_303"][s:s + T] for s in st]) * rng.lognormal(0, 0.25, (N, T))
    cfs = np.stack([p["dacf_303"][s:s + T] for s in st])
    cp = np.concatenate([lmp, np.full((N, 1), 2.5)], axis=1)
    cases.append((f"wind_battery_pem_T{T}_N{N}", t, cp, TP.wind_battery_rparams(T, cfs, 847.0, 100.0, pem_mw=200.0), S.KERNEL_AUTO))
out = {}
for name, t, cp, rp, kern in cases:
    sol = S.BatchLPSolver(t, kernel=kern)
    cpd = torch.tensor(cp, device=dev); rpd = None if rp is None else torch.tensor(rp, device=dev)
    o, ms = timed(sol, cpd, rpd)
    out[name] = dict(ms=ms, non_optimal=int((o.status != 0).sum()), iters=float(o.iters.float().mean()), launch=S.last_launch(), m=t.m, n=t.n, w=t.w,
                     obj_sum=float(o.obj.sum()), obj_hash=hash(o.obj.cpu().numpy().tobytes()) & 0xffffffff)
    print(name, json.dumps(out[name]), flush=True)
json.dump(out, open("gpurun_out/band_rw_%s.json" % ("inplace" if os.environ.get("DSP_BAND_NO_RW") else "rw"), "w"), indent=1)
