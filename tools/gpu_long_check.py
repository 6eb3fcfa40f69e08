"""long-horizon wind+battery LPs: the stage kernel's long variant (one warp per LP, workspace) vs the band kernel -- parity and timing,
up to the reference's full-year n_time_points = 8736 sweep (64 design points)"""
import sys, json, time
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
p = SC.pool()
out = {}
def timed(sol, cp, rp, reps=3):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return o, float(np.median(ts))
for T, N, band_too in ((168, 2000, True), (672, 256, True), (2184, 64, True), (8736, 64, False), (8736, 1, True)):
    t = TP.wind_battery(T)
    lam, cf = p["dalmp_303"][:T], p["dacf_303"][:T]
    rng = np.random.default_rng(T)
    lmp = np.tile(lam, (N, 1)) * (rng.lognormal(0, 0.1, (N, T)) if N > 1 else 1.0)
    wind = np.repeat(np.linspace(200, 1600, 8), 8)[:N] if N == 64 else np.full(N, 847.0)
    batt = (np.tile(np.linspace(0.05, 1.0, 8), 8)[:N] * wind) if N == 64 else np.full(N, 211.75)
    rp = TP.wind_battery_rparams(T, np.tile(cf, (N, 1)), wind, batt)
    cpd = torch.tensor(lmp, device=dev); rpd = torch.tensor(rp, device=dev)
    s2 = S.BatchLPSolver(t, kernel=S.KERNEL_STAGE)
    a, ms = timed(s2, cpd, rpd, 2); la = S.last_launch()
    rec = dict(stage_long_ms=ms, launch=la, non_optimal=int((a.status != 0).sum()), iters=float(a.iters.float().mean()), iters_max=int(a.iters.max()))
    if band_too:
        sb = S.BatchLPSolver(t, kernel=S.KERNEL_BAND)
        b, msb = timed(sb, cpd, rpd, 1)
        rec.update(band_ms=msb, speedup=msb / ms, band_non_optimal=int((b.status != 0).sum()),
                   rel=float(((a.obj - b.obj).abs() / b.obj.abs().clamp(min=1)).max()), iters_band=float(b.iters.float().mean()))
        sb.close()
    out[f"T{T}_N{N}"] = rec
    print(T, N, json.dumps(rec), flush=True)
    s2.close()
json.dump(out, open("gpurun_out/long_check.json", "w"), indent=1)
