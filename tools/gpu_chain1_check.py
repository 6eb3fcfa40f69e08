"""C3 (nuclear T=48, 5000 LPs): descriptor-driven chain stage kernel vs the band kernel -- parity and timing"""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
def timed(sol, cp, rp, reps=9):
    o = sol.solve(cp, rp); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return o, float(np.median(ts))
out = {}
for T, N in ((48, 5000), (48, 50000), (24, 20000), (96, 5000)):
    t = TP.nuclear(T)
    c1, band = S.BatchLPSolver(t), S.BatchLPSolver(t, kernel=S.KERNEL_BAND)
    p = SC.pool()["cluster_days"]; rng = np.random.default_rng(T)
    days = rng.integers(0, len(p) - 4, N)
    lmp = np.stack([np.concatenate([p[k + i] for i in range(4)])[:T] for k in days]) * rng.lognormal(0, 0.25, (N, T))
    cp = torch.tensor(lmp, device=dev)
    a, ms = timed(c1, cp, None); la = S.last_launch(); b, msb = timed(band, cp, None, 3)
    rel = float(((a.obj - b.obj).abs() / b.obj.abs().clamp(min=1)).max())
    out[f"nuclear_T{T}_N{N}"] = dict(chain1_ms=ms, band_ms=msb, speedup=msb / ms, lps=N / ms * 1e3, non_optimal=int((a.status != 0).sum()),
                                    iters=float(a.iters.float().mean()), iters_equal=float((a.iters == b.iters).float().mean()), rel=rel, launch=la)
    print(json.dumps(out[f"nuclear_T{T}_N{N}"]), flush=True)
json.dump(out, open("gpurun_out/chain1_check.json", "w"), indent=1)
