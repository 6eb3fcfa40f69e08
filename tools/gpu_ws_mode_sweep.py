"""band kernel: shared-memory work regions vs the global-workspace mode with more warps per SM (env DSP_BAND_WS_WARPS),
on the templates whose per-LP work region is large (C4 fossil T=168, bidder T=48, wind+battery T=168)"""
import sys, os, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, scenarios as SC, solver as S
dev = torch.device("cuda:0")
rng = np.random.default_rng(3)
cases = {}
if "--mid" in sys.argv:      # templates with 2-3 LPs per SM in shared memory
    T = 48; t = TP.wind_battery_operation(T, "bidder_da")
    N = 4000; cf = np.tile(SC.c2(1)[1], 2)
    da = rng.uniform(5, 80, (N, T)); rt = da + rng.normal(0, 10, (N, T))
    cases["bidder_da_T48"] = (t, torch.tensor(np.concatenate([da, rt, np.full((N, 1), 1e3)], 1), device=dev),
                              torch.tensor(np.repeat(TP.wind_battery_operation_rparams(T, cf, 200, 25, 100), N, 0), device=dev))
    T = 96; t = TP.wind_battery(T)
    cases["wind_battery_T96"] = (t, torch.tensor(np.tile(SC.c2(3000)[0], (1, 4)), device=dev),
                                 torch.tensor(TP.wind_battery_rparams(T, np.tile(SC.c2(1)[1], 4), 847.0, 211.75)[0], device=dev))
else:
    t = TP.fossil_surrogate(168); cases["C4_fossil_T168"] = (t, torch.tensor(SC.c4(2000), device=dev), None)
    T = 168; t = TP.wind_battery(T)
    lmp = np.tile(SC.c2(2000)[0], (1, 7)); cf = np.tile(SC.c2(1)[1], 7)
    cases["wind_battery_T168"] = (t, torch.tensor(lmp, device=dev), torch.tensor(TP.wind_battery_rparams(T, cf, 847.0, 211.75)[0], device=dev))
    t = TP.nuclear(48); cases["C3_nuclear_T48"] = (t, torch.tensor(SC.c3(5000), device=dev), None)
out = {}
for name, (t, cp, rp) in cases.items():
    sol = S.BatchLPSolver(t)
    ref = None
    for ws in ((0, 16) if "--mid" in sys.argv else (0, 2, 4, 8, 16)):
        if ws: os.environ["DSP_BAND_WS_WARPS"] = str(ws)
        else: os.environ.pop("DSP_BAND_WS_WARPS", None)
        o = sol.solve(cp, rp); torch.cuda.synchronize()
        ts = []
        for _ in range(2):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); sol.solve(cp, rp, out=o); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        obj = o.obj.cpu().numpy()
        if ref is None: ref = obj
        line = dict(ws_warps=ws, ms=min(ts), lps=cp.shape[0] / min(ts) * 1e3, nonopt=int((o.status != 0).sum()),
                    maxdiff=float((np.abs(obj - ref) / np.maximum(1, np.abs(ref))).max()), launch=S.last_launch())
        print(name, json.dumps(line)); out.setdefault(name, []).append(line)
json.dump(out, open("gpurun_out/ws_mode_sweep%s.json" % ("_mid" if "--mid" in sys.argv else ""), "w"), indent=1)
