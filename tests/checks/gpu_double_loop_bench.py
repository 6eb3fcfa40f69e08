"""throughput of the double-loop LP templates on one GPU (band kernel), with objective parity on a sample vs HiGHS"""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
from dispatches_b200 import templates as TP, solver as S
from oracle import double_loop as DL, highs as H
g = json.load(open("tests/golden/double_loop_golden.json")); CF = np.array(g["cf_309_rt"])
rng = np.random.default_rng(11)
dev = torch.device("cuda:0")
out = []
for mode, T, N in (("tracker", 4, 20000), ("tracker", 24, 10000), ("bidder_da", 48, 10000), ("bidder_rt", 4, 20000)):
    t = TP.wind_battery_operation(T, mode)
    sol = S.BatchLPSolver(t)
    wind = rng.uniform(100, 400, N); batt = rng.uniform(5, 60, N)
    cf = np.array([np.roll(np.tile(CF, 2), -int(k))[:T] for k in rng.integers(0, 48, N)])
    soc0 = np.round(rng.uniform(0, 1, N) * batt * 4e3 * 0.9, 2); thr0 = np.round(rng.uniform(0, 1e5, N), 2)
    sig = rng.uniform(0, 1, (N, T)) * wind[:, None] * 0.5
    rp = TP.wind_battery_operation_rparams(T, cf, wind, batt, 4 * batt, soc0, thr0, sig)
    da = rng.uniform(5, 80, (N, T)); rt = np.where(rng.uniform(size=(N, T)) < 0.3, da, da + rng.normal(0, 10, (N, T)))
    cp = np.full((N, 1), 1e3) if mode == "tracker" else np.concatenate([da, rt, np.full((N, 1), 1e3)], 1)
    cpd, rpd = torch.tensor(cp, device=dev), torch.tensor(rp, device=dev)
    r = sol.solve(cpd, rpd); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); sol.solve(cpd, rpd, out=r); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    obj = r.obj.cpu().numpy(); st = r.status.cpu().numpy(); it = r.iters.cpu().numpy()
    err = 0.0
    for k in rng.choice(N, 40, replace=False):
        if mode == "tracker":
            ref, _ = H.solve(DL.tracker_raw(sig[k], cf[k], wind[k], batt[k], 4 * batt[k], soc0[k], thr0[k]))
        else:
            ref, _ = H.solve(DL.bidder_raw(da[k], rt[k], cf[k], wind[k], batt[k], 4 * batt[k], soc0[k], thr0[k],
                                           da_dispatch=(sig[k] if mode == "bidder_rt" else None)))
            if mode == "bidder_rt":
                ref = ref + float(np.sum((da[k] - rt[k]) * sig[k]))
        err = max(err, abs(obj[k] - ref) / max(1.0, abs(ref)))
    line = dict(mode=mode, T=T, N=N, m=t.m, n=t.n, w=t.w, ms=float(np.median(ts)), lps_per_s=N / np.median(ts) * 1e3,
                non_optimal=int((st != 0).sum()), iters_mean=float(it.mean()), iters_max=int(it.max()), max_rel_err_vs_highs_sample=err,
                launch=S.last_launch())
    print(json.dumps(line)); out.append(line)
json.dump(out, open("gpurun_out/double_loop_bench.json", "w"), indent=1)
