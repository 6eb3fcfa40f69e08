"""CPU study (test infrastructure: uses the oracle's numpy mirror for the iteration counts): event-driven simulation of the stage
kernel's schedule -- 148 SMs x 8 warps x 4 LP groups, one global ticket counter, CTA-synchronised rounds -- on the eight per-rank C2
batches of bench.py, for the batch as given and for tickets ordered by max |price| descending.

    python tests/checks/sim_schedule.py            # ~10 min on 8 cores (the mirror solves 8 x 10 000 LPs)

Round time of an SM: t2 if two active warps share a scheduler, else t1.  Findings of round 2: the eight batches
have the same mean iteration count (11.75-11.80) but their simulated kernel times spread 556-608 us, because a 10 000-LP batch is
2.1 waves of the 4 736 group slots and the step ends with the longest LP of the thin last wave; LPs with the highest price spikes
take ~14 iterations (max 20-23) against 11.8 on average, so handing those out FIRST shortens the tail: 534-549 us for every batch
(max over ranks -10 %, mean -4 %); an oracle longest-first order would give 479-505 us."""
import heapq
import sys
from multiprocessing import Pool
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))


def simulate(iters, t2=15.3, t1=11.0, n_sm=148, warps=8, gpw=4):
    N, nxt = len(iters), 0
    rem = np.zeros((n_sm, warps, gpw), int)
    for w in range(warps):
        for g in range(gpw):
            for s in range(n_sm):
                if nxt < N:
                    rem[s, w, g] = iters[nxt]; nxt += 1
    heap = [(0.0, s) for s in range(n_sm)]
    heapq.heapify(heap)
    end = 0.0
    while heap:
        t, s = heapq.heappop(heap)
        r = rem[s]
        act = (r > 0).any(axis=1)
        if not act.any():
            end = max(end, t)
            continue
        t += t2 if (act[:4].astype(int) + act[4:].astype(int)).max() >= 2 else t1
        r[r > 0] -= 1
        for w in range(warps):
            for g in range(gpw):
                if r[w, g] == 0 and nxt < N:
                    r[w, g] = iters[nxt]; nxt += 1
        heapq.heappush(heap, (t, s))
    return end


def rank_batch(rank):
    from dispatches_b200 import scenarios as SC, templates as TP
    from oracle import ipm_stage_numpy as ST
    st = TP.wind_battery(24).meta["stage_wb"]
    consts = {k: st[k] for k in ("a", "binv", "half", "delta", "dur", "k_rev")}
    lmp, cf, W, P = SC.c2(10000, seed=20240101 + rank)
    it = ST.solve_batch(lmp, W * 1e3 * cf, P * 1e3, consts)["iters"]
    order = np.argsort(-np.abs(lmp).max(1), kind="stable")
    return (rank, float(it.mean()), int(it.max()), simulate(it), simulate(it[order]), simulate(np.sort(it)[::-1]))


if __name__ == "__main__":
    with Pool(8) as pool:
        rows = sorted(pool.map(rank_batch, range(8)))
    print("rank  iters mean/max   as given   max|price| first   longest first (oracle)   [us]")
    for r in rows:
        print("%4d  %6.2f / %2d   %8.0f   %16.0f   %22.0f" % r)
    print("max over ranks: %.0f -> %.0f us" % (max(r[3] for r in rows), max(r[4] for r in rows)))
