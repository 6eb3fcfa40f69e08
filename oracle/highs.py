"""CPU ORACLE (test infrastructure -- NOT product code).

Solves oracle.lp_models.RawLP instances with scipy's bundled HiGHS.  The reference hands the same
LPs to CBC (wind_battery_LMP.py:266-267, wind_battery_PEM_LMP.py:296-298), Gurobi
(nuclear_case/report/price_taker_analysis.py:365,403) or IPOPT; none of those is installable here
(SURVEY.md §8c).  Any exact LP solver returns the same optimal objective, so HiGHS dual simplex
(vertex solution, ~1e-16 relative objective agreement with HiGHS-IPM) is the accuracy oracle and the
timed CPU baseline ("port").
"""
from __future__ import annotations

import dataclasses
import multiprocessing as mp
import os
import time

import numpy as np
from scipy.optimize import linprog

from . import lp_models


# HiGHS' default 1e-7 feasibility tolerances leave objective errors up to 2.5e-6 relative on these LPs (columns
# reach 1e6 kW); with 1e-10 (the tightest HiGHS accepts) dual simplex, HiGHS-IPM and the reduced template agree to
# 1e-11.  The oracle is the accuracy reference, so it runs tight.
TIGHT = dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10)


def host_cores():
    """cores this process may really use: scheduler affinity capped by the cgroup CPU quota (containers often expose
    all host CPUs in the affinity mask but throttle to a few through cpu.max)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        txt = open("/sys/fs/cgroup/cpu.max").read().split()
        if txt[0] != "max":
            n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(q / p + 0.5)))
        except Exception:
            pass
    return n


def solve(lp: lp_models.RawLP, method="highs-ds"):
    bounds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(lp.lb, lp.ub)]
    kw = dict(A_ub=lp.A_ub if lp.A_ub.shape[0] else None, b_ub=lp.b_ub if lp.A_ub.shape[0] else None,
              A_eq=lp.A_eq if lp.A_eq.shape[0] else None, b_eq=lp.b_eq if lp.A_eq.shape[0] else None, bounds=bounds)
    # tight dual simplex first; HiGHS occasionally mis-reports a badly scaled LP (fossil surrogate, 1e6-kg inventories)
    # as unbounded under 1e-10 tolerances -> fall back to its interior point, then to default tolerances
    for m, opts in ((method, TIGHT), ("highs-ipm", TIGHT), (method, {})):
        res = linprog(lp.c, method=m, options=dict(opts), **kw)
        if res.status == 0:
            return float(res.fun + lp.c0), res.x
    raise RuntimeError(f"HiGHS status {res.status}: {res.message}")


# ---- batched loop used as the CPU baseline: the constraint data are built once per worker, only the
# ---- cost vector (and, for design sweeps, rhs/bounds) is swapped per LP -- a generous baseline, the
# ---- reference rebuilds the whole Pyomo model per LP.
_W = {}


def _init_worker(kind, kwargs):
    os.environ["OMP_NUM_THREADS"] = "1"
    _W["kind"], _W["kwargs"] = kind, kwargs


def _build(kind, lmp, extra, kwargs):
    if kind == "wind_battery":
        if extra is not None:           # per-LP design point / capacity factors (design sweeps)
            cf, wind_mw, batt_mw = extra
            kw = {k: v for k, v in kwargs.items() if k not in ("cf", "wind_mw", "batt_mw")}
            return lp_models.wind_battery_raw(lmp, cf, wind_mw, batt_mw, **kw)
        return lp_models.wind_battery_raw(lmp, **kwargs)
    if kind == "nuclear":
        return lp_models.nuclear_raw(lmp, **kwargs)
    if kind == "fossil_surrogate":
        return lp_models.fossil_surrogate_raw(lmp, **kwargs)
    raise ValueError(kind)


def _solve_chunk(args):
    lmps, extras = args
    kind, kwargs = _W["kind"], _W["kwargs"]
    out = np.empty(len(lmps))
    base = None
    for k, lmp in enumerate(lmps):
        extra = extras[k] if extras is not None else None
        if base is None or extra is not None:
            base, base_lmp = _build(kind, lmp, extra, kwargs), lmp
            lp = base
        else:       # same constraints, new cost vector only
            lp = dataclasses.replace(base, c=lp_models.swap_lmp(base, base_lmp, lmp))
        out[k], _ = solve(lp)
    return out


def solve_batch(kind, lmps, extras=None, kwargs=None, procs=None):
    """Objective of every LP of a batch; ``procs`` worker processes (default: all host cores).
    Returns (obj[N], seconds, procs)."""
    kwargs = kwargs or {}
    procs = procs or host_cores()
    lmps = np.asarray(lmps, float)
    N = lmps.shape[0]
    chunks = np.array_split(np.arange(N), max(1, min(N, procs * 4)))
    jobs = [(lmps[ix], None if extras is None else [extras[i] for i in ix]) for ix in chunks if ix.size]
    if procs == 1:
        _init_worker(kind, kwargs)
        t0 = time.perf_counter()
        outs = [_solve_chunk(j) for j in jobs]
        dt = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(procs, initializer=_init_worker, initargs=(kind, kwargs)) as pool:
            pool.map(_noop, range(procs))          # workers up before the clock starts (generous to the CPU arm)
            t0 = time.perf_counter()
            outs = pool.map(_solve_chunk, jobs)
            dt = time.perf_counter() - t0
    return np.concatenate(outs), dt, procs


def _noop(_):
    return 0
