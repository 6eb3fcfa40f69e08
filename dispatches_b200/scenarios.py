"""Seeded synthetic LMP scenario batches for the BASELINE.json configs (definitions: SURVEY.md §8(d)).

The pools are REAL price / capacity-factor series of the reference's committed data files, extracted once by
tests/golden/make_golden.py into dispatches_b200/data/lmp_pool.npz (RTS-GMLC day-ahead / real-time LMPs at
buses 122/303/309/317 from renewables_case/data/Wind_Thermal_Dispatch.csv, the 3100 cluster days of
nuclear_case/lmp_signal.json, the 8736-h 303_DALMP / 303_WIND_1-DACF series of load_parameters.py:82-112).
Exact zeros (25 % of hours) and the 10 000 $/MWh scarcity spikes are kept: they are the numerically hard cases.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

_POOL = None
FIXED_WIND_MW = 847.0          # load_parameters.py:63


def pool():
    global _POOL
    if _POOL is None:
        with np.load(Path(__file__).resolve().parent / "data" / "lmp_pool.npz") as d:
            _POOL = {k: d[k] for k in d.files}
    return _POOL


def _noisy(base, rng, sigma=0.25):
    return base * rng.lognormal(0.0, sigma, base.shape)


def c1():
    """C1 plumbing case: the first day of the 303 DA series, W = 847 MW, P = 0.25 W."""
    p = pool()
    return p["dalmp_303"][:24].copy(), p["dacf_303"][:24].copy(), FIXED_WIND_MW, 0.25 * FIXED_WIND_MW


def c2(N=10000, seed=20240101):
    """C2: N 24-h LMP vectors (2912 real day windows U 3100 cluster days, resampled, x lognormal(0,0.25));
    capacity factors fixed to the first day.  Returns lmp [N,24], cf [24], wind_mw, batt_mw."""
    p = pool()
    rng = np.random.default_rng(seed)
    base = np.concatenate([p["day_windows"], p["cluster_days"]])
    lmp = _noisy(base[rng.integers(0, len(base), N)], rng)
    return lmp, p["dacf_303"][:24].copy(), FIXED_WIND_MW, 0.25 * FIXED_WIND_MW


def c3(N=5000, seed=20240102):
    """C3: N 48-h vectors = two consecutive cluster days + noise (nuclear template)."""
    p = pool()
    rng = np.random.default_rng(seed)
    cl = p["cluster_days"]
    k = rng.integers(0, len(cl) - 1, N)
    return _noisy(np.concatenate([cl[k], cl[k + 1]], axis=1), rng)


def c4(N=2000, seed=20240103):
    """C4: N 168-h weekly vectors from the 303 DA series + noise (fossil surrogate template)."""
    p = pool()
    rng = np.random.default_rng(seed)
    weeks = p["dalmp_303"].reshape(52, 168)
    return _noisy(weeks[rng.integers(0, 52, N)], rng)


def c5(n_wind=8, n_ratio=8, n_hours=8760):
    """C5 design sweep: (wind size x battery ratio) x start hour -> 24-h windows of price AND capacity factor.
    Returns lmp [D*H,24], cf [D*H,24], wind_mw [D*H], batt_mw [D*H] with D = n_wind*n_ratio."""
    p = pool()
    lam, cf = p["dalmp_303"], p["dacf_303"]
    L = len(lam)
    idx = (np.arange(n_hours)[:, None] + np.arange(24)[None, :]) % L
    lam_w, cf_w = lam[idx], cf[idx]
    wind = np.linspace(200.0, 1600.0, n_wind)
    ratio = np.linspace(0.05, 1.0, n_ratio)
    W, R = np.meshgrid(wind, ratio, indexing="ij")
    W, R = W.ravel(), R.ravel()
    D = W.size
    lmp = np.broadcast_to(lam_w[None], (D, n_hours, 24)).reshape(-1, 24)
    cfs = np.broadcast_to(cf_w[None], (D, n_hours, 24)).reshape(-1, 24)
    wind_mw = np.repeat(W, n_hours)
    batt_mw = np.repeat(W * R, n_hours)
    return np.ascontiguousarray(lmp), np.ascontiguousarray(cfs), wind_mw, batt_mw
