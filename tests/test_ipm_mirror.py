"""The interior-point ALGORITHM (numpy mirror of the CUDA kernel, oracle/ipm_numpy.py) against HiGHS on the
hard price signals: exact zeros, 10 000 $/MWh spikes."""
import numpy as np
import pytest
from scipy.optimize import linprog

from dispatches_b200 import scenarios as SC
from dispatches_b200 import templates as TP
from oracle import ipm_numpy as I

TOL_OBJ = 1e-6      # north_star: objective within 1e-6 relative


def _batch(t, cps, rp):
    cs, bs, us, ks = zip(*(t.instantiate(cp, rp) for cp in cps))
    return np.array(cs), np.array(bs), np.array(us), np.array(ks)


def test_mirror_wind_battery_200():
    t = TP.wind_battery(24)
    lmp, cf, W, P = SC.c2(200)
    rp = TP.wind_battery_rparams(24, cf, W, P)[0]
    c, b, u, k = _batch(t, lmp, rp)
    r = I.solve_batch(t.A.toarray(), b, c, u, tol=1e-8, eta=0.9995)
    assert (r["status"] == I.OPTIMAL).all() and r["iters"].max() <= 30
    ref = np.array([linprog(c[i], A_eq=t.A, b_eq=b[i], bounds=[(0, None if not np.isfinite(v) else v) for v in u[i]],
                            method="highs-ds").fun for i in range(len(c))])
    err = np.abs(r["obj"] - ref) / np.maximum(1.0, np.abs(ref + k))
    assert err.max() < TOL_OBJ
    assert (np.abs(r["obj"] - ref) / np.maximum(1e-3, np.abs(ref))).max() < 1e-6   # LP part alone


def test_mirror_nuclear_50():
    t = TP.nuclear(48)
    lmp = SC.c3(50)
    c, b, u, k = _batch(t, lmp, np.zeros(0))
    r = I.solve_batch(t.A.toarray(), b, c, u, tol=1e-8, eta=0.9995)
    assert (r["status"] == I.OPTIMAL).all()
    ref = np.array([linprog(c[i], A_eq=t.A, b_eq=b[i], bounds=[(0, None if not np.isfinite(v) else v) for v in u[i]],
                            method="highs-ds").fun for i in range(len(c))])
    assert (np.abs(r["obj"] - ref) / np.maximum(1.0, np.abs(ref + k))).max() < TOL_OBJ


def test_stage_mirror_matches_generic_mirror_and_highs():
    """The stage kernel's algebra (local elimination + twisted block LDL', oracle/ipm_stage_numpy.py) gives the
    same iterates as the generic normal-equations mirror, and stays optimal on the degenerate C5 design sweep
    (hours with zero capacity factor, tiny batteries)."""
    from oracle import ipm_stage_numpy as ST
    t = TP.wind_battery(24)
    st = t.meta["stage_wb"]
    consts = {k: st[k] for k in ("a", "binv", "half", "delta", "dur", "k_rev")}
    lmp, cf, W, P = SC.c2(120)
    rp = TP.wind_battery_rparams(24, cf, W, P)[0]
    c, b, u, k = _batch(t, lmp, rp)
    g = I.solve_batch(t.A.toarray(), b, c, u)
    s = ST.solve_batch(lmp, W * 1e3 * cf, P * 1e3, consts)
    assert (s["status"] == 0).all() and (g["status"] == 0).all()
    assert (s["iters"] == g["iters"]).mean() > 0.95
    assert (np.abs(s["obj_lp"] - g["obj"]) / np.maximum(1.0, np.abs(g["obj"] + k))).max() < 1e-7
    lmp5, cf5, w5, b5 = SC.c5(4, 4, 40)
    s5 = ST.solve_batch(lmp5, (w5 * 1e3)[:, None] * cf5, b5 * 1e3, consts)
    assert (s5["status"] == 0).all() and s5["iters"].max() <= 30
    from oracle import highs as H, lp_models as L
    for i in range(0, len(lmp5), 97):
        ref = H.solve(L.wind_battery_raw(lmp5[i], cf5[i], w5[i], b5[i]))[0]
        ki = t.instantiate(lmp5[i], TP.wind_battery_rparams(24, cf5[i], w5[i], b5[i])[0])[3]
        assert abs(s5["obj_lp"][i] + ki - ref) / max(1.0, abs(ref)) < TOL_OBJ


def test_stage_mirror_deferred_reciprocal_step(monkeypatch):
    """INV="defer" restates the kernel's elimination step (the neighbour's block is passed on, C adj(R) C' is formed while
    1/det(R) is in flight): same iteration counts and objectives as the LDL' inverse on the C2 sample and the C5 slice."""
    from oracle import ipm_stage_numpy as ST
    t = TP.wind_battery(24)
    st = t.meta["stage_wb"]
    consts = {k: st[k] for k in ("a", "binv", "half", "delta", "dur", "k_rev")}
    lmp, cf, W, P = SC.c2(120)
    lmp5, cf5, w5, b5 = SC.c5(4, 4, 40)
    ref = ST.solve_batch(lmp, W * 1e3 * cf, P * 1e3, consts)
    ref5 = ST.solve_batch(lmp5, (w5 * 1e3)[:, None] * cf5, b5 * 1e3, consts)
    monkeypatch.setattr(ST, "INV", "defer")
    s = ST.solve_batch(lmp, W * 1e3 * cf, P * 1e3, consts)
    s5 = ST.solve_batch(lmp5, (w5 * 1e3)[:, None] * cf5, b5 * 1e3, consts)
    assert (s["status"] == 0).all() and (s5["status"] == 0).all()
    assert (s["iters"] == ref["iters"]).mean() > 0.95 and (s5["iters"] == ref5["iters"]).mean() > 0.9
    assert (np.abs(s["obj_lp"] - ref["obj_lp"]) / np.maximum(1e-3, np.abs(ref["obj_lp"]))).max() < 1e-6
