"""The pyomo-free half of the Pyomo walker (dispatches_b200/lp_template.py::standard_form) and the C-side symbolic setup
(dsp_lp_analyze_csr): the reference-shaped RAW LPs of the oracle -- every Var of every period block a column, fixed Vars with
lb == ub, free Vars (pem.electricity is Reals), explicit arc / link / periodic equalities -- go through the same path a walked
Pyomo model takes and must reproduce the oracle objective and a feasible model-space solution."""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import linprog

from dispatches_b200 import lp_template as LT, scenarios as SC
from oracle import highs as H, lp_models as L


def rows_of(lp):
    rows, lo, hi = [], [], []
    for A, b, eq in ((lp.A_eq, lp.b_eq, True), (lp.A_ub, lp.b_ub, False)):
        A = A.tocsr()
        for i in range(A.shape[0]):
            sl = slice(A.indptr[i], A.indptr[i + 1])
            rows.append({int(j): float(v) for j, v in zip(A.indices[sl], A.data[sl])})
            hi.append(b[i]); lo.append(b[i] if eq else -np.inf)
    return rows, lo, hi


def template_of(lp, lmp, **kw):
    rows, lo, hi = rows_of(lp)
    dcost = np.zeros((lp.n, len(lmp)))
    for j, t, coef in lp.meta["lmp_terms"]:
        dcost[j, t] += coef
    return LT.standard_form(rows, lo, hi, lp.lb, lp.ub, lp.c, lp.c0, 1.0, var_names=lp.names, p0=np.asarray(lmp, float), dcost=dcost, **kw)


def highs_template(t, cp, rp):
    c, b, u, k = t.instantiate(cp, rp)
    r = linprog(c, A_eq=t.A, b_eq=b, bounds=[(0, None if not np.isfinite(v) else v) for v in u], method="highs-ds",
                options=dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10))
    assert r.status == 0, r.message
    return r.fun + k, r.x, -r.eqlin.marginals if hasattr(r, "eqlin") else None


def feasible(lp, x, tol=1e-6):
    sc = max(1.0, np.abs(lp.b_eq).max(), np.abs(lp.b_ub).max() if lp.b_ub.size else 0.0)
    ok = np.abs(lp.A_eq @ x - lp.b_eq).max() <= tol * sc
    if lp.A_ub.shape[0]:
        ok &= (lp.A_ub @ x - lp.b_ub).max() <= tol * sc
    return ok and np.all(x >= lp.lb - tol * sc) and np.all(x <= lp.ub + tol * sc)


@pytest.mark.parametrize("kw", [dict(), dict(pem_mw=200.0, h2_price=2.5)], ids=["wind_battery", "wind_battery_pem_free_vars"])
def test_raw_reference_shaped_lp_through_the_walker_path(kw):
    lmp, cf, W, P = SC.c2(5)
    lp0 = L.wind_battery_raw(lmp[0], cf, W, P, **kw)
    t = template_of(lp0, lmp[0])
    assert t.w <= 32 and t.m < lp0.A_eq.shape[0] + lp0.A_ub.shape[0]        # arcs / links presolved away, band fits the kernel
    for k in range(5):                                                        # the batch: the LMP Params only
        obj, x, _ = highs_template(t, lmp[k], lmp[k])
        lp = L.wind_battery_raw(lmp[k], cf, W, P, **kw)
        ref, _ = H.solve(lp)
        assert obj == pytest.approx(ref, rel=1e-12)
        xm = LT.model_values(t, x[None], lmp[k][None])[0]
        assert feasible(lp, xm) and lp.c @ xm + lp.c0 == pytest.approx(ref, rel=1e-9)


def test_ranged_rows_maximise_upper_bounded_and_param_dependent_bounds():
    """max 3x + 2y - z,  1 <= x + y <= 4 + p,  x - z = 0.5,  x <= 3 (no lower bound),  0 <= y <= 2 p,  z free"""
    rows = [{0: 1.0, 1: 1.0}, {0: 1.0, 2: -1.0}]
    p0 = np.array([1.0])
    t = LT.standard_form(rows, [1.0, 0.5], [5.0, 0.5], [-np.inf, 0.0, -np.inf], [3.0, 2.0, np.inf], [3.0, 2.0, -1.0], c0=7.0, sense=-1.0,
                         p0=p0, dhi=[[1.0], [0.0]], dub=[[0.0], [2.0], [0.0]], presolve=False, equilibrate=False)
    for p in (1.0, 0.25, 2.0):
        obj, x, _ = highs_template(t, [p], [p])
        r = linprog([-3.0, -2.0, 1.0], A_ub=[[1, 1, 0], [-1, -1, 0]], b_ub=[4 + p, -1], A_eq=[[1, 0, -1]], b_eq=[0.5],
                    bounds=[(None, 3), (0, 2 * p), (None, None)], method="highs-ds")
        assert -obj == pytest.approx(-r.fun + 7.0, rel=1e-12)
        assert LT.model_values(t, x[None], np.array([[p]]))[0] == pytest.approx(r.x, abs=1e-9)


def test_c_side_symbolic_setup_matches_the_python_one(cuda_solver_lib):
    """dsp_lp_analyze_csr (host only): bounded columns first, row order of minimal band (RCM vs natural) -- the band it finds
    is never wider than what lp_template.finalize() finds, for shuffled rows / columns as a C caller might present them."""
    from dispatches_b200 import solver as S, templates as TP
    lib = cuda_solver_lib
    for t in (TP.wind_battery(24), TP.nuclear(48), TP.wind_battery_pem(24), TP.wind_battery_operation(48, "bidder_da")):
        rng = np.random.default_rng(1)
        rperm, cperm = rng.permutation(t.m), rng.permutation(t.n)
        A = t.A.tocsr()[rperm][:, cperm].tocsr(); A.sort_indices()
        u0 = np.where(np.isfinite(t.u0), t.u0, 1e300)[cperm]
        keep = dict(ptr=A.indptr.astype(np.int32), idx=A.indices.astype(np.int32), val=A.data.astype(float), u0=np.ascontiguousarray(u0))
        d = S._LpDesc(m=t.m, n=t.n, Pc=0, Pr=0, A_ptr=keep["ptr"].ctypes.data, A_idx=keep["idx"].ctypes.data, A_val=keep["val"].ctypes.data,
                      u0=keep["u0"].ctypes.data)
        nb, w, wn, wr = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        cp = np.zeros(t.n, np.int32); rp = np.zeros(t.m, np.int32)
        rc = lib.dsp_lp_analyze_csr(C.byref(d), C.byref(nb), C.byref(w), C.byref(wn), C.byref(wr), cp.ctypes.data_as(C.c_void_p), rp.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert nb.value == t.nb and sorted(cp) == list(range(t.n)) and sorted(rp) == list(range(t.m))
        assert np.all(np.isfinite(t.u0[cperm][cp[:nb.value]])) and not np.any(np.isfinite(t.u0[cperm][cp[nb.value:]]))
        assert w.value <= max(t.w, 1) + 1, (t.name, w.value, t.w)
        B = A[rp]
        Pm = (abs(B) @ abs(B).T).tocoo()
        assert np.abs(Pm.row - Pm.col).max() == w.value                      # the reported band is the band of the permuted A A'
