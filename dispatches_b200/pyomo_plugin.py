"""Pyomo solver plugin ``SolverFactory("b200ipm")`` -- SURVEY.md §8(f)-1, import-guarded.

Replaces ``pyo.SolverFactory("cbc")`` at the reference's call sites (wind_battery_LMP.py:266-267,
wind_battery_PEM_LMP.py:296-298, price_taker_analysis.py:365,403) without touching the model builders: the built
Pyomo model is walked ONCE into an LPTemplate, then solved on the GPU -- for the model as it stands (batch of 1, values
written back into the Vars like any Pyomo solver) or for a whole batch of values of its mutable Params
(``batch_params={param: array[N]}``, e.g. every block's ``lmp_signal``, wind_battery_LMP.py:234,243-244).

pyomo is NOT installed in the build image, so this module is exercised only where pyomo exists;
``import dispatches_b200.pyomo_plugin`` never fails, ``available()`` reports the truth.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .lp_template import INF, LPTemplate
from .solver import MAX_ITER, NUMERICAL, OPTIMAL, BatchLPSolver

try:                                            # pragma: no cover - pyomo absent in the build image
    import pyomo.environ as pyo
    from pyomo.core.base.var import _GeneralVarData  # noqa: F401
    from pyomo.opt import SolverFactory, SolverResults, SolverStatus, TerminationCondition
    from pyomo.repn import generate_standard_repn
    HAVE_PYOMO = True
except Exception:                               # noqa: BLE001
    HAVE_PYOMO = False


def walk_model(model, batch_params=None):
    """Pyomo model -> (LPTemplate, var list, param list).  Linear models only.

    Every active Constraint / the active Objective is reduced with generate_standard_repn (fixed Vars fold into the
    constant).  Mutable Params listed in ``batch_params`` are handled by finite differencing the (affine) dependence
    of the repn constants / coefficients on them: cost coefficients may depend on them (-> Cmap), right-hand sides and
    bounds may (-> Bmap / Umap); constraint-matrix coefficients may not (A is shared by the batch)."""
    if not HAVE_PYOMO:
        raise RuntimeError("pyomo is not installed: the b200ipm plugin needs it to walk the model")
    batch_params = list(batch_params or [])
    vars_, vid = [], {}

    def col(v):
        k = id(v)
        if k not in vid:
            vid[k] = len(vars_)
            vars_.append(v)
        return vid[k]

    def snapshot():
        rows, rhs_lo, rhs_hi = [], [], []
        for c in model.component_data_objects(pyo.Constraint, active=True, descend_into=True):
            r = generate_standard_repn(c.body, compute_values=True)
            if not r.is_linear():
                raise ValueError(f"constraint {c.name} is not linear: b200ipm solves LPs only")
            rows.append({col(v): float(a) for v, a in zip(r.linear_vars, r.linear_coefs)})
            lo = -INF if c.lower is None else float(pyo.value(c.lower)) - float(r.constant)
            hi = INF if c.upper is None else float(pyo.value(c.upper)) - float(r.constant)
            rhs_lo.append(lo); rhs_hi.append(hi)
        objs = list(model.component_data_objects(pyo.Objective, active=True, descend_into=True))
        if len(objs) != 1:
            raise ValueError("exactly one active Objective expected")
        r = generate_standard_repn(objs[0].expr, compute_values=True)
        if not r.is_linear():
            raise ValueError("objective is not linear")
        sign = 1.0 if objs[0].sense == pyo.minimize else -1.0
        cost = {col(v): sign * float(a) for v, a in zip(r.linear_vars, r.linear_coefs)}
        return rows, np.array(rhs_lo), np.array(rhs_hi), cost, sign * float(r.constant), sign

    base = snapshot()
    rows, lo, hi, cost, c0, sign = base
    n0 = len(vars_)
    lb = np.array([(-INF if v.lb is None else float(v.lb)) for v in vars_])
    ub = np.array([(INF if v.ub is None else float(v.ub)) for v in vars_])
    if np.any(~np.isfinite(lb)):
        raise ValueError("free / lower-unbounded Vars are not supported yet: give them a lower bound")
    # affine dependence on the batch params by one finite difference each (the dependence is affine by construction)
    dcost = np.zeros((n0, len(batch_params))); dlo = np.zeros((len(rows), len(batch_params)))
    dhi = np.zeros_like(dlo); dc0 = np.zeros(len(batch_params))
    for k, p in enumerate(batch_params):
        old = pyo.value(p)
        p.set_value(old + 1.0)
        rows1, lo1, hi1, cost1, c01, _ = snapshot()
        p.set_value(old)
        if any(r1 != r0 for r1, r0 in zip(rows1, rows)):
            raise ValueError(f"Param {p.name} enters the constraint matrix: it cannot be batched")
        for j, v in cost1.items():
            dcost[j, k] = v - cost.get(j, 0.0)
        fin = np.isfinite(lo); dlo[fin, k] = lo1[fin] - lo[fin]
        fin = np.isfinite(hi); dhi[fin, k] = hi1[fin] - hi[fin]
        dc0[k] = c01 - c0
    p0 = np.array([pyo.value(p) for p in batch_params], float)
    # standard form: shift lower bounds to 0, slack column per inequality row, equality rows as they are
    A_rows, b0, bmap, slack_ub = [], [], [], []
    n = n0
    extra_cols = []
    for i, row in enumerate(rows):
        shift = sum(a * lb[j] for j, a in row.items())
        if lo[i] == hi[i]:
            A_rows.append(dict(row)); b0.append(hi[i] - shift); bmap.append(dhi[i])
        else:
            if np.isfinite(lo[i]) and np.isfinite(hi[i]):
                raise ValueError("ranged constraints are not supported: split them")
            r2 = dict(row); r2[n] = 1.0 if np.isfinite(hi[i]) else -1.0
            extra_cols.append(n); n += 1
            A_rows.append(r2)
            b0.append((hi[i] if np.isfinite(hi[i]) else lo[i]) - shift)
            bmap.append(dhi[i] if np.isfinite(hi[i]) else dlo[i])
    m = len(A_rows)
    ri, ci, vv = [], [], []
    for i, row in enumerate(A_rows):
        for j, a in row.items():
            if a != 0.0:
                ri.append(i); ci.append(j); vv.append(a)
    A = sp.csr_matrix((vv, (ri, ci)), shape=(m, n))
    c_vec = np.zeros(n); Cmap = np.zeros((n, len(batch_params)))
    for j, v in cost.items():
        c_vec[j] = v
    Cmap[:n0] = dcost
    u0 = np.full(n, INF); u0[:n0] = ub - lb
    P = len(batch_params)
    b0 = np.array(b0) - (np.array(bmap).reshape(m, P) @ p0 if P else 0.0)
    c0_vec = c_vec - (Cmap @ p0 if P else 0.0)
    o0 = c0 + float(c_vec[:n0] @ lb) - (float(dc0 @ p0) if P else 0.0)
    # one parameter vector serves as both cparams and rparams (costs and right-hand sides may share Params)
    t = LPTemplate("pyomo:" + str(model.name), A, b0, sp.csr_matrix(np.array(bmap).reshape(m, P)), c0_vec,
                   sp.csr_matrix(Cmap), u0, sp.csr_matrix((n, P)), o0 - 0.0, np.zeros(P), dc0 + (Cmap[:n0].T @ lb if P else 0.0),
                   np.concatenate([lb, np.zeros(n - n0)]), np.ones(n),
                   [v.name for v in vars_] + [f"slack[{k}]" for k in range(n - n0)], [f"row[{i}]" for i in range(m)],
                   dict(kind="pyomo", sign=sign))
    return t.finalize(equilibrate=True), vars_, batch_params, p0


class B200IPM:
    """Object with the part of Pyomo's solver interface the reference uses."""

    def __init__(self, **kw):
        self.options = dict(tol=1e-9, feas_tol=1e-9, max_iter=60)
        self.options.update(kw)

    def available(self, exception_flag=True):
        try:
            from .solver import load_library
            import torch
            load_library()
            ok = HAVE_PYOMO and torch.cuda.is_available()
        except Exception:                       # noqa: BLE001
            ok = False
        if not ok and exception_flag:
            raise RuntimeError("b200ipm needs pyomo, libdsp_lp.so and a CUDA device (there is no CPU fallback)")
        return ok

    def solve(self, model, tee=False, batch_params=None, **_):
        t, vars_, params, p0 = walk_model(model, batch_params=list((batch_params or {}).keys()))
        sol = BatchLPSolver(t, tol=self.options["tol"], feas_tol=self.options["feas_tol"], max_iter=self.options["max_iter"])
        if batch_params:
            pv = np.column_stack([np.asarray(batch_params[p], float) for p in params])
        else:
            pv = p0[None, :] if p0.size else np.zeros((1, 0))
        r = sol.solve_host(pv, pv if t.Pr else None, want_x=True, want_y=True)
        xm = sol.to_model_space(r.x)
        sign = t.meta["sign"]
        if pv.shape[0] == 1:                    # plain solver behaviour: write back into the Vars
            for j, v in enumerate(vars_):
                v.set_value(float(xm[0, j]), skip_validation=True)
                v.stale = False
        res = SolverResults()
        worst = int(r.status.max())
        res.solver.status = {OPTIMAL: SolverStatus.ok, MAX_ITER: SolverStatus.warning, NUMERICAL: SolverStatus.error}[worst]
        res.solver.termination_condition = {OPTIMAL: TerminationCondition.optimal,
                                            MAX_ITER: TerminationCondition.maxIterations,
                                            NUMERICAL: TerminationCondition.error}[worst]
        res.problem.lower_bound = res.problem.upper_bound = float(sign * r.obj[0])
        res.batch = dict(obj=sign * r.obj, status=r.status, iters=r.iters, x=xm, var_names=[v.name for v in vars_])
        if tee:
            print(f"b200ipm: {pv.shape[0]} LP(s), m={t.m} n={t.n} w={t.w}, iterations max {int(r.iters.max())}")
        return res


if HAVE_PYOMO:                                  # pragma: no cover
    try:
        SolverFactory.register("b200ipm", doc="B200 batched interior-point LP solver")(B200IPM)
    except Exception:                           # noqa: BLE001
        pass
