"""Host-side plumbing of the reference-shaped APIs on a box WITHOUT a GPU: the bodies of the API-level gpu tests are run
with BatchLPSolver replaced by a HiGHS-backed stand-in (tests only -- the product has no CPU path).  This checks
argument handling, template parameter layouts, post-processing (NPV, record_results, result files, bids) against the
oracle; the CUDA solve itself is what the gpu-marked originals test."""
import numpy as np
import pytest
from scipy.optimize import linprog

from dispatches_b200 import double_loop as DLH
from dispatches_b200 import pricetaker as PT
from dispatches_b200 import solver as S


class HighsStandIn:
    def __init__(self, t, **kw):
        self.t = t

    def solve_host(self, cp, rp=None, want_x=False, want_y=False, out=None):
        cp = np.atleast_2d(cp)
        objs, xs = [], []
        for k in range(cp.shape[0]):
            r_ = np.zeros(0) if rp is None else (rp if np.ndim(rp) == 1 else rp[k])
            c, b, u, kk = self.t.instantiate(cp[k], r_)
            for opts in (dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10), {}):
                r = linprog(c, A_eq=self.t.matrix(r_), b_eq=b, bounds=[(0, None if not np.isfinite(v) else v) for v in u],
                            method="highs-ds", options=opts)
                if r.status == 0:
                    break
            assert r.status == 0, r.message
            objs.append(r.fun + kk); xs.append(r.x)
        n = len(objs)
        return S.LPResult(np.array(objs), np.zeros(n, np.int32), np.full(n, 10, np.int32), np.array(xs),
                          np.zeros((n, self.t.m)) if want_y else None)       # (duals are what the gpu tests check)

    def to_model_space(self, x):
        return x * self.t.col_scale + self.t.col_shift


@pytest.fixture
def stand_in(monkeypatch):
    for mod in (S, PT, DLH):
        monkeypatch.setattr(mod, "BatchLPSolver", HighsStandIn)
    monkeypatch.setattr(PT, "_SOLVERS", {})
    monkeypatch.setattr(DLH, "_SOLVERS", {})


def test_pricetaker_api_plumbing(stand_in, tmp_path):
    import test_gpu_parity as G
    G.test_reference_shaped_api()
    G.test_design_opt_free_wind()
    G.test_sweep_drivers_write_reference_shaped_results(tmp_path)


def test_solar_api_plumbing(stand_in):
    import test_solar_battery_hydrogen as SB
    SB.check_reference_shaped_api()


def test_multiperiod_builder_api_plumbing(stand_in, monkeypatch):
    """reference-shaped MultiPeriodModel code through SolverFactory("b200ipm").solve(m): Var write-back into the cloned period blocks"""
    import test_multiperiod_api as MP
    from dispatches_b200 import pyomo_plugin as PP
    monkeypatch.setattr(PP, "BatchLPSolver", HighsStandIn)
    MP.check_reference_shaped_optimize()


def test_double_loop_api_plumbing(stand_in):
    import test_double_loop as D
    D._check_tracker(None)
    D._check_nuclear(3)
