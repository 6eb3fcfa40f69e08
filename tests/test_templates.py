"""Host logic: the product's reduced templates (dispatches_b200/templates.py) against the raw oracle LPs."""
import numpy as np
import pytest
from scipy.optimize import linprog

from dispatches_b200 import scenarios as SC
from dispatches_b200 import templates as TP
from oracle import highs as H
from oracle import lp_models as L


def solve_template(t, cp, rp):
    c, b, u, k = t.instantiate(cp, rp)
    r = linprog(c, A_eq=t.A, b_eq=b, bounds=[(0, None if not np.isfinite(v) else v) for v in u], method="highs-ds")
    assert r.status == 0, r.message
    return r.fun + k, r.x


@pytest.fixture(scope="module")
def wb24():
    return TP.wind_battery(24)


def test_wind_battery_structure(wb24):
    t = wb24
    assert (t.m, t.n, t.nb) == (96, 167, 48)          # 4 rows/period; 5 cols + 2 slacks per period - s[T-1]
    assert t.w == 4                                   # block tridiagonal in time -> half bandwidth 4
    assert np.all(np.isfinite(t.u0[: t.nb])) and np.all(~np.isfinite(t.u0[t.nb:]))
    # the assembly list reproduces A D A' for a random D
    d = np.random.default_rng(0).uniform(0.1, 2.0, t.n)
    M = (t.A @ np.diag(d) @ t.A.T)
    M = np.asarray(M)
    for i in range(t.m):
        for k in range(t.w + 1):
            e = i * (t.w + 1) + k
            val = sum(t.asm_val[q] * d[t.asm_col[q]] for q in range(t.asm_ptr[e], t.asm_ptr[e + 1]))
            ref = M[i, i - k] if i - k >= 0 else 0.0
            assert val == pytest.approx(ref, abs=1e-12)
    # nothing outside the band
    ii, jj = np.nonzero(M)
    assert np.max(np.abs(ii - jj)) <= t.w


def test_wind_battery_matches_raw_oracle(wb24):
    lmp, cf, W, P = SC.c2(8)
    rp = TP.wind_battery_rparams(24, cf, W, P)[0]
    for k in range(8):
        a, _ = solve_template(wb24, lmp[k], rp)
        b, _ = H.solve(L.wind_battery_raw(lmp[k], cf, W, P))
        assert a == pytest.approx(b, rel=1e-11, abs=1e-9)


def test_wind_battery_c1_and_design_points(wb24):
    lmp, cf, W, P = SC.c1()
    for (w_mw, p_mw) in ((W, P), (200.0, 10.0), (1600.0, 1600.0)):
        rp = TP.wind_battery_rparams(24, cf, w_mw, p_mw)[0]
        a, _ = solve_template(wb24, lmp, rp)
        b, _ = H.solve(L.wind_battery_raw(lmp, cf, w_mw, p_mw))
        assert a == pytest.approx(b, rel=1e-11, abs=1e-9)


@pytest.mark.parametrize("with_battery", [True, False])
def test_wind_battery_pem_matches_raw_oracle(with_battery):
    t = TP.wind_battery_pem(24, with_battery=with_battery)
    lmp, cf, W, P = SC.c2(4)
    batt = 100.0 if with_battery else 0.0
    rp = TP.wind_battery_rparams(24, cf, W, batt, pem_mw=200.0)[0]
    for k in range(4):
        a, _ = solve_template(t, np.append(lmp[k], 2.5), rp)
        b, _ = H.solve(L.wind_battery_raw(lmp[k], cf, W, batt, pem_mw=200.0, h2_price=2.5))
        assert a == pytest.approx(b, rel=1e-11, abs=1e-9)


def test_nuclear_matches_raw_oracle():
    t = TP.nuclear(48)
    assert t.w == 1 and t.m == 48
    lmp = SC.c3(4)
    for k in range(4):
        a, _ = solve_template(t, lmp[k], np.zeros(0))
        b, _ = H.solve(L.nuclear_raw(lmp[k]))
        assert a == pytest.approx(b, rel=1e-11, abs=1e-9)


def test_fossil_surrogate_matches_raw_oracle():
    t = TP.fossil_surrogate(168)
    assert t.w <= 5
    lmp = SC.c4(2)
    for k in range(2):
        a, x = solve_template(t, lmp[k], np.zeros(0))
        b, _ = H.solve(L.fossil_surrogate_raw(lmp[k]))
        assert a == pytest.approx(b, rel=1e-11, abs=1e-9)
        xm = x * t.col_scale + t.col_shift
        j = t.col_names.index("blk[3].fs.plant_power_out[0]")
        assert 283.0 <= xm[j] <= 436.0 + 1e-9


def test_scenarios_are_seeded_and_keep_hard_cases():
    a, *_ = SC.c2(2000)
    b, *_ = SC.c2(2000)
    assert np.array_equal(a, b) and a.shape == (2000, 24)
    assert (a == 0).mean() > 0.05 and a.max() > 5000.0      # exact zeros and scarcity spikes survive
    lmp, cf, w, p = SC.c5(2, 2, 48)
    assert lmp.shape == (2 * 2 * 48, 24) and cf.shape == lmp.shape and w.shape == (192,)


def test_wind_battery_design_opt_template_matches_raw_oracle():
    """design_opt=True, extant_wind=True: per-period nameplate_power columns + link rows keep the matrix banded."""
    t = TP.wind_battery_design(24)
    assert t.w <= 16
    lmp, cf, W, P = SC.c2(6)
    lmp[0] *= 40.0                                   # a scarcity day: the optimal battery is not zero
    rp = TP.wind_battery_rparams(24, cf, W, 0.0)[0]
    jP = t.col_names.index("blk[0].fs.battery.nameplate_power")
    sizes = []
    for k in range(6):
        a, x = solve_template(t, lmp[k], rp)
        b, xr = H.solve(L.wind_battery_raw(lmp[k], cf, W, 0.0, design_opt=True, extant_wind=True))
        assert a == pytest.approx(b, rel=1e-10, abs=1e-8)
        sizes.append(x[jP])
    assert max(sizes) > 1.0                          # kW


def test_nuclear_report_enumeration_presolve_equals_the_full_lp():
    """run_exhaustive_enumeration (price_taker_analysis.py:353-425): the per-hour closed form of the product equals the
    full multi-period LP (tank and turbine capacity 0) solved by the oracle, on two weeks of the report's RT prices."""
    from dispatches_b200 import run_pricetaker as RP
    lmp = SC.pool()["nuc_report_lmp_rt"][2000:2336]
    res = RP.run_exhaustive_enumeration(lmp, pem_capex=400.0, h2_prices=(0.75, 1.5), pem_fractions=(0.05, 0.5))
    for i1, hp in enumerate((0.75, 1.5)):
        for i2, pc in enumerate((0.05, 0.5)):
            obj, x = H.solve(L.nuclear_report_raw(lmp, hp, pc * 400.0, pem_capex=400.0))
            assert res["net_npv"][f"{i1}{i2}"] == pytest.approx(-obj / 1e6, rel=1e-9)
    assert 0.0 <= res["pem_cap_factor"]["11"] <= 1.0 and res["solver_stat"]["00"] == "optimal"
    # schedule files in the format of _write_results (price_taker_analysis.py:325-350)
    import tempfile, pandas as pd
    with tempfile.TemporaryDirectory() as td:
        RP.run_exhaustive_enumeration(lmp, pem_capex=400.0, h2_prices=(1.5,), pem_fractions=(0.5,), schedule_csv_dir=td)
        df = pd.read_csv(f"{td}/results_00_schedule.csv", index_col=0)
        assert list(df.columns)[:3] == ["LMP [$/MWh]", "np_to_grid", "np_to_electrolyzer"] and len(df) == lmp.size
        assert np.allclose(df["net_cash_inflow"], df["h2_revenue"] + df["electricity_revenue"] - df["vom"])
        assert res["h2_rev"]["11"] * 1e6 == pytest.approx(df["h2_revenue"].sum(), rel=1e-12)


def test_design_free_wind_matches_raw_oracle():
    """design_opt=True, extant_wind=False: battery and wind size free (wind_battery_LMP.py:209-219); cf_t sits in A."""
    lmp, cf, W, P = SC.c2(6)
    t = TP.wind_battery_design_free_wind(24, cf)
    assert t.w <= 16 and t.Pr == 0
    for k in range(6):
        for scale in (1.0, 20.0):                       # ordinary prices: build nothing; 20x prices: build to the 10 GW cap
            c, b, u, kk = t.instantiate(lmp[k] * scale, np.zeros(0))
            r = None
            for opts in (dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10), {}):
                r = linprog(c, A_eq=t.A, b_eq=b, bounds=[(0, None if not np.isfinite(v) else v) for v in u], method="highs-ds", options=opts)
                if r.status == 0:
                    break
            assert r.status == 0, r.message
            ref, _ = H.solve(L.wind_battery_raw(lmp[k] * scale, cf, W, P, design_opt=True, extant_wind=False))
            assert r.fun + kk == pytest.approx(ref, rel=1e-9, abs=1e-6)


@pytest.mark.parametrize("tank,turb,demand", [(30000.0, 0.0, 3000.0), (50000.0, 40.0, 2000.0), (0.0, 25.0, 8000.0)])
def test_nuclear_report_with_tank_and_turbine_matches_raw_oracle(tank, turb, demand):
    """the report LP with a storage tank / hydrogen turbine (price_taker_analysis.py:116-222): reduced template == raw oracle LP"""
    from test_double_loop import solve_template
    lmp = SC.pool()["nuc_report_lmp_rt"][3000:3168]
    t = TP.nuclear_report(168, pem_capex=400.0, demand=demand)
    for hp, pem in ((0.75, 40.0), (2.0, 200.0), (1.25, 120.0)):
        obj, x = solve_template(t, np.r_[lmp, hp], np.array([pem, tank, turb]))
        ref, xr = H.solve(L.nuclear_report_raw(lmp, hp, pem, pem_capex=400.0, tank_cap=tank, turbine_cap=turb, demand=demand))
        assert obj == pytest.approx(ref, rel=1e-10)


def test_free_wind_with_a_capacity_factor_series_per_problem_matches_raw_oracle():
    """design_opt=True, extant_wind=False with a DIFFERENT capacity-factor series per batch member: cf_t multiplies the wind-capacity
    column (wind_power.py:120-122), i.e. it is a matrix coefficient -- LPTemplate.amap / matrix()"""
    from scipy.optimize import linprog
    T = 24
    t = TP.wind_battery_design_free_wind(T)
    assert t.amap is not None and len(t.amap[0]) == T and t.Pr == T
    lmp, cf, W, P = SC.c2(3)
    rng = np.random.default_rng(0)
    for k in range(3):
        cfk = np.clip(cf * rng.uniform(0.5, 1.5, T), 0, 1)
        c, b, u, kc = t.instantiate(lmp[k] * 20, cfk - TP.CF_NOMINAL)
        for opts in (dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10), dict()):
            r = linprog(c, A_eq=t.matrix(cfk - TP.CF_NOMINAL), b_eq=b, bounds=[(0, None if not np.isfinite(v) else v) for v in u],
                        method="highs-ds", options=opts)
            if r.status == 0:
                break
        ref, _ = H.solve(L.wind_battery_raw(lmp[k] * 20, cfk, W, P, design_opt=True, extant_wind=False))
        assert r.fun + kc == pytest.approx(ref, rel=1e-10)
