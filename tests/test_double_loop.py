"""Double-loop (tracking / bidding) LPs, SURVEY.md §8(f)-2.

CPU part: the oracle restatement (oracle/double_loop.py) against the reference's known answers
(tests/golden/double_loop_golden.json <- case_studies/renewables_case/tests/test_multiperiod_wind_battery_doubleloop.py),
the reduced product template against the raw oracle LP, and the host logic of dispatches_b200/double_loop.py with the
LP solve replaced by HiGHS on the instantiated template (the CUDA solve itself is covered by the gpu tests below).
"""
import json
from pathlib import Path

import numpy as np
import pytest
from scipy.optimize import linprog

from dispatches_b200 import double_loop as DLH
from dispatches_b200 import templates as TP
from oracle import double_loop as DL
from oracle import highs as H

G = json.load(open(Path(__file__).parent / "golden" / "double_loop_golden.json"))
CF = np.array(G["cf_309_rt"])
SIZES = dict(wind_mw=G["wind_pmax_mw"], batt_mw=G["battery_pmax_mw"], energy_mwh=G["battery_energy_capacity_mwh"])


def solve_template(t, cp, rp):
    c, b, u, k = t.instantiate(cp, rp)
    r = linprog(c, A_eq=t.A, b_eq=b, bounds=[(0, None if not np.isfinite(v) else v) for v in u], method="highs-ds",
                options=dict(primal_feasibility_tolerance=1e-10, dual_feasibility_tolerance=1e-10))
    assert r.status == 0, r.message
    return r.fun + k, r.x


def highs_lp_solve(mode, T, cparams, rparams, n_tracking_hour=1, options=None, family="wind_battery"):
    """Stand-in for dispatches_b200.double_loop._lp_solve on a box without a GPU (tests only)."""
    t = DLH._FAMILIES[family](T, mode, n_tracking_hour)
    objs, xs = zip(*(solve_template(t, cp, rp) for cp, rp in zip(cparams, rparams)))
    xm = np.array(xs) * t.col_scale + t.col_shift
    return np.array(objs), np.zeros(len(objs), np.int32), DLH._columns(t, xm, T)


# ---------------------------------------------------------------- oracle vs the reference's known answers
def test_oracle_tracker_known_answer():
    g = G["tracker"]
    lp = DL.tracker_raw(g["market_dispatch"], CF[:4], **SIZES, n_tracking_hour=g["n_tracking_hour"])
    _, x = H.solve(lp)
    p = DL.read_profile(lp, x)
    assert p["wind"] == pytest.approx(g["expected_wind_power"], rel=g["rel"])                 # reference test :92-98
    assert p["P_T"] == pytest.approx(g["market_dispatch"], abs=g["abs_power"])                # :100-106
    assert p["batt_in"] == pytest.approx(np.array(g["expected_wind_power"]) - 1e3 * np.array(g["market_dispatch"]), rel=g["rel"])


def _forecasts():
    return DL.backcast(G["da_309"], 0, 48, 1)[0], DL.backcast(G["rt_309"], 0, 48, 1)[0]


def test_oracle_self_schedule_known_answer():
    da, rt = _forecasts()
    lp = DL.bidder_raw(da, rt, CF, **SIZES)
    _, x = H.solve(lp)
    bids = np.round(DL.read_profile(lp, x)["da"], 4)
    known = np.array(G["self_schedule"]["known_solution"])
    assert np.allclose(bids, known, rtol=G["self_schedule"]["reltol"], atol=1e-4)            # reference test :175


def test_oracle_thermal_bid_known_answer():
    da, rt = _forecasts()
    lp = DL.bidder_raw(da, rt, CF, **SIZES)
    _, x = H.solve(lp)
    power = np.round(DL.read_profile(lp, x)["da"], 4)
    known = np.array(G["thermal_bid"]["known_solution"])
    assert np.allclose(power * np.round(da, 4), known, rtol=G["thermal_bid"]["reltol"], atol=1e-3)   # :233-252


# ---------------------------------------------------------------- template vs raw oracle
@pytest.mark.parametrize("T", [4, 24, 48])
def test_operation_templates_match_raw_oracle(T):
    rng = np.random.default_rng(T)
    cf = np.roll(CF, -T)[:T]
    da = rng.uniform(0, 60, T); rt = da + rng.normal(0, 10, T)
    soc0, thr0 = 31234.57, 5021.5
    rp = TP.wind_battery_operation_rparams(T, cf, 200, 25, 100, soc0, thr0)[0]
    a, _ = solve_template(TP.wind_battery_operation(T, "bidder_da"), np.r_[da, rt, 1e3], rp)
    b, _ = H.solve(DL.bidder_raw(da, rt, cf, **SIZES, soc0=soc0, thr0=thr0))
    assert a == pytest.approx(b, rel=1e-10, abs=1e-7)
    disp = rng.uniform(0, 120, T)
    rp = TP.wind_battery_operation_rparams(T, cf, 200, 25, 100, soc0, thr0, disp)[0]
    a, _ = solve_template(TP.wind_battery_operation(T, "tracker"), np.array([1e3]), rp)
    b, _ = H.solve(DL.tracker_raw(disp, cf, **SIZES, soc0=soc0, thr0=thr0))
    assert a == pytest.approx(b, rel=1e-10, abs=1e-7)
    a, _ = solve_template(TP.wind_battery_operation(T, "bidder_rt"), np.r_[da, rt, 1e3], rp)
    b, _ = H.solve(DL.bidder_raw(da, rt, cf, **SIZES, soc0=soc0, thr0=thr0, da_dispatch=disp))
    assert a - float(np.sum((da - rt) * disp)) == pytest.approx(b, rel=1e-10, abs=1e-6)


def test_operation_template_band():
    t = TP.wind_battery_operation(48, "bidder_da")
    assert (t.m, t.w) == (240, 5) and t.nb == 96


# ---------------------------------------------------------------- host logic (HiGHS stand-in for the CUDA solve)
def _model(n=1):
    md = DLH.RenewableGeneratorModelData(gen_name="309_WIND_1", bus="Carter", p_min=0, p_max=200, p_cost=0)
    series = np.tile(CF, 3)
    return DLH.MultiPeriodWindBattery(model_data=md, wind_capacity_factors=series if n == 1 else np.tile(series, (n, 1)),
                                      wind_pmax_mw=200, battery_pmax_mw=25, battery_energy_capacity_mwh=100)


def _check_tracker(solve):
    g = G["tracker"]
    tr = DLH.Tracker(tracking_model_object=_model(), tracking_horizon=4, n_tracking_hour=1)
    tr.track_market_dispatch(market_dispatch=g["market_dispatch"], date="2020-01-02", hour="00:00")
    wind = (tr.fs.sol["grid"] + tr.fs.sol["batt_in"])[0]
    assert wind == pytest.approx(g["expected_wind_power"], rel=g["rel"])
    assert tr.power_output[0] == pytest.approx(g["market_dispatch"], abs=g["abs_power"])
    assert tr.fs.sol["batt_in"][0] == pytest.approx(np.array(g["expected_wind_power"]) - 1e3 * np.array(g["market_dispatch"]), rel=g["rel"])
    # update_model: the realised SoC of hour 0 (rounded to 2 decimals) starts the next horizon, capacity factors shift by 1 h
    assert tr.fs.soc0[0] == pytest.approx(round(0.95 * (g["expected_wind_power"][0] - 0.0), 2), rel=1e-3)
    assert tr.fs._time_idx == 1 and np.allclose(tr.fs.cf[0], CF[1:5])
    assert tr.get_last_delivered_power()[0] == pytest.approx(0.0, abs=1e-3)
    df = tr.tracking_model_object.result_list[0]
    assert list(df["Total Power Output [MW]"]) == pytest.approx(g["market_dispatch"], abs=0.011)


def _check_bidders(solve):
    bc = DLH.Backcaster({"Carter": G["da_309"]}, {"Carter": G["rt_309"]})
    ss = DLH.SelfScheduler(bidding_model_object=_model(), day_ahead_horizon=48, real_time_horizon=4, n_scenario=1, forecaster=bc)
    bids = ss.compute_day_ahead_bids(date="2020-01-02")
    energies = np.array([b["309_WIND_1"]["p_max"] for b in bids.values()])
    known = np.array(G["self_schedule"]["known_solution"])
    lp = DL.bidder_raw(*_forecasts(), CF, **SIZES)
    ref_obj, _ = H.solve(lp)
    assert -ss.day_ahead_model.objective[0] == pytest.approx(ref_obj, rel=1e-8)
    # the LP has alternative optima: where da == rt exactly the day-ahead quantity has zero cost, and at hour 5 the
    # battery can charge in either of two equally priced hours (HiGHS simplex and HiGHS IPM+crossover already return
    # different vertices there).  An interior-point solution sits between them, so bids are compared where unique.
    da_f, rt_f = _forecasts()
    uniq = np.abs(da_f - rt_f) > 1e-9; uniq[5] = False
    assert uniq.sum() >= 30
    assert np.allclose(energies[uniq], known[uniq], rtol=G["self_schedule"]["reltol"], atol=2e-3)
    md = DLH.ThermalGeneratorModelData(gen_name="309_WIND_1", bus="Carter", p_min=0, p_max=200,
                                       production_cost_bid_pairs=[(0, 0), (200, 0)], startup_cost_pairs=[(0, 0)])
    mb = _model(); mb.model_data = md
    bd = DLH.Bidder(bidding_model_object=mb, day_ahead_horizon=48, real_time_horizon=4, n_scenario=1, forecaster=bc)
    bids = bd.compute_day_ahead_bids(date="2020-01-02")
    prices = np.array([b["309_WIND_1"]["p_cost"][-1][1] for b in bids.values()])
    knownp = np.array(G["thermal_bid"]["known_solution"])
    assert np.allclose(prices[uniq], knownp[uniq], rtol=G["thermal_bid"]["reltol"], atol=0.1)
    # real-time problem: day-ahead dispatch fixed, 4-hour horizon; underbid only where the schedule cannot be met
    rt_bids = ss.compute_real_time_bids(date="2020-01-02", hour=0, realized_day_ahead_prices=_forecasts()[0],
                                        realized_day_ahead_dispatches=known)
    assert len(rt_bids) == 4 and np.all(ss.real_time_underbid_power < 1e-6)
    lp = DL.bidder_raw(_forecasts()[0][:4], _forecasts()[1][:4], CF[:4], **SIZES, da_dispatch=known[:4])
    ref_obj, _ = H.solve(lp)
    assert -ss.real_time_model.objective[0] == pytest.approx(ref_obj, rel=1e-8, abs=1e-6)


def test_host_tracker_logic(monkeypatch):
    monkeypatch.setattr(DLH, "_lp_solve", highs_lp_solve)
    _check_tracker(highs_lp_solve)


def test_host_bidder_logic(monkeypatch):
    monkeypatch.setattr(DLH, "_lp_solve", highs_lp_solve)
    _check_bidders(highs_lp_solve)


def test_host_rolling_horizon_batch(monkeypatch, tmp_path):
    """three simulations with different battery sizes advanced four tracking steps: state hand-over and the result table"""
    monkeypatch.setattr(DLH, "_lp_solve", highs_lp_solve)
    md = DLH.RenewableGeneratorModelData("309_WIND_1", "Carter", 0, 200)
    m = DLH.MultiPeriodWindBattery(md, np.tile(CF, 2), 200.0, np.array([5.0, 25.0, 50.0]), np.array([20.0, 100.0, 200.0]))
    assert m.N == 3
    tr = DLH.Tracker(m, tracking_horizon=4, n_tracking_hour=1)
    disp = np.array([0.5, 1.0, 10.0, 20.0])
    for h in range(4):
        soc_before = tr.fs.soc0.copy()
        prof = tr.track_market_dispatch(disp, date="2020-01-02", hour=h)
        # the SoC fixed for the next horizon is this horizon's first-hour SoC rounded to 2 decimals (double_loop.py:189-191)
        assert np.allclose(tr.fs.soc0, np.round(prof["realized_soc"][-1], 2))
        assert tr.fs._time_idx == h + 1 and np.allclose(tr.fs.cf[0], np.tile(CF, 2)[h + 1:h + 5])
        first = tr.tracking_model_object.result_list[-1]
        assert len(first) == 3 * 4 and set(first["Simulation"]) == {0, 1, 2}
        # energy balance of the implemented hour: soc = soc0 + 0.95 in - out/0.95
        step = tr.fs.sol["soc"][:, 0] * 0 + prof["realized_soc"][-1] - soc_before
        assert np.all(np.abs(step) <= 0.95 * m._battery_pmax_mw * 1e3 + 1e-6)
    m.write_results(tmp_path / "tracker.csv")
    import pandas as pd
    df = pd.read_csv(tmp_path / "tracker.csv")
    assert list(df.columns[:4]) == ["Generator", "Date", "Hour", "Horizon [hr]"] and len(df) == 4 * 12
    assert {"Total Wind Generation [MW]", "Total Power Output [MW]", "Wind Power Output [MW]", "Wind Curtailment [MW]",
            "Battery Power Output [MW]", "Wind Power to Battery [MW]", "State of Charge [MWh]", "Total Cost [$]"} <= set(df.columns)


@pytest.mark.parametrize("T", [4, 48])
def test_nuclear_operation_templates_match_raw_oracle(T):
    rng = np.random.default_rng(T)
    da = rng.uniform(5, 60, T); rt = da + rng.normal(0, 8, T); h0 = 812345.0; disp = rng.uniform(380, 520, T)
    rp = np.r_[h0, disp]
    a, _ = solve_template(TP.nuclear_operation(T, "tracker"), np.array([4.0]), rp)
    b, _ = H.solve(DL.nuclear_tracker_raw(disp, h0))
    assert a == pytest.approx(b, rel=1e-10, abs=1e-7)
    a, _ = solve_template(TP.nuclear_operation(T, "bidder_da"), np.r_[da, rt, 4.0], rp)
    b, _ = H.solve(DL.nuclear_bidder_raw(da, rt, h0))
    assert a == pytest.approx(b, rel=1e-10, abs=1e-7)
    a, _ = solve_template(TP.nuclear_operation(T, "bidder_rt"), np.r_[da, rt, 4.0], rp)
    b, _ = H.solve(DL.nuclear_bidder_raw(da, rt, h0, da_dispatch=disp))
    assert a - float(np.sum((da - rt) * disp)) == pytest.approx(b, rel=1e-10, abs=1e-6)


def _check_nuclear(n_sim):
    """MultiPeriodNuclear under the Tracker and the SelfScheduler: objectives against the raw oracle, state hand-over."""
    rng = np.random.default_rng(5)
    md = DLH.ThermalGeneratorModelData(gen_name="121_NUCLEAR_1", bus="Attlee", p_min=400, p_max=500)
    m = DLH.MultiPeriodNuclear(md, n_sim=n_sim)
    tr = DLH.Tracker(m, tracking_horizon=4, n_tracking_hour=1)
    holdup = np.zeros(n_sim)
    for h in range(3):
        disp = rng.uniform(400, 500, (n_sim, 4))
        prof = tr.track_market_dispatch(disp, date="2020-07-10", hour=h)
        for k in range(min(n_sim, 4)):
            ref, _ = H.solve(DL.nuclear_tracker_raw(disp[k], holdup[k]))
            assert tr.objective[k] == pytest.approx(ref, rel=1e-7, abs=1e-5)
        assert np.allclose(tr.power_output[:, 0], disp[:, 0], atol=1e-3)                  # hard-tracked hour
        holdup = np.round(prof["implemented_tank_holdup"][-1])
        assert np.array_equal(tr.fs.holdup0, holdup)                                       # :232-235 integer rounding
    bc = DLH.Backcaster({"Attlee": rng.uniform(5, 60, 48)}, {"Attlee": rng.uniform(5, 60, 48)})
    ss = DLH.SelfScheduler(DLH.MultiPeriodNuclear(md, n_sim=n_sim), day_ahead_horizon=48, real_time_horizon=4, n_scenario=1, forecaster=bc)
    bids = ss.compute_day_ahead_bids(date="2020-07-10")
    da = bc.forecast_day_ahead_prices("d", 0, "Attlee", 48, 1)[0]; rt = bc.forecast_real_time_prices("d", 0, "Attlee", 48, 1)[0]
    ref, _ = H.solve(DL.nuclear_bidder_raw(da, rt, 0.0))
    assert -ss.day_ahead_model.objective[0] == pytest.approx(ref, rel=1e-8)
    pmax = np.array([np.atleast_1d(b["121_NUCLEAR_1"]["p_max"])[0] for b in bids.values()])
    # day-ahead quantity: all of P_T where da > rt, nothing where da < rt; P_T in [400, 500] (NPP 500 MW, PEM <= 100 MW)
    assert np.all((pmax < 1e-3) | ((pmax >= 400 - 1e-3) & (pmax <= 500 + 1e-3)))
    assert np.all(pmax[da > rt + 1e-9] >= 400 - 1e-3) and np.all(pmax[da < rt - 1e-9] < 1e-3)
    df = ss.bidding_model_object.result_list[0]
    assert {"Power to Grid [MW]", "Power to PEM [MW]", "Initial holdup [kg]", "Final holdup [kg]", "Hydrogen Market [kg/hr]",
            "Total Cost [$]"} <= set(df.columns)


def test_host_nuclear_logic(monkeypatch):
    monkeypatch.setattr(DLH, "_lp_solve", highs_lp_solve)
    _check_nuclear(2)


def test_oracle_and_host_wind_pem_tracker_known_answer(monkeypatch):
    """wind + PEM tracker against the reference's known answers (tests/test_wind_PEM_double_loop.py:55-121)."""
    g = G["wind_pem_tracker"]
    lp = DL.wind_pem_tracker_raw(g["market_dispatch"], CF[:4], G["wind_pmax_mw"], g["pem_pmax_mw"])
    ref_obj, x = H.solve(lp)
    v = lp.meta["v"]
    wind = np.array([x[v["w", t]] for t in range(4)]); pem = np.array([x[v["pe", t]] for t in range(4)])
    assert CF[0] == pytest.approx(0.00562, rel=1e-3)                                                     # :77-78
    assert wind == pytest.approx(g["expected_wind_power"], rel=g["rel"])                                 # :93-98
    assert np.array([x[v["g", t]] for t in range(4)]) * 1e-3 == pytest.approx(g["market_dispatch"], abs=g["abs_power"])
    assert pem == pytest.approx(np.array(g["expected_wind_power"]) - 1e3 * np.array(g["market_dispatch"]), rel=g["rel"])   # :112-119
    # reduced template and host class (HiGHS stand-in for the CUDA solve)
    monkeypatch.setattr(DLH, "_lp_solve", highs_lp_solve)
    md = DLH.RenewableGeneratorModelData("309_WIND_1", "Carter", 0, 200)
    m = DLH.MultiPeriodWindPEM(md, np.tile(CF, 2), wind_pmax_mw=200, pem_pmax_mw=25)
    tr = DLH.Tracker(m, tracking_horizon=4, n_tracking_hour=1)
    tr.track_market_dispatch(g["market_dispatch"], date="2020-01-02", hour="00:00")
    assert tr.objective[0] == pytest.approx(ref_obj, rel=1e-9)
    assert tr.fs.sol["pem"][0] == pytest.approx(pem, rel=1e-6, abs=1e-3)
    assert np.all(np.abs(tr.fs.wind_waste) < 1e-3)                                                       # :99-102
    assert tr.power_output[0] == pytest.approx(g["market_dispatch"], abs=g["abs_power"])
    assert tr.fs._time_idx == 1 and np.allclose(tr.fs.cf[0], CF[1:5])
    assert "Wind to PEM [MW]" in m.result_list[0].columns


def test_backcaster_order():
    bc = DLH.Backcaster({"b": np.arange(72.0)}, {"b": np.arange(72.0)})
    f = bc.forecast_day_ahead_prices("d", 0, "b", 48, 2)
    assert f.shape == (2, 48) and f[0, 0] == 48 and f[0, 24] == 24 and f[1, 0] == 24 and f[1, 24] == 0
    assert np.array_equal(f, DL.backcast(np.arange(72.0), 0, 48, 2))
    with pytest.raises(ValueError):
        DLH.Backcaster({"b": [1.0] * 5}, {"b": [1.0] * 30})


# ---------------------------------------------------------------- CUDA path
@pytest.mark.gpu
def test_gpu_tracker_known_answer():
    _check_tracker(None)


@pytest.mark.gpu
def test_gpu_wind_pem_tracker_known_answer():
    """wind + PEM tracker (wind_PEM_double_loop.py:103-330) through the CUDA path against the reference's known answers
    (tests/test_wind_PEM_double_loop.py:55-121) and the raw oracle LP."""
    g = G["wind_pem_tracker"]
    ref_obj, x = H.solve(DL.wind_pem_tracker_raw(g["market_dispatch"], CF[:4], G["wind_pmax_mw"], g["pem_pmax_mw"]))
    md = DLH.RenewableGeneratorModelData("309_WIND_1", "Carter", 0, 200)
    m = DLH.MultiPeriodWindPEM(md, np.tile(CF, 2), wind_pmax_mw=200, pem_pmax_mw=25)
    tr = DLH.Tracker(m, tracking_horizon=4, n_tracking_hour=1)
    tr.track_market_dispatch(g["market_dispatch"], date="2020-01-02", hour="00:00")
    assert tr.objective[0] == pytest.approx(ref_obj, rel=1e-7, abs=1e-6)
    pem = np.array(g["expected_wind_power"]) - 1e3 * np.array(g["market_dispatch"])                  # :112-119
    assert tr.fs.sol["pem"][0] == pytest.approx(pem, rel=g["rel"], abs=1.0)
    assert (tr.fs.sol["pem"][0] + 1e3 * tr.power_output[0]) == pytest.approx(g["expected_wind_power"], rel=g["rel"])   # :93-98
    assert np.all(np.abs(tr.fs.wind_waste) < 1.0)                                                    # kW; :99-102
    assert tr.power_output[0] == pytest.approx(g["market_dispatch"], abs=g["abs_power"])
    # a batch of 64 lock-step simulations with different sizes / series: objective against the raw oracle LP
    rng = np.random.default_rng(3)
    N = 64
    series = np.array([np.roll(np.tile(CF, 2), -int(k)) for k in rng.integers(0, 48, N)])
    wind = rng.uniform(100, 400, N); pemmw = rng.uniform(5, 60, N)
    mb = DLH.MultiPeriodWindPEM(md, series, wind_pmax_mw=wind, pem_pmax_mw=pemmw)
    trb = DLH.Tracker(mb, tracking_horizon=4, n_tracking_hour=1)
    cf0 = trb.fs.cf.copy()
    disp = rng.uniform(0, 1, (N, 4)) * wind[:, None] * cf0[:, :4] * 0.9
    trb.track_market_dispatch(disp, date="d", hour=0)
    for k in rng.choice(N, 16, replace=False):
        ref, _ = H.solve(DL.wind_pem_tracker_raw(disp[k], cf0[k, :4], wind[k], pemmw[k]))
        assert trb.objective[k] == pytest.approx(ref, rel=1e-7, abs=1e-5)


@pytest.mark.gpu
def test_gpu_bidders_known_answers():
    _check_bidders(None)


@pytest.mark.gpu
def test_gpu_nuclear_double_loop():
    _check_nuclear(64)


@pytest.mark.gpu
def test_gpu_double_loop_batch_parity():
    """A batch of 256 rolling-horizon simulations (different sizes and wind series) advanced 3 tracking steps: every LP
    objective against HiGHS on the raw oracle LP."""
    rng = np.random.default_rng(7)
    N = 256
    series = np.array([np.roll(np.tile(CF, 2), -int(k)) for k in rng.integers(0, 48, N)])
    wind = rng.uniform(100, 400, N); batt = rng.uniform(5, 60, N); energy = 4 * batt
    md = DLH.RenewableGeneratorModelData("g", "b", 0, 400)
    m = DLH.MultiPeriodWindBattery(md, series, wind, batt, energy)
    tr = DLH.Tracker(m, tracking_horizon=24, n_tracking_hour=1)
    for step in range(3):
        soc0, thr0, cf = tr.fs.soc0.copy(), tr.fs.thr0.copy(), tr.fs.cf.copy()
        disp = rng.uniform(0, 1, (N, 24)) * wind[:, None] * 0.5
        tr.track_market_dispatch(disp, date="d", hour=step)
        for k in rng.choice(N, 12, replace=False):
            ref, _ = H.solve(DL.tracker_raw(disp[k], cf[k], wind[k], batt[k], energy[k], soc0[k], thr0[k]))
            assert tr.objective[k] == pytest.approx(ref, rel=1e-7, abs=1e-5)
    bc = [DLH.Backcaster({"b": rng.uniform(5, 80, 72)}, {"b": rng.uniform(5, 80, 72)}) for _ in range(N)]
    ss = DLH.SelfScheduler(m, day_ahead_horizon=48, real_time_horizon=4, n_scenario=1, forecaster=bc)
    ss.compute_day_ahead_bids(date="d")
    for k in rng.choice(N, 12, replace=False):
        da = bc[k].forecast_day_ahead_prices("d", 0, "b", 48, 1)[0]; rt = bc[k].forecast_real_time_prices("d", 0, "b", 48, 1)[0]
        ref, _ = H.solve(DL.bidder_raw(da, rt, series[k, :48], wind[k], batt[k], energy[k]))
        assert -ss.day_ahead_model.objective[k] == pytest.approx(ref, rel=1e-7, abs=1e-5)
