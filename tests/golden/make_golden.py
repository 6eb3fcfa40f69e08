"""Regenerates tests/golden/* from the reference's committed DATA files (run HERE, in the dev container;
/root/reference does not exist on the GPU box, so the outputs are committed).

    python tests/golden/make_golden.py

Writes
  ../../dispatches_b200/data/lmp_pool.npz -- real price/capacity-factor series the synthetic scenario batches are drawn from
                           (SURVEY.md §8(d)): day windows of Wind_Thermal_Dispatch.csv DA/RT LMPs at buses
                           122/303/309/317 (load_parameters.py:82-112), the 3100 cluster days of
                           nuclear_case/lmp_signal.json, the 8736-h 303_DALMP / 303_WIND_1-DACF series,
                           the fossil 24-h / 168-h LMP vectors are NOT copied (source literals).
  wind_pem_golden.json  -- the reference's own committed result tables for the wind+PEM price-taker sweep
                           (wind_PEM/wind_PEM_RT_1000.csv, design_wind_PEM_results.csv,
                           design_wind_PEM_RT_results.csv: h2_price x pem_ratio -> annual_rev_h2, NPV) (the 8784-h parquet
                           price / capacity-factor series they were computed on go into lmp_pool.npz as pq1000_* / pq500_*).
  unit_kats.json        -- known answers of the reference's unit-model tests for the battery rows
                           (unit_models/tests/test_battery.py:40-67, :95-119).
  double_loop_golden.json -- inputs and known answers of the reference's double-loop tests
                           (case_studies/renewables_case/tests/test_multiperiod_wind_battery_doubleloop.py:78-111 tracker,
                           :168-175 self-schedule bids, :245-252 thermal-generator bid prices): the first 48 h of
                           309_WIND_1-RTCF / 309_DALMP / 309_RTLMP and the asserted vectors (read with ``ast``); the wind + PEM
                           tracker's known answers (tests/test_wind_PEM_double_loop.py:55-121).
  solar_golden.json     -- known answers of the reference's PV + battery + hydrogen price-taker test
                           (case_studies/renewables_case/tests/test_solar_battery_hydrogen.py:20-48: capital cost, NPV, optimal sizes,
                           every ``pytest.approx`` with its tolerance, read with ``ast``) and the 24 prices the case runs on
                           (load_parameters.py:62-88: 303 DA LMPs after the reference's date filter, first day).
No reference SOURCE is copied: only numeric data and test constants.
"""
import ast
import json
from pathlib import Path

import numpy as np
import pandas as pd

REF = Path("/root/reference/dispatches")
OUT = Path(__file__).parent


def double_loop():
    rc = REF / "case_studies" / "renewables_case"
    df = pd.read_csv(rc / "data" / "Wind_Thermal_Dispatch.csv")
    tree = ast.parse(open(rc / "tests" / "test_multiperiod_wind_battery_doubleloop.py").read())
    lists = {}
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and isinstance(node.value, ast.List):
                try:
                    lists[fn.name, node.targets[0].id] = [float(v) for v in ast.literal_eval(node.value)]
                except (ValueError, TypeError):
                    pass
    tree2 = ast.parse(open(rc / "tests" / "test_wind_PEM_double_loop.py").read())
    for fn in [n for n in tree2.body if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and isinstance(node.value, ast.List):
                try:
                    lists["wind_pem", fn.name, node.targets[0].id] = [float(v) for v in ast.literal_eval(node.value)]
                except (ValueError, TypeError):
                    pass
    gold = dict(
        # test_wind_PEM_double_loop.py:55-121: tracker of the wind + PEM model (200 MW wind, PEM 25 MW)
        wind_pem_tracker=dict(market_dispatch=lists["wind_pem", "test_track_market_dispatch", "market_dispatch"],
                              expected_wind_power=lists["wind_pem", "test_track_market_dispatch", "expected_wind_power"],
                              rel=1e-3, abs_power=1e-3, tracking_horizon=4, n_tracking_hour=1, pem_pmax_mw=25.0),
        cf_309_rt=df["309_WIND_1-RTCF"].values[:48].tolist(),
        da_309=df["309_DALMP"].values[:48].tolist(), rt_309=df["309_RTLMP"].values[:48].tolist(),
        wind_pmax_mw=200.0, battery_pmax_mw=25.0, battery_energy_capacity_mwh=100.0,
        tracker=dict(market_dispatch=lists["test_track_market_dispatch", "market_dispatch"],
                     expected_wind_power=lists["test_track_market_dispatch", "expected_wind_power"],
                     rel=1e-3, abs_power=1e-3, tracking_horizon=4, n_tracking_hour=1),
        self_schedule=dict(known_solution=lists["test_compute_bids_self_schedule", "known_solution"], reltol=1e-2,
                           day_ahead_horizon=48, n_scenario=1),
        thermal_bid=dict(known_solution=lists["test_compute_bids_thermal_gen", "known_solution"], reltol=1e-2))
    json.dump(gold, open(OUT / "double_loop_golden.json", "w"))


def solar():
    rc = REF / "case_studies" / "renewables_case"
    tree = ast.parse(open(rc / "tests" / "test_solar_battery_hydrogen.py").read())
    gold = {}
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        checks = {}
        for node in ast.walk(fn):
            # assert des_res['key'] == pytest.approx(value, rel=.. | abs=..)
            if isinstance(node, ast.Assert) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Subscript):
                key = ast.literal_eval(node.test.left.slice)
                call = node.test.comparators[0]
                val = float(ast.literal_eval(call.args[0]))
                tol = {k.arg: float(ast.literal_eval(k.value)) for k in call.keywords}
                checks[key] = dict(value=val, **tol)
        overrides = {}
        for node in ast.walk(fn):
            if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Subscript) and isinstance(node.targets[0].value, ast.Name) \
                    and node.targets[0].value.id == "params":
                overrides[ast.literal_eval(node.targets[0].slice)] = ast.literal_eval(node.value)
        gold[fn.name] = dict(params=overrides, expect=checks)
    # the price series of the case: load_parameters.py:62-88 (date filter, bus 303, DA market), first 24 hours
    df = pd.read_csv(rc / "data" / "Wind_Thermal_Dispatch.csv")
    df.index = pd.to_datetime(df["DateTime"])
    start = pd.Timestamp("2020-01-02 00:00:00")
    ix = pd.date_range(start=start, end=start + pd.offsets.DateOffset(days=365) - pd.offsets.DateOffset(hours=1), freq="1h")
    ix = ix[(ix.day != 29) | (ix.month != 2)]
    df = df[df.index.isin(ix)]
    gold["lmp_24"] = [float(v) for v in df["303_DALMP"].values[:24]]
    json.dump(gold, open(OUT / "solar_golden.json", "w"), indent=1)
    return gold


def main():
    double_loop()
    solar()
    rc = REF / "case_studies" / "renewables_case"
    df = pd.read_csv(rc / "data" / "Wind_Thermal_Dispatch.csv")
    assert len(df) == 8736
    windows = []
    for bus in ("122", "303", "309", "317"):
        for mk in ("DA", "RT"):
            windows.append(df[f"{bus}_{mk}LMP"].values.reshape(364, 24))
    day_windows = np.concatenate(windows)                      # 2912 x 24
    j = json.load(open(REF / "case_studies" / "nuclear_case" / "lmp_signal.json"))
    cl = []
    for sc in sorted(j, key=int):
        for yr in sorted(j[sc], key=int):
            for d in sorted(j[sc][yr], key=int):
                cl.append([j[sc][yr][d][str(h)] for h in range(1, 25)])
    cluster_days = np.array(cl, float)                         # 3100 x 24
    assert cluster_days.shape == (3100, 24)
    extra = {}
    for shortfall in (1000, 500):
        p = pd.read_parquet(rc / "data" / f"303_LMPs_15_reserve_{shortfall}_shortfall.parquet")
        for col, key in (("LMP", "rt_lmp"), ("LMP DA", "da_lmp"), ("303_WIND_1-RTCF", "rt_cf"), ("303_WIND_1-DACF", "da_cf")):
            extra[f"pq{shortfall}_{key}"] = p[col].values
    nr = pd.read_csv(REF / "case_studies" / "nuclear_case" / "report" / "rts_gmlc_15_500.csv")
    extra["nuc_report_lmp_rt"] = nr["LMP"].values          # get_lmp_data, price_taker_analysis.py:45-113
    extra["nuc_report_lmp_da"] = nr["LMP DA"].values
    np.savez_compressed(OUT.parent.parent / "dispatches_b200" / "data" / "lmp_pool.npz", day_windows=day_windows, cluster_days=cluster_days,
                        dalmp_303=df["303_DALMP"].values, dacf_303=df["303_WIND_1-DACF"].values, **extra)

    gold = {}
    for name in ("wind_PEM_RT_1000", "design_wind_PEM_results", "design_wind_PEM_RT_results"):
        t = pd.read_csv(rc / "wind_PEM" / f"{name}.csv")
        keep = ["wind_mw", "batt_mw", "pem_mw", "h2_price_per_kg", "annual_rev_h2", "annual_rev_E", "NPV"]
        gold[name] = t[[k for k in keep if k in t.columns]].to_dict(orient="list")
    json.dump(gold, open(OUT / "wind_pem_golden.json", "w"))

    kats = {
        # test_battery.py:40-67: elec_in = 5 kW for dt = 1 h from empty -> SoC 4.75 kWh, throughput 2.5 kWh
        "battery_charge": dict(elec_in=5.0, elec_out=0.0, soc0=0.0, thr0=0.0, soc=4.75, throughput=2.5),
        # test_battery.py:95-119: soc0 = 5, thr0 = 5, elec_out fixed 5, state_of_charge fixed 0
        #   -> the reference asserts energy_throughput == approx(7.638, rel=1e-3)
        "battery_discharge": dict(soc0=5.0, thr0=5.0, elec_out=5.0, soc=0.0, throughput=7.638, rel=1e-3),
    }
    json.dump(kats, open(OUT / "unit_kats.json", "w"), indent=1)
    print("wrote", [p.name for p in OUT.iterdir()])


if __name__ == "__main__":
    import sys
    if sys.argv[1:] == ["double_loop"]:
        double_loop()
    elif sys.argv[1:] == ["solar"]:
        print(solar())
    else:
        main()
