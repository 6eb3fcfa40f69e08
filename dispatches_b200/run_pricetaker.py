"""Batched sweep drivers with the reference's result formats.

Mirrors  case_studies/renewables_case/run_pricetaker_wind_battery.py:37-70 (`run_design`: one JSON per design point,
skip if the file exists)  and  run_pricetaker_wind_PEM.py:27-110 (`run_design` per (h2_price, pem_ratio), results table
written with `pd.DataFrame(res).to_csv`).  Where the reference maps `run_design` over a `multiprocessing.Pool`, one
call here solves every (design point x LMP signal) on the GPU; the files it leaves behind have the same names/keys,
so downstream readers are untouched.  Any horizon works: n_time_points <= 32 runs on the stage kernel, longer ones on
the band kernel (shared memory up to a few hundred periods, a global workspace beyond -- the reference's full-year
n_time_points = 8736 LP takes about 2 s per warp, all design points in parallel).
"""
from __future__ import annotations

import json
from itertools import product
from pathlib import Path

import numpy as np

from . import pricetaker as PT


def run_wind_battery_sweep(wind_sizes, battery_ratios, lmp, cf, n_time_points=24, out_dir=None, market="DA",
                           extant_wind=True):
    """All design points of run_pricetaker_wind_battery.run_design in one batch.

    lmp [S, >=T] ($/MWh) and cf [>=T] or [S, >=T]: S price signals per design point.  Returns a list of result dicts
    (one per design point, NPV / annual revenue averaged over the S signals exactly as a caller averaging the
    reference's per-signal JSONs would) and writes result_<market>_wind_<w>_battery_<r>.json when out_dir is given;
    existing files are reused (the reference's resume behaviour, :44-54)."""
    T = int(n_time_points)
    lmp = np.atleast_2d(np.asarray(lmp, float))[:, :T]
    cf = np.asarray(cf, float)
    cf = np.broadcast_to(cf[..., :T], lmp.shape)
    S = lmp.shape[0]
    out_dir = Path(out_dir) if out_dir else None
    designs = list(product(wind_sizes, battery_ratios))
    todo, results = [], {}
    for w, r in designs:
        f = out_dir / f"result_{market}_wind_{w}_battery_{r}.json" if out_dir else None
        if f is not None and f.exists():
            results[(w, r)] = json.load(open(f))
        else:
            todo.append((w, r))
    if todo:
        W = np.repeat([w for w, _ in todo], S)
        B = np.repeat([w * r for w, r in todo], S)
        params = {"wind_mw": W, "batt_mw": B, "design_opt": False, "extant_wind": extant_wind,
                  "wind_resource": np.tile(cf, (len(todo), 1)), "DA_LMPs": np.tile(lmp, (len(todo), 1))}
        res = PT.wind_battery_optimize(T, params, want_solution=True)
        npv = res.NPV.reshape(len(todo), S); rev = res.annual_revenue.reshape(len(todo), S)
        ok = (res.status == 0).reshape(len(todo), S).all(1)
        for k, (w, r) in enumerate(todo):
            d = {"wind_mw": float(w), "batt_mw": float(w * r), "NPV": float(npv[k].mean()),
                 "annual revenue": float(rev[k].mean()), "n_signals": int(S),
                 "termination_condition": "optimal" if ok[k] else "non-optimal"}
            results[(w, r)] = d
            if out_dir:
                out_dir.mkdir(parents=True, exist_ok=True)
                json.dump(d, open(out_dir / f"result_{market}_wind_{w}_battery_{r}.json", "w"))
    return [results[d] for d in designs]


def run_wind_pem_sweep(h2_prices, pem_ratios, lmp, cf, wind_mw=847.0, n_time_points=24, csv_path=None):
    """run_pricetaker_wind_PEM.py:100-110 for fixed PEM sizes (design_opt False, batt_mw = 0): the table with the
    reference's columns wind_mw, batt_mw, pem_mw, h2_price_per_kg, annual_rev_h2, annual_rev_E, NPV."""
    T = int(n_time_points)
    lmp = np.atleast_2d(np.asarray(lmp, float))[:, :T]
    S = lmp.shape[0]
    cf = np.broadcast_to(np.asarray(cf, float)[..., :T], lmp.shape)
    designs = [(h, p) for h, p in product(h2_prices, pem_ratios) if p is not None and p > 0]
    D = len(designs)
    params = {"wind_mw": wind_mw, "batt_mw": 0.0, "pem_mw": np.repeat([p * wind_mw for _, p in designs], S),
              "h2_price_per_kg": np.repeat([h for h, _ in designs], S), "design_opt": False, "extant_wind": True,
              "wind_resource": np.tile(cf, (D, 1)), "DA_LMPs": np.tile(lmp, (D, 1))}
    res = PT.wind_battery_pem_optimize(T, params, want_solution=True)
    rows = []
    for k, (h, p) in enumerate(designs):
        sl = slice(k * S, (k + 1) * S)
        rows.append({"wind_mw": wind_mw, "batt_mw": 0.0, "pem_mw": p * wind_mw, "h2_price_per_kg": h,
                     "annual_rev_h2": float(res.annual_rev_h2[sl].mean()), "annual_rev_E": float(res.annual_elec_revenue[sl].mean()),
                     "NPV": float(res.NPV[sl].mean())})
    if csv_path:
        import pandas as pd
        pd.DataFrame(rows).to_csv(csv_path)
    return rows


def run_exhaustive_enumeration(lmp, pem_capex=400.0, h2_prices=(0.75, 1, 1.25, 1.5, 1.75, 2),
                               pem_fractions=tuple(i / 100 for i in range(5, 51, 5)), json_path=None,
                               plant_life=30, tax_rate=0.2, discount_rate=0.08, tank_capacity=0.0, turbine_capacity=0.0,
                               demand=400.0 * 20, schedule_csv_dir=None):
    """nuclear_case/report/price_taker_analysis.py:353-425 (market variants V1-V3: one LMP series).

    tank_capacity = turbine_capacity = 0 (the report's sweep, :377-379): the tank balance forces holdup = 0 in every hour,
    so the 8784-period LP of build_deterministic_model (:181-222) separates into one-variable LPs per hour,
        max over e in [0, pem_cap]:  (1 - tax) * (20 * h2_price - lmp_t) * e ,
    whose solution is e = pem_cap where 20 h2_price > lmp_t (20 kg/MWh, :42; demand bound 400*20 never binds, vom_pem=0) -- an
    exact presolve instead of 60 Gurobi solves of a 87 840-column LP.
    With a tank and / or turbine the hours couple through the holdup: all |h2_prices| x |pem_fractions| LPs (T = len(lmp)
    periods each) go to the CUDA solver in ONE batch (templates.nuclear_report; T = 8784 runs the band kernel in workspace mode).
    The result dict has the reference's keys (in M$, :405-425); ``schedule_csv_dir`` additionally writes the hourly schedule of
    every design point the way _write_results does (:325-350)."""
    lmp = np.asarray(lmp, float)
    T = lmp.size
    k = 1.0 - tax_rate
    cf = (1.0 - (1.0 + discount_rate) ** (-plant_life)) / discount_rate
    res = {"h2_price": list(h2_prices), "pem_cap": list(pem_fractions), "solver_stat": {}, "elec_rev": {}, "h2_rev": {},
           "net_npv": {}, "net_profit": {}, "pem_cap_factor": {}}
    designs = [(i1, hp, i2, pc) for i1, hp in enumerate(h2_prices) for i2, pc in enumerate(pem_fractions)]
    sched = {}
    if tank_capacity > 0.0 or turbine_capacity > 0.0:
        from . import templates as TP
        from .solver import BatchLPSolver, STATUS_NAMES
        t = TP.nuclear_report(T, pem_capex=pem_capex, demand=demand, plant_life=plant_life, tax_rate=tax_rate, discount_rate=discount_rate)
        sol = BatchLPSolver(t)
        D = len(designs)
        cp = np.concatenate([np.tile(lmp, (D, 1)), np.array([[hp] for _, hp, _, _ in designs])], axis=1)
        rp = np.array([[pc * 400.0, tank_capacity, turbine_capacity] for _, _, _, pc in designs])
        r = sol.solve_host(cp, rp, want_x=True)
        xm = sol.to_model_space(r.x)
        names = {nm: j for j, nm in enumerate(t.col_names)}
        col = lambda v: np.array([names[f"period[{h + 1}].fs.{v}"] for h in range(T)])
        E, U, TB, HH = xm[:, col("np_to_electrolyzer")], xm[:, col("h2_to_pipeline")], xm[:, col("h2_to_turbine")], xm[:, col("tank_holdup")]
        sol.close()
    for d, (i1, hp, i2, pc) in enumerate(designs):
        cap = pc * 400.0
        if tank_capacity > 0.0 or turbine_capacity > 0.0:
            e, u, tb, hold = E[d], U[d], TB[d], HH[d]
            stat = STATUS_NAMES[int(r.status[d])]
        else:
            e = np.where(20.0 * hp > lmp, cap, 0.0)
            u, tb, hold = 20.0 * e, np.zeros(T), np.zeros(T)
            stat = "optimal"
        net_power = 400.0 - e + 0.0125 * tb
        elec = float(np.sum(lmp * net_power)); h2 = float(np.sum(hp * u))
        cash = h2 + elec - float(np.sum(4.25 * 0.0125 * tb)) - 2.3 * 400.0 * T
        capex = pem_capex * 1000.0 * cap + 29.0 * 33.3 * tank_capacity + 947.0 * 1000.0 * turbine_capacity
        fom = 1000.0 * 0.03 * pem_capex * cap + 1000.0 * 7.0 * turbine_capacity + 120.0 * 1000.0 * 400.0
        dep = capex / plant_life
        profit = dep + k * (cash - fom - dep)
        key = str(i1) + str(i2)
        res["elec_rev"][key] = elec / 1e6; res["h2_rev"][key] = h2 / 1e6
        res["net_profit"][key] = profit / 1e6; res["net_npv"][key] = (profit - capex / cf) / 1e6
        res["solver_stat"][key] = stat; res["pem_cap_factor"][key] = float(e.sum() / (cap * T))
        vom = 0.0 * e + 4.25 * 0.0125 * tb + 2.3 * 400.0                      # :239-241 with vom_pem = 0 (:395)
        sched[key] = {"LMP [$/MWh]": lmp, "np_to_grid": 400.0 - e, "np_to_electrolyzer": e,
                      "tank_holdup_previous": np.r_[0.0, hold[:-1]], "tank_holdup": hold, "h2_to_pipeline": u, "h2_to_turbine": tb,
                      "h2_turbine_power": 0.0125 * tb, "h2_revenue": hp * u, "electricity_revenue": lmp * net_power, "vom": vom,
                      "net_cash_inflow": hp * u + lmp * net_power - vom}
    res["capex"] = capex / 1e6; res["fom"] = fom / 1e6
    if json_path:
        json.dump(res, open(json_path, "w"), indent=4)
    if schedule_csv_dir:
        write_schedules(sched, schedule_csv_dir)
    return res


def write_schedules(sched, out_dir, prefix="results"):
    """_write_results of the report (price_taker_analysis.py:325-350): `<filename>_schedule.csv` with the reference's columns
    (LMP [$/MWh], np_to_grid, np_to_electrolyzer, tank_holdup_previous, tank_holdup, h2_to_pipeline, h2_to_turbine,
    h2_turbine_power, h2_revenue, electricity_revenue, vom, net_cash_inflow), one file per design point of the enumeration."""
    import pandas as pd
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    paths = []
    for key, cols in sched.items():
        f = out_dir / f"{prefix}_{key}_schedule.csv"
        pd.DataFrame(cols).to_csv(f)
        paths.append(f)
    return paths


def write_simulation_data(dispatch_mw, inputs, dispatch_csv, input_h5=None, input_columns=None):
    """Sweep outputs in the layout the reference's surrogate-training reader consumes
    (workflow/train_market_surrogates/dynamic/Simulation_Data.py:138-220; sample: dynamic/tests/data/simdatatest.csv):

      dispatch_csv : one row per simulation, index ``run_<i>``, columns ``0 .. H-1`` = hourly dispatch in MW
                     (``_read_data_to_array`` drops the first column and parses the run number out of ``run_<i>``);
      input_h5     : the sweep's input parameters as a DataFrame whose first column is the run index (``read_data_to_dict`` does
                     ``pd.read_hdf(...).iloc[index, 1:]``).  HDF needs pytables, which this image lacks: without it the same
                     frame is written next to it as ``<name>.csv`` and the function says so in its return value.
    dispatch_mw [N, H] comes from the batched double-loop / price-taker runs (e.g. Tracker.power_output history)."""
    import pandas as pd
    dispatch_mw = np.atleast_2d(np.asarray(dispatch_mw, float))
    N, Hh = dispatch_mw.shape
    df = pd.DataFrame(dispatch_mw, index=[f"run_{i}" for i in range(N)], columns=[str(h) for h in range(Hh)])
    df.to_csv(dispatch_csv)
    out = {"dispatch_csv": str(dispatch_csv), "input_file": None, "input_format": None}
    if input_h5 is not None:
        inputs = np.atleast_2d(np.asarray(inputs, float))
        cols = list(input_columns) if input_columns else [f"x{k}" for k in range(inputs.shape[1])]
        dfi = pd.DataFrame(np.column_stack([np.arange(N), inputs]), columns=["index"] + cols)
        try:
            dfi.to_hdf(input_h5, key="df", mode="w")
            out.update(input_file=str(input_h5), input_format="hdf")
        except ImportError:
            alt = Path(str(input_h5)).with_suffix(".csv")
            dfi.to_csv(alt, index=False)
            out.update(input_file=str(alt), input_format="csv (pytables missing: convert with DataFrame.to_hdf where it exists)")
    return out
