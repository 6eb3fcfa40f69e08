"""CPU ORACLE (test infrastructure -- NOT product code; only tests/, bench.py's CPU legs and
__graft_entry__.smoke() may import this).

Raw restatements of the double-loop (bidding / tracking) LPs that sit next to the price-taker path
(SURVEY.md §8(f)-2).  Reference: case_studies/renewables_case/wind_battery_double_loop.py
  :27-51    create_multiperiod_wind_battery_model (the same period blocks and link pairs as the price-taker)
  :54-83    transform_design_model_to_operation_model (sizes fixed, initial SoC fixed, periodic row deactivated)
  :160-171  P_T, wind_waste, tot_cost expressions
  :175-206  update_model (realised SoC / throughput rounded to 2 decimals and fixed; capacity factors shifted)

The objective / extra rows come from idaes-pse 2.0 (``idaes.apps.grid_integration``: Tracker, SelfScheduler, Bidder,
Backcaster), which is NOT vendored under /root/reference; they are restated from its published formulation and
anchored on the reference's own known-answer tests
(case_studies/renewables_case/tests/test_multiperiod_wind_battery_doubleloop.py:78-111, :168-175, :245-252 ->
tests/golden/double_loop_golden.json, made by tests/golden/make_golden.py):

  Tracker:        min  sum_t [ tot_cost_t + pen_t (under_t + over_t) ]
                  s.t. P_T[t] + under_t == dispatch_t + over_t ;  pen_t = 1e4 (t < n_tracking_hour) else 1e4/(H - n_tracking_hour)
  Bidder (DA):    max  sum_t [ da_t Pda_t + rt_t (P_T[t] - Pda_t) - tot_cost_t - 1e4 underbid_t ]
                  s.t. Pda_t <= P_T[t] + underbid_t,  Pda, underbid >= 0,  underbid fixed 0 in the day-ahead problem
  Bidder (RT):    the same with Pda fixed to the cleared day-ahead dispatch and underbid free
  Backcaster:     scenario k = the historical days in reverse chronological order starting k days back (cyclic),
                  concatenated over the horizon -- pinned by the 48-h golden for k = 0 with two historical days
                  (hours 0-23 <- most recent day, hours 24-47 <- the day before); k > 0 is recalled, not pinned.
The Tracker golden is insensitive to the exact penalty values (any large value gives the same dispatch).
"""
from __future__ import annotations

import numpy as np

from .lp_models import (_Builder, BATT_CAP_COST_KW, DEGRADATION, ETA_C, ETA_D, PEM_ELEC_TO_MOL, PEM_OP_COST, PEM_VAR_COST, WIND_OP_COST,
                        nuclear_blocks, nuclear_operating_cost)

BATT_REP_COST_KWH = BATT_CAP_COST_KW * 0.5 / 4.0     # load_parameters.py:48
WASTE_PENALTY = 1e3                                  # wind_battery_double_loop.py:165
LARGE_PENALTY = 1e4                                  # idaes Tracker deviation penalty / Bidder underbid penalty


def _operation_blocks(B, v, T, cf, wind_mw, batt_mw, energy_mwh, soc0, thr0):
    """Period blocks of the operation model in the reference's kW units; returns the per-period linear forms
    P_T[t] (MW) and tot_cost[t] ($) as (coeff dict, constant)."""
    C = wind_mw * 1e3; P = batt_mw * 1e3; E = energy_mwh * 1e3
    PT, cost = [], []
    for t in range(T):
        p = f"blk[{t}].fs."
        v["w", t] = B.var(p + "windpower.electricity[0]")
        v["g", t] = B.var(p + "splitter.grid_elec[0]")
        v["i", t] = B.var(p + "battery.elec_in[0]")
        v["o", t] = B.var(p + "battery.elec_out[0]")
        v["s0", t] = B.var(p + "battery.initial_state_of_charge")
        v["e0", t] = B.var(p + "battery.initial_energy_throughput")
        v["s", t] = B.var(p + "battery.state_of_charge[0]")
        v["e", t] = B.var(p + "battery.energy_throughput[0]")
        B.le({v["w", t]: 1.0}, C * cf[t])                                                     # wind_power.py:120-122
        B.eq({v["w", t]: 1.0, v["g", t]: -1.0, v["i", t]: -1.0})                              # elec_splitter.py:115-117 + arcs
        B.eq({v["s", t]: 1.0, v["s0", t]: -1.0, v["i", t]: -ETA_C, v["o", t]: 1.0 / ETA_D})   # battery.py:145-149
        B.eq({v["e", t]: 1.0, v["e0", t]: -1.0, v["i", t]: -0.5, v["o", t]: -0.5})            # battery.py:151-153
        B.le({v["s", t]: 1.0, v["e", t]: DEGRADATION}, E)                                      # battery.py:155-157
        B.le({v["i", t]: 1.0}, P)                                                              # battery.py:159-161
        B.le({v["o", t]: 1.0}, P)                                                              # battery.py:163-165
        PT.append(({v["g", t]: 1e-3, v["o", t]: 1e-3}, 0.0))                                   # double_loop.py:168
        # tot_cost = op_total_cost + var_cost + penalty * wind_waste   (double_loop.py:169-171, wind_battery_LMP.py:57-71)
        cost.append(({v["e", t]: DEGRADATION * BATT_REP_COST_KWH, v["e0", t]: -DEGRADATION * BATT_REP_COST_KWH,
                      v["w", t]: -WASTE_PENALTY * 1e-3},
                     C * WIND_OP_COST / 8760.0 + WASTE_PENALTY * 1e-3 * C * cf[t]))
    for t in range(T - 1):                                                                     # wind_battery_LMP.py:32-34
        B.eq({v["s", t]: 1.0, v["s0", t + 1]: -1.0})
        B.eq({v["e", t]: 1.0, v["e0", t + 1]: -1.0})
    B.lb[v["s0", 0]] = B.ub[v["s0", 0]] = float(soc0)                                          # double_loop.py:76-77 / :190-191
    if thr0 is not None:                                                                       # :193-196 (free before the first update)
        B.lb[v["e0", 0]] = B.ub[v["e0", 0]] = float(thr0)
    return PT, cost


def tracker_raw(dispatch, cf, wind_mw=200.0, batt_mw=25.0, energy_mwh=100.0, soc0=0.0, thr0=None, n_tracking_hour=1):
    """Tracker.track_market_dispatch LP for one horizon (dispatch [MW], cf [-], both length H)."""
    dispatch = np.asarray(dispatch, float); H = dispatch.size
    B = _Builder(); v = {}
    PT, cost = _operation_blocks(B, v, H, cf, wind_mw, batt_mw, energy_mwh, soc0, thr0)
    for t in range(H):
        v["under", t] = B.var(f"power_underdelivered[{t}]")
        v["over", t] = B.var(f"power_overdelivered[{t}]")
        row = dict(PT[t][0]); row[v["under", t]] = 1.0; row[v["over", t]] = -1.0
        B.eq(row, dispatch[t])
        pen = LARGE_PENALTY if t < n_tracking_hour else LARGE_PENALTY / max(1, H - n_tracking_hour)
        B.cost(v["under", t], pen); B.cost(v["over", t], pen)
        for j, a in cost[t][0].items():
            B.cost(j, a)
        B.c0 += cost[t][1]
    return B.finish(dict(kind="tracker", T=H, v=v))


def bidder_raw(da, rt, cf, wind_mw=200.0, batt_mw=25.0, energy_mwh=100.0, soc0=0.0, thr0=None, da_dispatch=None):
    """One scenario block of the (Self)Scheduler / Bidder problem, written as a minimisation of -profit.
    da_dispatch=None: day-ahead problem (underbid fixed 0); else the real-time problem with Pda fixed."""
    da = np.asarray(da, float); rt = np.asarray(rt, float); H = da.size
    B = _Builder(); v = {}
    PT, cost = _operation_blocks(B, v, H, cf, wind_mw, batt_mw, energy_mwh, soc0, thr0)
    for t in range(H):
        v["da", t] = B.var(f"day_ahead_power[{t}]", fix=None if da_dispatch is None else float(da_dispatch[t]))
        v["ub", t] = B.var(f"real_time_underbid_power[{t}]", fix=0.0 if da_dispatch is None else None)
        row = {v["da", t]: 1.0, v["ub", t]: -1.0}
        for j, a in PT[t][0].items():
            row[j] = row.get(j, 0.0) - a
        B.le(row, 0.0)
        B.cost(v["da", t], -(da[t] - rt[t]))
        for j, a in PT[t][0].items():
            B.cost(j, -rt[t] * a)
        B.cost(v["ub", t], LARGE_PENALTY)
        for j, a in cost[t][0].items():
            B.cost(j, a)
        B.c0 += cost[t][1]
    return B.finish(dict(kind="bidder", T=H, v=v))


# ---------------------------------------------------------------------------------------------------------------
# nuclear + PEM + tank: MultiPeriodNuclear (case_studies/nuclear_case/nuclear_flowsheet_multiperiod_class.py:158-344)
#   :190-215 populate_model  (holdup_previous of block 0 fixed; P_T = np_to_grid * 1e-3; tot_cost = operating_cost)
#   :218-237 update_model    (implemented holdup rounded to an integer and fixed)
# No known answers exist in the reference for these LPs (nuclear_case/tests has no double-loop test): unpinned beyond
# the rows / constants shared with the price-taker LP.
# ---------------------------------------------------------------------------------------------------------------
def _nuclear_forms(B, v, T, holdup0, h2_price, **kw):
    nuclear_blocks(B, v, T, holdup0=holdup0, **kw)
    PT = [({v["xg", t]: 1e-3}, 0.0) for t in range(T)]
    cost = [(nuclear_operating_cost(v, t, h2_price), 0.0) for t in range(T)]
    return PT, cost


def nuclear_tracker_raw(dispatch, holdup0=0.0, h2_price=4.0, n_tracking_hour=1, **kw):
    dispatch = np.asarray(dispatch, float); H = dispatch.size
    B = _Builder(); v = {}
    PT, cost = _nuclear_forms(B, v, H, holdup0, h2_price, **kw)
    for t in range(H):
        v["under", t] = B.var(f"power_underdelivered[{t}]")
        v["over", t] = B.var(f"power_overdelivered[{t}]")
        row = dict(PT[t][0]); row[v["under", t]] = 1.0; row[v["over", t]] = -1.0
        B.eq(row, dispatch[t])
        pen = LARGE_PENALTY if t < n_tracking_hour else LARGE_PENALTY / max(1, H - n_tracking_hour)
        B.cost(v["under", t], pen); B.cost(v["over", t], pen)
        for j, a in cost[t][0].items():
            B.cost(j, a)
    return B.finish(dict(kind="nuclear_tracker", T=H, v=v))


def nuclear_bidder_raw(da, rt, holdup0=0.0, h2_price=4.0, da_dispatch=None, **kw):
    da = np.asarray(da, float); rt = np.asarray(rt, float); H = da.size
    B = _Builder(); v = {}
    PT, cost = _nuclear_forms(B, v, H, holdup0, h2_price, **kw)
    for t in range(H):
        v["da", t] = B.var(f"day_ahead_power[{t}]", fix=None if da_dispatch is None else float(da_dispatch[t]))
        v["ub", t] = B.var(f"real_time_underbid_power[{t}]", fix=0.0 if da_dispatch is None else None)
        row = {v["da", t]: 1.0, v["ub", t]: -1.0}
        for j, a in PT[t][0].items():
            row[j] = row.get(j, 0.0) - a
        B.le(row, 0.0)
        B.cost(v["da", t], -(da[t] - rt[t]))
        for j, a in PT[t][0].items():
            B.cost(j, -rt[t] * a)
        B.cost(v["ub", t], LARGE_PENALTY)
        for j, a in cost[t][0].items():
            B.cost(j, a)
    return B.finish(dict(kind="nuclear_bidder", T=H, v=v))


# ---------------------------------------------------------------------------------------------------------------
# wind + PEM: MultiPeriodWindPEM (case_studies/renewables_case/wind_PEM_double_loop.py:25-260), battery size 0
#   :56-82   transform_design_model_to_operation_model: wind size fixed, pem.electricity <= pem_system_capacity with
#            pem_system_capacity a (free, non-negative) Var of the model, periodic row deactivated
#   :163-172 P_T = grid_elec * 1e-3 ; wind_waste in kW ; tot_cost = wind O&M + pem_system_capacity * pem O&M / 8760
#            + pem var cost + wind_waste
# Its bidder (PEM_parametrized_bidder.py) computes bids from forecasts without an optimisation; the Tracker LP is the
# only LP and is pinned by tests/test_wind_PEM_double_loop.py:55-121 (wind, delivered power, PEM power).
# ---------------------------------------------------------------------------------------------------------------
def wind_pem_tracker_raw(dispatch, cf, wind_mw=200.0, pem_mw=25.0, n_tracking_hour=1):
    dispatch = np.asarray(dispatch, float); H = dispatch.size
    B = _Builder(); v = {}
    C = wind_mw * 1e3
    v["Pc"] = B.var("pem_system_capacity")                                   # :69 (initialised at pem_mw, not fixed)
    for t in range(H):
        p = f"blk[{t}].fs."
        v["w", t] = B.var(p + "windpower.electricity[0]")
        v["g", t] = B.var(p + "splitter.grid_elec[0]")
        v["pe", t] = B.var(p + "pem.electricity[0]")
        v["h", t] = B.var(p + "pem.outlet.flow_mol[0]")
        B.le({v["w", t]: 1.0}, C * cf[t])                                    # wind_power.py:120-122
        B.eq({v["w", t]: 1.0, v["g", t]: -1.0, v["pe", t]: -1.0})            # splitter + arcs (battery size 0)
        B.eq({v["h", t]: 1.0, v["pe", t]: -PEM_ELEC_TO_MOL})                 # pem_electrolyzer.py:111-114
        B.le({v["pe", t]: 1.0, v["Pc"]: -1.0})                               # :73
        v["under", t] = B.var(f"power_underdelivered[{t}]")
        v["over", t] = B.var(f"power_overdelivered[{t}]")
        B.eq({v["g", t]: 1e-3, v["under", t]: 1.0, v["over", t]: -1.0}, dispatch[t])
        pen = LARGE_PENALTY if t < n_tracking_hour else LARGE_PENALTY / max(1, H - n_tracking_hour)
        B.cost(v["under", t], pen); B.cost(v["over", t], pen)
        # tot_cost[t] (:169-172); wind_waste[t] = C*cf - w in kW enters with weight 1
        B.cost(v["Pc"], PEM_OP_COST / 8760.0); B.cost(v["pe", t], PEM_VAR_COST); B.cost(v["w", t], -1.0)
        B.c0 += C * WIND_OP_COST / 8760.0 + C * cf[t]
    return B.finish(dict(kind="wind_pem_tracker", T=H, v=v))


def backcast(historical, hour, horizon, n_samples):
    """Backcaster forecast [n_samples, horizon]: see the module docstring."""
    h = np.asarray(historical, float)
    days = h[: (h.size // 24) * 24].reshape(-1, 24)
    nd = days.shape[0]
    reps = (hour + horizon) // 24 + 1
    out = []
    for k in range(n_samples):
        seq = np.concatenate([days[(nd - 1 - k - r) % nd] for r in range(reps)])
        out.append(seq[hour:hour + horizon])
    return np.array(out)


def read_profile(lp, x):
    v = lp.meta["v"]; T = lp.meta["T"]
    col = lambda k: np.array([x[v[k, t]] for t in range(T)])
    out = dict(wind=col("w"), grid=col("g"), batt_in=col("i"), batt_out=col("o"), soc=col("s"), throughput=col("e"))
    out["P_T"] = (out["grid"] + out["batt_out"]) * 1e-3
    for k in ("under", "over", "da", "ub"):
        if (k, 0) in v:
            out[k] = col(k)
    return out
