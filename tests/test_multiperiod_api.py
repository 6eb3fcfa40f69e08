"""SURVEY 8(a) row "MultiPeriodModel builder API": code written THE WAY THE REFERENCE WRITES IT -- create_model / wind_battery_model /
wind_battery_mp_block (clone cache) / *_variable_pairs / MultiPeriodModel(...).build_multi_period_model(wind_resource) /
get_active_process_blocks() / model-level capacity Vars, Params, Expressions, Objective (wind_battery_LMP.py:22-50,106-169,172-267 on
RE_flowsheet.py:338-396) -- handed to SolverFactory("b200ipm") instead of "cbc".

idaes-pse and Pyomo are not installable in the build image: when the imports fail, tests/pyomo_stub (stand-ins for the API slice used
here, TEST INFRASTRUCTURE, see their docstrings) goes on sys.path; with the real packages present the same code runs against them.
The functions below restate the reference's model-building STEPS in the reference's order; nothing is presolved."""
import sys
from functools import partial
from pathlib import Path

import numpy as np
import pytest

try:
    import pyomo.environ as pyo                      # noqa: F401
    from idaes.apps.grid_integration.multiperiod.multiperiod import MultiPeriodModel
except Exception:                                     # noqa: BLE001
    stub = str(Path(__file__).parent / "pyomo_stub")
    if stub not in sys.path:
        sys.path.insert(0, stub)
    for k in [k for k in sys.modules if k.split(".")[0] in ("pyomo", "idaes") and not getattr(sys.modules[k], "__file__", "x").startswith(stub)]:
        del sys.modules[k]
    import pyomo.environ as pyo
    from idaes.apps.grid_integration.multiperiod.multiperiod import MultiPeriodModel
    import importlib
    import dispatches_b200.pyomo_plugin as _pp
    importlib.reload(_pp)

from dispatches_b200 import pyomo_plugin as PP, scenarios as SC
from oracle import highs as H, lp_models as L
from test_standard_form import highs_template

DURATION, BATTERY_RAMP_RATE = 4.0, 1e8


def create_model(wind_mw, batt_mw, resource_config):
    """RE_flowsheet.create_model(re_mw, None, batt_mw, None, None, None, resource_config) :338-396: wind + splitter(grid, battery) + battery"""
    m = pyo.ConcreteModel()
    m.fs = pyo.Block()
    fs = m.fs
    fs.windpower = pyo.Block(); wp = fs.windpower                                      # unit_models/wind_power.py:99-122
    wp.system_capacity = pyo.Var(within=pyo.NonNegativeReals, initialize=wind_mw * 1e3)
    wp.system_capacity.fix(wind_mw * 1e3)                                              # add_wind, RE_flowsheet.py:86
    wp.electricity = pyo.Var([0], within=pyo.NonNegativeReals, initialize=0.0)
    wp.capacity_factor = pyo.Param([0], mutable=True, initialize=float(resource_config["capacity_factor"][0]))
    wp.elec_from_capacity_factor = pyo.Constraint([0], rule=lambda b, t: b.electricity[t] <= b.system_capacity * b.capacity_factor[t])
    fs.splitter = pyo.Block(); sp = fs.splitter                                        # unit_models/elec_splitter.py:107-117
    sp.electricity = pyo.Var([0], within=pyo.NonNegativeReals, initialize=0.0)
    sp.grid_elec = pyo.Var([0], within=pyo.NonNegativeReals, initialize=0.0)
    sp.battery_elec = pyo.Var([0], within=pyo.NonNegativeReals, initialize=0.0)
    sp.sum_split = pyo.Constraint([0], rule=lambda b, t: b.electricity[t] == b.grid_elec[t] + b.battery_elec[t])
    fs.battery = pyo.Block(); bt = fs.battery                                          # unit_models/battery.py:69-165, add_battery :138-157
    bt.nameplate_power = pyo.Var(within=pyo.NonNegativeReals, bounds=(0, 1e8), initialize=batt_mw * 1e3)
    bt.nameplate_power.fix(batt_mw * 1e3)
    bt.nameplate_energy = pyo.Var(within=pyo.NonNegativeReals, bounds=(0, 1e9), initialize=DURATION * batt_mw * 1e3)
    bt.initial_state_of_charge = pyo.Var(within=pyo.NonNegativeReals, initialize=0.0)
    bt.initial_energy_throughput = pyo.Var(within=pyo.NonNegativeReals, initialize=0.0)
    for nm in ("elec_in", "elec_out", "state_of_charge", "energy_throughput"):
        setattr(bt, nm, pyo.Var([0], within=pyo.NonNegativeReals, initialize=0.0))
    bt.four_hr_battery = pyo.Constraint(expr=bt.nameplate_power * DURATION == bt.nameplate_energy)
    bt.state_evolution = pyo.Constraint([0], rule=lambda b, t: b.state_of_charge[t] == b.initial_state_of_charge + 0.95 * b.elec_in[t] - b.elec_out[t] / 0.95)
    bt.accumulate_energy_throughput = pyo.Constraint([0], rule=lambda b, t: b.energy_throughput[t] == b.initial_energy_throughput + (b.elec_in[t] + b.elec_out[t]) / 2.0)
    bt.state_of_charge_bounds = pyo.Constraint([0], rule=lambda b, t: b.state_of_charge[t] <= b.nameplate_energy - 1e-4 * b.energy_throughput[t])
    bt.power_bound_in = pyo.Constraint([0], rule=lambda b, t: b.elec_in[t] <= b.nameplate_power)
    bt.power_bound_out = pyo.Constraint([0], rule=lambda b, t: b.elec_out[t] <= b.nameplate_power)
    fs.wind_to_splitter = pyo.Constraint(expr=wp.electricity[0] == sp.electricity[0])              # arcs (expanded), :389-396
    fs.splitter_to_battery = pyo.Constraint(expr=sp.battery_elec[0] == bt.elec_in[0])
    return m


def wind_battery_variable_pairs(m1, m2):                                               # wind_battery_LMP.py:22-36
    return [(m1.fs.battery.state_of_charge[0], m2.fs.battery.initial_state_of_charge),
            (m1.fs.battery.energy_throughput[0], m2.fs.battery.initial_energy_throughput),
            (m1.fs.battery.nameplate_power, m2.fs.battery.nameplate_power)]


def wind_battery_periodic_variable_pairs(m1, m2):                                      # :39-50
    return [(m1.fs.battery.state_of_charge[0], m2.fs.battery.initial_state_of_charge),
            (m1.fs.battery.nameplate_power, m2.fs.battery.nameplate_power)]


def wind_battery_model(wind_resource_config, input_params, verbose=False):             # :75-122
    m = create_model(input_params["wind_mw"], input_params["batt_mw"], wind_resource_config)
    m.fs.windpower.op_cost = pyo.Param(initialize=L.WIND_OP_COST)                      # wind_battery_om_costs :52-71
    m.fs.battery.op_cost = pyo.Param(initialize=L.BATT_OP_COST)
    batt = m.fs.battery
    batt.energy_down_ramp = pyo.Constraint(expr=batt.initial_state_of_charge - batt.state_of_charge[0] <= BATTERY_RAMP_RATE)
    batt.energy_up_ramp = pyo.Constraint(expr=batt.state_of_charge[0] - batt.initial_state_of_charge <= BATTERY_RAMP_RATE)
    return m


def wind_battery_mp_block(wind_resource_config, input_params, verbose=False):          # :125-169: clone of the cached one-period model
    if "pyo_model" not in input_params:
        input_params["pyo_model"] = wind_battery_model(wind_resource_config, input_params, verbose=verbose)
    m = input_params["pyo_model"].clone()
    m.fs.windpower.capacity_factor[0].set_value(float(wind_resource_config["capacity_factor"][0]))
    return m


def wind_battery_optimize(n_time_points, input_params, solver_name, verbose=False, solve=True):
    """wind_battery_LMP.wind_battery_optimize :172-267, statement by statement; only the solver name differs"""
    mp_wind_battery = MultiPeriodModel(
        n_time_points=n_time_points,
        process_model_func=partial(wind_battery_mp_block, input_params=input_params, verbose=verbose),
        linking_variable_func=wind_battery_variable_pairs,
        periodic_variable_func=wind_battery_periodic_variable_pairs)
    mp_wind_battery.build_multi_period_model(input_params["wind_resource"])
    m = mp_wind_battery.pyomo_model
    blks = mp_wind_battery.get_active_process_blocks()
    blks[0].fs.battery.initial_state_of_charge.fix(0)
    blks[0].fs.battery.initial_energy_throughput.fix(0)
    m.wind_system_capacity = pyo.Var(domain=pyo.NonNegativeReals, initialize=input_params["wind_mw"] * 1e3, bounds=(0, input_params["wind_mw_ub"] * 1e3))
    m.battery_system_capacity = pyo.Var(domain=pyo.NonNegativeReals, initialize=input_params["batt_mw"] * 1e3)
    if input_params["design_opt"]:
        for blk in blks:
            if not input_params["extant_wind"]:
                blk.fs.windpower.system_capacity.unfix()
            blk.fs.battery.nameplate_power.unfix()
    m.wind_max_p = pyo.Constraint(m.TIME, rule=lambda b, t: blks[t].fs.windpower.system_capacity <= m.wind_system_capacity)
    m.battery_max_p = pyo.Constraint(m.TIME, rule=lambda b, t: blks[t].fs.battery.nameplate_power <= m.battery_system_capacity)
    for blk in blks:
        blk_wind, blk_battery = blk.fs.windpower, blk.fs.battery
        blk_wind.op_total_cost = pyo.Expression(expr=m.wind_system_capacity * blk_wind.op_cost / 8760)
        blk_battery.op_total_cost = pyo.Expression(expr=m.battery_system_capacity * blk_battery.op_cost / 8760)
        blk.lmp_signal = pyo.Param(default=0, mutable=True)
        blk.elec_output = blk.fs.splitter.grid_elec[0] + blk_battery.elec_out[0]
        blk.revenue = blk.lmp_signal * (blk.fs.splitter.grid_elec[0] + blk_battery.elec_out[0])
        blk.profit = pyo.Expression(expr=blk.revenue - blk_wind.op_total_cost - blk_battery.op_total_cost)
    for (i, blk) in enumerate(blks):
        blk.lmp_signal.set_value(input_params["DA_LMPs"][i] * 1e-3)
    m.wind_cap_cost = pyo.Param(default=L.WIND_CAP_COST, mutable=True)
    if input_params["extant_wind"]:
        m.wind_cap_cost.set_value(0.0)
    m.batt_cap_cost_kw = pyo.Param(default=L.BATT_CAP_COST_KW, mutable=True)
    m.batt_cap_cost_kwh = pyo.Param(default=L.BATT_CAP_COST_KWH, mutable=True)
    n_weeks = n_time_points / (7 * 24)
    m.annual_revenue = pyo.Expression(expr=sum([blk.profit for blk in blks]) * 52 / n_weeks)
    m.NPV = pyo.Expression(expr=-(m.wind_cap_cost * m.wind_system_capacity + m.batt_cap_cost_kw * m.battery_system_capacity
                                  + m.batt_cap_cost_kwh * m.battery_system_capacity * DURATION) + L.PA * m.annual_revenue)
    m.obj = pyo.Objective(expr=-m.NPV * 1e-5)
    if solve:
        opt = pyo.SolverFactory(solver_name)
        opt.solve(m, tee=verbose)
    return mp_wind_battery


def default_params(T, lmp, cf, W, P, design_opt=False):
    return dict(wind_mw=W, wind_mw_ub=10000, batt_mw=P, design_opt=design_opt, extant_wind=True, DA_LMPs=lmp,
                wind_resource={t: {"wind_resource_config": {"capacity_factor": [cf[t]]}} for t in range(T)})


def test_reference_shaped_multiperiod_build_walks_to_the_oracle_lp():
    T = 24
    lmp, cf, W, P = SC.c2(3)
    mp = wind_battery_optimize(T, default_params(T, lmp[0], cf, W, P), "b200ipm", solve=False)
    m = mp.pyomo_model
    blks = mp.get_active_process_blocks()
    assert len(blks) == T and len(m.blocks[0].link_constraints) == 3 and len(m.blocks[T - 1].periodic_constraints) == 2
    assert pyo.value(blks[5].fs.windpower.capacity_factor[0]) == pytest.approx(cf[5])          # the clone took its own capacity factor
    params = [blk.lmp_signal for blk in blks]
    t, vars_, cons, p0 = PP.walk_model(m, batch_params=params)
    assert np.allclose(p0, lmp[0] * 1e-3) and t.w <= 32
    for k in range(3):
        obj, _, _ = highs_template(t, lmp[k] * 1e-3, lmp[k] * 1e-3)
        ref, _ = H.solve(L.wind_battery_raw(lmp[k], cf, W, P))
        assert obj == pytest.approx(ref, rel=1e-11)


def check_reference_shaped_optimize():
    T = 24
    lmp, cf, W, P = SC.c2(2)
    mp = wind_battery_optimize(T, default_params(T, lmp[1], cf, W, P), "b200ipm")               # ... opt.solve(m) inside
    m, blks = mp.pyomo_model, mp.get_active_process_blocks()
    ref, _ = H.solve(L.wind_battery_raw(lmp[1], cf, W, P))
    assert pyo.value(m.NPV) == pytest.approx(-ref * 1e5, rel=1e-6)
    # record_results-style reads (wind_battery_LMP.py:285-305)
    soc = [pyo.value(blks[i].fs.battery.state_of_charge[0]) for i in range(T)]
    wind_gen = [pyo.value(blks[i].fs.windpower.electricity[0]) for i in range(T)]
    assert max(soc) <= 4 * P * 1e3 * (1 + 1e-6) and min(soc) >= -1e-3
    assert all(w <= W * 1e3 * cf[i] * (1 + 1e-6) + 1e-3 for i, w in enumerate(wind_gen))
    assert pyo.value(m.annual_revenue) == pytest.approx((pyo.value(m.NPV) + (L.BATT_CAP_COST_KW + 4 * L.BATT_CAP_COST_KWH) * P * 1e3) / L.PA, rel=1e-6)   # (the interior-point battery_system_capacity is tight to ~1e-9)
    names = [v.name for b in blks for v in (b.fs.battery.elec_out[0], b.fs.splitter.grid_elec[0])]
    assert len(set(names)) == 2 * T and names[0] == "blocks[0].process.fs.battery.elec_out[0]"       # hierarchical names of the cloned blocks


@pytest.mark.gpu
def test_reference_shaped_optimize_runs_on_the_gpu_solver():
    check_reference_shaped_optimize()
