"""The Pyomo solver plugin (dispatches_b200/pyomo_plugin.py) EXECUTED on a reference-shaped model.

Real Pyomo is not installable in the build image.  When `import pyomo` fails, tests/pyomo_stub (a stand-in for
the slice of Pyomo's API the plugin and this model builder touch -- test infrastructure, see its docstring) is put on sys.path;
where Pyomo exists the same tests run against it.  The model below is written the way the reference writes
wind_battery_optimize (wind_battery_LMP.py:106-267 on RE_flowsheet.py:338-464 and the unit models): one block per period with
every unit-model Var and Constraint, arcs, link / periodic equalities, fixed Vars, mutable lmp_signal Params, objective
-NPV*1e-5 -- nothing presolved.  Checked: walk (extract + standard_form) == the oracle's raw LP through HiGHS (CPU); the full
`SolverFactory("b200ipm").solve(m)` with write-back of Vars and duals, and the Param-batched solve (GPU)."""
import sys
from pathlib import Path

import numpy as np
import pytest

try:
    import pyomo.environ as pyo                      # noqa: F401
except Exception:                                     # noqa: BLE001
    sys.path.insert(0, str(Path(__file__).parent / "pyomo_stub"))
    for k in [k for k in sys.modules if k == "pyomo" or k.startswith("pyomo.")]:
        del sys.modules[k]
    import pyomo.environ as pyo
    import importlib
    import dispatches_b200.pyomo_plugin as _pp
    importlib.reload(_pp)

from dispatches_b200 import pyomo_plugin as PP, scenarios as SC
from oracle import highs as H, lp_models as L
from test_standard_form import highs_template


def build_wind_battery(T, lmp, cf, wind_mw, batt_mw):
    """reference-shaped wind + battery price-taker model (see module docstring); returns the model"""
    m = pyo.ConcreteModel()
    m.blocks = pyo.Block(range(T))
    wind_kw, batt_kw = wind_mw * 1e3, batt_mw * 1e3
    for t in range(T):
        b = m.blocks[t]
        b.fs = pyo.Block()
        fs = b.fs
        fs.windpower = pyo.Block(); wp = fs.windpower                                   # unit_models/wind_power.py:99-122
        wp.system_capacity = pyo.Var(within=pyo.NonNegativeReals, initialize=wind_kw)
        wp.system_capacity.fix(wind_kw)                                                  # RE_flowsheet.py:86
        wp.electricity = pyo.Var(within=pyo.NonNegativeReals, initialize=0.0)
        wp.capacity_factor = pyo.Param(mutable=True, initialize=float(cf[t]))
        wp.elec_from_capacity_factor = pyo.Constraint(expr=wp.electricity <= wp.system_capacity * wp.capacity_factor)
        fs.splitter = pyo.Block(); sp = fs.splitter                                      # unit_models/elec_splitter.py:107-117
        sp.electricity = pyo.Var(within=pyo.NonNegativeReals, initialize=0.0)
        sp.grid_elec = pyo.Var(within=pyo.NonNegativeReals, initialize=0.0)
        sp.battery_elec = pyo.Var(within=pyo.NonNegativeReals, initialize=0.0)
        sp.sum_split = pyo.Constraint(expr=sp.electricity == sp.grid_elec + sp.battery_elec)
        fs.battery = pyo.Block(); bt = fs.battery                                        # unit_models/battery.py:69-165
        bt.nameplate_power = pyo.Var(within=pyo.NonNegativeReals, bounds=(0, 1e8), initialize=batt_kw)
        bt.nameplate_power.fix(batt_kw)                                                  # RE_flowsheet.py:154
        bt.nameplate_energy = pyo.Var(within=pyo.NonNegativeReals, bounds=(0, 1e9), initialize=4 * batt_kw)
        for nm in ("initial_state_of_charge", "initial_energy_throughput", "elec_in", "elec_out", "state_of_charge", "energy_throughput"):
            setattr(bt, nm, pyo.Var(within=pyo.NonNegativeReals, initialize=0.0))
        bt.four_hr_battery = pyo.Constraint(expr=bt.nameplate_power * 4.0 == bt.nameplate_energy)     # RE_flowsheet.py:155-156
        bt.state_evolution = pyo.Constraint(expr=bt.state_of_charge == bt.initial_state_of_charge + 0.95 * bt.elec_in - bt.elec_out / 0.95)
        bt.accumulate_energy_throughput = pyo.Constraint(expr=bt.energy_throughput == bt.initial_energy_throughput + (bt.elec_in + bt.elec_out) / 2.0)
        bt.state_of_charge_bounds = pyo.Constraint(expr=bt.state_of_charge <= bt.nameplate_energy - 1e-4 * bt.energy_throughput)
        bt.power_bound_in = pyo.Constraint(expr=bt.elec_in <= bt.nameplate_power)
        bt.power_bound_out = pyo.Constraint(expr=bt.elec_out <= bt.nameplate_power)
        fs.wind_to_splitter = pyo.Constraint(expr=wp.electricity == sp.electricity)      # arcs, RE_flowsheet.py:389-396
        fs.splitter_to_battery = pyo.Constraint(expr=sp.battery_elec == bt.elec_in)
        b.energy_down_ramp = pyo.Constraint(expr=bt.initial_state_of_charge - bt.state_of_charge <= 1e8)   # wind_battery_LMP.py:139-142
        b.energy_up_ramp = pyo.Constraint(expr=bt.state_of_charge - bt.initial_state_of_charge <= 1e8)
        b.lmp_signal = pyo.Param(mutable=True, initialize=float(lmp[t]))                 # :234
    m.link = pyo.Block(range(T))
    for t in range(T):                                                                   # link / periodic pairs, :22-50
        nx = m.blocks[(t + 1) % T].fs.battery
        bt = m.blocks[t].fs.battery
        m.link[t].soc = pyo.Constraint(expr=bt.state_of_charge == nx.initial_state_of_charge)
        m.link[t].power = pyo.Constraint(expr=bt.nameplate_power == nx.nameplate_power)
        if t < T - 1:
            m.link[t].throughput = pyo.Constraint(expr=bt.energy_throughput == nx.initial_energy_throughput)
    m.blocks[0].fs.battery.initial_state_of_charge.fix(0.0)                              # :206-207
    m.blocks[0].fs.battery.initial_energy_throughput.fix(0.0)
    m.wind_system_capacity = pyo.Var(within=pyo.NonNegativeReals, bounds=(0, 1e7), initialize=wind_kw)   # :209-210
    m.battery_system_capacity = pyo.Var(within=pyo.NonNegativeReals, initialize=batt_kw)
    m.wind_max_p = pyo.Constraint(range(T), rule=lambda mm, t: mm.blocks[t].fs.windpower.system_capacity <= mm.wind_system_capacity)   # :218-219
    m.battery_max_p = pyo.Constraint(range(T), rule=lambda mm, t: mm.blocks[t].fs.battery.nameplate_power <= mm.battery_system_capacity)
    n_weeks = T / 168.0
    profit = 0.0
    for t in range(T):
        fs = m.blocks[t].fs
        rev = m.blocks[t].lmp_signal * 1e-3 * (fs.splitter.grid_elec + fs.battery.elec_out)          # :235-237
        profit = profit + rev - m.wind_system_capacity * (L.WIND_OP_COST / 8760.0) - m.battery_system_capacity * (L.BATT_OP_COST / 8760.0)
    m.annual_revenue = pyo.Expression(expr=profit * 52.0 / n_weeks)                                   # :255
    m.NPV = pyo.Expression(expr=-(L.BATT_CAP_COST_KW * m.battery_system_capacity + L.BATT_CAP_COST_KWH * 4.0 * m.battery_system_capacity)
                           + L.PA * m.annual_revenue)                                                # :256-263 with extant wind (:247-248)
    m.obj = pyo.Objective(expr=-m.NPV * 1e-5, sense=pyo.minimize)                                     # :264
    m.dual = pyo.Suffix(direction=pyo.Suffix.IMPORT)
    return m


def test_walker_runs_on_a_reference_shaped_model_and_matches_the_oracle():
    T = 24
    lmp, cf, W, P = SC.c2(4)
    m = build_wind_battery(T, lmp[0], cf, W, P)
    params = [m.blocks[t].lmp_signal for t in range(T)]
    t, vars_, cons, p0 = PP.walk_model(m, batch_params=params)
    assert np.allclose(p0, lmp[0]) and len(cons) > 12 * T and t.w <= 32
    for k in range(4):                                         # the batch: all lmp_signal Params at once
        obj, x, _ = highs_template(t, lmp[k], lmp[k])
        ref, _ = H.solve(L.wind_battery_raw(lmp[k], cf, W, P))
        assert obj == pytest.approx(ref, rel=1e-11)
    # a Param in the constraint matrix cannot be batched; a nonlinear model is refused
    with pytest.raises(ValueError):
        m2 = build_wind_battery(4, lmp[0], cf, W, P)
        m2.blocks[0].fs.windpower.system_capacity.unfix()
        PP.walk_model(m2, batch_params=[m2.blocks[0].fs.windpower.capacity_factor])
    with pytest.raises(ValueError):
        m3 = build_wind_battery(4, lmp[0], cf, W, P)
        m3.bad = pyo.Constraint(expr=m3.blocks[0].fs.battery.elec_in * m3.blocks[0].fs.battery.elec_out <= 1.0)
        PP.walk_model(m3)


@pytest.mark.gpu
def test_solver_factory_solve_writes_back_values_and_duals():
    T = 24
    lmp, cf, W, P = SC.c2(64)
    m = build_wind_battery(T, lmp[0], cf, W, P)
    opt = pyo.SolverFactory("b200ipm")
    assert opt.available(exception_flag=False)
    res = opt.solve(m, tee=False)                                                       # wind_battery_LMP.py:266-267
    assert res.solver.status == pyo.SolverStatus.ok and res.solver.termination_condition == pyo.TerminationCondition.optimal
    lp = L.wind_battery_raw(lmp[0], cf, W, P)
    ref, xref = H.solve(lp)
    assert pyo.value(m.obj.expr) == pytest.approx(ref, rel=1e-6)                         # post-solve reads, :285-305
    assert pyo.value(m.NPV) == pytest.approx(-ref * 1e5, rel=1e-6)
    soc = np.array([pyo.value(m.blocks[t].fs.battery.state_of_charge) for t in range(T)])
    assert soc.min() >= -1e-3 and soc.max() <= 4 * P * 1e3 * (1 + 1e-6) and abs(soc[-1]) < 1.0      # periodic: ends empty
    # duals: d objective / d rhs of the wind-availability rows = -(marginal value of wind energy) <= 0 where wind is scarce
    duals = np.array([m.dual[m.blocks[t].fs.windpower.elec_from_capacity_factor] for t in range(T)])
    assert np.all(duals <= 1e-9) and duals.min() < 0
    # finite-difference check of one dual -- in the windiest hour: where cf = 0 the row reads 0 <= x <= 0, its dual is not unique
    # (an interior-point method returns the centre of an unbounded dual face there)
    eps = 1.0
    t0 = int(np.argmax(cf))
    cf2 = cf.copy(); cf2[t0] += eps / (W * 1e3)
    ref2, _ = H.solve(L.wind_battery_raw(lmp[0], cf2, W, P))
    assert (ref2 - ref) / eps == pytest.approx(duals[t0], rel=1e-3, abs=1e-9)
    # the Param-batched solve: every block's lmp_signal for 64 scenarios in one launch
    res = opt.solve(m, batch_params={m.blocks[t].lmp_signal: lmp[:, t] for t in range(T)})
    refs, _, _ = H.solve_batch("wind_battery", lmp, kwargs=dict(cf=cf, wind_mw=W, batt_mw=P))
    assert np.abs(res.batch["obj"] - refs).max() <= 1e-6 * np.abs(refs).max()
