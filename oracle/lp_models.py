"""CPU ORACLE (test infrastructure -- NOT product code; only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import this).

Raw, *un-presolved* restatements of the reference's price-taker LPs, shaped the
way Pyomo would hand them to CBC/IPOPT: every Var of every cloned period block
is a column (fixed Vars keep their column with lb == ub), every Constraint a
row, linking / periodic equalities are explicit rows.  The product templates in
``dispatches_b200/templates.py`` are *reduced* forms written independently; the
parity tests compare one against the other through HiGHS and the CUDA solver.

Reference equations restated here (file:line under /root/reference/dispatches):
  unit_models/battery.py:69-165            battery Vars, bounds and 5 rows
  unit_models/wind_power.py:99-122,178-183 electricity <= system_capacity*cf
  unit_models/elec_splitter.py:107-117     electricity == sum(outlets)
  unit_models/pem_electrolyzer.py:90-114   flow_mol == electricity*electricity_to_mol
  case_studies/renewables_case/RE_flowsheet.py:69-87,138-157,387-396
  case_studies/renewables_case/wind_battery_LMP.py:22-50,139-142,206-264
  case_studies/renewables_case/wind_battery_PEM_LMP.py:217-294
  case_studies/renewables_case/load_parameters.py:24-140
  case_studies/nuclear_case/nuclear_flowsheet_multiperiod_class.py:36-155
  case_studies/nuclear_case/nuclear_flowsheet.py:119-157,263-291
  unit_models/hydrogen_tank_simplified.py:177-184
  case_studies/fossil_case/ultra_supercritical_plant/storage/
      multiperiod_integrated_storage_usc.py:49-54,75-164,334-342 (structure only)

Parity pinning: see oracle/README.md -- the wind+PEM restatement is pinned to the
reference's committed result tables (tests/golden/wind_pem_golden.json, made by
tests/golden/make_golden.py); battery row semantics to the reference's unit-test
known answers; the fossil surrogate is "parity unpinned".
"""
from __future__ import annotations

import dataclasses
import numpy as np
import scipy.sparse as sp

INF = float("inf")

# ---- constants restated from load_parameters.py:24-121 and wind_battery_cost_parameter.json
# (scenario "moderate", year 2023, 4-h battery column)
WIND_CAP_COST = 1308.0          # $/kW
WIND_OP_COST = 41.78            # $/kW-yr
BATT_OP_COST = 31.39            # $/kW-yr
BATT_CAP_COST_KW = 236.365      # $/kW
BATT_CAP_COST_KWH = 254.835     # $/kWh
PEM_CAP_COST = 1200.0           # $/kW
PEM_OP_COST = 0.03 * PEM_CAP_COST
PEM_VAR_COST = 0.0
H2_MOLS_PER_KG = 500.0
DURATION = 4.0                  # hours of storage
BATTERY_RAMP_RATE = 1e8
ETA_C = 0.95
ETA_D = 0.95
DEGRADATION = 1.0 / 10000.0
PEM_ELEC_TO_MOL = 0.00275984    # RE_flowsheet.py:131
DISCOUNT, YEARS = 0.08, 30
PA = ((1 + DISCOUNT) ** YEARS - 1) / (DISCOUNT * (1 + DISCOUNT) ** YEARS)


@dataclasses.dataclass
class RawLP:
    """min c@x + c0  s.t.  A_eq x = b_eq, A_ub x <= b_ub, lb <= x <= ub."""
    c: np.ndarray
    c0: float
    A_eq: sp.csr_matrix
    b_eq: np.ndarray
    A_ub: sp.csr_matrix
    b_ub: np.ndarray
    lb: np.ndarray
    ub: np.ndarray
    names: list
    meta: dict

    @property
    def n(self):
        return self.c.size


class _Builder:
    def __init__(self):
        self.names, self.lb, self.ub, self.c = [], [], [], []
        self.eq_rows, self.eq_rhs, self.ub_rows, self.ub_rhs = [], [], [], []
        self.c0 = 0.0
        self.lmp_terms = []

    def var(self, name, lb=0.0, ub=INF, fix=None):
        if fix is not None:
            lb = ub = float(fix)
        self.names.append(name)
        self.lb.append(lb)
        self.ub.append(ub)
        self.c.append(0.0)
        return len(self.names) - 1

    def eq(self, coeffs, rhs=0.0):
        self.eq_rows.append(coeffs)
        self.eq_rhs.append(rhs)

    def le(self, coeffs, rhs=0.0):
        self.ub_rows.append(coeffs)
        self.ub_rhs.append(rhs)

    def cost(self, j, v):
        self.c[j] += v

    def cost_lmp(self, j, t, coef, lmp):
        """cost entry coef*lmp[t] on column j; remembered so a batch loop can swap the LMP vector only."""
        self.c[j] += coef * lmp[t]
        self.lmp_terms.append((j, t, coef))

    def _mat(self, rows):
        n = len(self.names)
        ri, ci, vv = [], [], []
        for r, row in enumerate(rows):
            for j, v in row.items():
                ri.append(r); ci.append(j); vv.append(v)
        return sp.csr_matrix((vv, (ri, ci)), shape=(len(rows), n))

    def finish(self, meta):
        meta = dict(meta, lmp_terms=self.lmp_terms)
        return RawLP(np.array(self.c, float), float(self.c0), self._mat(self.eq_rows),
                     np.array(self.eq_rhs, float), self._mat(self.ub_rows), np.array(self.ub_rhs, float),
                     np.array(self.lb, float), np.array(self.ub, float), self.names, meta)


def swap_lmp(lp: "RawLP", lmp_old, lmp_new):
    """Cost vector of the same LP under another LMP signal (everything else is LMP-independent)."""
    c = lp.c.copy()
    d = np.asarray(lmp_new, float) - np.asarray(lmp_old, float)
    for j, t, coef in lp.meta["lmp_terms"]:
        c[j] += coef * d[t]
    return c


def _add(d, j, v):
    d[j] = d.get(j, 0.0) + v


# --------------------------------------------------------------------------------------
# A.1 / A.2  wind + battery (+ PEM) price-taker
# --------------------------------------------------------------------------------------
def wind_battery_raw(lmp, cf, wind_mw, batt_mw, pem_mw=None, h2_price=2.0,
                     design_opt=False, extant_wind=True, wind_mw_ub=10000.0):
    """Raw LP of wind_battery_optimize (wind_battery_LMP.py:172-267) or, when ``pem_mw`` is not
    None, of wind_battery_pem_optimize (wind_battery_PEM_LMP.py:180-298).

    lmp [$/MWh], cf [-] have length T.  Objective = -NPV*1e-5 (:264 / :294).
    design_opt: False | True | "PEM" (the reference's three modes).
    """
    lmp = np.asarray(lmp, float); cf = np.asarray(cf, float)
    T = lmp.size
    with_pem = pem_mw is not None
    B = _Builder()
    v = {}
    free_batt = (design_opt is True)
    free_wind = bool(design_opt) and not extant_wind
    for t in range(T):
        p = f"blk[{t}].fs."
        # wind_power.py:99-115 ; RE_flowsheet.py:86 fixes system_capacity
        v["C", t] = B.var(p + "windpower.system_capacity", fix=None if free_wind else wind_mw * 1e3)
        v["w", t] = B.var(p + "windpower.electricity[0]")
        # elec_splitter.py:107-111
        v["z", t] = B.var(p + "splitter.electricity[0]")
        v["g", t] = B.var(p + "splitter.grid_elec[0]")
        v["q", t] = B.var(p + "splitter.battery_elec[0]")
        # battery.py:69-136 ; RE_flowsheet.py:154-156
        batt_fix = None if free_batt else (0.0 if design_opt == "PEM" else batt_mw * 1e3)
        v["P", t] = B.var(p + "battery.nameplate_power", ub=1e8, fix=batt_fix)
        v["E", t] = B.var(p + "battery.nameplate_energy", ub=1e9)
        v["s0", t] = B.var(p + "battery.initial_state_of_charge")
        v["e0", t] = B.var(p + "battery.initial_energy_throughput")
        v["i", t] = B.var(p + "battery.elec_in[0]")
        v["o", t] = B.var(p + "battery.elec_out[0]")
        v["s", t] = B.var(p + "battery.state_of_charge[0]")
        v["e", t] = B.var(p + "battery.energy_throughput[0]")
        if with_pem:
            v["p", t] = B.var(p + "splitter.pem_elec[0]")
            v["pe", t] = B.var(p + "pem.electricity[0]", lb=-INF)       # Reals, pem_electrolyzer.py:96-100
            v["h", t] = B.var(p + "pem.outlet.flow_mol[0]", lb=-INF)
        # rows
        B.le({v["w", t]: 1.0, v["C", t]: -cf[t]})                       # wind_power.py:120-122
        B.eq({v["w", t]: 1.0, v["z", t]: -1.0})                         # arc wind_to_splitter
        row = {v["z", t]: 1.0, v["g", t]: -1.0, v["q", t]: -1.0}        # elec_splitter.py:115-117
        if with_pem:
            row[v["p", t]] = -1.0
        B.eq(row)
        B.eq({v["q", t]: 1.0, v["i", t]: -1.0})                         # arc splitter_to_battery
        if with_pem:
            B.eq({v["p", t]: 1.0, v["pe", t]: -1.0})                    # arc splitter_to_pem
            B.eq({v["h", t]: 1.0, v["pe", t]: -PEM_ELEC_TO_MOL})        # pem_electrolyzer.py:111-114
        B.eq({v["s", t]: 1.0, v["s0", t]: -1.0, v["i", t]: -ETA_C, v["o", t]: 1.0 / ETA_D})   # battery.py:145-149
        B.eq({v["e", t]: 1.0, v["e0", t]: -1.0, v["i", t]: -0.5, v["o", t]: -0.5})            # battery.py:151-153
        B.le({v["s", t]: 1.0, v["E", t]: -1.0, v["e", t]: DEGRADATION})                        # battery.py:155-157
        B.le({v["i", t]: 1.0, v["P", t]: -1.0})                                                 # battery.py:159-161
        B.le({v["o", t]: 1.0, v["P", t]: -1.0})                                                 # battery.py:163-165
        B.eq({v["P", t]: DURATION, v["E", t]: -1.0})                                            # RE_flowsheet.py:155-156
        B.le({v["s0", t]: 1.0, v["s", t]: -1.0}, BATTERY_RAMP_RATE)                             # wind_battery_LMP.py:139-140
        B.le({v["s", t]: 1.0, v["s0", t]: -1.0}, BATTERY_RAMP_RATE)                             # :141-142
    # linking t -> t+1 (wind_battery_LMP.py:32-36)
    for t in range(T - 1):
        B.eq({v["s", t]: 1.0, v["s0", t + 1]: -1.0})
        B.eq({v["e", t]: 1.0, v["e0", t + 1]: -1.0})
        B.eq({v["P", t]: 1.0, v["P", t + 1]: -1.0})
    # periodic (:48-49)
    B.eq({v["s", T - 1]: 1.0, v["s0", 0]: -1.0})
    if T > 1:
        B.eq({v["P", T - 1]: 1.0, v["P", 0]: -1.0})
    # fixed initial conditions: wind_battery_LMP.py:206-207 fixes both; the PEM variant only e0 (:217)
    B.lb[v["e0", 0]] = B.ub[v["e0", 0]] = 0.0
    if not with_pem:
        B.lb[v["s0", 0]] = B.ub[v["s0", 0]] = 0.0
    # model-level capacities (:209-219)
    Wc = B.var("wind_system_capacity", ub=wind_mw_ub * 1e3)
    Bc = B.var("battery_system_capacity")
    if with_pem and design_opt is not False and extant_wind:
        B.lb[Wc] = B.ub[Wc] = wind_mw * 1e3                             # wind_battery_PEM_LMP.py:230-231
    if with_pem:
        Pc = B.var("pem_system_capacity", fix=(pem_mw * 1e3 if design_opt is False else None))
    for t in range(T):
        B.le({v["C", t]: 1.0, Wc: -1.0})
        B.le({v["P", t]: 1.0, Bc: -1.0})
        if with_pem:
            B.le({v["pe", t]: 1.0, Pc: -1.0})
    # objective  -NPV*1e-5
    n_weeks = T / 168.0
    ann = 52.0 / n_weeks
    wind_cap = 0.0 if extant_wind else WIND_CAP_COST
    npv = {}           # NPV as a linear form over columns
    _add(npv, Wc, -wind_cap)
    _add(npv, Bc, -(BATT_CAP_COST_KW + BATT_CAP_COST_KWH * DURATION))
    if with_pem:
        _add(npv, Pc, -PEM_CAP_COST)
    for t in range(T):
        B.cost_lmp(v["g", t], t, -1e-5 * PA * ann * 1e-3, lmp)
        B.cost_lmp(v["o", t], t, -1e-5 * PA * ann * 1e-3, lmp)
        _add(npv, Wc, -PA * ann * WIND_OP_COST / 8760.0)
        _add(npv, Bc, -PA * ann * BATT_OP_COST / 8760.0)
        if with_pem:
            _add(npv, Pc, -PA * ann * PEM_OP_COST / 8760.0)
            _add(npv, v["pe", t], -PA * ann * PEM_VAR_COST)
            _add(npv, v["h", t], PA * ann * h2_price / H2_MOLS_PER_KG * 3600.0)
    for j, val in npv.items():
        B.cost(j, -val * 1e-5)
    meta = dict(kind="wind_battery_pem" if with_pem else "wind_battery", T=T, v=v, Wc=Wc, Bc=Bc,
                Pc=(Pc if with_pem else None), ann=ann, h2_price=h2_price)
    return B.finish(meta)


def wind_pem_closed_form(lmp, cf, wind_mw, pem_mw, h2_price, design_opt=False):
    """Exact optimum of the wind + PEM price-taker LP with batt_mw = 0 (the sweep of run_pricetaker_wind_PEM.py:37-41): without
    storage the LP of wind_battery_PEM_LMP.py:217-294 separates per hour --  max lam_t g + kh p  s.t.  g + p <= W cf_t, p <= Pc --
    so p = min(Pc, W cf_t) where the H2 value kh exceeds the price, and g takes the rest where the price is positive.
    With design_opt="PEM" the NPV is concave piecewise linear in Pc with breakpoints at the hourly wind outputs: the optimum is
    the best breakpoint.  Returns dict(NPV, annual_rev_h2, annual_rev_E, pem_kw).  Checked against HiGHS on the raw LP in
    tests/test_oracle_golden.py; lets the CPU suite pin EVERY PEM > 0 row of the reference's committed table in milliseconds."""
    lmp = np.asarray(lmp, float); cf = np.asarray(cf, float)
    T = lmp.size
    W = wind_mw * 1e3
    ann = 52.0 / (T / 168.0)
    kh = h2_price * PEM_ELEC_TO_MOL / H2_MOLS_PER_KG * 3600.0 - PEM_VAR_COST          # $ per kWh sent to the PEM (:276, :268)
    lam = lmp * 1e-3
    avail = W * cf

    def evaluate(Pc):
        Pc = np.atleast_1d(np.asarray(Pc, float))[:, None]
        pe = np.where((kh > lam) & (kh > 0), np.minimum(Pc, avail), 0.0)
        g = np.where(lam > 0, avail - pe, 0.0)
        h2 = (kh * pe).sum(1) * ann
        el = (lam * g).sum(1) * ann
        fixed = (W * WIND_OP_COST + Pc[:, 0] * PEM_OP_COST) / 8760.0 * T * ann
        return h2, el, -(PEM_CAP_COST * Pc[:, 0]) + PA * (h2 + el - fixed)

    if design_opt == "PEM":
        cand = np.unique(np.concatenate([[0.0], avail]))
        best = -np.inf
        for lo in range(0, cand.size, 512):                    # chunks keep the [candidates, T] temporaries small
            h2, el, npv = evaluate(cand[lo:lo + 512])
            k = int(np.argmax(npv))
            if npv[k] > best:
                best, out = npv[k], dict(NPV=float(npv[k]), annual_rev_h2=float(h2[k]), annual_rev_E=float(el[k]), pem_kw=float(cand[lo + k]))
        return out
    h2, el, npv = evaluate(pem_mw * 1e3)
    return dict(NPV=float(npv[0]), annual_rev_h2=float(h2[0]), annual_rev_E=float(el[0]), pem_kw=pem_mw * 1e3)


def wind_battery_report(lp: RawLP, x, lmp):
    """Quantities the reference reads back (wind_battery_LMP.py:252-263, record_results :272-325;
    wind_battery_PEM_LMP.py:300-330)."""
    m = lp.meta; v = m["v"]; T = m["T"]; ann = m["ann"]
    lam = np.asarray(lmp, float) * 1e-3
    g = np.array([x[v["g", t]] for t in range(T)]); o = np.array([x[v["o", t]] for t in range(T)])
    elec_rev = float(np.sum(lam * (g + o)))
    out = dict(NPV=-(lp.c @ x + lp.c0) * 1e5, annual_elec_revenue=elec_rev * ann,
               total_elec_output=float(np.sum(g + o)) * ann)
    fixed = x[m["Wc"]] * WIND_OP_COST / 8760.0 + x[m["Bc"]] * BATT_OP_COST / 8760.0
    h2_rev = 0.0
    if m["Pc"] is not None:
        h = np.array([x[v["h", t]] for t in range(T)])
        h2_rev = float(np.sum(h)) * m["h2_price"] / H2_MOLS_PER_KG * 3600.0
        fixed += x[m["Pc"]] * PEM_OP_COST / 8760.0
        out["annual_rev_h2"] = h2_rev * ann
        out["annual_rev_E"] = elec_rev * ann
    out["annual_revenue"] = (elec_rev + h2_rev - fixed * T) * ann
    return out


# --------------------------------------------------------------------------------------
# A.3  nuclear + PEM + tank, 48-h dispatch (nuclear_flowsheet_multiperiod_class.py:72-155)
# --------------------------------------------------------------------------------------
MW_H2 = 2.016e-3
NUC_PEM_ELEC_TO_MOL = 0.002527406    # nuclear_flowsheet.py:269


def nuclear_raw(lmp, np_capacity=500.0, pem_capacity=100.0, tank_capacity=5000.0,
                h2_demand=0.35, h2_price=4.0):
    """Raw LP: max sum_t [ lmp_t * np_to_grid_t * 1e-3 - operating_cost_t ]  written as a minimisation.

    Per block (nuclear_flowsheet.py:119-157, fixed values :263-291): splitter electricity fixed at
    np_capacity*1e3 kW with split-fraction Vars (elec_splitter.py:119-129, linear because the inlet is fixed),
    pem.electricity <= pem_capacity*1e3, flow = 0.002527406*electricity, simple tank balance with dt = 3600
    (hydrogen_tank_simplified.py:177-184), turbine outlet fixed 0, pipeline flow <= h2_demand/mw,
    tank_holdup_previous <= tank_capacity/mw, link holdup[t] -> holdup_previous[t+1], holdup_previous[0]=0.
    Operating cost (…_class.py:149-153).
    """
    lmp = np.asarray(lmp, float); T = lmp.size
    B = _Builder(); v = {}
    nuclear_blocks(B, v, T, np_capacity, pem_capacity, tank_capacity, h2_demand)
    for t in range(T):
        # objective (minimise cost - revenue)
        B.cost_lmp(v["xg", t], t, -1e-3, lmp)
        for j, a in nuclear_operating_cost(v, t, h2_price).items():
            B.cost(j, a)
    return B.finish(dict(kind="nuclear", T=T, v=v))


def nuclear_blocks(B, v, T, np_capacity=500.0, pem_capacity=100.0, tank_capacity=5000.0, h2_demand=0.35, holdup0=0.0):
    """The T period blocks + holdup links of create_multiperiod_nuclear_model (rows only, no objective)."""
    E = np_capacity * 1e3
    for t in range(T):
        p = f"blk[{t}].fs."
        v["E", t] = B.var(p + "np_power_split.electricity[0]", fix=E)
        v["fg", t] = B.var(p + "np_power_split.split_fraction[np_to_grid,0]", ub=1.0)
        v["fp", t] = B.var(p + "np_power_split.split_fraction[np_to_pem,0]", ub=1.0)
        v["xg", t] = B.var(p + "np_power_split.np_to_grid_elec[0]")
        v["xs", t] = B.var(p + "np_power_split.np_to_pem_elec[0]")
        v["xp", t] = B.var(p + "pem.electricity[0]", lb=-INF, ub=pem_capacity * 1e3)
        v["f", t] = B.var(p + "pem.outlet.flow_mol[0]", lb=-INF)
        v["fi", t] = B.var(p + "h2_tank.inlet.flow_mol[0]", lb=-INF)
        v["Hp", t] = B.var(p + "h2_tank.tank_holdup_previous[0]", ub=tank_capacity / MW_H2)
        v["H", t] = B.var(p + "h2_tank.tank_holdup[0]")
        v["u", t] = B.var(p + "h2_tank.outlet_to_pipeline.flow_mol[0]", ub=h2_demand / MW_H2)
        v["vt", t] = B.var(p + "h2_tank.outlet_to_turbine.flow_mol[0]", fix=0.0)
        B.eq({v["E", t]: 1.0, v["xg", t]: -1.0, v["xs", t]: -1.0})            # sum of outlets
        B.eq({v["xg", t]: 1.0, v["fg", t]: -E})                               # outlet = sf * E (E fixed)
        B.eq({v["xs", t]: 1.0, v["fp", t]: -E})
        B.eq({v["xs", t]: 1.0, v["xp", t]: -1.0})                             # arc np -> pem
        B.eq({v["f", t]: 1.0, v["xp", t]: -NUC_PEM_ELEC_TO_MOL})
        B.eq({v["f", t]: 1.0, v["fi", t]: -1.0})                              # arc pem -> tank
        B.eq({v["H", t]: 1.0, v["Hp", t]: -1.0, v["fi", t]: -3600.0,
              v["u", t]: 3600.0, v["vt", t]: 3600.0})
    for t in range(T - 1):
        B.eq({v["H", t]: 1.0, v["Hp", t + 1]: -1.0})
    B.lb[v["Hp", 0]] = B.ub[v["Hp", 0]] = float(holdup0)


def nuclear_operating_cost(v, t, h2_price=4.0):
    """fs.operating_cost of period t as a linear form (nuclear_flowsheet_multiperiod_class.py:149-153)."""
    return {v["E", t]: 1e-3 * 2.3, v["xp", t]: 1e-3 * 1.3, v["H", t]: MW_H2 * 0.01, v["u", t]: -MW_H2 * 3600.0 * h2_price}


# --------------------------------------------------------------------------------------
# A.4  fossil USC + molten-salt storage: STRUCTURE-ONLY LP SURROGATE  (parity unpinned)
# --------------------------------------------------------------------------------------
FOSSIL = dict(
    p_lo=283.0, p_hi=436.0,            # plant power bounds, multiperiod_integrated_storage_usc.py:49-50,75-80
    pprev_lo=284.0, pprev_hi=466.0,    # previous_power bounds :51-54,89-94
    hx_lo=10.0, hx_hi=200.0,           # storage duty bounds :82-86
    ramp=60.0,                         # :125-135
    salt_total=6739292.0,              # :98
    hot_init=75000.0 + 1103053.48,     # hot tank starts near-empty + min level (:111-123, surrogate constant)
    pprev0=447.66,                     # :114
    # --- surrogate constants (OURS, not the reference's NLP): linear duty->salt-flow, duty->power, heat-rate cost
    kc=6.5, kd=7.0,                    # kg/s of salt per MW of charge / discharge duty
    eta_c=0.40, eta_d=0.38,            # MW of net power lost / gained per MW of duty
    fuel=22.0, fixed=6.0,              # $/MWh of plant power ; $/h constant
)


def fossil_surrogate_raw(lmp, par=None):
    """Linear surrogate of the USC + TES price-taker (pricetaker_with_multiperiod_integrated_storage_usc.py:69-156).
    Keeps every *linear* inter-period piece of the reference and replaces the per-period steam-cycle NLP by
    three linear maps (documented in FOSSIL).  Not comparable with the reference's IPOPT objective."""
    P = dict(FOSSIL); P.update(par or {})
    lmp = np.asarray(lmp, float); T = lmp.size
    B = _Builder(); v = {}
    for t in range(T):
        p = f"blk[{t}].fs."
        v["P", t] = B.var(p + "plant_power_out[0]", lb=P["p_lo"], ub=P["p_hi"])
        v["Pp", t] = B.var(p + "previous_power", lb=P["pprev_lo"], ub=P["pprev_hi"])
        v["c", t] = B.var(p + "hxc.heat_duty[0]", lb=P["hx_lo"], ub=P["hx_hi"])
        v["d", t] = B.var(p + "hxd.heat_duty[0]", lb=P["hx_lo"], ub=P["hx_hi"])
        v["hp", t] = B.var(p + "previous_salt_inventory_hot", ub=P["salt_total"])
        v["h", t] = B.var(p + "salt_inventory_hot", ub=P["salt_total"])
        v["cp", t] = B.var(p + "previous_salt_inventory_cold", ub=P["salt_total"])
        v["cl", t] = B.var(p + "salt_inventory_cold", ub=P["salt_total"])
        B.le({v["P", t]: 1.0, v["Pp", t]: -1.0}, P["ramp"])                    # :125-135
        B.le({v["Pp", t]: 1.0, v["P", t]: -1.0}, P["ramp"])
        B.eq({v["h", t]: 1.0, v["hp", t]: -1.0, v["c", t]: -3600.0 * P["kc"], v["d", t]: 3600.0 * P["kd"]})   # :137-144
        B.le({v["d", t]: 3600.0 * P["kd"], v["hp", t]: -1.0})                 # :146-151
        B.le({v["c", t]: 3600.0 * P["kc"], v["cp", t]: -1.0})                 # :153-158
        B.eq({v["h", t]: 1.0, v["cl", t]: 1.0}, P["salt_total"])              # :160-164
        B.eq({v["hp", t]: 1.0, v["cp", t]: 1.0}, P["salt_total"])
        # minimise -(lmp*net_power - opcost)
        B.cost(v["P", t], P["fuel"])
        B.cost_lmp(v["P", t], t, -1.0, lmp)
        B.cost_lmp(v["c", t], t, P["eta_c"], lmp)
        B.cost_lmp(v["d", t], t, -P["eta_d"], lmp)
        B.c0 += P["fixed"]
    for t in range(T - 1):                                                     # :334-342
        B.eq({v["h", t]: 1.0, v["hp", t + 1]: -1.0})
        B.eq({v["P", t]: 1.0, v["Pp", t + 1]: -1.0})
    B.eq({v["h", T - 1]: 1.0, v["hp", 0]: -1.0})                               # periodic, pricetaker…:88-90
    B.lb[v["hp", 0]] = B.ub[v["hp", 0]] = P["hot_init"]
    B.lb[v["Pp", 0]] = B.ub[v["Pp", 0]] = P["pprev0"]
    return B.finish(dict(kind="fossil_surrogate", T=T, v=v))


# --------------------------------------------------------------------------------------
# A.3 (report variant)  nuclear + PEM (+tank, +turbine) price-taker, price_taker_analysis.py:116-322
# --------------------------------------------------------------------------------------
def nuclear_report_raw(lmp, h2_price, pem_cap_mw, pem_capex=400.0, tank_cap=0.0, turbine_cap=0.0, demand=400.0 * 20,
                       vom_pem=0.0, plant_life=30, tax_rate=0.2, discount_rate=0.08):
    """Raw LP of run_exhaustive_enumeration's inner solve (price_taker_analysis.py:353-403) for T = len(lmp) hours:
    build_ne_flowsheet rows (:143-170), capacity rows (:199-213), demand bound (:219-220), cash flow (:239-254),
    NPV pieces (:274-308), annualised objective (:318-322, maximise -> minimise the negative)."""
    lmp = np.asarray(lmp, float); T = lmp.size
    B = _Builder(); v = {}
    fom_pem = 0.03 * pem_capex
    capex = pem_capex * 1000.0 * pem_cap_mw + 29.0 * 33.3 * tank_cap + 947.0 * 1000.0 * turbine_cap
    fom = 1000.0 * fom_pem * pem_cap_mw + 1000.0 * 7.0 * turbine_cap + 120.0 * 1000.0 * 400.0
    dep = capex / plant_life
    cf = (1.0 - (1.0 + discount_rate) ** (-plant_life)) / discount_rate
    # objective: net_profit - capex/cf = dep + (1-tax)(sum cash - fom - dep) - capex/cf
    k = 1.0 - tax_rate
    for t in range(T):
        p = f"period[{t + 1}].fs."
        v["np", t] = B.var(p + "np_power", fix=400.0)
        v["g", t] = B.var(p + "np_to_grid")
        v["e", t] = B.var(p + "np_to_electrolyzer")
        v["h", t] = B.var(p + "h2_production")
        v["H", t] = B.var(p + "tank_holdup")
        v["Hp", t] = B.var(p + "tank_holdup_previous")
        v["u", t] = B.var(p + "h2_to_pipeline", ub=demand)
        v["tb", t] = B.var(p + "h2_to_turbine")
        v["tp", t] = B.var(p + "h2_turbine_power")
        v["n", t] = B.var(p + "net_power")
        B.eq({v["np", t]: 1.0, v["g", t]: -1.0, v["e", t]: -1.0})
        B.eq({v["h", t]: 1.0, v["e", t]: -20.0})
        B.eq({v["H", t]: 1.0, v["Hp", t]: -1.0, v["h", t]: -1.0, v["u", t]: 1.0, v["tb", t]: 1.0})
        B.eq({v["tp", t]: 1.0, v["tb", t]: -0.0125})
        B.eq({v["n", t]: 1.0, v["g", t]: -1.0, v["tp", t]: -1.0})
        B.le({v["e", t]: 1.0}, pem_cap_mw)
        B.le({v["H", t]: 1.0}, tank_cap)
        B.le({v["tp", t]: 1.0}, turbine_cap)
        B.cost(v["u", t], -k * h2_price)
        B.cost_lmp(v["n", t], t, -k, lmp)
        B.cost(v["e", t], k * vom_pem); B.cost(v["tp", t], k * 4.25); B.cost(v["np", t], k * 2.3)
    for t in range(T - 1):
        B.eq({v["H", t]: 1.0, v["Hp", t + 1]: -1.0})
    B.lb[v["Hp", 0]] = B.ub[v["Hp", 0]] = 0.0
    B.c0 = -(dep + k * (-fom - dep) - capex / cf)
    return B.finish(dict(kind="nuclear_report", T=T, v=v, capex=capex, fom=fom))


# --------------------------------------------------------------------------------------
# A.6  PV + battery + PEM + hydrogen tank + hydrogen turbine, load following with reserves
# --------------------------------------------------------------------------------------
# constants of case_studies/renewables_case/solar_battery_hydrogen_inputs.py:22-70 (they override load_parameters.py where the
# names coincide; "pem_cap_cost" keeps load_parameters' 1200 because the inputs file defines pem_cap_cost_kw instead)
SOLAR = dict(
    h2_blend_ratio=1.0, s_per_ts=3600.0, timestep_hrs=1.0,
    pv_cap_cost=WIND_CAP_COST, pv_op_cost=WIND_OP_COST,            # inputs.py:96-97 reuse the wind numbers
    batt_cap_cost_kw=236.36 * 0.5, batt_cap_cost_kwh=254.83 * 0.5,  # tax incentive 0.5
    pem_cap_cost=PEM_CAP_COST, pem_op_cost=47.9, pem_var_cost=1.3 / 1000.0,
    tank_cap_cost_per_kg=500.0, tank_op_cost=0.17 * 500.0,
    turbine_cap_cost=1320.0, turbine_op_cost=11.65, turbine_var_cost=3.0 / 1000.0,
    h2_price_per_kg=2.5, capacity_requirement=100.0, capacity_credit_battery=0.33,
    turbine_min_mw=0.0, turbine_ramp_mw_per_min=100.0,
    h2_turb_conv=0.39 * 33.391, ng_turb_conv=0.33 * 13.09, mmbtu_to_ng_kg=20.133,
    pv_mw=200.0, batt_mw=0.0, batt_mwh=0.0, pem_mw=0.0, turb_mw=100.0,
    flow_mol_ub=1e5,                                                 # properties/h2_ideal_vap.py:86 state bound
    kg_to_tons=0.00110231,
)
SOLAR["tank_size"] = SOLAR["capacity_requirement"] * 1e3 / SOLAR["h2_turb_conv"]       # inputs.py:86 (blend ratio 1)


def solar_default_series():
    """inputs.py:57-61: sinusoidal PV capacity factors, 100 MW load and reserve, 3 $/MMBtu gas, for one 24-hour day"""
    return dict(pv_cfs=np.sin(np.deg2rad(np.linspace(0, 180, 24))), load_mw=np.full(24, 100.0), reserve_mw=np.full(24, 100.0),
                ng_prices=np.full(24, 3.0))


def solar_battery_hydrogen_raw(lmp, design_opt, par=None, pv_cfs=None, load_mw=None, reserve_mw=None, max_sales=1000.0, max_purchases=1000.0):
    """Raw LP of pv_battery_hydrogen_optimize (case_studies/renewables_case/solar_battery_hydrogen.py:375-457): per-period flowsheet
    :123-174 (create_model with re_type='pv', RE_flowsheet.py:338-396; unit_models/solar_pv.py:82-84,
    hydrogen_tank_simplified.py:177-184), linking / periodic pairs :33-61, size_constraints :205-236, capital / fixed / variable
    costs :238-305, add_load_following_obj :308-372.  h2_blend_ratio = 1: no natural gas (ng_kg is a zero Param, :149-150).
    Objective = -NPV * 1e-3 (:372).  Pinned to the reference's known answers tests/test_solar_battery_hydrogen.py:20-48."""
    P = dict(SOLAR); P.update(par or {})
    d = solar_default_series()
    cf = np.asarray(d["pv_cfs"] if pv_cfs is None else pv_cfs, float)
    load = np.asarray(d["load_mw"] if load_mw is None else load_mw, float)
    reserve = np.asarray(d["reserve_mw"] if reserve_mw is None else reserve_mw, float)
    lmp = np.asarray(lmp, float)
    T = lmp.size
    assert cf.size == T and load.size == T and reserve.size == T
    B = _Builder()
    v = {}
    batt_mwh = P["batt_mwh"] if "batt_hr" not in P else P["batt_mw"] * P["batt_hr"]
    fmax = P["flow_mol_ub"]
    k_turb = P["s_per_ts"] / H2_MOLS_PER_KG * P["h2_turb_conv"]              # kW of turbine output per mol/s of hydrogen (:147,:157)
    for t in range(T):
        p = f"blk[{t}].fs."
        fx = (lambda val: None) if design_opt else (lambda val: val)
        v["C", t] = B.var(p + "pv.system_capacity", fix=fx(P["pv_mw"] * 1e3))              # :166-170
        v["w", t] = B.var(p + "pv.electricity[0]")
        v["z", t] = B.var(p + "splitter.electricity[0]")
        v["g", t] = B.var(p + "splitter.grid_elec[0]")
        v["p", t] = B.var(p + "splitter.pem_elec[0]")
        v["q", t] = B.var(p + "splitter.battery_elec[0]")
        v["P", t] = B.var(p + "battery.nameplate_power", ub=1e8, fix=fx(P["batt_mw"] * 1e3))
        v["E", t] = B.var(p + "battery.nameplate_energy", ub=1e9, fix=fx(batt_mwh * 1e3))
        v["s0", t] = B.var(p + "battery.initial_state_of_charge")
        v["e0", t] = B.var(p + "battery.initial_energy_throughput")
        v["i", t] = B.var(p + "battery.elec_in[0]")
        v["o", t] = B.var(p + "battery.elec_out[0]")
        v["s", t] = B.var(p + "battery.state_of_charge[0]")
        v["e", t] = B.var(p + "battery.energy_throughput[0]")
        v["pe", t] = B.var(p + "pem.electricity[0]", lb=-INF)
        v["h", t] = B.var(p + "pem.outlet.flow_mol[0]", ub=fmax)
        v["ti", t] = B.var(p + "h2_tank.inlet.flow_mol[0]", ub=fmax)
        v["tt", t] = B.var(p + "h2_tank.outlet_to_turbine.flow_mol[0]", ub=fmax)
        v["tp", t] = B.var(p + "h2_tank.outlet_to_pipeline.flow_mol[0]", ub=fmax)
        v["hp", t] = B.var(p + "h2_tank.tank_holdup_previous[0]")
        v["hd", t] = B.var(p + "h2_tank.tank_holdup[0]")
        v["gp", t] = B.var(f"blk[{t}].grid_purchase", ub=max_purchases * 1e3)               # :358-360
        v["gs", t] = B.var(f"blk[{t}].grid_sales", ub=max_sales * 1e3)                      # :355-357
        v["tr", t] = B.var(f"blk[{t}].turbine_reserve")
        v["br", t] = B.var(f"blk[{t}].battery_reserve")
        B.le({v["w", t]: 1.0, v["C", t]: -cf[t]})                                           # solar_pv.py:82-84
        B.eq({v["w", t]: 1.0, v["z", t]: -1.0})                                             # arc pv_to_splitter
        B.eq({v["z", t]: 1.0, v["g", t]: -1.0, v["p", t]: -1.0, v["q", t]: -1.0})           # elec_splitter.py:115-117
        B.eq({v["p", t]: 1.0, v["pe", t]: -1.0})                                            # arc splitter_to_pem
        B.eq({v["q", t]: 1.0, v["i", t]: -1.0})                                             # arc splitter_to_battery
        B.eq({v["h", t]: 1.0, v["pe", t]: -PEM_ELEC_TO_MOL})                                # pem_electrolyzer.py:111-114
        B.eq({v["ti", t]: 1.0, v["h", t]: -1.0})                                            # arc pem_to_tank
        B.eq({v["hd", t]: 1.0, v["hp", t]: -1.0, v["ti", t]: -P["s_per_ts"], v["tp", t]: P["s_per_ts"], v["tt", t]: P["s_per_ts"]})
        B.eq({v["s", t]: 1.0, v["s0", t]: -1.0, v["i", t]: -ETA_C, v["o", t]: 1.0 / ETA_D})  # battery.py:145-149
        B.eq({v["e", t]: 1.0, v["e0", t]: -1.0, v["i", t]: -0.5, v["o", t]: -0.5})           # battery.py:151-153
        B.le({v["s", t]: 1.0, v["E", t]: -1.0})                                              # battery.py:155-157, degradation_rate = 0 (:173)
        B.le({v["i", t]: 1.0, v["P", t]: -1.0})
        B.le({v["o", t]: 1.0, v["P", t]: -1.0})
        B.le({v["tt", t]: -k_turb}, -P["turbine_min_mw"] * 1e3)                              # h2_turbine_pmin :159
    for t in range(T - 1):                                                                   # :42-47
        B.eq({v["hd", t]: 1.0, v["hp", t + 1]: -1.0})
        B.eq({v["s", t]: 1.0, v["s0", t + 1]: -1.0})
        B.eq({v["e", t]: 1.0, v["e0", t + 1]: -1.0})
        B.eq({v["P", t]: 1.0, v["P", t + 1]: -1.0})
    B.eq({v["hd", T - 1]: 1.0, v["hp", 0]: -1.0})                                            # :58-61
    B.eq({v["s", T - 1]: 1.0, v["s0", 0]: -1.0})
    if T > 1:
        B.eq({v["P", T - 1]: 1.0, v["P", 0]: -1.0})
    # ---- size_constraints :205-236
    fx = (lambda val: None) if design_opt else (lambda val: val)
    pv_base = P["pv_mw"] * 1e3
    Va = B.var("pv_add_system_capacity", ub=1e7, fix=fx(0.0))
    Bc = B.var("battery_system_capacity", ub=1e7, fix=fx(P["batt_mw"] * 1e3))
    Be = B.var("battery_system_energy", fix=fx(batt_mwh * 1e3))
    Pc = B.var("pem_system_capacity", ub=1e7, fix=fx(P["pem_mw"] * 1e3))
    Ts = B.var("h2_tank_size", ub=1e7, fix=fx(P["tank_size"]))
    Tc = B.var("turb_system_capacity", lb=P["turb_mw"] * 1e3, ub=1e8, fix=fx(P["turb_mw"] * 1e3))
    for t in range(T):
        B.le({v["C", t]: 1.0, Va: -1.0}, pv_base)
        B.le({v["P", t]: 1.0, Bc: -1.0})
        B.le({v["E", t]: 1.0, Be: -1.0})
        B.le({v["pe", t]: 1.0, Pc: -1.0})
        B.le({v["hd", t]: 1.0 / H2_MOLS_PER_KG, Ts: -1.0})
        B.le({v["tt", t]: k_turb, Tc: -1.0})
    B.le({Bc: 0.5, Be: -1.0})                                                                # battery between 0.5 and 8 hours
    B.le({Be: 1.0, Bc: -8.0})
    # ---- add_load_following_obj :308-372
    ramp = P["turbine_ramp_mw_per_min"] * 1e3
    for t in range(T):
        tprev = (t - 1) % T
        if T > 1:
            B.le({v["tt", tprev]: k_turb, v["tt", t]: -k_turb}, ramp)                        # :323-324
            B.le({v["tt", t]: k_turb, v["tt", tprev]: -k_turb}, ramp)
        B.eq({v["g", t]: 1.0, v["o", t]: 1.0, v["tt", t]: k_turb, v["gp", t]: 1.0, v["gs", t]: -1.0}, load[t] * 1e3)    # meet_load :328-331
        B.le({v["tr", t]: 1.0, v["hd", t]: -P["h2_turb_conv"] / H2_MOLS_PER_KG})            # turbine_reserve_lb1 (blend ratio 1: no gas reserve)
        B.le({v["tr", t]: 1.0, Tc: -1.0, v["tt", t]: k_turb})                                # turbine_reserve_lb2
        B.le({v["br", t]: 1.0, Bc: -1.0})
        B.le({v["br", t]: 1.0, v["s", t]: -1.0})
        r1 = (max(reserve[max(t - int(1 / P["timestep_hrs"]), 0):t]) if t > 0 else reserve[0]) * 1e3                  # :347-348
        # total_reserve = battery_reserve + turbine_reserve + (C cf - w) + pem.electricity >= reserve_over_1hr
        B.le({v["br", t]: -1.0, v["tr", t]: -1.0, v["C", t]: -cf[t], v["w", t]: 1.0, v["pe", t]: -1.0}, -r1)
        B.le({Bc: -P["capacity_credit_battery"], Tc: -1.0}, -P["capacity_requirement"] * 1e3)                         # cap_requirement :352
        B.le({v["gs", t]: 1.0, v["gp", t]: -1.0}, max_sales * 1e3)                            # :356
        B.le({v["gp", t]: 1.0, v["gs", t]: -1.0}, max_purchases * 1e3)                        # :359
    n_weeks = T / 168.0
    ann = 52.143 / n_weeks
    npv = {}
    cap_const = P["turbine_cap_cost"] * P["turb_mw"] * 1e3            # - turb_cap_cost * (Tc - turb_mw*1e3): the constant part
    _add(npv, Va, -P["pv_cap_cost"]); _add(npv, Bc, -P["batt_cap_cost_kw"]); _add(npv, Be, -P["batt_cap_cost_kwh"])
    _add(npv, Pc, -P["pem_cap_cost"]); _add(npv, Ts, -P["tank_cap_cost_per_kg"]); _add(npv, Tc, -P["turbine_cap_cost"])
    _add(npv, Va, -PA * P["pv_op_cost"]); _add(npv, Pc, -PA * P["pem_op_cost"]); _add(npv, Ts, -PA * P["tank_op_cost"])
    _add(npv, Tc, -PA * P["turbine_op_cost"])
    fixed_const = -PA * pv_base * P["pv_op_cost"]
    for t in range(T):
        B.cost_lmp(v["gs", t], t, -1e-3 * PA * ann * 1e-3, lmp)                               # grid_cost = LMP (purchase - sales) 1e-3
        B.cost_lmp(v["gp", t], t, +1e-3 * PA * ann * 1e-3, lmp)
        _add(npv, v["pe", t], -PA * ann * P["pem_var_cost"])
        _add(npv, v["tt", t], -PA * ann * P["turbine_var_cost"] * k_turb)
        _add(npv, v["tp", t], PA * ann * P["h2_price_per_kg"] / H2_MOLS_PER_KG * P["s_per_ts"])
    for j, val in npv.items():
        B.cost(j, -val * 1e-3)
    B.c0 = -(cap_const + fixed_const) * 1e-3
    meta = dict(kind="solar_battery_hydrogen", T=T, v=v, Va=Va, Bc=Bc, Be=Be, Pc=Pc, Ts=Ts, Tc=Tc, ann=ann, par=P, k_turb=k_turb,
                pv_base=pv_base, design_opt=bool(design_opt))
    return B.finish(meta)


def solar_report(lp: RawLP, x):
    """the entries of design_res the reference's test asserts (solar_battery_hydrogen.py:513-548)"""
    m, P = lp.meta, lp.meta["par"]
    cap = (P["pv_cap_cost"] * x[m["Va"]] + P["batt_cap_cost_kw"] * x[m["Bc"]] + P["batt_cap_cost_kwh"] * x[m["Be"]]
           + P["pem_cap_cost"] * x[m["Pc"]] + P["tank_cap_cost_per_kg"] * x[m["Ts"]]
           + P["turbine_cap_cost"] * (x[m["Tc"]] - P["turb_mw"] * 1e3))
    return dict(pv_mw=(m["pv_base"] + x[m["Va"]]) * 1e-3, batt_mw=x[m["Bc"]] * 1e-3, batt_mwh=x[m["Be"]] * 1e-3, pem_mw=x[m["Pc"]] * 1e-3,
                tank_tonH2=x[m["Ts"]] * P["kg_to_tons"], turb_mw=x[m["Tc"]] * 1e-3, capital_cost=cap,
                NPV=-(lp.c @ x + lp.c0) * 1e3)
