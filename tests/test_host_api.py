"""Host-side mirror of the reference interface (no GPU): argument handling of dispatches_b200/pricetaker.py."""
import numpy as np
import pytest

from dispatches_b200 import pricetaker as PT
from dispatches_b200 import scenarios as SC
from dispatches_b200 import templates as TP


def reference_params(lmp, cf, W, P, **kw):
    d = {"wind_mw": W, "wind_mw_ub": 10000, "batt_mw": P, "design_opt": False, "extant_wind": True,
         "wind_resource": {t: {"wind_resource_config": {"capacity_factor": [cf[t]]}} for t in range(len(cf))},
         "DA_LMPs": lmp}
    d.update(kw)
    return d


def test_design_opt_is_refused_not_silently_fixed():
    lmp, cf, W, P = SC.c2(2)
    # a free wind size with one capacity-factor series per scenario is a per-problem MATRIX coefficient (supported since round 2);
    # what is still refused is a series count that matches neither 1 nor the number of scenarios
    params = reference_params(lmp, cf, W, P, design_opt=True, extant_wind=False)
    params["wind_resource"] = np.stack([cf, 0.5 * cf, 0.25 * cf])
    with pytest.raises(ValueError):
        PT.wind_battery_optimize(24, params)
    with pytest.raises(NotImplementedError):
        PT.wind_battery_pem_optimize(24, reference_params(lmp, cf, W, P, design_opt=True, pem_mw=100, h2_price_per_kg=2))


def test_capacity_factor_dict_and_lmp_shapes():
    lmp, cf, W, P = SC.c2(3)
    params = reference_params(lmp, cf, W, P)
    assert np.array_equal(PT._capacity_factors(params, 24), cf)          # the reference's {t: {...}} layout
    assert PT._lmps(params, 24).shape == (3, 24)
    assert PT._lmps(dict(params, DA_LMPs=np.arange(30.0)), 24).shape == (1, 24)   # 1-D signal longer than T, like DA_LMPs[:T]
    with pytest.raises(ValueError):
        PT._lmps(dict(params, DA_LMPs=np.arange(10.0)), 24)


def test_rparams_layout_matches_the_stage_descriptor():
    t = TP.wind_battery(24)
    st = t.meta["stage_wb"]
    rp = TP.wind_battery_rparams(24, np.linspace(0, 1, 24), 100.0, 25.0)[0]
    assert rp[st["p_off"]] == 25e3 and rp[st["wcf_off"] + 5] == pytest.approx(100e3 * 5 / 23)
    c, b, u, k = t.instantiate(np.ones(24), rp)
    g0 = t.col_names.index("blk[0].fs.splitter.grid_elec[0]")
    assert c[g0] == pytest.approx(st["k_rev"])                              # cost of g_t = k_rev * lmp_t
    assert st["col_idx"][0, 0] == g0 and st["col_idx"][23, 3] == -1         # s[T-1] presolved away


def test_simulation_data_files_parse_like_the_reference_reader(tmp_path):
    """write_simulation_data -> the parsing steps of Simulation_Data._read_data_to_array / read_data_to_dict (:138-220)"""
    import re
    import pandas as pd
    from dispatches_b200 import run_pricetaker as RP
    rng = np.random.default_rng(0)
    disp = rng.uniform(0, 200, (4, 48)); inp = rng.uniform(0, 1, (4, 3))
    out = RP.write_simulation_data(disp, inp, tmp_path / "sim.csv", tmp_path / "inputs.h5", ["pmax", "battery_ratio", "threshold"])
    df = pd.read_csv(out["dispatch_csv"], nrows=3)                            # the reader: nrows = num_sims
    run_index = df.iloc[:, 0].to_numpy(dtype=str)
    index = [int(re.split(r"_|\.", r)[1]) for r in run_index]
    assert index == [0, 1, 2] and np.allclose(df.iloc[:, 1:].to_numpy(dtype=float), disp[:3])
    dfi = pd.read_hdf(out["input_file"]) if out["input_format"] == "hdf" else pd.read_csv(out["input_file"])
    assert np.allclose(dfi.iloc[index, list(range(1, dfi.shape[1]))].to_numpy(), inp[:3])
