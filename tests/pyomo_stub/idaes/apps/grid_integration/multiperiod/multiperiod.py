"""TEST INFRASTRUCTURE -- NOT idaes-pse.  A stand-in for ``idaes.apps.grid_integration.multiperiod.multiperiod.MultiPeriodModel``
(idaes-pse 2.0, not installable here) restated from its published behaviour and from the way the reference uses it
(wind_battery_LMP.py:195-205, wind_battery_double_loop.py:69-81, price_taker_analysis.py:190-196; API surface listed in SURVEY.md 8a):

    mp = MultiPeriodModel(n_time_points, process_model_func, linking_variable_func, periodic_variable_func=None)
    mp.build_multi_period_model(model_data_kwargs)      # {t: kwargs of process_model_func}
    mp.pyomo_model                                      # ConcreteModel with .TIME and .blocks[t].process
    mp.get_active_process_blocks()                      # [blocks[t].process for t in TIME]

One block per time point holds the flowsheet ``process_model_func(**model_data_kwargs[t])``; ``linking_variable_func(b_t, b_t+1)``
returns pairs of Vars that are equated by ``blocks[t].link_constraints[i]``; ``periodic_variable_func(b_last, b_first)`` likewise by
``blocks[last].periodic_constraints[i]``.  Put on sys.path by tests only when ``import idaes`` fails.
"""
import pyomo.environ as pyo


class MultiPeriodModel:
    def __init__(self, n_time_points, process_model_func, linking_variable_func, periodic_variable_func=None,
                 use_stochastic_build=False, **unused):
        if use_stochastic_build:
            raise NotImplementedError("stub: only the legacy (deterministic) build the price-taker scripts use")
        self.n_time_points = n_time_points
        self.create_process_model = process_model_func
        self.get_linking_variable_pairs = linking_variable_func
        self.get_periodic_variable_pairs = periodic_variable_func
        self._pyomo_model = None
        self._first_active_time = None

    def build_multi_period_model(self, model_data_kwargs=None):
        m = pyo.ConcreteModel()
        m.TIME = pyo.Set(initialize=range(self.n_time_points))
        if model_data_kwargs is None:
            model_data_kwargs = {t: {} for t in m.TIME}
        m.blocks = pyo.Block(m.TIME)
        for t in m.TIME:
            m.blocks[t].process = self.create_process_model(**model_data_kwargs[t])
        for t in m.TIME:
            if t == m.TIME.last():
                continue
            pairs = self.get_linking_variable_pairs(m.blocks[t].process, m.blocks[t + 1].process)
            m.blocks[t].link_constraints = pyo.Constraint(range(len(pairs)), rule=lambda b, i, pairs=pairs: pairs[i][0] == pairs[i][1])
        if self.get_periodic_variable_pairs is not None:
            last, first = m.TIME.last(), m.TIME.first()
            pairs = self.get_periodic_variable_pairs(m.blocks[last].process, m.blocks[first].process)
            m.blocks[last].periodic_constraints = pyo.Constraint(range(len(pairs)), rule=lambda b, i, pairs=pairs: pairs[i][0] == pairs[i][1])
        self._pyomo_model = m
        self._first_active_time = m.TIME.first()
        return m

    @property
    def pyomo_model(self):
        return self._pyomo_model

    @property
    def current_time(self):
        return self._first_active_time

    def get_active_process_blocks(self):
        m = self._pyomo_model
        return [m.blocks[t].process for t in m.TIME if m.blocks[t].active]
