"""TEST INFRASTRUCTURE -- NOT Pyomo.  A ~300-line stand-in for the slice of Pyomo's API that the b200ipm solver plugin
(dispatches_b200/pyomo_plugin.py) and reference-style model builders touch, so that the walker can be EXECUTED in an image
where Pyomo cannot be installed (no network, not in /opt/wheelhouse).  tests/test_pyomo_plugin.py puts this directory on
sys.path only when `import pyomo` fails; with the real package present the same tests run against it.

Covered: ConcreteModel / Block (scalar + indexed, clone()) / Set / RangeSet / Var / Param(mutable) / Constraint (expr= or indexed rule=) / Objective /
Expression / Suffix, linear expression algebra with lazily evaluated Params, component_data_objects, value(),
repn.generate_standard_repn, opt.SolverFactory / SolverResults / SolverStatus / TerminationCondition.
"""
__stub__ = True
