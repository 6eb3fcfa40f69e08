import enum
import types


class SolverStatus(enum.Enum):
    ok = "ok"; warning = "warning"; error = "error"; aborted = "aborted"


class TerminationCondition(enum.Enum):
    optimal = "optimal"; infeasible = "infeasible"; unbounded = "unbounded"; maxIterations = "maxIterations"; error = "error"


class SolverResults:
    def __init__(self):
        self.solver = types.SimpleNamespace(status=None, termination_condition=None)
        self.problem = types.SimpleNamespace(lower_bound=None, upper_bound=None)


class _Factory:
    def __init__(self): self._reg = {}

    def register(self, name, doc=None):
        def deco(cls):
            self._reg[name] = cls
            return cls
        return deco

    def __call__(self, name, **kw):
        return self._reg[name](**kw)


SolverFactory = _Factory()
