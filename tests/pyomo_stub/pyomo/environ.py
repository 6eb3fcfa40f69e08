from .core import (Block, ConcreteModel, Constraint, Expression, NonNegativeReals, Objective, Param, RangeSet, Reals, Set, Suffix, Var, inequality,  # noqa: F401
                   maximize, minimize, value)
from .opt import SolverFactory, SolverStatus, TerminationCondition  # noqa: F401
