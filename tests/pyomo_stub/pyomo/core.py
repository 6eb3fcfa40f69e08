"""expression algebra + components of the Pyomo stand-in (see package docstring: test infrastructure, not Pyomo)"""
import math

minimize, maximize = 1, -1


class Expr:
    def __add__(self, o): return Sum([self, as_expr(o)])
    def __radd__(self, o): return Sum([as_expr(o), self])
    def __sub__(self, o): return Sum([self, Neg(as_expr(o))])
    def __rsub__(self, o): return Sum([as_expr(o), Neg(self)])
    def __mul__(self, o): return Prod(self, as_expr(o))
    def __rmul__(self, o): return Prod(as_expr(o), self)
    def __truediv__(self, o): return Div(self, as_expr(o))
    def __rtruediv__(self, o): return Div(as_expr(o), self)
    def __neg__(self): return Neg(self)
    def __le__(self, o): return Relational(self, as_expr(o), "<=")
    def __ge__(self, o): return Relational(self, as_expr(o), ">=")
    def __eq__(self, o): return Relational(self, as_expr(o), "==")
    __hash__ = object.__hash__


class Num(Expr):
    def __init__(self, v): self.v = float(v)


class Sum(Expr):
    def __init__(self, args): self.args = args


class Prod(Expr):
    def __init__(self, a, b): self.a, self.b = a, b


class Div(Expr):
    def __init__(self, a, b): self.a, self.b = a, b


class Neg(Expr):
    def __init__(self, a): self.a = a


class Relational:
    def __init__(self, lhs, rhs, op): self.lhs, self.rhs, self.op = lhs, rhs, op


def as_expr(o):
    if isinstance(o, Expr):
        return o
    if isinstance(o, ExpressionData):
        return o.expr
    return Num(o)


def linear(e):
    """(constant, {VarData: coef}) of a linear expression at the current Param values; fixed Vars fold into the constant"""
    if isinstance(e, Num):
        return e.v, {}
    if isinstance(e, ParamData):
        return float(e.value), {}
    if isinstance(e, VarData):
        return (float(e.value), {}) if e.fixed else (0.0, {e: 1.0})
    if isinstance(e, ExpressionData):
        return linear(e.expr)
    if isinstance(e, Neg):
        c, t = linear(e.a)
        return -c, {v: -a for v, a in t.items()}
    if isinstance(e, Sum):
        c, t = 0.0, {}
        for a in e.args:
            ca, ta = linear(a)
            c += ca
            for v, k in ta.items():
                t[v] = t.get(v, 0.0) + k
        return c, t
    if isinstance(e, Prod):
        (ca, ta), (cb, tb) = linear(e.a), linear(e.b)
        if ta and tb:
            raise NonLinear()
        t = {v: k * cb for v, k in ta.items()}
        for v, k in tb.items():
            t[v] = t.get(v, 0.0) + k * ca
        return ca * cb, t
    if isinstance(e, Div):
        (ca, ta), (cb, tb) = linear(e.a), linear(e.b)
        if tb:
            raise NonLinear()
        return ca / cb, {v: k / cb for v, k in ta.items()}
    raise TypeError(type(e))


class NonLinear(Exception):
    pass


def value(e):
    if isinstance(e, (int, float)):
        return e
    if isinstance(e, Relational):
        raise TypeError("value() of a relational expression")
    c, t = linear(as_expr(e))
    return c + sum(k * v.value for v, k in t.items())


# ---------------------------------------------------------------------------------------------------------------- components
class Domain:
    def __init__(self, lb, ub): self.lb, self.ub = lb, ub


NonNegativeReals, Reals = Domain(0.0, None), Domain(None, None)


class Component:
    ctype = None

    def __init__(self):
        self.name, self.active, self._parent, self._local_name = None, True, None, None

    def _attach(self, parent, name):
        self._parent = parent
        self._local_name = name
        self._rename()

    def _rename(self):
        """hierarchical name from the parent chain; re-run for a whole subtree when a pre-built model is attached as a sub-block
        (MultiPeriodModel: blocks[t].process = <clone of the one-period flowsheet>), as Pyomo's lazily computed names would give"""
        parent, name = self._parent, self._local_name
        self.name = name if parent is None or parent.name in (None, "") else f"{parent.name}.{name}"
        for i, d in (getattr(self, "_data", None) or {}).items():
            if not isinstance(d, BlockData):
                d.name = f"{self.name}[{i}]"

    def deactivate(self): self.active = False
    def activate(self): self.active = True

    def __deepcopy__(self, memo):
        # several component classes dispatch in __new__ (scalar vs indexed): copy the instance without going through it
        import copy
        cls = type(self)
        if isinstance(self, list):
            new = list.__new__(cls); memo[id(self)] = new
            list.extend(new, [copy.deepcopy(v, memo) for v in self])
        elif isinstance(self, dict):
            new = dict.__new__(cls); memo[id(self)] = new
            dict.update(new, {copy.deepcopy(k, memo): copy.deepcopy(v, memo) for k, v in self.items()})
        else:
            new = object.__new__(cls); memo[id(self)] = new
        for k, v in self.__dict__.items():
            object.__setattr__(new, k, copy.deepcopy(v, memo))
        return new


class Set(Component, list):
    """an ordered index set: ``m.TIME = Set(initialize=range(T))``; usable wherever the stub takes an index"""
    def __init__(self, initialize=(), doc=None, ordered=True):
        Component.__init__(self)
        list.__init__(self, initialize)
    __hash__ = object.__hash__
    def first(self): return self[0]
    def last(self): return self[-1]


def RangeSet(*a):
    return Set(initialize=range(*a) if len(a) > 1 else range(1, a[0] + 1))


class IndexedMixin:
    def _make_index(self, index):
        self._index = list(index) if index is not None else None
        self._data = {}

    def __getitem__(self, i): return self._data[i]
    def __iter__(self): return iter(self._data)
    def keys(self): return self._data.keys()
    def values(self): return self._data.values()
    def items(self): return self._data.items()
    def __len__(self): return len(self._data)


class VarData(Expr):
    def __init__(self, domain, bounds, initialize):
        self.lb = domain.lb if bounds is None or bounds[0] is None else bounds[0]
        self.ub = domain.ub if bounds is None or bounds[1] is None else bounds[1]
        if bounds is not None and domain.lb is not None and (bounds[0] is None or bounds[0] < domain.lb):
            self.lb = domain.lb
        self.value = initialize
        self.fixed, self.stale, self.name, self.active = False, True, None, True

    def fix(self, v=None):
        if v is not None:
            self.value = float(v)
        self.fixed = True

    def unfix(self): self.fixed = False
    def setub(self, v): self.ub = v
    def setlb(self, v): self.lb = v

    def set_value(self, v, skip_validation=False):
        self.value = v


class Var(Component, IndexedMixin):
    def __new__(cls, *index, within=None, domain=None, bounds=None, initialize=None, doc=None, units=None):
        if not index:
            v = ScalarVar(within or domain or Reals, bounds, initialize)
            return v
        return super().__new__(cls)

    def __init__(self, *index, within=None, domain=None, bounds=None, initialize=None, doc=None, units=None):
        Component.__init__(self)
        self._make_index(index[0])
        for i in self._index:
            self._data[i] = VarData(within or domain or Reals, bounds, initialize)

    def _attach(self, parent, name):
        Component._attach(self, parent, name)
        for i, d in self._data.items():
            d.name = f"{self.name}[{i}]"


class ScalarVar(VarData, Component):
    ctype = Var

    def __init__(self, domain, bounds, initialize):
        Component.__init__(self)
        VarData.__init__(self, domain, bounds, initialize)


Var.ctype = Var


class ParamData(Expr):
    def __init__(self, v): self.value, self.name = v, None
    def set_value(self, v): self.value = v


class Param(Component, IndexedMixin):
    def __new__(cls, *index, default=None, initialize=None, mutable=False, doc=None, units=None, within=None):
        if not index:
            return ScalarParam(initialize if initialize is not None else default)
        return super().__new__(cls)

    def __init__(self, *index, default=None, initialize=None, mutable=False, doc=None, units=None, within=None):
        Component.__init__(self)
        self._make_index(index[0])
        for i in self._index:
            v = initialize[i] if isinstance(initialize, dict) else (initialize if initialize is not None else default)
            self._data[i] = ParamData(v)

    def _attach(self, parent, name):
        Component._attach(self, parent, name)
        for i, d in self._data.items():
            d.name = f"{self.name}[{i}]"


class ScalarParam(ParamData, Component):
    ctype = Param

    def __init__(self, v):
        Component.__init__(self)
        ParamData.__init__(self, v)


Param.ctype = Param


class ExpressionData:
    def __init__(self, expr): self.expr = as_expr(expr)
    def __add__(self, o): return self.expr + o
    def __radd__(self, o): return o + self.expr
    def __sub__(self, o): return self.expr - o
    def __rsub__(self, o): return o - self.expr
    def __mul__(self, o): return self.expr * o
    def __rmul__(self, o): return o * self.expr
    def __truediv__(self, o): return self.expr / o
    def __neg__(self): return -self.expr


class Expression(Component, ExpressionData):
    def __init__(self, expr=None, doc=None):
        Component.__init__(self)
        ExpressionData.__init__(self, expr)


Expression.ctype = Expression


class ConstraintData:
    def __init__(self, rel):
        if isinstance(rel, tuple):                       # (lo, body, hi)
            lo, body, hi = rel
            self.body, self.lower, self.upper = as_expr(body), (None if lo is None else as_expr(lo)), (None if hi is None else as_expr(hi))
        else:
            self.body = rel.lhs - rel.rhs
            self.lower = Num(0.0) if rel.op in (">=", "==") else None
            self.upper = Num(0.0) if rel.op in ("<=", "==") else None
        self.active, self.name = True, None

    @property
    def equality(self): return self.lower is not None and self.upper is not None and value(self.lower) == value(self.upper)
    def deactivate(self): self.active = False


class Constraint(Component, IndexedMixin):
    Skip = object()

    def __new__(cls, *index, expr=None, rule=None, doc=None):
        if not index:
            return ScalarConstraint(expr)
        return super().__new__(cls)

    def __init__(self, *index, expr=None, rule=None, doc=None):
        Component.__init__(self)
        self._make_index(index[0])
        self._rule = rule

    def _attach(self, parent, name):
        self._parent, self._local_name = parent, name
        for i in self._index:
            rel = self._rule(parent, i)
            if rel is Constraint.Skip:
                continue
            self._data[i] = ConstraintData(rel)
        self._rename()


class ScalarConstraint(ConstraintData, Component):
    ctype = Constraint

    def __init__(self, expr):
        Component.__init__(self)
        ConstraintData.__init__(self, expr)


Constraint.ctype = Constraint


def inequality(lo, body, hi):
    return (lo, body, hi)


class Objective(Component):
    def __init__(self, expr=None, sense=minimize, doc=None):
        Component.__init__(self)
        self.expr, self.sense = as_expr(expr), sense


Objective.ctype = Objective


class Suffix(Component, dict):
    IMPORT, EXPORT, LOCAL = 1, 2, 0

    def __init__(self, direction=0):
        Component.__init__(self)
        dict.__init__(self)
        self.direction = direction

    def import_enabled(self): return self.direction == Suffix.IMPORT
    __hash__ = object.__hash__


Suffix.ctype = Suffix


class BlockData(Component):
    def __init__(self):
        Component.__init__(self)
        object.__setattr__(self, "_components", [])

    def __setattr__(self, k, v):
        if isinstance(v, Component) and k not in ("_parent",):
            v._attach(self, k)
            self._components.append(v)
        object.__setattr__(self, k, v)

    def _rename(self):
        Component._rename(self)
        for comp in self._components:
            comp._rename()

    def Constraint(self, *index, doc=None):              # decorator form  @m.Constraint(m.set)
        def deco(f):
            c = Constraint(*index, rule=f)
            setattr(self, f.__name__, c)
            return c
        return deco

    def clone(self):
        """deep copy of the block tree (what wind_battery_mp_block does with its cached one-period model, wind_battery_LMP.py:159-161)"""
        import copy
        c = copy.deepcopy(self)
        object.__setattr__(c, "_parent", None)
        return c

    def component_data_objects(self, ctype, active=None, descend_into=True):
        for comp in self._components:
            if isinstance(comp, BlockData):
                if descend_into and (active is None or comp.active == active):
                    yield from comp.component_data_objects(ctype, active, descend_into)
                continue
            if isinstance(comp, Block):
                if descend_into:
                    for b in comp.values():
                        if active is None or b.active == active:
                            yield from b.component_data_objects(ctype, active, descend_into)
                continue
            if getattr(comp, "ctype", None) is not ctype:
                continue
            datas = comp.values() if isinstance(comp, IndexedMixin) and comp._index is not None else [comp]
            for d in datas:
                if active is None or getattr(d, "active", True) == active:
                    yield d


class Block(Component, IndexedMixin):
    def __new__(cls, *index, **kw):
        if not index:
            return BlockData()
        return super().__new__(cls)

    def __init__(self, *index, **kw):
        Component.__init__(self)
        self._make_index(index[0])
        for i in self._index:
            self._data[i] = BlockData()

    def _rename(self):
        Component._rename(self)
        for i, b in self._data.items():
            b._parent = self._parent
            b.name = f"{self.name}[{i}]"
            for comp in b._components:
                comp._rename()


class ConcreteModel(BlockData):
    def __init__(self, name="model"):
        BlockData.__init__(self)
        self.name = ""
        object.__setattr__(self, "_model_name", name)
