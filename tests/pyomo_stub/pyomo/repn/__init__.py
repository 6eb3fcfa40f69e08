from ..core import NonLinear, as_expr, linear


class _Repn:
    def __init__(self, constant, linear_vars, linear_coefs, lin):
        self.constant, self.linear_vars, self.linear_coefs, self._lin = constant, linear_vars, linear_coefs, lin

    def is_linear(self): return self._lin


def generate_standard_repn(expr, compute_values=True):
    try:
        c, t = linear(as_expr(expr))
    except NonLinear:
        return _Repn(0.0, (), (), False)
    vs = tuple(t.keys())
    return _Repn(c, vs, tuple(t[v] for v in vs), True)
