"""Shared standard-form LP template for a batch of structurally identical dispatch LPs.

One template = one (case-study flowsheet, horizon T, fixed/free design mode).  Every LP of a batch is

    min  c'x + obj_const      s.t.  A x = b,   0 <= x <= u            (u_j = +inf allowed)

with the constraint matrix ``A`` (inequality rows already carry a slack column) SHARED by the whole
batch, and the per-problem data affine in two small parameter vectors:

    c = c0 + Cmap @ cparams        (cparams: the LMP signal, what changes per price scenario)
    b = b0 + Bmap @ rparams        (rparams: capacity-factor / design scalars, what changes per design)
    u = u0 + Umap @ rparams
    obj_const = o0 + omap @ rparams

This is the data model behind the C-ABI (include/dsp_lp.h: dsp_lp_template_create) and replaces, for the
hot path, what Pyomo's LP writer + CBC's reader rebuild per LP in the reference
(wind_battery_LMP.py:195-267: model build + SolverFactory("cbc").solve per scenario).

``finalize()`` does the once-per-template symbolic work the GPU kernel relies on: a bandwidth-reducing
row order for the normal matrix  M = A D A'  (block-tridiagonal in time for every multi-period flowsheet,
SURVEY.md §0.5), the band half-width ``w`` and the assembly list  M[i, i-k] = sum_p coef_p * d[col_p].
"""
from __future__ import annotations

import dataclasses
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee

INF = float("inf")


@dataclasses.dataclass
class LPTemplate:
    name: str
    A: sp.csr_matrix                 # m x n (slacks included)
    b0: np.ndarray
    Bmap: sp.csr_matrix              # m x Pr
    c0: np.ndarray
    Cmap: sp.csr_matrix              # n x Pc
    u0: np.ndarray                   # +inf where unbounded above
    Umap: sp.csr_matrix              # n x Pr (rows of unbounded columns are empty)
    o0: float
    omap: np.ndarray                 # Pr
    ocmap: np.ndarray                # Pc   (objective constant = o0 + omap@rparams + ocmap@cparams)
    col_shift: np.ndarray            # x_model = col_shift + col_scale * x_template   (lower-bound shift, equilibration)
    col_scale: np.ndarray
    col_names: list
    row_names: list
    meta: dict = dataclasses.field(default_factory=dict)
    # filled by finalize()
    row_perm: np.ndarray | None = None
    w: int = 0
    asm_ptr: np.ndarray | None = None
    asm_col: np.ndarray | None = None
    asm_val: np.ndarray | None = None

    @property
    def m(self):
        return self.A.shape[0]

    @property
    def n(self):
        return self.A.shape[1]

    @property
    def Pc(self):
        return self.Cmap.shape[1]

    @property
    def Pr(self):
        return self.Bmap.shape[1]

    @property
    def nb(self):
        """number of upper-bounded columns (ordered first)."""
        return int(np.isfinite(self.u0).sum())

    # ------------------------------------------------------------------
    def instantiate(self, cparams, rparams):
        """(c, b, u, obj_const) of one problem -- host-side helper for tests / plumbing."""
        cparams = np.asarray(cparams, float); rparams = np.asarray(rparams, float)
        c = self.c0 + self.Cmap @ cparams
        b = self.b0 + self.Bmap @ rparams
        u = self.u0.copy()
        fin = np.isfinite(u)
        u[fin] = u[fin] + (self.Umap @ rparams)[fin]
        return c, b, u, self.o0 + float(self.omap @ rparams) + float(self.ocmap @ cparams)

    # ------------------------------------------------------------------
    def finalize(self, equilibrate=False):
        """Column order (bounded first), row order (min bandwidth of A A'), band assembly list."""
        A = self.A.tocsr()
        m, n = A.shape
        # 1. bounded columns first (keeps the upper-bound vectors short in the kernel)
        fin = np.isfinite(self.u0)
        cperm = np.concatenate([np.flatnonzero(fin), np.flatnonzero(~fin)])
        A = A[:, cperm]
        self.c0 = self.c0[cperm]; self.Cmap = self.Cmap.tocsr()[cperm]
        self.u0 = self.u0[cperm]; self.Umap = self.Umap.tocsr()[cperm]
        self.col_shift = self.col_shift[cperm]; self.col_scale = self.col_scale[cperm]
        self.col_names = [self.col_names[j] for j in cperm]
        # 1b. geometric-mean equilibration  A <- R A C  (x = C x~, b~ = R b, c~ = C c, u~ = u / C)
        if equilibrate:
            A = A.tocsr().astype(float)
            R = np.ones(m); C = np.ones(n)
            for _ in range(6):
                B = sp.diags(R) @ A @ sp.diags(C)
                B = B.tocsr(); absB = abs(B)
                rmax = absB.max(axis=1).toarray().ravel(); rmin = _rowmin(absB)
                R /= np.sqrt(np.maximum(rmax * rmin, 1e-300))
                B = (sp.diags(R) @ A @ sp.diags(C)).tocsc(); absB = abs(B)
                cmax = absB.max(axis=0).toarray().ravel(); cmin = _rowmin(absB.T.tocsr())
                C /= np.sqrt(np.maximum(cmax * cmin, 1e-300))
            # round to powers of two: scaling is then exact in floating point
            R = 2.0 ** np.round(np.log2(R)); C = 2.0 ** np.round(np.log2(C))
            A = (sp.diags(R) @ A @ sp.diags(C)).tocsr()
            self.b0 = self.b0 * R; self.Bmap = (sp.diags(R) @ self.Bmap).tocsr()
            self.c0 = self.c0 * C; self.Cmap = (sp.diags(C) @ self.Cmap).tocsr()
            self.u0 = self.u0 / C; self.Umap = (sp.diags(1.0 / C) @ self.Umap).tocsr()
            self.col_scale = self.col_scale * C
            self.meta["row_scale"] = R
        # 2. row order: natural vs reverse Cuthill-McKee on the pattern of A A'
        P = (abs(A) @ abs(A).T).tocsr()
        P.data[:] = 1.0

        def bandwidth(perm):
            inv = np.empty(m, int); inv[perm] = np.arange(m)
            coo = P.tocoo()
            return int(np.max(np.abs(inv[coo.row] - inv[coo.col]))) if coo.nnz else 0

        nat = np.arange(m)
        rcm = np.asarray(reverse_cuthill_mckee(P, symmetric_mode=True))
        perm = nat if bandwidth(nat) <= bandwidth(rcm) else rcm
        self.w = bandwidth(perm)
        self.row_perm = perm
        A = A[perm]
        self.b0 = self.b0[perm]; self.Bmap = self.Bmap.tocsr()[perm]
        self.row_names = [self.row_names[i] for i in perm]
        self.A = A.tocsr(); self.A.sort_indices()
        # 3. assembly list of the lower band of M = A D A'
        w = self.w
        Acsc = self.A.tocsc()
        ent = [[] for _ in range(m * (w + 1))]
        for j in range(n):
            lo, hi = Acsc.indptr[j], Acsc.indptr[j + 1]
            rows, vals = Acsc.indices[lo:hi], Acsc.data[lo:hi]
            for a in range(len(rows)):
                for bq in range(len(rows)):
                    i, i2 = rows[a], rows[bq]
                    if i2 <= i:
                        ent[i * (w + 1) + (i - i2)].append((j, vals[a] * vals[bq]))
        ptr = np.zeros(m * (w + 1) + 1, np.int32)
        cols, vals = [], []
        for e, lst in enumerate(ent):
            ptr[e + 1] = ptr[e] + len(lst)
            for j, v in lst:
                cols.append(j); vals.append(v)
        self.asm_ptr, self.asm_col, self.asm_val = ptr, np.array(cols, np.int32), np.array(vals, float)
        return self

    # ------------------------------------------------------------------
    def column(self, name):
        return self.col_names.index(name)


def _rowmin(absB):
    """min over the stored nonzeros of each row of a CSR matrix (rows without entries -> 1)."""
    out = np.ones(absB.shape[0])
    for i in range(absB.shape[0]):
        d = absB.data[absB.indptr[i]:absB.indptr[i + 1]]
        d = d[d > 0]
        if d.size:
            out[i] = d.min()
    return out


class TemplateBuilder:
    """Tiny algebra for writing a template: columns, (in)equality rows, parameter-affine data.

    A "parameter-affine" quantity is  const + sum_k coef_k * param[k]  given as ``(const, {k: coef})``
    or a plain float.
    """

    def __init__(self, name, Pc, Pr):
        self.name, self.Pc, self.Pr = name, Pc, Pr
        self.cols, self.rows = [], []          # names
        self.u = []                            # (const, {k: coef}) or None
        self.lbs, self.fixed = [], []          # constant lower bound ; True if the column is a constant
        self.ocmap = np.zeros(Pc)
        self.c = []                            # (const, {k: coef})
        self.arows, self.rhs = [], []          # dict col->val ; (const, {k:coef})
        self.o0, self.omap = 0.0, np.zeros(Pr)
        self.meta = {}

    @staticmethod
    def _aff(q):
        if q is None:
            return None
        if isinstance(q, tuple):
            return float(q[0]), dict(q[1])
        return float(q), {}

    def var(self, name, ub=None, lb=0.0, fix=None):
        """column with lb <= x <= ub (lb constant, ub parameter-affine) or a constant (fix=value)."""
        if fix is not None:
            lb, ub = float(fix), None
        self.cols.append(name); self.u.append(self._aff(ub)); self.c.append((0.0, {}))
        self.lbs.append(float(lb)); self.fixed.append(fix is not None)
        return len(self.cols) - 1

    def cost(self, j, q):
        c0, cm = self.c[j]; a0, am = self._aff(q)
        for k, v in am.items():
            cm[k] = cm.get(k, 0.0) + v
        self.c[j] = (c0 + a0, cm)

    def eq(self, name, coeffs, rhs=0.0):
        self.rows.append(name); self.arows.append(dict(coeffs)); self.rhs.append(self._aff(rhs))

    def le(self, name, coeffs, rhs=0.0):
        s = self.var("slack:" + name)
        coeffs = dict(coeffs); coeffs[s] = 1.0
        self.eq(name, coeffs, rhs)

    def obj_const(self, q):
        a0, am = self._aff(q)
        self.o0 += a0
        for k, v in am.items():
            self.omap[k] += v

    def build(self, equilibrate=False) -> LPTemplate:
        # substitute x = lb + x' (and drop constant columns)
        lbs = np.array(self.lbs)
        for r, row in enumerate(self.arows):
            shift = sum(v * lbs[j] for j, v in row.items())
            if shift != 0.0:
                a0, am = self.rhs[r]; self.rhs[r] = (a0 - shift, am)
        for j in range(len(self.cols)):
            if lbs[j] != 0.0:
                c0, cm = self.c[j]
                self.o0 += c0 * lbs[j]
                for k, v in cm.items():
                    self.ocmap[k] += v * lbs[j]
                if self.u[j] is not None:
                    u0, um = self.u[j]; self.u[j] = (u0 - lbs[j], um)
        keep = [j for j in range(len(self.cols)) if not self.fixed[j]]
        remap = {j: k for k, j in enumerate(keep)}
        self.arows = [{remap[j]: v for j, v in row.items() if j in remap} for row in self.arows]
        self.cols = [self.cols[j] for j in keep]; self.u = [self.u[j] for j in keep]
        self.c = [self.c[j] for j in keep]; shifts = lbs[keep]
        n, m = len(self.cols), len(self.rows)

        def affine_rows(items, P, default=0.0):
            base = np.full(len(items), default)
            ri, ci, vv = [], [], []
            for r, it in enumerate(items):
                if it is None:
                    continue
                base[r] = it[0]
                for k, v in it[1].items():
                    ri.append(r); ci.append(k); vv.append(v)
            return base, sp.csr_matrix((vv, (ri, ci)), shape=(len(items), P))

        ri, ci, vv = [], [], []
        for r, row in enumerate(self.arows):
            for j, v in row.items():
                if v != 0.0:
                    ri.append(r); ci.append(j); vv.append(v)
        A = sp.csr_matrix((vv, (ri, ci)), shape=(m, n))
        b0, Bmap = affine_rows(self.rhs, self.Pr)
        c0, Cmap = affine_rows(self.c, self.Pc)
        u0, Umap = affine_rows(self.u, self.Pr, default=INF)
        t = LPTemplate(self.name, A, b0, Bmap, c0, Cmap, u0, Umap, self.o0, self.omap.copy(), self.ocmap.copy(),
                       shifts.copy(), np.ones(n), list(self.cols), list(self.rows), dict(self.meta))
        return t.finalize(equilibrate=equilibrate)
