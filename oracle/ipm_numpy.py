"""CPU ORACLE (test infrastructure -- NOT product code).

Batched numpy mirror of the interior-point algorithm the CUDA kernel implements
(dispatches_b200/csrc/dsp_lp_kernels.cu): Mehrotra predictor-corrector on

    min c'x   s.t.  A x = b,  0 <= x <= u

via the normal equations  (A D A') dy = r.  Same scaling, start point, step rule and stopping rule as
the kernel, but dense LAPACK factorisations instead of the on-chip band LDL'.  It is the numerics
test-bed (does the ALGORITHM reach 1e-6 on the hard price signals?) and a second, independent check
of the kernel's iteration counts; the accuracy oracle proper is HiGHS (oracle/highs.py).
"""
from __future__ import annotations

import numpy as np

OPTIMAL, MAXITER, NUMERR = 0, 1, 2


def solve_batch(A, b, c, u, tol=1e-9, max_iter=60, eta=0.9995, verbose=False, start="simple", gap_floor=1e-4, feas_tol=None, rho=1e-8):
    """A: dense [m,n] shared;  b [N,m], c [N,n], u [N,n] (inf = none).  Returns dict(obj,x,y,status,iters)."""
    A = np.asarray(A, float)
    feas_tol = 1e-9 if feas_tol is None else feas_tol
    m, n = A.shape
    b = np.atleast_2d(b).astype(float); c = np.atleast_2d(c).astype(float); u = np.atleast_2d(u).astype(float)
    N = b.shape[0]
    bd = np.isfinite(u[0])                       # bounded columns (same pattern for the whole batch)
    nbnd = int(bd.sum())
    ub = np.where(bd, u, 1.0)
    # ---- per-problem scaling
    beta_b = np.maximum(1e-300, np.maximum(np.abs(b).max(1), np.where(bd, u, 0.0).max(1)))
    beta_b = np.where(beta_b > 0, beta_b, 1.0)
    beta_c = np.abs(c).max(1); beta_c = np.where(beta_c > 0, beta_c, 1.0)
    b = b / beta_b[:, None]; ub = ub / beta_b[:, None]; c = c / beta_c[:, None]
    nb_ = 1.0 + np.abs(b).max(1); nc_ = 1.0 + np.abs(c).max(1)
    # ---- start point
    x = np.ones((N, n)); x[:, bd] = np.minimum(1.0, 0.5 * ub[:, bd])
    s = np.where(bd, ub - x, 1.0)
    z = np.ones((N, n)); w = np.tile(np.where(bd, 1.0, 0.0), (N, 1))
    y = np.zeros((N, m))
    if start == "mehrotra":
        # Mehrotra's least-squares start; (A A')^-1 is problem independent (precomputed per template)
        M0 = A @ A.T
        x = np.linalg.solve(M0, b.T).T @ A      # least-norm A x = b
        y = np.linalg.solve(M0, (c @ A.T).T).T
        z = c - y @ A
        x[:, bd] = np.minimum(x[:, bd], 0.5 * ub[:, bd])
        dx = np.maximum(-1.5 * x.min(1), 0.0); dz = np.maximum(-1.5 * z.min(1), 0.0)
        x = x + dx[:, None] + 1e-3; z = z + dz[:, None] + 1e-3
        xz = (x * z).sum(1)
        x = x + (0.5 * xz / z.sum(1))[:, None]; z = z + (0.5 * xz / x.sum(1))[:, None]
        x[:, bd] = np.minimum(x[:, bd], 0.9 * ub[:, bd])
        s = np.where(bd, ub - x, 1.0)
        w = np.where(bd, 0.5 * z, 0.0); z = np.where(bd, 1.5 * z, z)
    status = np.full(N, MAXITER); iters = np.zeros(N, int)
    active = np.ones(N, bool)
    ntot = n + nbnd
    for it in range(max_iter):
        rp = b - x @ A.T
        ru = np.where(bd, ub - x - s, 0.0)
        rd = c - y @ A - z + w
        mu = ((x * z).sum(1) + (s * w).sum(1)) / ntot
        pobj = (c * x).sum(1); dobj = (b * y).sum(1) - (np.where(bd, ub, 0.0) * w).sum(1)
        pres = np.maximum(np.abs(rp).max(1), np.abs(ru).max(1)) / nb_
        dres = np.abs(rd).max(1) / nc_
        gap = np.abs(pobj - dobj) / np.maximum(gap_floor, np.abs(pobj))
        res = np.maximum(pres, dres)
        den = np.maximum(gap_floor, np.abs(pobj))
        cgap = ntot * mu / den                      # complementarity gap (what further iterations can still reduce)
        done = (res < feas_tol) & (gap < tol)
        # complementarity has converged but residuals / objective gap sit at the rounding floor of the
        # ill-conditioned normal equations: iterating further only loses accuracy -> accept what is there
        done |= (cgap < tol) & (res < 10.0 * feas_tol) & (gap < 10.0 * tol)
        giveup = (cgap < 1e-3 * tol) & ~done
        done |= giveup & (res < 100.0 * feas_tol) & (gap < 1000.0 * tol)
        failed = active & giveup & ~done
        status[failed] = NUMERR; iters[failed] = it
        active &= ~failed
        newly = active & done
        status[newly] = OPTIMAL; iters[newly] = it
        active &= ~done
        if verbose:
            print(it, "act", active.sum(), "pres %.2e dres %.2e gap %.2e mu %.2e" % (pres[active].max() if active.any() else 0,
                  dres[active].max() if active.any() else 0, gap[active].max() if active.any() else 0, mu.max()))
        if not active.any():
            break
        ix = np.flatnonzero(active)
        xa, sa, za, wa, ya = x[ix], s[ix], z[ix], w[ix], y[ix]
        d = 1.0 / (za / xa + np.where(bd, wa / sa, 0.0) + rho / np.maximum(1.0, xa * xa))
        M = np.einsum("ij,nj,kj->nik", A, d, A, optimize=True)
        M[:, np.arange(m), np.arange(m)] *= (1.0 + 1e-14)
        try:
            Lc = np.linalg.cholesky(M)
        except np.linalg.LinAlgError:
            M[:, np.arange(m), np.arange(m)] += 1e-12 * np.abs(M[:, np.arange(m), np.arange(m)]).max(1)[:, None]
            Lc = np.linalg.cholesky(M)

        def newton(rxz, rsw):
            h = rd[ix] - rxz / xa + np.where(bd, (rsw - wa * ru[ix]) / sa, 0.0)
            rhs = rp[ix] + (d * h) @ A.T
            t = np.linalg.solve(Lc, rhs[:, :, None])
            dy = np.linalg.solve(np.swapaxes(Lc, 1, 2), t)[:, :, 0]
            dx = d * (dy @ A - h)
            ds = np.where(bd, ru[ix] - dx, 0.0)
            dz = (rxz - za * dx) / xa
            dw = np.where(bd, (rsw - wa * ds) / sa, 0.0)
            return dx, ds, dy, dz, dw

        def maxstep(v, dv, mask=None):
            r = np.where(dv < 0, -v / np.where(dv < 0, dv, -1.0), np.inf)
            if mask is not None:
                r = np.where(mask, r, np.inf)
            return r.min(1)

        dx, ds, dy, dz, dw = newton(-xa * za, -sa * wa)
        ap = np.minimum(1.0, np.minimum(maxstep(xa, dx), maxstep(sa, ds, bd)))
        ad = np.minimum(1.0, np.minimum(maxstep(za, dz), maxstep(wa, dw, bd)))
        mu_a = (((xa + ap[:, None] * dx) * (za + ad[:, None] * dz)).sum(1)
                + ((sa + ap[:, None] * ds) * (wa + ad[:, None] * dw)).sum(1)) / ntot
        sigma = (mu_a / mu[ix]) ** 3
        sm = (sigma * mu[ix])[:, None]
        dx, ds, dy, dz, dw = newton(sm - xa * za - dx * dz, np.where(bd, sm - sa * wa - ds * dw, 0.0))
        ap = np.minimum(1.0, eta * np.minimum(maxstep(xa, dx), maxstep(sa, ds, bd)))
        ad = np.minimum(1.0, eta * np.minimum(maxstep(za, dz), maxstep(wa, dw, bd)))
        x[ix] = xa + ap[:, None] * dx; s[ix] = np.where(bd, sa + ap[:, None] * ds, 1.0)
        y[ix] = ya + ad[:, None] * dy; z[ix] = za + ad[:, None] * dz; w[ix] = np.where(bd, wa + ad[:, None] * dw, 0.0)
        bad = ~np.isfinite(x[ix]).all(1)
        if bad.any():
            status[ix[bad]] = NUMERR; active[ix[bad]] = False
            x[ix[bad]] = 1.0; z[ix[bad]] = 1.0; s[ix[bad]] = 1.0
    iters[active] = max_iter
    scale = beta_b * beta_c
    return dict(obj=(c * x).sum(1) * scale, x=x * beta_b[:, None], y=y * beta_c[:, None], z=z * beta_c[:, None], w=w * beta_c[:, None],
                status=status, iters=iters)
