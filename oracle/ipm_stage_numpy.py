"""CPU ORACLE (test infrastructure -- NOT product code).

numpy mirror of the STAGE kernel (dispatches_b200/csrc/dsp_stage_wb.cuh): the same Mehrotra predictor-corrector
as oracle/ipm_numpy.py, but with the linear algebra the stage kernel uses for the wind+battery flowsheet:

  * one "lane" per period t, all vectors shaped [N, T];
  * rows per period: r1 state-of-charge evolution, r2 throughput accumulation (both couple to t-1),
    r3 SoC bound (+slack p), r4 wind balance (+slack q);  columns g,i,o,s,e,p,q  (s[T-1] == 0 is absent);
  * the normal matrix M = A D A' is reduced per period by eliminating r4 and r3 (scalar pivots), which leaves a
    block-tridiagonal system with 2x2 blocks in (dy1, dy2);
  * that system is solved by parallel cyclic reduction across the period axis (log2(32) = 5 strides).

Used by tests to check (a) the algebra against the generic mirror / HiGHS and (b) the CUDA stage kernel's
iteration counts.  Constants (a, binv, delta, dur) come from the template, not from this file.
"""
from __future__ import annotations

import numpy as np

OPTIMAL, MAXITER, NUMERR = 0, 1, 2
INV = "ldl"         # "adj" (adjugate / determinant) | "defer" (the kernel's deferred-reciprocal elimination step)
MODE = "twisted"     # "pcr" | "refine" (one step of iterative refinement) | "dense" (LAPACK, for diagnosis)


def _shift_down(v, k):
    """value of lane t-k at lane t (zeros shifted in)."""
    out = np.zeros_like(v)
    if k < v.shape[1]:
        out[:, k:] = v[:, :-k]
    return out


def _shift_up(v, k):
    out = np.zeros_like(v)
    if k < v.shape[1]:
        out[:, :-k] = v[:, k:]
    return out


def _mm(A, B):   # [N,T,2,2] @ [N,T,2,2]
    return np.einsum("ntij,ntjk->ntik", A, B)


def _mv(A, v):
    return np.einsum("ntij,ntj->nti", A, v)


def _inv2(D):
    """inverse of (nearly) symmetric positive definite 2x2 blocks through their LDL' factors -- the adjugate formula
    loses the small pivot to cancellation in det = D00*D11 - D01*D10 when the block is ill conditioned."""
    if INV == "adj":
        det = D[..., 0, 0] * D[..., 1, 1] - D[..., 0, 1] * D[..., 1, 0]
        r = 1.0 / det
        out = np.empty_like(D)
        out[..., 0, 0] = D[..., 1, 1] * r; out[..., 1, 1] = D[..., 0, 0] * r
        out[..., 0, 1] = -D[..., 0, 1] * r; out[..., 1, 0] = -D[..., 1, 0] * r
        return out
    i1 = 1.0 / D[..., 0, 0]
    l10 = D[..., 1, 0] * i1; u01 = D[..., 0, 1] * i1
    i2 = 1.0 / (D[..., 1, 1] - l10 * D[..., 0, 1])
    out = np.empty_like(D)
    out[..., 1, 1] = i2
    out[..., 0, 1] = -u01 * i2
    out[..., 1, 0] = -l10 * i2
    out[..., 0, 0] = i1 + u01 * l10 * i2
    return out


class PCR:
    """Block-tridiagonal (2x2 blocks) solve by parallel cyclic reduction; factor once, solve many."""

    def __init__(self, L, D, U, lanes=32):
        N, T = D.shape[:2]
        self.orig = (L.copy(), D.copy(), U.copy())

        def padm(M, ident=False):
            P = np.zeros((N, lanes, 2, 2))
            P[:, :T] = M
            if ident:
                P[:, T:, 0, 0] = 1.0; P[:, T:, 1, 1] = 1.0
            return P
        L, D, U = padm(L), padm(D, True), padm(U)
        self.T, self.lanes, self.steps = T, lanes, []
        k = 1
        while k < lanes:
            Di = _inv2(D)
            al = -_mm(L, _shift_down(Di, k))
            ga = -_mm(U, _shift_up(Di, k))
            D = D + _mm(al, _shift_down(U, k)) + _mm(ga, _shift_up(L, k))
            L, U = _mm(al, _shift_down(L, k)), _mm(ga, _shift_up(U, k))
            self.steps.append((k, al, ga))
            k *= 2
        self.Dinv = _inv2(D)

    def solve(self, f):
        if MODE == "dense":
            return self.dense_solve(f)
        if MODE == "thomas":
            return self.thomas(f)
        if MODE == "twisted":
            return self.twisted(f)
        u = self._solve(f)
        if MODE == "refine":
            r = f - self.apply(u)
            u = u + self._solve(r)
        return u

    def apply(self, u):
        L, D, U = self.orig
        return _mv(D, u) + _mv(L, _shift_down(u, 1)) + _mv(U, _shift_up(u, 1))

    def thomas(self, f):
        """sequential block LDL' (what a Cholesky of the reduced system does)"""
        L, D, U = self.orig
        N, T = f.shape[:2]
        if not hasattr(self, "_th"):
            Dh = D.copy(); G = np.zeros_like(D)          # G_t = L_t Dh_{t-1}^-1
            for t in range(1, T):
                G[:, t] = np.einsum("nij,njk->nik", L[:, t], _inv2(Dh[:, t - 1]))
                Dh[:, t] = D[:, t] - np.einsum("nij,njk->nik", G[:, t], U[:, t - 1])
            self._th = (G, np.stack([_inv2(Dh[:, t]) for t in range(T)], axis=1))
        G, Dhi = self._th
        g = f.copy()
        for t in range(1, T):
            g[:, t] = g[:, t] - np.einsum("nij,nj->ni", G[:, t], g[:, t - 1])
        u = np.zeros_like(g)
        u[:, T - 1] = np.einsum("nij,nj->ni", Dhi[:, T - 1], g[:, T - 1])
        for t in range(T - 2, -1, -1):
            u[:, t] = np.einsum("nij,nj->ni", Dhi[:, t], g[:, t] - np.einsum("nij,nj->ni", U[:, t], u[:, t + 1]))
        return u

    def twisted(self, f):
        """block LDL' eliminated from both ends towards the root period r = T//2 (what the CUDA kernel does:
        half the sequential depth of a one-way sweep, same stability -- it is a Cholesky in another order)."""
        L, D, U = self.orig
        N, T = f.shape[:2]
        r = T // 2
        if not hasattr(self, "_tw"):
            Dh = D.copy(); G = np.zeros_like(D); G2 = np.zeros((N, 2, 2))
            mm = lambda A, B: np.einsum("nij,njk->nik", A, B)
            def step(C, R, Ct, Dt):
                """G = C R^-1, Dh = D - G C'.  INV == "defer" mirrors the kernel's elimination step: the neighbour's block R
                itself is passed on, X = C adj(R) and X C' are formed while 1/det(R) is in flight, one FMA finishes."""
                if INV != "defer":
                    Gt = mm(C, _inv2(R))
                    return Gt, Dt - mm(Gt, Ct)
                rd = (1.0 / (R[:, 0, 0] * R[:, 1, 1] - R[:, 0, 1] * R[:, 1, 0]))[:, None, None]
                adj = np.empty_like(R)
                adj[:, 0, 0] = R[:, 1, 1]; adj[:, 1, 1] = R[:, 0, 0]; adj[:, 0, 1] = -R[:, 0, 1]; adj[:, 1, 0] = -R[:, 1, 0]
                X = mm(C, adj)
                return X * rd, Dt - mm(X, Ct) * rd
            for t in range(1, r):                       # chain A, downwards
                G[:, t], Dh[:, t] = step(L[:, t], Dh[:, t - 1], U[:, t - 1], D[:, t])
            for t in range(T - 2, r, -1):               # chain B, upwards
                G[:, t], Dh[:, t] = step(U[:, t], Dh[:, t + 1], L[:, t + 1], D[:, t])
            if r >= 1:
                G[:, r] = mm(L[:, r], _inv2(Dh[:, r - 1])); Dh[:, r] = D[:, r] - mm(G[:, r], U[:, r - 1])
            if r + 1 <= T - 1:
                G2 = mm(U[:, r], _inv2(Dh[:, r + 1])); Dh[:, r] = Dh[:, r] - mm(G2, L[:, r + 1])
            self._tw = (G, G2, np.stack([_inv2(Dh[:, t]) for t in range(T)], axis=1))
        G, G2, Dhi = self._tw
        mv = lambda A, v: np.einsum("nij,nj->ni", A, v)
        g = f.copy()
        for t in range(1, r):
            g[:, t] -= mv(G[:, t], g[:, t - 1])
        for t in range(T - 2, r, -1):
            g[:, t] -= mv(G[:, t], g[:, t + 1])
        if r >= 1:
            g[:, r] -= mv(G[:, r], g[:, r - 1])
        if r + 1 <= T - 1:
            g[:, r] -= mv(G2, g[:, r + 1])
        u = np.zeros_like(g)
        u[:, r] = mv(Dhi[:, r], g[:, r])
        for t in range(r - 1, -1, -1):
            u[:, t] = mv(Dhi[:, t], g[:, t] - mv(U[:, t], u[:, t + 1]))
        for t in range(r + 1, T):
            u[:, t] = mv(Dhi[:, t], g[:, t] - mv(L[:, t], u[:, t - 1]))
        return u

    def dense_solve(self, f):
        L, D, U = self.orig
        N, T = f.shape[:2]
        M = np.zeros((N, 2 * T, 2 * T))
        for t in range(T):
            M[:, 2*t:2*t+2, 2*t:2*t+2] = D[:, t]
            if t > 0: M[:, 2*t:2*t+2, 2*t-2:2*t] = L[:, t]
            if t < T-1: M[:, 2*t:2*t+2, 2*t+2:2*t+4] = U[:, t]
        return np.linalg.solve(M, f.reshape(N, 2*T, 1)).reshape(N, T, 2)

    def _solve(self, f):
        N = f.shape[0]
        F = np.zeros((N, self.lanes, 2)); F[:, :self.T] = f
        for k, al, ga in self.steps:
            F = F + _mv(al, _shift_down(F, k)) + _mv(ga, _shift_up(F, k))
        return _mv(self.Dinv, F)[:, :self.T]


def solve_batch(lmp, wcf, P, consts, tol=1e-9, feas_tol=1e-9, max_iter=60, eta=0.9995, gap_floor=1e-4, verbose=False, rho=1e-8, start=None, stop_mu=None, start_mode=0, rho_rule=None, rho_rel=True):
    """lmp [N,T] $/MWh; wcf [N,T] = wind_kw*cf (kW); P [N] battery kW.
    consts: dict(a, binv, half, delta, dur, k_rev) taken from the LP template.
    Returns dict(obj_lp [N] (= c'x, without the design constant), status, iters, g,i,o,s,e [N,T])."""
    lmp = np.atleast_2d(np.asarray(lmp, float)); N, T = lmp.shape
    wcf = np.broadcast_to(np.atleast_2d(np.asarray(wcf, float)), (N, T)).copy()
    P = np.broadcast_to(np.asarray(P, float), (N,)).copy()
    a, binv, hf, dl, dur, k_rev = (consts[k] for k in ("a", "binv", "half", "delta", "dur", "k_rev"))
    c = k_rev * lmp                                       # cost of g and o
    b3 = (dur * P)[:, None] * np.ones((1, T)); b4 = wcf
    beta_b = np.maximum(np.maximum(np.abs(b3).max(1), np.abs(b4).max(1)), P); beta_b = np.where(beta_b > 0, beta_b, 1.0)
    beta_c = np.abs(c).max(1); beta_c = np.where(beta_c > 0, beta_c, 1.0)
    c = c / beta_c[:, None]; b3 = b3 / beta_b[:, None]; b4 = b4 / beta_b[:, None]
    u = np.maximum(P / beta_b, 1e-10)[:, None] * np.ones((1, T))
    nrm_b = 1.0 + np.maximum(np.abs(b3).max(1), np.abs(b4).max(1)); nrm_c = 1.0 + (np.abs(c).max(1) > 0)
    hs = np.ones((1, T)); hs[0, T - 1] = 0.0              # mask of the s column (absent in the last period)
    one = np.ones((N, T))
    x = {k: one.copy() for k in "giosepq"}; z = {k: one.copy() for k in "giosepq"}
    for k in "io":
        x[k] = np.minimum(1.0, 0.5 * u)
    x["s"] = x["s"] * hs; z["s"] = z["s"] * hs
    sb = {k: u - x[k] for k in "io"}; wb = {k: one.copy() for k in "io"}
    y = {k: np.zeros((N, T)) for k in (1, 2, 3, 4)}
    if start_mode >= 1:          # experiment: primal start that satisfies the two local rows exactly
        th = 0.5
        x["g"] = np.maximum(th * b4, 1e-2); x["i"] = np.minimum(np.minimum(1.0, 0.5 * u), np.maximum(0.25 * b4, 1e-2)); x["o"] = x["i"].copy()
        x["q"] = np.maximum(b4 - x["g"] - x["i"], 1e-2)
        x["s"] = np.maximum(0.5 * b3, 1e-2) * hs; x["e"] = np.maximum(0.5 * np.cumsum(x["i"] + x["o"], axis=1), 1e-2)
        x["p"] = np.maximum(b3 - x["s"] - dl * x["e"], 1e-2)
        sb = {k: u - x[k] for k in "io"}
    if start_mode >= 2:
        mu0 = 0.1
        for k in "giosepq":
            z[k] = np.where(x[k] > 0, mu0 / np.where(x[k] > 0, x[k], 1.0), 0.0)
        for k in "io":
            wb[k] = mu0 / sb[k]
    if start is not None:        # experiment: common warm start (scaled iterate of a representative LP)
        x = {k: np.repeat(start["x"][k], N, 0) for k in "giosepq"}; z = {k: np.repeat(start["z"][k], N, 0) for k in "giosepq"}
        sb = {k: np.repeat(start["sb"][k], N, 0) for k in "io"}; wb = {k: np.repeat(start["wb"][k], N, 0) for k in "io"}
        y = {k: np.repeat(start["y"][k], N, 0) for k in (1, 2, 3, 4)}
    ntot = 7 * T - 1 + 2 * T
    status = np.full(N, MAXITER); iters = np.full(N, max_iter)
    active = np.ones(N, bool)
    pobj = np.zeros(N)
    for it in range(max_iter + 1):
        y1n, y2n = _shift_up(y[1], 1), _shift_up(y[2], 1)
        rp = {1: -(x["s"] - _shift_down(x["s"], 1) - a * x["i"] + binv * x["o"]),
              2: -(x["e"] - _shift_down(x["e"], 1) - hf * x["i"] - hf * x["o"]),
              3: b3 - (x["s"] + dl * x["e"] + x["p"]),
              4: b4 - (x["g"] + x["i"] + x["q"])}
        rd = {"g": c - y[4] - z["g"],
              "i": -(-a * y[1] - hf * y[2] + y[4]) - z["i"] + wb["i"],
              "o": c - (binv * y[1] - hf * y[2]) - z["o"] + wb["o"],
              "s": (-(y[1] - y1n + y[3]) - z["s"]) * hs,
              "e": -(y[2] - y2n + dl * y[3]) - z["e"],
              "p": -y[3] - z["p"], "q": -y[4] - z["q"]}
        ru = {k: u - x[k] - sb[k] for k in "io"}
        mu = (sum((x[k] * z[k]).sum(1) for k in "giosepq") + sum((sb[k] * wb[k]).sum(1) for k in "io")) / ntot
        po = (c * (x["g"] + x["o"])).sum(1)
        dobj = (b3 * y[3] + b4 * y[4]).sum(1) - (u * (wb["i"] + wb["o"])).sum(1)
        pmax = np.max([np.abs(v).max(1) for v in rp.values()] + [np.abs(v).max(1) for v in ru.values()], axis=0)
        dmax = np.max([np.abs(v).max(1) for v in rd.values()], axis=0)
        res = np.maximum(pmax / nrm_b, dmax / nrm_c)
        den = np.maximum(gap_floor, np.abs(po))
        gap = np.abs(po - dobj) / den; cgap = ntot * mu / den
        done = (res < feas_tol) & (gap < tol)
        done |= (cgap < tol) & (res < 10 * feas_tol) & (gap < 10 * tol)
        give = (cgap < 1e-3 * tol) & ~done
        done |= give & (res < 100 * feas_tol) & (gap < 1000 * tol)
        fail = active & give & ~done
        newly = active & done
        pobj = np.where(active, po, pobj)
        status[newly] = OPTIMAL; iters[newly] = it; status[fail] = NUMERR; iters[fail] = it
        active &= ~(done | fail)
        if verbose:
            print(it, "active", active.sum(), "res %.2e gap %.2e mu %.2e" % (res.max(), gap.max(), mu.max()),
                  "rp", " ".join("%.1e" % np.abs(v).max() for v in rp.values()), "ru", " ".join("%.1e" % np.abs(v).max() for v in ru.values()),
                  "rd", " ".join("%s %.1e" % (k, np.abs(v).max()) for k, v in rd.items()))
        if not active.any() or it == max_iter:
            break
        if stop_mu is not None and mu.max() < stop_mu:
            return dict(x=x, z=z, sb=sb, wb=wb, y=y, it=it)
        # ---- scaling matrix (rho_rule = (c, lo, hi): proximal weight clip(c*mu, lo, hi) per LP instead of a constant)
        if rho_rule is not None:
            rho = np.clip(rho_rule[0] * mu, rho_rule[1], rho_rule[2])[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            rr = (lambda k: rho / np.maximum(1.0, x[k] * x[k])) if rho_rel else (lambda k: rho)
            d = {k: 1.0 / (z[k] / x[k] + rr(k)) for k in "gepq"}
            d["s"] = np.where(hs > 0, 1.0 / (np.where(hs > 0, z["s"] / np.where(hs > 0, x["s"], 1.0), 1.0) + rr("s")), 0.0)
            for k in "io":
                d[k] = 1.0 / (z[k] / x[k] + wb[k] / sb[k] + rr(k))
        # ---- per-period blocks of M after eliminating r4 (pivot m44) and r3 (pivot m33), written in the
        # ---- cancellation-free form  d - d^2/m = d (m - d)/m  (the d's span 20+ orders of magnitude near the end)
        kap = 1.0 / (d["s"] + dl * dl * d["e"] + d["p"])
        s11 = d["s"] * (dl * dl * d["e"] + d["p"]) * kap
        s22 = d["e"] * (d["s"] + d["p"]) * kap
        s12 = dl * d["s"] * d["e"] * kap
        iot = 1.0 / (d["g"] + d["i"] + d["q"])
        tau = d["i"] * (d["g"] + d["q"]) * iot
        s11p, s22p, s12p = _shift_down(s11, 1), _shift_down(s22, 1), _shift_down(s12, 1)
        D = np.zeros((N, T, 2, 2))
        D[..., 0, 0] = s11 + s11p + a * a * tau + binv * binv * d["o"]
        D[..., 1, 1] = s22 + s22p + hf * hf * tau + hf * hf * d["o"]
        D[..., 0, 1] = D[..., 1, 0] = a * hf * tau - hf * binv * d["o"] - s12 - s12p
        Bm = np.zeros((N, T, 2, 2))                       # coupling of period t+1 (rows) with t (cols), symmetric
        Bm[..., 0, 0] = -s11; Bm[..., 1, 1] = -s22
        Bm[..., 0, 1] = Bm[..., 1, 0] = s12
        Bm[:, T - 1] = 0.0
        Lb = np.zeros_like(Bm); Lb[:, 1:] = Bm[:, :-1]
        pcr = PCR(Lb, D, Bm)
        dsk, dek, dii = d["s"] * kap, dl * d["e"] * kap, d["i"] * iot

        def newton(ax, as_):
            h = {k: rd[k] + z[k] - (ax[k] / x[k] if ax is not None else 0.0) for k in "gepq"}
            with np.errstate(divide="ignore", invalid="ignore"):
                h["s"] = np.where(hs > 0, rd["s"] + z["s"] - (ax["s"] / np.where(hs > 0, x["s"], 1.0) if ax is not None else 0.0), 0.0)
            for k in "io":
                aa = (as_[k] if as_ is not None else 0.0) - wb[k] * ru[k]
                h[k] = rd[k] + z[k] - (ax[k] / x[k] if ax is not None else 0.0) + aa / sb[k] - wb[k]
            w3 = rp[3] + d["p"] * h["p"]
            ph1 = s11 * h["s"] - s12 * h["e"] - dsk * w3
            ph2 = s22 * h["e"] - s12 * h["s"] - dek * w3
            w4 = rp[4] + d["g"] * h["g"] + d["q"] * h["q"]
            psi = tau * h["i"] - dii * w4
            doh = d["o"] * h["o"]
            f = np.stack([rp[1] + ph1 - _shift_down(ph1, 1) - a * psi + binv * doh,
                          rp[2] + ph2 - _shift_down(ph2, 1) - hf * psi - hf * doh], axis=-1)
            uu = pcr.solve(f)
            dy1, dy2 = uu[..., 0], uu[..., 1]
            e1 = dy1 - _shift_up(dy1, 1) - h["s"]; e2 = dy2 - _shift_up(dy2, 1) - h["e"]
            v = a * dy1 + hf * dy2
            dx = {}
            dx["s"] = (s11 * e1 - s12 * e2 + dsk * w3) * hs
            dx["e"] = s22 * e2 - s12 * e1 + dek * w3
            dx["p"] = None
            dx["i"] = -tau * (v + h["i"]) + dii * w4
            dx["o"] = d["o"] * (binv * dy1 - hf * dy2 - h["o"])
            dx["g"] = d["g"] * iot * (rp[4] + d["i"] * (h["i"] - h["g"] + v) + d["q"] * (h["q"] - h["g"]))
            dy3 = kap * (w3 - d["s"] * e1 - dl * d["e"] * e2)
            dy4 = iot * (w4 + d["i"] * (h["i"] + v))
            dx["p"] = d["p"] * (dy3 - h["p"])
            dx["q"] = d["q"] * (dy4 - h["q"])
            return dx, {1: dy1, 2: dy2, 3: dy3, 4: dy4}

        def dual_steps(dx, ax, as_):
            dz, ds_, dw = {}, {}, {}
            for k in "giosepq":
                with np.errstate(divide="ignore", invalid="ignore"):
                    xx = np.where(x[k] > 0, x[k], 1.0)
                    dz[k] = ((ax[k] if ax is not None else 0.0) / xx - z[k] - z[k] * dx[k] / xx) * (hs if k == "s" else 1.0)
            for k in "io":
                ds_[k] = ru[k] - dx[k]
                dw[k] = (as_[k] if as_ is not None else 0.0) / sb[k] - wb[k] - wb[k] * ds_[k] / sb[k]
            return dz, ds_, dw

        def maxstep(pairs):
            r = np.full(N, np.inf)
            for v, dv in pairs:
                with np.errstate(divide="ignore", invalid="ignore"):
                    q = np.where(dv < 0, -v / np.where(dv < 0, dv, -1.0), np.inf)
                r = np.minimum(r, q.min(1))
            return r

        dx, dy = newton(None, None)
        dz, ds_, dw = dual_steps(dx, None, None)
        ap = np.minimum(1.0, maxstep([(x[k], dx[k]) for k in "giosepq"] + [(sb[k], ds_[k]) for k in "io"]))
        ad = np.minimum(1.0, maxstep([(z[k], dz[k]) for k in "giosepq"] + [(wb[k], dw[k]) for k in "io"]))
        mua = (sum(((x[k] + ap[:, None] * dx[k]) * (z[k] + ad[:, None] * dz[k])).sum(1) for k in "giosepq")
               + sum(((sb[k] + ap[:, None] * ds_[k]) * (wb[k] + ad[:, None] * dw[k])).sum(1) for k in "io")) / ntot
        smu = ((mua / mu) ** 3 * mu)[:, None]
        ax = {k: smu - dx[k] * dz[k] for k in "giosepq"}
        as_ = {k: smu - ds_[k] * dw[k] for k in "io"}
        dx, dy = newton(ax, as_)
        dz, ds_, dw = dual_steps(dx, ax, as_)
        ap = np.minimum(1.0, eta * maxstep([(x[k], dx[k]) for k in "giosepq"] + [(sb[k], ds_[k]) for k in "io"]))
        ad = np.minimum(1.0, eta * maxstep([(z[k], dz[k]) for k in "giosepq"] + [(wb[k], dw[k]) for k in "io"]))
        A_ = active[:, None]
        for k in "giosepq":
            x[k] = np.where(A_, x[k] + ap[:, None] * dx[k], x[k]); z[k] = np.where(A_, z[k] + ad[:, None] * dz[k], z[k])
        for k in "io":
            sb[k] = np.where(A_, sb[k] + ap[:, None] * ds_[k], sb[k]); wb[k] = np.where(A_, wb[k] + ad[:, None] * dw[k], wb[k])
        for k in (1, 2, 3, 4):
            y[k] = np.where(A_, y[k] + ad[:, None] * dy[k], y[k])
    return dict(obj_lp=pobj * beta_b * beta_c, status=status, iters=iters,
                **{k: x[k] * beta_b[:, None] for k in "giose"})
