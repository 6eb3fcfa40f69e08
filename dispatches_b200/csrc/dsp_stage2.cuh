// dsp_stage2.cuh -- stage-structured IPM kernel, generation 2: SEVERAL LPs PER WARP, several periods per lane.
//
// Hot path of BASELINE configs C1/C2/C5: wind_battery_optimize with design_opt=False (wind_battery_LMP.py:172-267) in the
// reduced form of dispatches_b200/templates.py::wind_battery, horizons T <= 32*P (instantiated for P <= 3: T <= 96).
//
// Why (ncu, round 1): the lane-per-period kernel (dsp_stage_wb.cuh) spends 56 % of its cycles in three sequential block-LDL'
// sweeps in which 1-2 of 32 lanes do useful work, and T = 24 leaves 8 lanes idle throughout.  Here
//   * an LP occupies a GROUP of L lanes (L = 8 for T <= 24), each lane owns P CONSECUTIVE periods (P = 3): 32/L LPs per warp,
//     32/32 lanes busy in the element-wise passes;
//   * the iterate (22 doubles per period) stays in REGISTERS (period loops are unrolled); what must survive from one pass to
//     the next (scaling blocks, reciprocals, second-order products, the local factor) is parked in SHARED MEMORY by design,
//     laid out [array][period slot][lane] (conflict free); cheap things (residuals, right-hand-side pieces) are recomputed:
//     on this chip one shared-memory access costs as much SM throughput as ~4 FP64 instructions;
//   * the block-tridiagonal (2x2 blocks) reduced normal equations are factorised by a PARTITIONED elimination: every lane
//     eliminates its first P-1 periods locally (all lanes busy; the fill is one 2x2 "spike" block towards the left
//     neighbour's last period), the L remaining separator periods form a short chain that is eliminated from both ends with
//     width-L shuffles (L/2 steps instead of T/2).  It is a block Cholesky in nested-dissection order: same stability;
//   * every group fetches its next LP from the ticket counter as soon as ITS LP has converged (iteration counts differ,
//     8...20): the warp never waits for its slowest LP.
// Algorithm (Mehrotra predictor-corrector, scaling, start point, stopping rules, proximal term): identical to dsp_stage_wb.cuh
// and oracle/ipm_stage_numpy.py; only the elimination ORDER of the reduced system differs (results agree to rounding).
//
// The warp body is plain C++ over the warp-collective builtins, so tests/emu compiles THIS FILE with g++ on a lock-step
// SIMT emulator and checks it against the oracle without a GPU (test infrastructure; the product path is the CUDA build).
#pragma once

namespace stage2 {

#ifndef DSP_OPTIMAL
#define DSP_OPTIMAL 0
#define DSP_MAX_ITER 1
#define DSP_NUMERICAL 2
#define DSP_INFEASIBLE 3
#endif

#define S2D __device__ __forceinline__
#ifndef DSP_S2_SYNCMASK          // which of the five phase boundaries of a round carry a CTA barrier (experiments: tools/build_variants.py)
#define DSP_S2_SYNCMASK 0       // measured: the exit vote at the top of the round alone keeps the warps in step (profiles/stage2_variants_r2.log)
#endif
constexpr unsigned FULL = 0xffffffffu;
constexpr double kGapFloor2 = 1e-4;

struct Params {
    // batch
    long long N;
    const double *cparams, *rparams;
    long long rstride;
    int Pc, Pr;
    const double *omap, *ocmap;
    double o0;
    double tol, feas_tol, step_frac, reg;
    int max_iter;
    double *obj, *x_out, *y_out;
    int *status, *iters;
    int n, m;
    unsigned long long *ticket;
    // wind+battery stage structure (include/dsp_lp.h: dsp_stage_wb_desc)
    int T;
    double a, binv, hf, dl, dur, krev;
    int wcf_off, p_off;
    const int *col_idx, *row_idx;
    int ahead;                 // number of group slots of the launch (L2 prefetch distance in LPs); 0 in the emulator
};

S2D double frcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
    double r = (double)(1.0f / (float)x);        // emulation: any ~20-bit seed; the Newton steps below do the rest
    if (!(r == r) || r == 0.0 || r > 1e300 || r < -1e300) r = 1.0 / x;
#endif
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
#if !defined(DSP_S2_RCP_NEWTON) || DSP_S2_RCP_NEWTON >= 2
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
#endif
    return r;
}
S2D double dmax(double a, double b) { return a > b ? a : b; }
// L2 prefetch of `bytes` starting at p, one 128-byte line per lane `l` of the group (inputs of an LP are read once, at its refill,
// by a group that then waits for them: a DRAM access on the critical path of every round of the CTA unless the lines are in L2)
S2D void prefetch_l2(const void *p, int bytes, int l) {
#if defined(__CUDA_ARCH__)
    int off = l * 128;
    if (off < bytes + 128) {                       // (every prefetched address lies inside the row: the last lane takes its last byte)
        if (off > bytes - 1) off = bytes - 1;
        asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)p + off));
    }
#endif
}

// ---- collectives over the L lanes of an LP group (every lane of the warp executes them)
template <int L>
S2D double gsum(double v) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o, L);
    return v;
}
template <int L>
S2D double gmax(double v) {
#pragma unroll
    for (int o = L / 2; o; o >>= 1) v = dmax(v, __shfl_xor_sync(FULL, v, o, L));
    return v;
}
template <int L>
S2D double gup1(double v, int gl) {            // value of the previous lane of the group, 0 for the first
    const double r = __shfl_up_sync(FULL, v, 1, L);
    return gl == 0 ? 0.0 : r;
}
template <int L>
S2D double gdown1(double v, int gl) {          // value of the next lane of the group, 0 for the last
    const double r = __shfl_down_sync(FULL, v, 1, L);
    return gl == L - 1 ? 0.0 : r;
}
template <int L>
S2D double gfrom(double v, int src) { return __shfl_sync(FULL, v, src, L); }

struct Sym2 { double a, b, c; };            // [[a, b], [b, c]]
struct Mat2 { double a, b, c, d; };         // [[a, b], [c, d]]

S2D Sym2 inv_spd(const Sym2 &D) {
    const double i = frcp(fma(D.a, D.c, -(D.b * D.b)));
    Sym2 r;
    r.a = D.c * i; r.b = -D.b * i; r.c = D.a * i;
    return r;
}
S2D Mat2 mul_ms(const Mat2 &A, const Sym2 &B) {       // A * B
    Mat2 r;
    r.a = fma(A.a, B.a, A.b * B.b); r.b = fma(A.a, B.b, A.b * B.c);
    r.c = fma(A.c, B.a, A.d * B.b); r.d = fma(A.c, B.b, A.d * B.c);
    return r;
}
S2D Mat2 mul_sm(const Sym2 &A, const Mat2 &B) {       // A * B
    Mat2 r;
    r.a = fma(A.a, B.a, A.b * B.c); r.b = fma(A.a, B.b, A.b * B.d);
    r.c = fma(A.b, B.a, A.c * B.c); r.d = fma(A.b, B.b, A.c * B.d);
    return r;
}
S2D Mat2 mul_mm(const Mat2 &A, const Mat2 &B) {
    Mat2 r;
    r.a = fma(A.a, B.a, A.b * B.c); r.b = fma(A.a, B.b, A.b * B.d);
    r.c = fma(A.c, B.a, A.d * B.c); r.d = fma(A.c, B.b, A.d * B.d);
    return r;
}
S2D Mat2 transp(const Mat2 &A) { Mat2 r; r.a = A.a; r.b = A.c; r.c = A.b; r.d = A.d; return r; }
// D -= X * C'   (the product is symmetric in exact arithmetic: X = C K with K symmetric)
S2D void sub_xct(Sym2 &D, const Mat2 &X, const Mat2 &C) {
    D.a -= fma(X.a, C.a, X.b * C.b);
    D.b -= fma(X.a, C.c, X.b * C.d);
    D.c -= fma(X.c, C.c, X.d * C.d);
}

struct Per {            // one period of the iterate + its data (registers)
    double xg, xi, xo, xs, xe, xp, xq;
    double zg, zi, zo, zs, ze, zp, zq;
    double si, so, wi, wo;
    double y1, y2, y3, y4;
};
S2D double comp_sum(const Per &q) {           // sum of the 9 complementarity products of a period
    return q.xg * q.zg + q.xi * q.zi + q.xo * q.zo + q.xs * q.zs + q.xe * q.ze + q.xp * q.zp + q.xq * q.zq + q.si * q.wi + q.so * q.wo;
}
struct Res { double rp1, rp2, rp3, rp4, rdg, rdi, rdo, rds, rde, rdp, rdq, rui, ruo; };
struct H7 { double hg, hi, ho, hs, he, hp, hq; };
struct Cst { double a, binv, hf, dl; };

S2D void residuals(const Per &q, double c, double b4, double xsp, double xep, double y1n, double y2n, double b3, double u, const Cst &K,
                   bool act, bool has_s, Res &r) {
    if (act) {
        r.rp1 = -(q.xs - xsp - K.a * q.xi + K.binv * q.xo);
        r.rp2 = -(q.xe - xep - K.hf * q.xi - K.hf * q.xo);
        r.rp3 = b3 - (q.xs + K.dl * q.xe + q.xp);
        r.rp4 = b4 - (q.xg + q.xi + q.xq);
        r.rdg = c - q.y4 - q.zg;
        r.rdi = K.a * q.y1 + K.hf * q.y2 - q.y4 - q.zi + q.wi;
        r.rdo = c - K.binv * q.y1 + K.hf * q.y2 - q.zo + q.wo;
        r.rds = has_s ? -(q.y1 - y1n + q.y3) - q.zs : 0.0;
        r.rde = -(q.y2 - y2n + K.dl * q.y3) - q.ze;
        r.rdp = -q.y3 - q.zp;
        r.rdq = -q.y4 - q.zq;
        r.rui = u - q.xi - q.si;
        r.ruo = u - q.xo - q.so;
    } else {
        r.rp1 = r.rp2 = r.rp3 = r.rp4 = r.rdg = r.rdi = r.rdo = r.rds = r.rde = r.rdp = r.rdq = r.rui = r.ruo = 0.0;
    }
}

// shared-memory arrays of a warp: [array][period slot j][lane]
enum { A_DS = 0, A_DE, A_DP, A_KAP, A_DG, A_DI, A_DQ, A_IOT, A_DO,   // 9 scaling values; the blocks s11.. are re-derived (12 flops)
       A_RX = 9,       // 7: 1/x of g,i,o,s,e,p,q
       A_PR = 16,      // 9: second-order products dx dz (7), ds dw (2) of the predictor
       A_F = 25,       // 2: forward-eliminated right-hand side of the reduced system
       A_C = 27, A_B4 = 28,   // period data: scaled cost of g / o, scaled wind availability
       NA_FULL = 29,
       // after the corrector's direction recovery the scaling values of a period are dead: slots 0..8 then hold dx (7), dy3, dy4
       A_DX = 0, A_DY3 = 7, A_DY4 = 8,
       // interior factor (periods 0..P-2 of a lane): K (3), G = C K (4), H = E' K (4)
       I_K = 0, I_G = 3, I_H = 7, NA_INT = 11 };

struct Scal { double s11, s22, s12, dsk, dek, kap, tau, dii, iot, dO, dg, dq, dp, di; };
// the per-period blocks after eliminating the two local rows, from the stored scaling values (cancellation-free forms)
S2D Scal make_scal(double ds, double de, double dp, double kap, double dg, double di, double dq, double iot, double dO, double dl) {
    Scal r;
    r.dsk = ds * kap; r.dek = dl * de * kap;
    r.s11 = r.dsk * fma(dl * dl, de, dp);
    r.s22 = de * (ds + dp) * kap;
    r.s12 = r.dsk * dl * de;
    r.dii = di * iot;
    r.tau = r.dii * (dg + dq);
    r.kap = kap; r.iot = iot; r.dO = dO; r.dg = dg; r.dq = dq; r.dp = dp; r.di = di;
    return r;
}

template <int P>
struct SmemDoubles { static constexpr int value = (NA_FULL * P + NA_INT * (P > 1 ? P - 1 : 0)) * 32; };
template <int P>
constexpr int smem_doubles_per_warp() { return SmemDoubles<P>::value; }

// CTA_SYNC: the warps of a CTA pass the phases of an IPM round together (bar.sync at the phase boundaries).  They all execute the
// same ~6.7k instructions per round; unsynchronised they spread over the loop body and each streams it through the 32 KB
// instruction cache on its own (ncu, round 2: 41 % of the stall samples were no_instructions) -- in step, one warp's fetch
// serves the others.
template <bool CTA_SYNC>
S2D void cta_sync() {
#if defined(__CUDA_ARCH__)
    if (CTA_SYNC) __syncthreads();
#endif
}
template <bool CTA_SYNC>
S2D bool cta_all(bool pred) {
#if defined(__CUDA_ARCH__)
    if (CTA_SYNC) return __syncthreads_and(pred) != 0;
#endif
    return __all_sync(FULL, pred) != 0;
}

template <int L, int P, bool CTA_SYNC = false>
__device__ void warp_body(const Params &Q, double *smw, int lane) {
#define SMF(arr, j) sm[((arr) * P + (j)) * 32]
#define SMI(arr, j) smi[((arr) * (P - 1) + (j)) * 32]
#define LOAD_SCAL(j) make_scal(SMF(A_DS, j), SMF(A_DE, j), SMF(A_DP, j), SMF(A_KAP, j), SMF(A_DG, j), SMF(A_DI, j), SMF(A_DQ, j), SMF(A_IOT, j), SMF(A_DO, j), dl)
    const int gl = lane & (L - 1);
    double *sm = smw + lane;
    double *smi = smw + NA_FULL * P * 32 + lane;
    Cst K;
    K.a = Q.a; K.binv = Q.binv; K.hf = Q.hf; K.dl = Q.dl;
    const double a = K.a, binv = K.binv, hf = K.hf, dl = K.dl;
    const int T = Q.T;
    constexpr int r_root = L / 2;
    constexpr int kmax = (r_root - 1 > L - 2 - r_root) ? r_root - 1 : L - 2 - r_root;
    constexpr int smax = (r_root > L - 1 - r_root) ? r_root : L - 1 - r_root;

    Per pr[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        Per &q = pr[j];
        q.xg = q.xi = q.xo = q.xs = q.xe = q.xp = q.xq = 0.0;
        q.zg = q.zi = q.zo = q.zs = q.ze = q.zp = q.zq = 0.0;
        q.si = q.so = q.wi = q.wo = q.y1 = q.y2 = q.y3 = q.y4 = 0.0;
    }
    double b3 = 0.0, u = 1.0, nrm_b = 1.0, nrm_c = 1.0, ntot = 1.0, beta_b = 1.0, beta_c = 1.0, kconst = 0.0;
    double step_frac = Q.step_frac, reg = Q.reg;
    long long p = -1;
    int it = 0, it0 = 0, attempt = 0, Tg = 0;
    int mode = 1;                     // 0 running, 1 needs a new LP, 2 retries its LP with safer parameters, 3 out of work

    double mu_keep = 0.0;             // complementarity measure of the group's current iterate (set by the check / the refill)
    double xs_left = 0.0, xe_left = 0.0, y1_right = 0.0, y2_right = 0.0;      // neighbours of the lane's block
#define XSP(j) ((j) == 0 ? xs_left : pr[(j) > 0 ? (j) - 1 : 0].xs)
#define XEP(j) ((j) == 0 ? xe_left : pr[(j) > 0 ? (j) - 1 : 0].xe)
#define Y1N(j) ((j) == P - 1 ? y1_right : pr[(j) < P - 1 ? (j) + 1 : 0].y1)
#define Y2N(j) ((j) == P - 1 ? y2_right : pr[(j) < P - 1 ? (j) + 1 : 0].y2)
#define ACT(j) (gl * P + (j) < Tg)
#define HAS_S(j) (gl * P + (j) < Tg - 1)
#define NEIGHBOURS()                                                                                      \
    {                                                                                                     \
        xs_left = gup1<L>(pr[P - 1].xs, gl); xe_left = gup1<L>(pr[P - 1].xe, gl);                        \
        y1_right = gdown1<L>(pr[0].y1, gl); y2_right = gdown1<L>(pr[0].y2, gl);                          \
    }

    for (;;) {
        // =========================================================================================== convergence check
        // residual norms, duality gap and complementarity of the iterate every running group holds (one cheap evaluation of the
        // residuals).  It runs BEFORE the refill, so a group whose LP has just converged starts its next LP in this very round
        // (until round 2 the test sat inside pass 1 and a finished group idled through the rest of that round: 1 round in 13)
        if (__any_sync(FULL, mode == 0)) {
            NEIGHBOURS();
            double pm = 0.0, dm = 0.0, mus = 0.0, po = 0.0, dob = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const Per &q = pr[j];
                if (ACT(j)) {
                    Res r;
                    residuals(q, SMF(A_C, j), SMF(A_B4, j), XSP(j), XEP(j), Y1N(j), Y2N(j), b3, u, K, true, HAS_S(j), r);
                    pm = dmax(pm, dmax(dmax(dmax(fabs(r.rp1), fabs(r.rp2)), dmax(fabs(r.rp3), fabs(r.rp4))), dmax(fabs(r.rui), fabs(r.ruo))));
                    dm = dmax(dm, dmax(dmax(dmax(fabs(r.rdg), fabs(r.rdi)), dmax(fabs(r.rdo), fabs(r.rds))),
                                       dmax(dmax(fabs(r.rde), fabs(r.rdp)), fabs(r.rdq))));
                    mus += comp_sum(q);
                    po += SMF(A_C, j) * (q.xg + q.xo);
                    dob += b3 * q.y3 + SMF(A_B4, j) * q.y4 - u * (q.wi + q.wo);
                }
            }
            const double res = gmax<L>(dmax(pm / nrm_b, dm / nrm_c));
            mus = gsum<L>(mus); po = gsum<L>(po); dob = gsum<L>(dob);
            const double mu = mus / ntot;
            const double den = dmax(kGapFloor2, fabs(po));
            const double gap = fabs(po - dob) / den, cgap = ntot * mu / den;
            if (mode == 0) {
                mu_keep = mu;
                int status = -1;
                if (!(mu == mu) || !(po == po) || mu > 1e100) status = DSP_NUMERICAL;
                else if (res < Q.feas_tol && gap < Q.tol) status = DSP_OPTIMAL;
                else if (cgap < Q.tol && res < 10.0 * Q.feas_tol && gap < 10.0 * Q.tol) status = DSP_OPTIMAL;
                else if (cgap < 1e-3 * Q.tol) status = (res < 100.0 * Q.feas_tol && gap < 1000.0 * Q.tol) ? DSP_OPTIMAL : DSP_NUMERICAL;
                else if (it == Q.max_iter) status = DSP_MAX_ITER;
                if (status >= 0) {
                    if (gl == 0) {
                        Q.obj[p] = po * beta_b * beta_c + kconst;
                        Q.status[p] = status;
                        Q.iters[p] = it + it0;
                    }
                    if (Q.x_out) {
                        double *xo_ = Q.x_out + p * (long long)Q.n;
#pragma unroll
                        for (int j = 0; j < P; ++j) {
                            const int t = gl * P + j;
                            if (t < T) {
                                const int *ci_ = Q.col_idx + t * 7;
                                const Per &q = pr[j];
                                const double vals[7] = {q.xg, q.xi, q.xo, q.xs, q.xe, q.xp, q.xq};
#pragma unroll
                                for (int k = 0; k < 7; ++k)
                                    if (ci_[k] >= 0) xo_[ci_[k]] = vals[k] * beta_b;
                            }
                        }
                    }
                    if (Q.y_out) {
                        double *yo_ = Q.y_out + p * (long long)Q.m;
#pragma unroll
                        for (int j = 0; j < P; ++j) {
                            const int t = gl * P + j;
                            if (t < T) {
                                const int *ri_ = Q.row_idx + t * 4;
                                yo_[ri_[0]] = pr[j].y1 * beta_c; yo_[ri_[1]] = pr[j].y2 * beta_c;
                                yo_[ri_[2]] = pr[j].y3 * beta_c; yo_[ri_[3]] = pr[j].y4 * beta_c;
                            }
                        }
                    }
                    // second attempt (shorter step, stronger proximal term) for the rare LP whose first attempt ends non-optimal
                    if (status != DSP_OPTIMAL && attempt == 0) { mode = 2; attempt = 1; it0 = it + it0; }
                    else mode = 1;
                }
            }
        }
        // =========================================================================================== (re)fill groups
        if (__any_sync(FULL, mode == 1 || mode == 2)) {
            unsigned long long tk = 0;
            if (mode == 1 && gl == 0) tk = atomicAdd(Q.ticket, 1ULL);
            tk = __shfl_sync(FULL, tk, 0, L);
            if (mode == 1) {
                if ((long long)tk >= Q.N) { mode = 3; p = -1; Tg = 0; }
                else { p = (long long)tk; attempt = 0; it0 = 0; }
            }
            const bool ld = (mode == 1 || mode == 2);
            double kc = 0.0, b4m = 0.0, cm = 0.0, Pw = 0.0;
            if (ld) {
                const double *cp = Q.cparams + p * (long long)Q.Pc;
                const double *rp = Q.rparams + p * Q.rstride;
                {   // the LP one wave of group slots ahead: in L2 by the time a group asks for it
                    const long long pa = p + (long long)Q.ahead;
                    if (mode == 1 && pa < Q.N) {
                        prefetch_l2(Q.cparams + pa * (long long)Q.Pc, Q.Pc * 8, gl);
                        if (Q.rstride) prefetch_l2(Q.rparams + pa * Q.rstride, Q.Pr * 8, gl);
                    }
                }
                for (int r = gl; r < Q.Pr; r += L) kc += Q.omap[r] * rp[r];
                for (int r = gl; r < Q.Pc; r += L) kc += Q.ocmap[r] * cp[r];
                Pw = rp[Q.p_off];
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const int t = gl * P + j;
                    const bool act = t < T;
                    const double cj = act ? Q.krev * cp[t] : 0.0, bj = act ? rp[Q.wcf_off + t] : 0.0;
                    SMF(A_C, j) = cj; SMF(A_B4, j) = bj;
                    b4m = dmax(b4m, fabs(bj));
                    cm = dmax(cm, fabs(cj));
                }
            }
            kc = gsum<L>(kc);
            b4m = gmax<L>(b4m);
            cm = gmax<L>(cm);
            double mu0 = 0.0;
            if (ld) {
                kconst = kc + Q.o0;
                if (Pw < 0.0) {       // negative battery power bound: infeasible (not silently clamped)
                    if (gl == 0) { Q.obj[p] = __longlong_as_double(0x7ff8000000000000LL); Q.status[p] = DSP_INFEASIBLE; Q.iters[p] = 0; }
                    mode = 1; Tg = 0;                      // fetches the next LP at the top of the next round
#pragma unroll
                    for (int j = 0; j < P; ++j) {           // (an all-inactive group must not carry the finished LP's iterate)
                        Per &q = pr[j];
                        q.xg = q.xi = q.xo = q.xs = q.xe = q.xp = q.xq = 0.0;
                        q.zg = q.zi = q.zo = q.zs = q.ze = q.zp = q.zq = 0.0;
                        q.si = q.so = q.wi = q.wo = q.y1 = q.y2 = q.y3 = q.y4 = 0.0;
                    }
                } else {
                    step_frac = attempt ? 0.99 : Q.step_frac;
                    reg = attempt ? 10.0 * Q.reg : Q.reg;
                    double b3u = Q.dur * Pw;
                    beta_b = dmax(dmax(fabs(b3u), b4m), Pw);
                    beta_b = beta_b > 0.0 ? beta_b : 1.0;
                    beta_c = cm > 0.0 ? cm : 1.0;
                    b3 = b3u / beta_b;
                    u = dmax(Pw / beta_b, 1e-10);
                    nrm_b = 1.0 + dmax(fabs(b3), b4m / beta_b);
                    nrm_c = 1.0 + (cm > 0.0 ? 1.0 : 0.0);
                    ntot = (double)(9 * T - 1);
                    Tg = T;
                    const double x0 = fmin(1.0, 0.5 * u);
#pragma unroll
                    for (int j = 0; j < P; ++j) {
                        const int t = gl * P + j;
                        const bool act = t < T, has_s = t < T - 1;
                        Per &q = pr[j];
                        SMF(A_C, j) = SMF(A_C, j) / beta_c; SMF(A_B4, j) = SMF(A_B4, j) / beta_b;
                        const double one = act ? 1.0 : 0.0;
                        q.xg = one; q.xi = act ? x0 : 0.0; q.xo = q.xi; q.xs = has_s ? 1.0 : 0.0; q.xe = one; q.xp = one; q.xq = one;
                        q.zg = one; q.zi = one; q.zo = one; q.zs = has_s ? 1.0 : 0.0; q.ze = one; q.zp = one; q.zq = one;
                        q.si = act ? u - x0 : 0.0; q.so = q.si; q.wi = one; q.wo = one;
                        q.y1 = q.y2 = q.y3 = q.y4 = 0.0;
                        if (act) mu0 += comp_sum(q);
                    }
                    it = 0;
                    mode = 0;
                }
            }
            mu0 = gsum<L>(mu0);
            if (ld) mu_keep = mu0 / ntot;      // (the start point is never optimal: its own convergence check is skipped)
        }
        if (cta_all<CTA_SYNC>(mode == 3)) break;
        if (__all_sync(FULL, mode == 3)) {
            // this warp is out of work while others of its CTA still iterate: it must not compete for their issue slots --
            // it only keeps the CTA's barriers of the round balanced and waits at the next exit vote
#pragma unroll
            for (int b = 0; b < 5; ++b)
                if (DSP_S2_SYNCMASK & (1 << b)) cta_sync<CTA_SYNC>();
            continue;
        }

        // =========================================================================================== neighbours of the lane's block
        NEIGHBOURS();

        // =========================================================================================== pass 1
        // residuals, scaling matrix, local elimination of the wind-balance / SoC-bound rows, predictor right-hand side
        Sym2 Dd[P];
        double f1[P], f2[P];
        double s11l, s22l, s12l;       // scaling blocks of the left neighbour's last period
        {
            double s11c = 0.0, s22c = 0.0, s12c = 0.0, ph1c = 0.0, ph2c = 0.0;      // carried from period j-1
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const Per &q = pr[j];
                const bool act = ACT(j), has_s = HAS_S(j);
                Res r;
                residuals(q, SMF(A_C, j), SMF(A_B4, j), XSP(j), XEP(j), Y1N(j), Y2N(j), b3, u, K, act, has_s, r);
                if (act) {
                    // ---- scaling matrix and reciprocals.  d = 1 / (z/x [+ w/s] + reg / max(1, x^2))
                    const double rxg = frcp(q.xg), rxi = frcp(q.xi), rxo = frcp(q.xo), rxe = frcp(q.xe), rxp = frcp(q.xp), rxq = frcp(q.xq);
                    const double rxs = has_s ? frcp(q.xs) : 0.0;
                    const double rsi = frcp(q.si), rso = frcp(q.so);
                    const double qg = q.xg > 1.0 ? reg * rxg * rxg : reg, qe = q.xe > 1.0 ? reg * rxe * rxe : reg;
                    const double qp = q.xp > 1.0 ? reg * rxp * rxp : reg, qq = q.xq > 1.0 ? reg * rxq * rxq : reg;
                    const double qs = q.xs > 1.0 ? reg * rxs * rxs : reg;
                    const double dg = q.xg * frcp(fma(qg, q.xg, q.zg)), de = q.xe * frcp(fma(qe, q.xe, q.ze));
                    const double dp = q.xp * frcp(fma(qp, q.xp, q.zp)), dq = q.xq * frcp(fma(qq, q.xq, q.zq));
                    const double ds = has_s ? q.xs * frcp(fma(qs, q.xs, q.zs)) : 0.0;
                    const double di = frcp(fma(q.zi, rxi, q.wi * rsi) + reg), dO = frcp(fma(q.zo, rxo, q.wo * rso) + reg);
                    // ---- per-period blocks after eliminating the two local rows (cancellation-free form d - d^2/m = d (m - d)/m)
                    const double kap = frcp(ds + dl * dl * de + dp);
                    const double iot = frcp(dg + di + dq);
                    const Scal sc = make_scal(ds, de, dp, kap, dg, di, dq, iot, dO, dl);       // same rounding as every later re-derivation
                    const double s11 = sc.s11, s22 = sc.s22, s12 = sc.s12, tau = sc.tau, dsk = sc.dsk, dek = sc.dek, dii = sc.dii;
                    Dd[j].a = s11 + s11c + a * a * tau + binv * binv * dO;
                    Dd[j].c = s22 + s22c + hf * hf * (tau + dO);
                    Dd[j].b = a * hf * tau - hf * binv * dO - s12 - s12c;
                    // ---- predictor right-hand side (complementarity targets 0)
                    const double hg = r.rdg + q.zg, he = r.rde + q.ze, hp = r.rdp + q.zp, hq = r.rdq + q.zq;
                    const double hs = has_s ? r.rds + q.zs : 0.0;
                    const double hi = r.rdi + q.zi + (-q.wi * r.rui) * rsi - q.wi;
                    const double ho = r.rdo + q.zo + (-q.wo * r.ruo) * rso - q.wo;
                    const double w3 = r.rp3 + dp * hp;
                    const double ph1 = s11 * hs - s12 * he - dsk * w3;
                    const double ph2 = s22 * he - s12 * hs - dek * w3;
                    const double w4 = r.rp4 + dg * hg + dq * hq;
                    const double psi = tau * hi - dii * w4;
                    const double doh = dO * ho;
                    f1[j] = r.rp1 + ph1 - ph1c - a * psi + binv * doh;
                    f2[j] = r.rp2 + ph2 - ph2c - hf * psi - hf * doh;
                    SMF(A_DS, j) = ds; SMF(A_DE, j) = de; SMF(A_DP, j) = dp; SMF(A_KAP, j) = kap;
                    SMF(A_DG, j) = dg; SMF(A_DI, j) = di; SMF(A_DQ, j) = dq; SMF(A_IOT, j) = iot; SMF(A_DO, j) = dO;
                    SMF(A_RX + 0, j) = rxg; SMF(A_RX + 1, j) = rxi; SMF(A_RX + 2, j) = rxo; SMF(A_RX + 3, j) = rxs;
                    SMF(A_RX + 4, j) = rxe; SMF(A_RX + 5, j) = rxp; SMF(A_RX + 6, j) = rxq;
                    // the coupling with the next period exists only while the state of charge does
                    s11c = has_s ? s11 : 0.0; s22c = has_s ? s22 : 0.0; s12c = has_s ? s12 : 0.0; ph1c = ph1; ph2c = ph2;
                } else {
                    Dd[j].a = 1.0; Dd[j].b = 0.0; Dd[j].c = 1.0;
                    f1[j] = 0.0; f2[j] = 0.0;
#pragma unroll
                    for (int k = 0; k < 16; ++k) SMF(k, j) = 0.0;          // scaling values and 1/x
                    s11c = s22c = s12c = ph1c = ph2c = 0.0;
                }
            }
            // contributions of the left neighbour's last period to this lane's first period
            s11l = gup1<L>(s11c, gl); s22l = gup1<L>(s22c, gl); s12l = gup1<L>(s12c, gl);
            const double ph1l = gup1<L>(ph1c, gl), ph2l = gup1<L>(ph2c, gl);
            if (ACT(0)) {
                Dd[0].a += s11l; Dd[0].c += s22l; Dd[0].b -= s12l;
                f1[0] -= ph1l; f2[0] -= ph2l;
            }
        }
        const double mu = mu_keep;

        if (DSP_S2_SYNCMASK & 1) cta_sync<CTA_SYNC>();
        // =========================================================================================== factorisation + predictor solve
        // local elimination of periods 0..P-2 (forward part of the solve rides along), separator chain across the lanes
        Mat2 Wc;                       // coupling of the lane's separator (last period) with the left neighbour's separator
        Sym2 Asep;                     // separator diagonal block
        double g1, g2;                 // separator right-hand side
        Sym2 Ainv;                     // inverse of the eliminated separator block
        Mat2 Mout, Mout2;              // multipliers of the separator chain (root: two)
        Mat2 Cin;                      // coupling towards the root
        int fo, bo, fsrc, bsrc;
        const bool is_root = (gl == r_root);
        {
            // E: coupling block (period j, left separator); starts as the plain coupling of period 0 with the left lane's last period
            Mat2 E;
            {
                const bool cpl = (gl > 0) && ACT(0);       // (the left lane's last period has its s column whenever period 0 here is active)
                E.a = cpl ? -s11l : 0.0; E.d = cpl ? -s22l : 0.0; E.b = cpl ? s12l : 0.0; E.c = E.b;
            }
            Sym2 dS; dS.a = dS.b = dS.c = 0.0;            // update of the LEFT separator's diagonal block
            double dg1 = 0.0, dg2 = 0.0;                  // update of the LEFT separator's right-hand side
#pragma unroll
            for (int j = 0; j < P - 1; ++j) {
                const Sym2 Kj = inv_spd(Dd[j]);
                Mat2 C;                                   // coupling (j, j+1): symmetric
                const bool cn = HAS_S(j);
                const Scal sc = LOAD_SCAL(j);
                C.a = cn ? -sc.s11 : 0.0; C.d = cn ? -sc.s22 : 0.0; C.b = cn ? sc.s12 : 0.0; C.c = C.b;
                const Mat2 G = mul_ms(C, Kj);             // C K
                sub_xct(Dd[j + 1], G, C);                 // D_{j+1} -= G C'
                const Mat2 Hm = mul_ms(transp(E), Kj);    // E' K
                sub_xct(dS, Hm, transp(E));               // dS -= H E
                dg1 -= fma(Hm.a, f1[j], Hm.b * f2[j]); dg2 -= fma(Hm.c, f1[j], Hm.d * f2[j]);
                f1[j + 1] -= fma(G.a, f1[j], G.b * f2[j]); f2[j + 1] -= fma(G.c, f1[j], G.d * f2[j]);
                const Mat2 GE = mul_mm(G, E);
                SMI(I_K + 0, j) = Kj.a; SMI(I_K + 1, j) = Kj.b; SMI(I_K + 2, j) = Kj.c;
                SMI(I_G + 0, j) = G.a; SMI(I_G + 1, j) = G.b; SMI(I_G + 2, j) = G.c; SMI(I_G + 3, j) = G.d;
                SMI(I_H + 0, j) = Hm.a; SMI(I_H + 1, j) = Hm.b; SMI(I_H + 2, j) = Hm.c; SMI(I_H + 3, j) = Hm.d;
                SMF(A_F + 0, j) = f1[j]; SMF(A_F + 1, j) = f2[j];
                E.a = -GE.a; E.b = -GE.b; E.c = -GE.c; E.d = -GE.d;
            }
            Wc = E;
            Asep = Dd[P - 1];
            g1 = f1[P - 1]; g2 = f2[P - 1];
            // the right neighbour's eliminations updated this lane's separator
            Asep.a += gdown1<L>(dS.a, gl); Asep.b += gdown1<L>(dS.b, gl); Asep.c += gdown1<L>(dS.c, gl);
            g1 += gdown1<L>(dg1, gl); g2 += gdown1<L>(dg2, gl);
            Mat2 Wn;                                       // the right neighbour's coupling with this separator
            Wn.a = gdown1<L>(Wc.a, gl); Wn.b = gdown1<L>(Wc.b, gl); Wn.c = gdown1<L>(Wc.c, gl); Wn.d = gdown1<L>(Wc.d, gl);
            // ---- twisted block LDL' over the L separators: chains 0 -> root and L-1 -> root
            const bool low = gl < r_root;
            fsrc = low ? (gl > 0 ? gl - 1 : 0) : (gl < L - 1 ? gl + 1 : L - 1);
            bsrc = low ? gl + 1 : gl - 1;
            fo = is_root ? (1 << 20) : (low ? gl : L - 1 - gl);
            bo = is_root ? (1 << 20) : (low ? r_root - gl : gl - r_root);
            const Mat2 Cout = low ? Wc : transp(Wn);       // block (this, outer neighbour)
            Cin = low ? transp(Wn) : Wc;                   // block (this, inner neighbour)
            Mout.a = Mout.b = Mout.c = Mout.d = 0.0; Mout2 = Mout;
#pragma unroll
            for (int k = 1; k <= kmax; ++k) {
                Sym2 R;
                R.a = gfrom<L>(Asep.a, fsrc); R.b = gfrom<L>(Asep.b, fsrc); R.c = gfrom<L>(Asep.c, fsrc);
                const double q1 = gfrom<L>(g1, fsrc), q2 = gfrom<L>(g2, fsrc);
                if (fo == k) {
                    const Mat2 X = mul_ms(Cout, inv_spd(R));
                    sub_xct(Asep, X, Cout);
                    g1 -= fma(X.a, q1, X.b * q2); g2 -= fma(X.c, q1, X.d * q2);
                    Mout = X;
                }
            }
            {
                constexpr int la = r_root > 0 ? r_root - 1 : 0, lb = r_root + 1 < L ? r_root + 1 : L - 1;
                Sym2 Ra, Rb;
                Ra.a = gfrom<L>(Asep.a, la); Ra.b = gfrom<L>(Asep.b, la); Ra.c = gfrom<L>(Asep.c, la);
                Rb.a = gfrom<L>(Asep.a, lb); Rb.b = gfrom<L>(Asep.b, lb); Rb.c = gfrom<L>(Asep.c, lb);
                const double a1 = gfrom<L>(g1, la), a2 = gfrom<L>(g2, la), b1 = gfrom<L>(g1, lb), b2 = gfrom<L>(g2, lb);
                if (is_root) {
                    if (r_root >= 1) {
                        Mout = mul_ms(Wc, inv_spd(Ra)); sub_xct(Asep, Mout, Wc);
                        g1 -= fma(Mout.a, a1, Mout.b * a2); g2 -= fma(Mout.c, a1, Mout.d * a2);
                    }
                    if (r_root + 1 <= L - 1) {
                        const Mat2 Wt = transp(Wn);
                        Mout2 = mul_ms(Wt, inv_spd(Rb)); sub_xct(Asep, Mout2, Wt);
                        g1 -= fma(Mout2.a, b1, Mout2.b * b2); g2 -= fma(Mout2.c, b1, Mout2.d * b2);
                    }
                }
            }
            Ainv = inv_spd(Asep);
        }
        // backward half over the separators: on entry (g1, g2) is the forward-eliminated right-hand side, on exit the solution
#define SEP_BACK()                                                                                    \
        {                                                                                             \
            double u1 = 0.0, u2 = 0.0;                                                                \
            if (is_root) { u1 = fma(Ainv.a, g1, Ainv.b * g2); u2 = fma(Ainv.b, g1, Ainv.c * g2); }    \
            _Pragma("unroll")                                                                         \
            for (int s = 1; s <= smax; ++s) {                                                         \
                const double r1 = gfrom<L>(u1, bsrc), r2 = gfrom<L>(u2, bsrc);                        \
                if (bo == s) {                                                                        \
                    const double t1 = g1 - fma(Cin.a, r1, Cin.b * r2);                                \
                    const double t2 = g2 - fma(Cin.c, r1, Cin.d * r2);                                \
                    u1 = fma(Ainv.a, t1, Ainv.b * t2); u2 = fma(Ainv.b, t1, Ainv.c * t2);             \
                }                                                                                     \
            }                                                                                         \
            g1 = u1; g2 = u2;                                                                         \
        }
        // local back substitution: u_j = K g_j - G' u_{j+1} - H' u_left
#define LOCAL_BACK(dy1, dy2)                                                                          \
        {                                                                                             \
            const double ul1 = gup1<L>(g1, gl), ul2 = gup1<L>(g2, gl);                                \
            dy1[P - 1] = g1; dy2[P - 1] = g2;                                                         \
            _Pragma("unroll")                                                                         \
            for (int j = P - 2; j >= 0; --j) {                                                        \
                const double ka = SMI(I_K + 0, j), kb = SMI(I_K + 1, j), kc_ = SMI(I_K + 2, j);       \
                const double ga = SMI(I_G + 0, j), gb = SMI(I_G + 1, j), gc = SMI(I_G + 2, j), gd = SMI(I_G + 3, j); \
                const double ha = SMI(I_H + 0, j), hb = SMI(I_H + 1, j), hc = SMI(I_H + 2, j), hd = SMI(I_H + 3, j); \
                const double e1 = SMF(A_F + 0, j), e2 = SMF(A_F + 1, j);                              \
                dy1[j] = fma(ka, e1, kb * e2) - fma(ga, dy1[j + 1], gc * dy2[j + 1]) - fma(ha, ul1, hc * ul2);  \
                dy2[j] = fma(kb, e1, kc_ * e2) - fma(gb, dy1[j + 1], gd * dy2[j + 1]) - fma(hb, ul1, hd * ul2); \
            }                                                                                         \
        }
        double dy1[P], dy2[P];
        SEP_BACK();
        LOCAL_BACK(dy1, dy2);

        if (DSP_S2_SYNCMASK & 2) cta_sync<CTA_SYNC>();
        // =========================================================================================== pass 2: predictor direction
        // recovery of dx, dz; step lengths; sums for the centring parameter; second-order products
        double smu;
        {
            const double dy1_right = gdown1<L>(dy1[0], gl), dy2_right = gdown1<L>(dy2[0], gl);
            double ip = 0.0, id = 0.0, S1 = 0.0, S3 = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const Per &q = pr[j];
                const bool act = ACT(j), has_s = HAS_S(j);
                if (act) {
                    Res r;
                    residuals(q, SMF(A_C, j), SMF(A_B4, j), XSP(j), XEP(j), Y1N(j), Y2N(j), b3, u, K, act, has_s, r);
                    const double rxg = SMF(A_RX + 0, j), rxi = SMF(A_RX + 1, j), rxo = SMF(A_RX + 2, j), rxs = SMF(A_RX + 3, j);
                    const double rxe = SMF(A_RX + 4, j), rxp = SMF(A_RX + 5, j), rxq = SMF(A_RX + 6, j);
                    const double rsi = frcp(q.si), rso = frcp(q.so);
                    const Scal sc = LOAD_SCAL(j);
                    const double s11 = sc.s11, s22 = sc.s22, s12 = sc.s12, dsk = sc.dsk, dek = sc.dek;
                    const double kap = sc.kap, tau = sc.tau, dii = sc.dii, iot = sc.iot, dO = sc.dO;
                    const double dg = sc.dg, dq = sc.dq, dp = sc.dp, di = sc.di;
                    const double hg = r.rdg + q.zg, he = r.rde + q.ze, hp = r.rdp + q.zp, hq = r.rdq + q.zq;
                    const double hs = has_s ? r.rds + q.zs : 0.0;
                    const double hi = r.rdi + q.zi + (-q.wi * r.rui) * rsi - q.wi;
                    const double ho = r.rdo + q.zo + (-q.wo * r.ruo) * rso - q.wo;
                    const double w3 = r.rp3 + dp * hp, w4 = r.rp4 + dg * hg + dq * hq;
                    const double d1n = (j == P - 1) ? dy1_right : dy1[j < P - 1 ? j + 1 : 0];
                    const double d2n = (j == P - 1) ? dy2_right : dy2[j < P - 1 ? j + 1 : 0];
                    const double e1 = dy1[j] - d1n - hs, e2 = dy2[j] - d2n - he;
                    const double v = a * dy1[j] + hf * dy2[j];
                    const double dxs = has_s ? s11 * e1 - s12 * e2 + dsk * w3 : 0.0;
                    const double dxe = s22 * e2 - s12 * e1 + dek * w3;
                    const double dxi = -tau * (v + hi) + dii * w4;
                    const double dxo = dO * (binv * dy1[j] - hf * dy2[j] - ho);
                    const double dxg = dg * iot * (r.rp4 + di * (hi - hg + v) + dq * (hq - hg));
                    const double dy3 = kap * w3 - dsk * e1 - dek * e2;
                    const double dy4 = iot * w4 + dii * (hi + v);
                    const double dxp = dp * (dy3 - hp), dxq = dq * (dy4 - hq);
                    // dz = -z - z dx / x ;  ds = ru - dx ;  dw = -w - w ds / s
                    const double tg = dxg * rxg, ti = dxi * rxi, to = dxo * rxo, ts = dxs * rxs, te = dxe * rxe, tp = dxp * rxp, tq = dxq * rxq;
                    const double dzg = -q.zg - q.zg * tg, dzi = -q.zi - q.zi * ti, dzo = -q.zo - q.zo * to, dzs = has_s ? -q.zs - q.zs * ts : 0.0;
                    const double dze = -q.ze - q.ze * te, dzp = -q.zp - q.zp * tp, dzq = -q.zq - q.zq * tq;
                    const double dsi = r.rui - dxi, dso = r.ruo - dxo;
                    const double tsi = dsi * rsi, tso = dso * rso;
                    const double dwi = -q.wi - q.wi * tsi, dwo = -q.wo - q.wo * tso;
                    // 1/alpha: primal max(-dx/x), dual max(-dz/z) = max(1 + dx/x) for the affine direction (no 1/z needed)
                    ip = dmax(ip, dmax(dmax(dmax(-tg, -ti), dmax(-to, -ts)), dmax(dmax(-te, -tp), dmax(-tq, dmax(-tsi, -tso)))));
                    id = dmax(id, dmax(dmax(dmax(1.0 + tg, 1.0 + ti), dmax(1.0 + to, has_s ? 1.0 + ts : 0.0)),
                                       dmax(dmax(1.0 + te, 1.0 + tp), dmax(1.0 + tq, dmax(1.0 + tsi, 1.0 + tso)))));
                    S1 += q.zg * dxg + q.zi * dxi + q.zo * dxo + q.zs * dxs + q.ze * dxe + q.zp * dxp + q.zq * dxq + q.wi * dsi + q.wo * dso;
                    const double cg = dxg * dzg, ci = dxi * dzi, co = dxo * dzo, cs = dxs * dzs, ce = dxe * dze, cpp = dxp * dzp, cq = dxq * dzq;
                    const double csi = dsi * dwi, cso = dso * dwo;
                    S3 += cg + ci + co + cs + ce + cpp + cq + csi + cso;
                    SMF(A_PR + 0, j) = cg; SMF(A_PR + 1, j) = ci; SMF(A_PR + 2, j) = co; SMF(A_PR + 3, j) = cs; SMF(A_PR + 4, j) = ce;
                    SMF(A_PR + 5, j) = cpp; SMF(A_PR + 6, j) = cq; SMF(A_PR + 7, j) = csi; SMF(A_PR + 8, j) = cso;
                }
            }
            ip = gmax<L>(ip); id = gmax<L>(id);
            S1 = gsum<L>(S1); S3 = gsum<L>(S3);
            const double ap = ip > 1.0 ? 1.0 / ip : 1.0, ad = id > 1.0 ? 1.0 / id : 1.0;
            // sum (x + ap dx)(z + ad dz) with  sum(x dz + z dx) = -sum(x z)  for the affine direction
            const double musum = mu * ntot;
            const double S2 = -musum - S1;
            const double mua = (musum + ap * S1 + ad * S2 + ap * ad * S3) / ntot;
            const double sg = mua / mu;
            smu = sg * sg * sg * mu;
        }

        if (DSP_S2_SYNCMASK & 4) cta_sync<CTA_SYNC>();
        // =========================================================================================== pass 3: corrector right-hand side
        {
            double ph1c = 0.0, ph2c = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const Per &q = pr[j];
                const bool act = ACT(j), has_s = HAS_S(j);
                if (act) {
                    Res r;
                    residuals(q, SMF(A_C, j), SMF(A_B4, j), XSP(j), XEP(j), Y1N(j), Y2N(j), b3, u, K, act, has_s, r);
                    const double rsi = frcp(q.si), rso = frcp(q.so);
                    const Scal sc = LOAD_SCAL(j);
                    const double s11 = sc.s11, s22 = sc.s22, s12 = sc.s12, dsk = sc.dsk, dek = sc.dek;
                    const double tau = sc.tau, dii = sc.dii, dO = sc.dO;
                    const double dg = sc.dg, dq = sc.dq, dp = sc.dp;
                    const double hg = r.rdg + q.zg - (smu - SMF(A_PR + 0, j)) * SMF(A_RX + 0, j);
                    const double hs = has_s ? r.rds + q.zs - (smu - SMF(A_PR + 3, j)) * SMF(A_RX + 3, j) : 0.0;
                    const double he = r.rde + q.ze - (smu - SMF(A_PR + 4, j)) * SMF(A_RX + 4, j);
                    const double hp = r.rdp + q.zp - (smu - SMF(A_PR + 5, j)) * SMF(A_RX + 5, j);
                    const double hq = r.rdq + q.zq - (smu - SMF(A_PR + 6, j)) * SMF(A_RX + 6, j);
                    const double asi = -q.wi * r.rui + smu - SMF(A_PR + 7, j), aso = -q.wo * r.ruo + smu - SMF(A_PR + 8, j);
                    const double hi = r.rdi + q.zi - (smu - SMF(A_PR + 1, j)) * SMF(A_RX + 1, j) + asi * rsi - q.wi;
                    const double ho = r.rdo + q.zo - (smu - SMF(A_PR + 2, j)) * SMF(A_RX + 2, j) + aso * rso - q.wo;
                    const double w3 = r.rp3 + dp * hp;
                    const double ph1 = s11 * hs - s12 * he - dsk * w3;
                    const double ph2 = s22 * he - s12 * hs - dek * w3;
                    const double w4 = r.rp4 + dg * hg + dq * hq;
                    const double psi = tau * hi - dii * w4;
                    const double doh = dO * ho;
                    f1[j] = r.rp1 + ph1 - ph1c - a * psi + binv * doh;
                    f2[j] = r.rp2 + ph2 - ph2c - hf * psi - hf * doh;
                    ph1c = ph1; ph2c = ph2;
                } else {
                    f1[j] = 0.0; f2[j] = 0.0; ph1c = 0.0; ph2c = 0.0;
                }
            }
            const double ph1l = gup1<L>(ph1c, gl), ph2l = gup1<L>(ph2c, gl);
            if (ACT(0)) { f1[0] -= ph1l; f2[0] -= ph2l; }
            // forward elimination with the stored factor
            double dg1 = 0.0, dg2 = 0.0;
#pragma unroll
            for (int j = 0; j < P - 1; ++j) {
                const double ga = SMI(I_G + 0, j), gb = SMI(I_G + 1, j), gc = SMI(I_G + 2, j), gd = SMI(I_G + 3, j);
                const double ha = SMI(I_H + 0, j), hb = SMI(I_H + 1, j), hc = SMI(I_H + 2, j), hd = SMI(I_H + 3, j);
                dg1 -= fma(ha, f1[j], hb * f2[j]); dg2 -= fma(hc, f1[j], hd * f2[j]);
                f1[j + 1] -= fma(ga, f1[j], gb * f2[j]); f2[j + 1] -= fma(gc, f1[j], gd * f2[j]);
                SMF(A_F + 0, j) = f1[j]; SMF(A_F + 1, j) = f2[j];
            }
            g1 = f1[P - 1] + gdown1<L>(dg1, gl); g2 = f2[P - 1] + gdown1<L>(dg2, gl);
#pragma unroll
            for (int k = 1; k <= kmax; ++k) {
                const double q1 = gfrom<L>(g1, fsrc), q2 = gfrom<L>(g2, fsrc);
                if (fo == k) { g1 -= fma(Mout.a, q1, Mout.b * q2); g2 -= fma(Mout.c, q1, Mout.d * q2); }
            }
            {
                constexpr int la = r_root > 0 ? r_root - 1 : 0, lb = r_root + 1 < L ? r_root + 1 : L - 1;
                const double a1 = gfrom<L>(g1, la), a2 = gfrom<L>(g2, la), b1 = gfrom<L>(g1, lb), b2 = gfrom<L>(g2, lb);
                if (is_root) {
                    if (r_root >= 1) { g1 -= fma(Mout.a, a1, Mout.b * a2); g2 -= fma(Mout.c, a1, Mout.d * a2); }
                    if (r_root + 1 <= L - 1) { g1 -= fma(Mout2.a, b1, Mout2.b * b2); g2 -= fma(Mout2.c, b1, Mout2.d * b2); }
                }
            }
        }
        SEP_BACK();
        LOCAL_BACK(dy1, dy2);

        if (DSP_S2_SYNCMASK & 8) cta_sync<CTA_SYNC>();
        // =========================================================================================== pass 4: corrector direction
        double ap, ad;
        {
            const double dy1_right = gdown1<L>(dy1[0], gl), dy2_right = gdown1<L>(dy2[0], gl);
            double ip = 0.0, id = 0.0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const Per &q = pr[j];
                const bool act = ACT(j), has_s = HAS_S(j);
                if (act) {
                    Res r;
                    residuals(q, SMF(A_C, j), SMF(A_B4, j), XSP(j), XEP(j), Y1N(j), Y2N(j), b3, u, K, act, has_s, r);
                    const double rxg = SMF(A_RX + 0, j), rxi = SMF(A_RX + 1, j), rxo = SMF(A_RX + 2, j), rxs = SMF(A_RX + 3, j);
                    const double rxe = SMF(A_RX + 4, j), rxp = SMF(A_RX + 5, j), rxq = SMF(A_RX + 6, j);
                    const double rsi = frcp(q.si), rso = frcp(q.so);
                    const Scal sc = LOAD_SCAL(j);
                    const double s11 = sc.s11, s22 = sc.s22, s12 = sc.s12, dsk = sc.dsk, dek = sc.dek;
                    const double kap = sc.kap, tau = sc.tau, dii = sc.dii, iot = sc.iot, dO = sc.dO;
                    const double dg = sc.dg, dq = sc.dq, dp = sc.dp, di = sc.di;
                    const double ag = smu - SMF(A_PR + 0, j), ai = smu - SMF(A_PR + 1, j), ao = smu - SMF(A_PR + 2, j), as_ = smu - SMF(A_PR + 3, j);
                    const double ae = smu - SMF(A_PR + 4, j), app = smu - SMF(A_PR + 5, j), aq = smu - SMF(A_PR + 6, j);
                    const double asi_ = smu - SMF(A_PR + 7, j), aso_ = smu - SMF(A_PR + 8, j);
                    const double hg = r.rdg + q.zg - ag * rxg, he = r.rde + q.ze - ae * rxe, hp = r.rdp + q.zp - app * rxp, hq = r.rdq + q.zq - aq * rxq;
                    const double hs = has_s ? r.rds + q.zs - as_ * rxs : 0.0;
                    const double hi = r.rdi + q.zi - ai * rxi + (-q.wi * r.rui + asi_) * rsi - q.wi;
                    const double ho = r.rdo + q.zo - ao * rxo + (-q.wo * r.ruo + aso_) * rso - q.wo;
                    const double w3 = r.rp3 + dp * hp, w4 = r.rp4 + dg * hg + dq * hq;
                    const double d1n = (j == P - 1) ? dy1_right : dy1[j < P - 1 ? j + 1 : 0];
                    const double d2n = (j == P - 1) ? dy2_right : dy2[j < P - 1 ? j + 1 : 0];
                    const double e1 = dy1[j] - d1n - hs, e2 = dy2[j] - d2n - he;
                    const double v = a * dy1[j] + hf * dy2[j];
                    const double dxs = has_s ? s11 * e1 - s12 * e2 + dsk * w3 : 0.0;
                    const double dxe = s22 * e2 - s12 * e1 + dek * w3;
                    const double dxi = -tau * (v + hi) + dii * w4;
                    const double dxo = dO * (binv * dy1[j] - hf * dy2[j] - ho);
                    const double dxg = dg * iot * (r.rp4 + di * (hi - hg + v) + dq * (hq - hg));
                    const double dy3 = kap * w3 - dsk * e1 - dek * e2;
                    const double dy4 = iot * w4 + dii * (hi + v);
                    const double dxp = dp * (dy3 - hp), dxq = dq * (dy4 - hq);
                    const double dzg = ag * rxg - q.zg - q.zg * dxg * rxg, dzi = ai * rxi - q.zi - q.zi * dxi * rxi;
                    const double dzo = ao * rxo - q.zo - q.zo * dxo * rxo, dzs = has_s ? as_ * rxs - q.zs - q.zs * dxs * rxs : 0.0;
                    const double dze = ae * rxe - q.ze - q.ze * dxe * rxe, dzp = app * rxp - q.zp - q.zp * dxp * rxp;
                    const double dzq = aq * rxq - q.zq - q.zq * dxq * rxq;
                    const double dsi = r.rui - dxi, dso = r.ruo - dxo;
                    const double dwi = asi_ * rsi - q.wi - q.wi * dsi * rsi, dwo = aso_ * rso - q.wo - q.wo * dso * rso;
                    ip = dmax(ip, dmax(dmax(dmax(-dxg * rxg, -dxi * rxi), dmax(-dxo * rxo, -dxs * rxs)),
                                       dmax(dmax(-dxe * rxe, -dxp * rxp), dmax(-dxq * rxq, dmax(-dsi * rsi, -dso * rso)))));
                    id = dmax(id, dmax(dmax(dmax(-dzg * frcp(q.zg), -dzi * frcp(q.zi)), dmax(-dzo * frcp(q.zo), has_s ? -dzs * frcp(q.zs) : 0.0)),
                                       dmax(dmax(-dze * frcp(q.ze), -dzp * frcp(q.zp)),
                                            dmax(-dzq * frcp(q.zq), dmax(-dwi * frcp(q.wi), -dwo * frcp(q.wo))))));
                    // the scaling blocks of this period are dead now: park the direction in their slots
                    SMF(A_DX + 0, j) = dxg; SMF(A_DX + 1, j) = dxi; SMF(A_DX + 2, j) = dxo; SMF(A_DX + 3, j) = dxs; SMF(A_DX + 4, j) = dxe;
                    SMF(A_DX + 5, j) = dxp; SMF(A_DX + 6, j) = dxq; SMF(A_DY3, j) = dy3; SMF(A_DY4, j) = dy4;
                }
            }
            ip = gmax<L>(ip); id = gmax<L>(id);
            ap = step_frac < ip ? step_frac / ip : 1.0;      // min(1, step_frac / ip)
            ad = step_frac < id ? step_frac / id : 1.0;
        }

        if (DSP_S2_SYNCMASK & 16) cta_sync<CTA_SYNC>();
        // =========================================================================================== pass 5: step
#pragma unroll
        for (int j = 0; j < P; ++j) {
            Per &q = pr[j];
            if (ACT(j)) {
                const bool has_s = HAS_S(j);
                const double rxg = SMF(A_RX + 0, j), rxi = SMF(A_RX + 1, j), rxo = SMF(A_RX + 2, j), rxs = SMF(A_RX + 3, j);
                const double rxe = SMF(A_RX + 4, j), rxp = SMF(A_RX + 5, j), rxq = SMF(A_RX + 6, j);
                const double rsi = frcp(q.si), rso = frcp(q.so);
                const double dxg = SMF(A_DX + 0, j), dxi = SMF(A_DX + 1, j), dxo = SMF(A_DX + 2, j), dxs = SMF(A_DX + 3, j);
                const double dxe = SMF(A_DX + 4, j), dxp = SMF(A_DX + 5, j), dxq = SMF(A_DX + 6, j);
                const double dsi = (u - q.xi - q.si) - dxi, dso = (u - q.xo - q.so) - dxo;
                const double dzg = (smu - SMF(A_PR + 0, j)) * rxg - q.zg - q.zg * dxg * rxg;
                const double dzi = (smu - SMF(A_PR + 1, j)) * rxi - q.zi - q.zi * dxi * rxi;
                const double dzo = (smu - SMF(A_PR + 2, j)) * rxo - q.zo - q.zo * dxo * rxo;
                const double dzs = (smu - SMF(A_PR + 3, j)) * rxs - q.zs - q.zs * dxs * rxs;
                const double dze = (smu - SMF(A_PR + 4, j)) * rxe - q.ze - q.ze * dxe * rxe;
                const double dzp = (smu - SMF(A_PR + 5, j)) * rxp - q.zp - q.zp * dxp * rxp;
                const double dzq = (smu - SMF(A_PR + 6, j)) * rxq - q.zq - q.zq * dxq * rxq;
                const double dwi = (smu - SMF(A_PR + 7, j)) * rsi - q.wi - q.wi * dsi * rsi;
                const double dwo = (smu - SMF(A_PR + 8, j)) * rso - q.wo - q.wo * dso * rso;
                q.xg += ap * dxg; q.xi += ap * dxi; q.xo += ap * dxo; q.xe += ap * dxe; q.xp += ap * dxp; q.xq += ap * dxq;
                q.zg += ad * dzg; q.zi += ad * dzi; q.zo += ad * dzo; q.ze += ad * dze; q.zp += ad * dzp; q.zq += ad * dzq;
                if (has_s) { q.xs += ap * dxs; q.zs += ad * dzs; }
                q.si += ap * dsi; q.so += ap * dso; q.wi += ad * dwi; q.wo += ad * dwo;
                q.y1 += ad * dy1[j]; q.y2 += ad * dy2[j]; q.y3 += ad * SMF(A_DY3, j); q.y4 += ad * SMF(A_DY4, j);
            }
        }
        ++it;
    }
#undef SMF
#undef LOAD_SCAL
#undef SMI
#undef XSP
#undef XEP
#undef Y1N
#undef Y2N
#undef ACT
#undef HAS_S
#undef NEIGHBOURS
#undef SEP_BACK
#undef LOCAL_BACK
}

}  // namespace stage2
