// emu_stage2.cpp -- TEST INFRASTRUCTURE: compiles the CUDA warp body of dispatches_b200/csrc/dsp_stage2.cuh with g++ on the
// lock-step SIMT emulator (simt_emu.h) and exposes it through a C entry point for tests/test_stage2_emulation.py.
// Never linked into the product library.
#include "simt_emu.h"

#include "../../dispatches_b200/csrc/dsp_stage2.cuh"
#include "../../dispatches_b200/csrc/dsp_stage2_long.cuh"

#include <vector>

template <int L, int P>
static void run(const stage2::Params &Q, int warps) {
    const int nd = stage2::smem_doubles_per_warp<P>();
    for (int w = 0; w < warps; ++w) {
        std::vector<double> smem(nd, 0.0);
        emu::run_warp([&](int lane) { stage2::warp_body<L, P>(Q, smem.data(), lane); });
    }
}

extern "C" int emu_stage2_solve(int L, int P, int warps, long long N, const double *cparams, const double *rparams, long long rstride,
                                int Pc, int Pr, const double *omap, const double *ocmap, double o0, double tol, double feas_tol,
                                double step_frac, double reg, int max_iter, double *obj, double *x_out, double *y_out, int *status,
                                int *iters, int n, int m, int T, double a, double binv, double hf, double dl, double dur, double krev,
                                int wcf_off, int p_off, const int *col_idx, const int *row_idx) {
    unsigned long long ticket = 0;
    stage2::Params Q;
    Q.N = N; Q.cparams = cparams; Q.rparams = rparams; Q.rstride = rstride; Q.Pc = Pc; Q.Pr = Pr; Q.omap = omap; Q.ocmap = ocmap;
    Q.o0 = o0; Q.tol = tol; Q.feas_tol = feas_tol; Q.step_frac = step_frac; Q.reg = reg; Q.max_iter = max_iter;
    Q.obj = obj; Q.x_out = x_out; Q.y_out = y_out; Q.status = status; Q.iters = iters; Q.n = n; Q.m = m; Q.ticket = &ticket;
    Q.T = T; Q.a = a; Q.binv = binv; Q.hf = hf; Q.dl = dl; Q.dur = dur; Q.krev = krev; Q.wcf_off = wcf_off; Q.p_off = p_off;
    Q.col_idx = col_idx; Q.row_idx = row_idx; Q.ahead = 0;
    if (T > L * P) return -1;
#define CASE(l, p) if (L == l && P == p) { run<l, p>(Q, warps); return 0; }
    CASE(8, 3) CASE(4, 3) CASE(16, 2) CASE(32, 1) CASE(16, 3) CASE(32, 3) CASE(8, 2) CASE(2, 3) CASE(32, 2)
#undef CASE
    return -2;
}


// the long-horizon variant (dsp_stage2_long.cuh): one emulated warp per LP at a time, workspace on the heap
extern "C" int emu_stage2_long_solve(int warps, long long N, const double *cparams, const double *rparams, long long rstride,
                                     int Pc, int Pr, const double *omap, const double *ocmap, double o0, double tol, double feas_tol,
                                     double step_frac, double reg, int max_iter, double *obj, double *x_out, double *y_out, int *status,
                                     int *iters, int n, int m, int T, double a, double binv, double hf, double dl, double dur, double krev,
                                     int wcf_off, int p_off, const int *col_idx, const int *row_idx) {
    unsigned long long ticket = 0;
    stage2long::LongParams LQ;
    stage2::Params &Q = LQ.q;
    Q.N = N; Q.cparams = cparams; Q.rparams = rparams; Q.rstride = rstride; Q.Pc = Pc; Q.Pr = Pr; Q.omap = omap; Q.ocmap = ocmap;
    Q.o0 = o0; Q.tol = tol; Q.feas_tol = feas_tol; Q.step_frac = step_frac; Q.reg = reg; Q.max_iter = max_iter;
    Q.obj = obj; Q.x_out = x_out; Q.y_out = y_out; Q.status = status; Q.iters = iters; Q.n = n; Q.m = m; Q.ticket = &ticket;
    Q.T = T; Q.a = a; Q.binv = binv; Q.hf = hf; Q.dl = dl; Q.dur = dur; Q.krev = krev; Q.wcf_off = wcf_off; Q.p_off = p_off;
    Q.col_idx = col_idx; Q.row_idx = row_idx; Q.ahead = 0;
    LQ.P = (T + 31) / 32;
    std::vector<double> ws((size_t)warps * stage2long::NW * LQ.P * 32, 0.0);
    LQ.ws = ws.data();
    for (int w = 0; w < warps; ++w)
        emu::run_warp([&](int lane) { stage2long::warp_body_long(LQ, w, lane); });
    return 0;
}
