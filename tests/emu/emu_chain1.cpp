// emu_chain1.cpp -- TEST INFRASTRUCTURE: the CUDA warp body of dispatches_b200/csrc/dsp_stage_chain1.cuh compiled with g++ on the
// lock-step SIMT emulator (simt_emu.h) for tests/test_chain1_emulation.py.  Never linked into the product library.
#include "simt_emu.h"

#include "../../dispatches_b200/csrc/dsp_stage_chain1.cuh"

#include <vector>

template <int L, int P, int NF>
static void run(const chain1::Params &Q, int warps) {
    const int nd = chain1::Smem<NF, P>::doubles_per_warp;
    for (int w = 0; w < warps; ++w) {
        std::vector<double> smem(nd, 0.0);
        emu::run_warp([&](int lane) { chain1::warp_body<L, P, NF, false>(Q, smem.data(), lane); });
    }
}

extern "C" int emu_chain1_solve(int L, int P, int NF, int warps, chain1::Params *Qin) {
    unsigned long long ticket = 0;
    chain1::Params Q = *Qin;
    Q.ticket = &ticket;
    if (Q.T > L * P) return -1;
#define CASE(l, p, nf) if (L == l && P == p && NF == nf) { run<l, p, nf>(Q, warps); return 0; }
    CASE(16, 3, 2) CASE(8, 3, 2) CASE(32, 3, 2) CASE(16, 3, 3) CASE(8, 3, 3) CASE(32, 3, 3) CASE(4, 3, 2) CASE(4, 3, 3)
#undef CASE
    return -2;
}
