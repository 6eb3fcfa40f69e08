// emu_band.cpp -- TEST INFRASTRUCTURE: the warp-level band factorisation / substitution sweeps of dispatches_b200/csrc/dsp_band.cuh
// compiled with g++ on the lock-step SIMT emulator (simt_emu.h) for tests/test_band_emulation.py.  Never linked into the product.
#include "simt_emu.h"

#include "../../dispatches_b200/csrc/dsp_band.cuh"

// Mb: (m + 2W) * (W + 1) doubles, v: m + 2W doubles, both INCLUDING the W padding rows / entries in front and behind
template <int W>
static void run(int m, double *Mb, double *v, int rw) {
    double *M0 = Mb + W * (W + 1), *v0 = v + W;
    emu::run_warp([&](int lane) {
        band::band_factor<W>(M0, m, lane); band::band_solve<W>(M0, v0, m, lane);
    });
}

extern "C" int emu_band_factor_solve(int W, int m, double *Mb, double *v, int rw) {
    switch (W) {
        case 1: run<1>(m, Mb, v, rw); return 0;
        case 2: run<2>(m, Mb, v, rw); return 0;
        case 4: run<4>(m, Mb, v, rw); return 0;
        case 8: run<8>(m, Mb, v, rw); return 0;
        case 16: run<16>(m, Mb, v, rw); return 0;
        case 32: run<32>(m, Mb, v, rw); return 0;
    }
    return -1;
}
