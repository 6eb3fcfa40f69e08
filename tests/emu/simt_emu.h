// simt_emu.h -- TEST INFRASTRUCTURE (not product code): a tiny lock-step SIMT emulator so that the warp-level CUDA source
// of the stage kernels (dispatches_b200/csrc/dsp_stage2.cuh) can be compiled with g++ and executed on CPU "lanes".
//
// One OS thread, 32 ucontext coroutines per emulated warp.  Every warp-collective (__shfl_*_sync, __syncwarp, __any_sync,
// __ballot_sync) is a barrier: a lane deposits its operand, yields round-robin until all 32 lanes of the warp have arrived,
// then reads its partner's slot.  This is exactly the semantics the device code relies on (full-mask collectives in
// warp-uniform control flow); a lane that skips a collective deadlocks the emulation, which is the bug it would be on the GPU.
// Shared memory is a plain array owned by the warp; atomicAdd is a plain add (single thread).
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __device__
#define __host__
#define __forceinline__ inline
#define __global__
#define __launch_bounds__(...)
#define __restrict__

namespace emu {

struct Warp {
    ucontext_t main_ctx;
    ucontext_t ctx[32];
    std::vector<char> stacks[32];
    bool done[32];
    int cur = 0;
    // collective state
    int arrived = 0;
    unsigned long long gen = 0;
    double dslot[32];
    unsigned long long uslot[32];
    std::function<void(int)> body;
};

inline Warp *&W() {
    static Warp *w = nullptr;
    return w;
}
inline int lane_id() { return W()->cur; }

inline void yield_next() {
    Warp *w = W();
    int me = w->cur;
    for (int k = 1; k <= 32; ++k) {
        int nx = (me + k) & 31;
        if (!w->done[nx]) {
            if (nx == me) return;
            w->cur = nx;
            swapcontext(&w->ctx[me], &w->ctx[nx]);
            return;
        }
    }
}

// barrier over the 32 lanes of the warp (all lanes must be alive: device code never exits a lane early)
inline void barrier() {
    Warp *w = W();
    const unsigned long long g = w->gen;
    if (++w->arrived == 32) {
        w->arrived = 0;
        w->gen++;
        return;
    }
    while (w->gen == g) yield_next();
}

inline void trampoline() {
    Warp *w = W();
    int me = w->cur;
    w->body(me);
    w->done[me] = true;
    // switch to any live lane, else back to main
    for (int k = 1; k <= 32; ++k) {
        int nx = (me + k) & 31;
        if (!w->done[nx]) {
            w->cur = nx;
            setcontext(&w->ctx[nx]);
        }
    }
    setcontext(&w->main_ctx);
}

inline void run_warp(const std::function<void(int)> &body) {
    Warp w;
    W() = &w;
    w.body = body;
    for (int l = 0; l < 32; ++l) {
        w.done[l] = false;
        w.stacks[l].resize(1 << 20);
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks[l].data();
        w.ctx[l].uc_stack.ss_size = w.stacks[l].size();
        w.ctx[l].uc_link = nullptr;
        makecontext(&w.ctx[l], (void (*)())trampoline, 0);
    }
    w.cur = 0;
    swapcontext(&w.main_ctx, &w.ctx[0]);
    W() = nullptr;
}

}  // namespace emu

// ---- the CUDA builtins the device code uses
inline double __shfl_sync(unsigned, double v, int src, int width = 32) {
    emu::Warp *w = emu::W();
    const int me = w->cur;
    w->dslot[me] = v;
    emu::barrier();
    const int base = me & ~(width - 1);
    const double r = w->dslot[base + (src & (width - 1))];
    emu::barrier();
    return r;
}
inline unsigned long long __shfl_sync(unsigned, unsigned long long v, int src, int width = 32) {
    emu::Warp *w = emu::W();
    const int me = w->cur;
    w->uslot[me] = v;
    emu::barrier();
    const int base = me & ~(width - 1);
    const unsigned long long r = w->uslot[base + (src & (width - 1))];
    emu::barrier();
    return r;
}
inline long long __shfl_sync(unsigned m, long long v, int src, int width = 32) {
    return (long long)__shfl_sync(m, (unsigned long long)v, src, width);
}
inline int __shfl_sync(unsigned m, int v, int src, int width = 32) {
    return (int)__shfl_sync(m, (unsigned long long)(unsigned)v, src, width);
}
inline double __shfl_up_sync(unsigned, double v, unsigned delta, int width = 32) {
    emu::Warp *w = emu::W();
    const int me = w->cur;
    w->dslot[me] = v;
    emu::barrier();
    const int pos = me & (width - 1);
    const double r = (pos >= (int)delta) ? w->dslot[me - delta] : v;
    emu::barrier();
    return r;
}
inline double __shfl_down_sync(unsigned, double v, unsigned delta, int width = 32) {
    emu::Warp *w = emu::W();
    const int me = w->cur;
    w->dslot[me] = v;
    emu::barrier();
    const int pos = me & (width - 1);
    const double r = (pos + (int)delta < width) ? w->dslot[me + delta] : v;
    emu::barrier();
    return r;
}
inline double __shfl_xor_sync(unsigned, double v, int mask, int width = 32) {
    emu::Warp *w = emu::W();
    const int me = w->cur;
    w->dslot[me] = v;
    emu::barrier();
    const int other = me ^ mask;
    const double r = ((other & ~(width - 1)) == (me & ~(width - 1))) ? w->dslot[other] : v;
    emu::barrier();
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier(); }
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Warp *w = emu::W();
    const int me = w->cur;
    w->uslot[me] = pred ? 1ull : 0ull;
    emu::barrier();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(w->uslot[l] << l);
    emu::barrier();
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long o = *p;
    *p = o + v;
    return o;
}
inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, 8);
    return d;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
