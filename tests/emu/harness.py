"""TEST INFRASTRUCTURE: builds tests/emu/emu_stage2.cpp (the CUDA warp body of dispatches_b200/csrc/dsp_stage2.cuh compiled
with g++ on the lock-step SIMT emulator) and calls it through ctypes.  Lets the CPU suite execute the stage kernel's source."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE / "libemu_stage2.so"
_lib = None


def build(force=False):
    deps = [HERE / "emu_stage2.cpp", HERE / "simt_emu.h", ROOT / "dispatches_b200" / "csrc" / "dsp_stage2.cuh",
            ROOT / "dispatches_b200" / "csrc" / "dsp_stage2_long.cuh"]
    if force or not LIB.exists() or any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps):
        cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", str(LIB), str(HERE / "emu_stage2.cpp")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stderr[-3000:])
    return LIB


def solve(t, cparams, rparams, L, P, warps=1, want_xy=True, tol=1e-9, feas_tol=1e-9, step_frac=0.9995, reg=1e-8, max_iter=60):
    """runs `warps` emulated warps (one after the other, sharing the ticket counter) of stage2::warp_body<L, P>"""
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.emu_stage2_solve.argtypes = ([C.c_int] * 3 + [C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p] + [C.c_double] * 5 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_double] * 6
                                          + [C.c_int] * 2 + [C.c_void_p] * 2)
    st = t.meta["stage_wb"]
    cp = np.ascontiguousarray(np.atleast_2d(cparams), float)
    N = cp.shape[0]
    rp = np.ascontiguousarray(rparams, float)
    rstride = 0 if rp.ndim == 1 else rp.shape[1]
    obj = np.zeros(N); status = np.full(N, -9, np.int32); iters = np.zeros(N, np.int32)
    x = np.zeros((N, t.n)); y = np.zeros((N, t.m))
    ci = np.ascontiguousarray(st["col_idx"], np.int32); ri = np.ascontiguousarray(st["row_idx"], np.int32)
    omap = np.ascontiguousarray(t.omap, float); ocmap = np.ascontiguousarray(t.ocmap, float)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib.emu_stage2_solve(L, P, warps, N, vp(cp), vp(rp), rstride, t.Pc, t.Pr, vp(omap), vp(ocmap), float(t.o0), tol, feas_tol,
                               step_frac, reg, max_iter, vp(obj), vp(x) if want_xy else None, vp(y) if want_xy else None, vp(status),
                               vp(iters), t.n, t.m, st["T"], st["a"], st["binv"], st["half"], st["delta"], st["dur"], st["k_rev"],
                               st["wcf_off"], st["p_off"], vp(ci), vp(ri))
    if rc != 0:
        raise RuntimeError(f"emu_stage2_solve: unsupported geometry (rc {rc})")
    return obj, status, iters, x, y


def solve_long(t, cparams, rparams, warps=1, want_xy=True, tol=1e-9, feas_tol=1e-9, step_frac=0.9995, reg=1e-8, max_iter=60):
    """the long-horizon variant stage2long::warp_body_long (one warp per LP, ceil(T / 32) periods per lane, state in a workspace)"""
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
    _lib.emu_stage2_long_solve.argtypes = ([C.c_int] + [C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p] + [C.c_double] * 5 + [C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_double] * 6
                                           + [C.c_int] * 2 + [C.c_void_p] * 2)
    st = t.meta["stage_wb"]
    cp = np.ascontiguousarray(np.atleast_2d(cparams), float)
    N = cp.shape[0]
    rp = np.ascontiguousarray(rparams, float)
    rstride = 0 if rp.ndim == 1 else rp.shape[1]
    obj = np.zeros(N); status = np.full(N, -9, np.int32); iters = np.zeros(N, np.int32)
    x = np.zeros((N, t.n)); y = np.zeros((N, t.m))
    ci = np.ascontiguousarray(st["col_idx"], np.int32); ri = np.ascontiguousarray(st["row_idx"], np.int32)
    omap = np.ascontiguousarray(t.omap, float); ocmap = np.ascontiguousarray(t.ocmap, float)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _lib.emu_stage2_long_solve(warps, N, vp(cp), vp(rp), rstride, t.Pc, t.Pr, vp(omap), vp(ocmap), float(t.o0), tol, feas_tol,
                                    step_frac, reg, max_iter, vp(obj), vp(x) if want_xy else None, vp(y) if want_xy else None, vp(status),
                                    vp(iters), t.n, t.m, st["T"], st["a"], st["binv"], st["half"], st["delta"], st["dur"], st["k_rev"],
                                    st["wcf_off"], st["p_off"], vp(ci), vp(ri))
    assert rc == 0
    return obj, status, iters, x, y
