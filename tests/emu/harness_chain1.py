"""TEST INFRASTRUCTURE: builds tests/emu/emu_chain1.cpp (the CUDA warp body of dispatches_b200/csrc/dsp_stage_chain1.cuh compiled
with g++ on the lock-step SIMT emulator) and calls it through ctypes with a chain descriptor from lp_template.detect_chain1."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE / "libemu_chain1.so"
_lib = None


class Params(C.Structure):          # chain1::Params
    _fields_ = [("N", C.c_longlong), ("cparams", C.c_void_p), ("rparams", C.c_void_p), ("rstride", C.c_longlong), ("Pc", C.c_int), ("Pr", C.c_int),
                ("omap", C.c_void_p), ("ocmap", C.c_void_p), ("o0", C.c_double), ("tol", C.c_double), ("feas_tol", C.c_double),
                ("step_frac", C.c_double), ("reg", C.c_double), ("max_iter", C.c_int), ("obj", C.c_void_p), ("x_out", C.c_void_p),
                ("y_out", C.c_void_p), ("status", C.c_void_p), ("iters", C.c_void_p), ("n", C.c_int), ("m", C.c_int), ("nb", C.c_int),
                ("ticket", C.c_void_p), ("c0", C.c_void_p), ("b0", C.c_void_p), ("u0", C.c_void_p),
                ("cm_ptr", C.c_void_p), ("cm_idx", C.c_void_p), ("bm_ptr", C.c_void_p), ("bm_idx", C.c_void_p), ("um_ptr", C.c_void_p),
                ("um_idx", C.c_void_p), ("cm_val", C.c_void_p), ("bm_val", C.c_void_p), ("um_val", C.c_void_p),
                ("T", C.c_int), ("col_idx", C.c_void_p), ("row_idx", C.c_void_p), ("coef", C.c_void_p), ("coef_next", C.c_void_p),
                ("x_perm", C.c_void_p), ("y_perm", C.c_void_p)]


def build(force=False):
    deps = [HERE / "emu_chain1.cpp", HERE / "simt_emu.h", ROOT / "dispatches_b200" / "csrc" / "dsp_stage_chain1.cuh",
            ROOT / "dispatches_b200" / "csrc" / "dsp_stage2.cuh"]
    if force or not LIB.exists() or any(d.stat().st_mtime > LIB.stat().st_mtime for d in deps):
        cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", str(LIB), str(HERE / "emu_chain1.cpp")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stderr[-3000:])
    return LIB


def solve(t, desc, cparams, rparams, L, P, warps=1, tol=1e-9, feas_tol=1e-9, step_frac=0.9995, reg=1e-8, max_iter=60):
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.emu_chain1_solve.argtypes = [C.c_int] * 4 + [C.POINTER(Params)]
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    f64 = lambda a: np.ascontiguousarray(a, np.float64)
    cp = f64(np.atleast_2d(cparams)); N = cp.shape[0]
    rp = f64(rparams if rparams is not None else np.zeros(1))
    rstride = 0 if rp.ndim == 1 else rp.shape[1]
    nb = t.nb
    Cm, Bm, Um = t.Cmap.tocsr(), t.Bmap.tocsr(), t.Umap.tocsr()[:nb]
    keep = dict(cp=cp, rp=rp, omap=f64(t.omap if t.Pr else np.zeros(1)), ocmap=f64(t.ocmap if t.Pc else np.zeros(1)),
                obj=np.zeros(N), x=np.zeros((N, t.n)), y=np.zeros((N, t.m)), status=np.full(N, -9, np.int32), iters=np.zeros(N, np.int32),
                c0=f64(t.c0), b0=f64(t.b0), u0=f64(np.where(np.isfinite(t.u0[:max(nb, 1)]), t.u0[:max(nb, 1)], 0.0)),
                cm_ptr=i32(Cm.indptr), cm_idx=i32(Cm.indices), cm_val=f64(Cm.data), bm_ptr=i32(Bm.indptr), bm_idx=i32(Bm.indices),
                bm_val=f64(Bm.data), um_ptr=i32(Um.indptr), um_idx=i32(Um.indices), um_val=f64(Um.data),
                col_idx=i32(desc["col_idx"]), row_idx=i32(desc["row_idx"]), coef=f64(desc["coef"]), coef_next=f64(desc["coef_next"]))
    p = lambda k: keep[k].ctypes.data_as(C.c_void_p)
    Q = Params(N=N, cparams=p("cp"), rparams=p("rp"), rstride=rstride, Pc=t.Pc, Pr=t.Pr, omap=p("omap"), ocmap=p("ocmap"), o0=float(t.o0),
               tol=tol, feas_tol=feas_tol, step_frac=step_frac, reg=reg, max_iter=max_iter, obj=p("obj"), x_out=p("x"), y_out=p("y"),
               status=p("status"), iters=p("iters"), n=t.n, m=t.m, nb=nb, ticket=None, c0=p("c0"), b0=p("b0"), u0=p("u0"),
               cm_ptr=p("cm_ptr"), cm_idx=p("cm_idx"), bm_ptr=p("bm_ptr"), bm_idx=p("bm_idx"), um_ptr=p("um_ptr"), um_idx=p("um_idx"),
               cm_val=p("cm_val"), bm_val=p("bm_val"), um_val=p("um_val"), T=desc["T"], col_idx=p("col_idx"), row_idx=p("row_idx"),
               coef=p("coef"), coef_next=p("coef_next"), x_perm=None, y_perm=None)
    rc = _lib.emu_chain1_solve(L, P, desc["NF"], warps, C.byref(Q))
    if rc != 0:
        raise RuntimeError(f"emu_chain1_solve: unsupported geometry (rc {rc})")
    return keep["obj"], keep["status"], keep["iters"], keep["x"], keep["y"]
