"""The band kernel's factorisation / substitution sweeps (dispatches_b200/csrc/dsp_band.cuh) executed on CPU lanes (tests/emu compiles
the CUDA source with g++ on the lock-step SIMT emulator): every instantiated half bandwidth against a dense solve."""
import ctypes as C
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
HERE = Path(__file__).resolve().parent / "emu"
ROOT = HERE.parent.parent


@pytest.fixture(scope="module")
def lib():
    so = HERE / "libemu_band.so"
    deps = [HERE / "emu_band.cpp", HERE / "simt_emu.h", ROOT / "dispatches_b200" / "csrc" / "dsp_band.cuh"]
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        r = subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", str(so), str(HERE / "emu_band.cpp")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    L = C.CDLL(str(so))
    L.emu_band_factor_solve.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    return L


def banded_spd(m, W, rng, cond=1e6):
    """M = A D A' with a banded A (like the normal matrix of an LP iterate: D spans `cond`)"""
    A = np.zeros((m, m + W))
    for i in range(m):
        A[i, i:i + W + 1] = rng.normal(size=W + 1)
    D = np.exp(rng.uniform(0, np.log(cond), m + W))
    return A @ np.diag(D) @ A.T


def pack(M, W):
    m = M.shape[0]
    Mb = np.zeros((m + 2 * W, W + 1))
    for i in range(m):
        for k in range(min(W, i) + 1):
            Mb[W + i, k] = M[i, i - k]
    return Mb


@pytest.mark.parametrize("W,m", [(1, 1), (1, 7), (2, 5), (4, 3), (4, 171), (8, 9), (8, 257), (16, 90), (32, 70)])
def test_band_sweeps_match_a_dense_solve(lib, W, m):
    rng = np.random.default_rng(100 * W + m)
    M = banded_spd(m, W, rng)
    rhs = rng.normal(size=m)
    Mb = pack(M, W)
    v = np.zeros(m + 2 * W); v[W:W + m] = rhs
    assert lib.emu_band_factor_solve(W, m, Mb.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), 0) == 0
    x = v[W:W + m]
    ref = np.linalg.solve(M, rhs)
    assert np.abs(x - ref).max() <= 1e-7 * max(1.0, np.abs(ref).max())
    assert np.abs(M @ x - rhs).max() <= 1e-9 * max(1.0, np.abs(M).max() * np.abs(x).max())
    assert not v[:W].any() and not v[W + m:].any()          # the paddings stay zero


def test_non_positive_pivot_is_skipped(lib):
    """a zero pivot (an all-zero row of A: e.g. a period without its state column) gives 1/d = 0 instead of a NaN"""
    rng = np.random.default_rng(3)
    W, m = 4, 30
    M = banded_spd(m, W, rng, cond=10.0)
    M[7, :] = 0.0; M[:, 7] = 0.0
    Mb = pack(M, W)
    v = np.zeros(m + 2 * W); v[W:W + m] = rng.normal(size=m)
    lib.emu_band_factor_solve(W, m, Mb.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), 0)
    assert np.isfinite(v).all() and Mb[W + 7, 0] == 0.0
