"""ctypes driver of oracle/pdhg_port.c (TEST INFRASTRUCTURE; see the header of that file).

``build()`` compiles the C restatement with gcc + OpenMP into oracle/_build/ (git-ignored, but it
travels to the GPU box like every other built .so).  ``run_iterations`` is what bench.py times as
the multi-core CPU baseline: the same two fused passes per iteration the GPU runs."""
from __future__ import annotations

import ctypes as C
import subprocess
import time
from pathlib import Path

import numpy as np

from . import pdhg_ref

HERE = Path(__file__).resolve().parent
SRC = HERE / "pdhg_port.c"
LIB = HERE / "_build" / "libpdhg_port.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", str(LIB),
                        str(SRC), "-lm"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        L.port_num_threads.restype = C.c_int
        L.port_set_threads.argtypes = [C.c_int]
        L.port_set_threads.restype = None
        L.port_touch_copy_csr.argtypes = [i64] + [vp] * 6
        L.port_touch_copy_csr.restype = None
        L.port_touch_copy_vec.argtypes = [i64, vp, vp]
        L.port_touch_copy_vec.restype = None
        L.port_spmv.argtypes = [i64, vp, vp, vp, vp, vp]
        L.port_spmv.restype = None
        L.port_row_step.argtypes = [i64] + [vp] * 9 + [dbl, dbl]
        L.port_row_step.restype = dbl
        L.port_col_step.argtypes = [i64] + [vp] * 12 + [dbl, dbl]
        L.port_col_step.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().port_num_threads()


def use_all_threads() -> int:
    """Ignore an inherited OMP_NUM_THREADS (torchrun sets it to 1) and use every host thread."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().port_set_threads(n)
    return num_threads()


class PortState:
    """Scaled problem + iterate buffers laid out for the C port (int64 ptr, int32 idx)."""

    def __init__(self, S: pdhg_ref.Scaled):
        self.S = S
        K, KT = S.K, S.KT
        self.rp, self.ri, self.rv = self._touch_csr(K)
        self.cp, self.ci, self.cv = self._touch_csr(KT)
        self.m, self.n = K.shape
        n, m = self.n, self.m
        tv = self._touch_vec
        self.x0 = tv(np.clip(np.zeros(n), S.l, S.u))
        self.y0 = tv(np.zeros(m))
        self.aty0 = tv(np.zeros(n))
        self.aty = tv(np.zeros(n))
        self.y = tv(np.zeros(m))
        self.yhat = tv(np.zeros(m))
        # the constant vectors the two passes read (same first-touch placement as the iterates)
        self.lo, self.hi = tv(S.lo), tv(S.hi)
        self.c, self.l, self.u = tv(S.c), tv(S.l), tv(S.u)
        tau = S.eta / S.w0
        xhat = np.clip(self.x0 - tau * (S.c - self.aty0), S.l, S.u)
        self.xbar = tv(2.0 * xhat - self.x0)
        self.k = 0

    @staticmethod
    def _touch_csr(M):
        """Copies of the CSR arrays whose pages are first written by the OpenMP thread that will
        stream them (NUMA first touch; the SciPy originals were written by one thread)."""
        ptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
        idx = np.ascontiguousarray(M.indices, dtype=np.int32)
        val = np.ascontiguousarray(M.data, dtype=np.float64)
        p2, i2, v2 = np.empty_like(ptr), np.empty_like(idx), np.empty_like(val)
        lib().port_touch_copy_csr(M.shape[0], _p(ptr), _p(idx), _p(val), _p(p2), _p(i2), _p(v2))
        return p2, i2, v2

    @staticmethod
    def _touch_vec(v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.empty_like(v)
        lib().port_touch_copy_vec(v.size, _p(v), _p(out))
        return out

    def iterate(self, iters: int) -> float:
        """`iters` PDHG iterations from the current state (no restart logic); returns seconds."""
        L = lib()
        S = self.S
        tau, sigma = S.eta / S.w0, S.eta * S.w0
        t0 = time.perf_counter()
        for _ in range(iters):
            a = (self.k + 1.0) / (self.k + 2.0)
            L.port_row_step(self.m, _p(self.rp), _p(self.ri), _p(self.rv), _p(self.xbar), _p(self.lo),
                            _p(self.hi), _p(self.y0), _p(self.y), _p(self.yhat), sigma, a)
            L.port_col_step(self.n, _p(self.cp), _p(self.ci), _p(self.cv), _p(self.yhat), _p(self.x0),
                            _p(self.aty0), _p(self.c), _p(self.l), _p(self.u), _p(self.xbar), _p(self.aty),
                            None, tau, a)
            self.k += 1
        return time.perf_counter() - t0
