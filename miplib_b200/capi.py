"""ctypes binding of include/mipcu.h (tests and bench.py drive the engine through this).

Nothing here computes: every method is one C-ABI call on host NumPy buffers.  If the shared
library is missing the import fails loudly - there is no eager / CPU path to fall back to.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import re
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_LIB = None

PARAM = {
    "eps_rel": 0, "iter_limit": 1, "time_limit": 2, "check_period": 3, "verbose": 4,
    "step_size": 5, "use_graph": 6, "scaling": 7, "threads_per_row": 8, "obj_cutoff": 9, "use_tma": 10, "shard_exchange": 11, "use_pdl": 12, "multicast": 13, "row_sort": 14, "peer_timeout": 15, "validate": 16,
}
STATUS_NAMES = {0: "optimal", 1: "infeasible", 2: "infeasible_or_unbounded", 3: "unbounded",
                4: "iteration_limit", 5: "time_limit", 6: "numerical_error", 7: "cutoff", 8: "branch"}
UNIQUE_ID_BYTES = 128


class MipcuError(RuntimeError):
    pass


class _Stats(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("restarts", C.c_int32), ("iterations", C.c_int64),
        ("kernel_launches", C.c_int64), ("primal_obj", C.c_double), ("dual_obj", C.c_double),
        ("rel_primal_res", C.c_double), ("rel_dual_res", C.c_double), ("rel_gap", C.c_double),
        ("step_size", C.c_double), ("primal_weight", C.c_double), ("solve_seconds", C.c_double),
        ("iter_seconds", C.c_double), ("setup_seconds", C.c_double), ("row_kernel_us", C.c_double),
        ("col_kernel_us", C.c_double), ("row_kernel_bytes", C.c_int64),
        ("col_kernel_bytes", C.c_int64), ("col_partial_us", C.c_double), ("col_exchange_us", C.c_double),
    ]


@dataclasses.dataclass
class Stats:
    status: str
    restarts: int
    iterations: int
    kernel_launches: int
    primal_obj: float
    dual_obj: float
    rel_primal_res: float
    rel_dual_res: float
    rel_gap: float
    step_size: float
    primal_weight: float
    solve_seconds: float
    iter_seconds: float
    setup_seconds: float
    row_kernel_us: float
    col_kernel_us: float
    row_kernel_bytes: int
    col_kernel_bytes: int
    col_partial_us: float = 0.0
    col_exchange_us: float = 0.0

    @classmethod
    def from_c(cls, s: _Stats) -> "Stats":
        d = {f[0]: getattr(s, f[0]) for f in _Stats._fields_}
        d["status"] = STATUS_NAMES.get(d["status"], str(d["status"]))
        return cls(**d)


def library_path() -> Path:
    return _PKG / "libmiplib_cuda.so"


def declared_symbols() -> list[str]:
    """Every function include/mipcu.h declares (used by the CPU test that checks the exports)."""
    text = (_PKG.parent / "include" / "mipcu.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mipcu_[a-z_0-9]+)\s*\(", text)))


def load_library() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    p = library_path()
    if not p.exists():
        raise MipcuError(f"{p} is missing: build it with `python -m miplib_b200.build` "
                         "(there is no CPU fallback for the Cuda backend)")
    lib = C.CDLL(str(p))
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "mipcu_device_count": (C.c_int, []),
        "mipcu_create": (C.c_int, [P(vp), C.c_int]),
        "mipcu_destroy": (None, [vp]),
        "mipcu_reset": (C.c_int, [vp]),
        "mipcu_last_error": (C.c_char_p, [vp]),
        "mipcu_dist_unique_id": (C.c_int, [vp]),
        "mipcu_create_sharded": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_int, vp]),
        "mipcu_load_lp": (C.c_int, [vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]),
        "mipcu_set_bounds": (C.c_int, [vp, i64, vp, vp, vp]),
        "mipcu_set_objective": (C.c_int, [vp, vp, C.c_int]),
        "mipcu_set_param": (C.c_int, [vp, C.c_int, dbl]),
        "mipcu_get_param": (C.c_int, [vp, C.c_int, P(dbl)]),
        "mipcu_warm_start": (C.c_int, [vp, vp, vp]),
        "mipcu_solve_lp": (C.c_int, [vp, P(_Stats)]),
        "mipcu_solve_lp_batch": (C.c_int, [vp, i32, vp, vp, vp, P(_Stats), vp]),
        "mipcu_solve_lp_batch_warm": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, P(_Stats), vp, vp]),
        "mipcu_get_x": (C.c_int, [vp, vp]),
        "mipcu_get_y": (C.c_int, [vp, vp]),
        "mipcu_spmv": (C.c_int, [vp, C.c_int, vp, vp]),
        "mipcu_prepare": (C.c_int, [vp]),
        "mipcu_get_scaling": (C.c_int, [vp, vp, vp]),
        "mipcu_debug_iterate": (C.c_int, [vp, i64, vp, vp]),
        "mipcu_time_kernel": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, P(dbl), P(i64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _f64(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} entries, got {a.size}")
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One mipcu context = one GPU.  Mirrors the C ABI one call per method."""

    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, unique_id: bytes | None = None):
        self.lib = load_library()
        self.h = C.c_void_p()
        if world > 1:
            buf = C.create_string_buffer(unique_id, UNIQUE_ID_BYTES)
            rc = self.lib.mipcu_create_sharded(C.byref(self.h), device, rank, world, buf)
        else:
            rc = self.lib.mipcu_create(C.byref(self.h), device)
        if rc:
            raise MipcuError(self.lib.mipcu_last_error(None).decode())
        self.m = self.n = self.nnz = 0

    @staticmethod
    def device_count() -> int:
        return load_library().mipcu_device_count()

    @staticmethod
    def unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(UNIQUE_ID_BYTES)
        if lib.mipcu_dist_unique_id(buf):
            raise MipcuError(lib.mipcu_last_error(None).decode())
        return buf.raw

    def _check(self, rc):
        if rc:
            raise MipcuError(self.lib.mipcu_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.mipcu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- model --------------------------------------------------------------------------------
    def load(self, indptr, indices, data, c, lb, ub, row_lo, row_hi, maximize=False, n=None):
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = _f64(data)
        m = indptr.size - 1
        c = _f64(c)
        n = c.size if n is None else n
        self._keep = (indptr, indices, data, c, _f64(lb, n), _f64(ub, n), _f64(row_lo, m), _f64(row_hi, m))
        self._check(self.lib.mipcu_load_lp(self.h, m, n, data.size, *[_ptr(a) for a in self._keep],
                                           int(bool(maximize))))
        self._keep = None
        self.m, self.n, self.nnz = m, n, int(data.size)

    def load_lp(self, lpm):
        """Load an oracle.gen.LP-shaped object (A csr, c, lb, ub, row_lo, row_hi, maximize)."""
        A = lpm.A.tocsr()
        A.sort_indices()
        self.load(A.indptr, A.indices, A.data, lpm.c, lpm.lb, lpm.ub, lpm.row_lo, lpm.row_hi,
                  lpm.maximize, n=A.shape[1])

    def set_bounds(self, cols, lb, ub):
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        lb, ub = _f64(lb, cols.size), _f64(ub, cols.size)
        self._check(self.lib.mipcu_set_bounds(self.h, cols.size, _ptr(cols), _ptr(lb), _ptr(ub)))

    def set_objective(self, c, maximize=False):
        c = _f64(c, self.n)
        self._check(self.lib.mipcu_set_objective(self.h, _ptr(c), int(bool(maximize))))

    def set_param(self, name, value):
        self._check(self.lib.mipcu_set_param(self.h, PARAM[name], float(value)))

    def get_param(self, name) -> float:
        v = C.c_double()
        self._check(self.lib.mipcu_get_param(self.h, PARAM[name], C.byref(v)))
        return v.value

    def warm_start(self, x=None, y=None):
        x = None if x is None else _f64(x, self.n)
        y = None if y is None else _f64(y, self.m)
        self._check(self.lib.mipcu_warm_start(self.h, _ptr(x), _ptr(y)))

    # ---- solve --------------------------------------------------------------------------------
    def solve(self) -> Stats:
        s = _Stats()
        self._check(self.lib.mipcu_solve_lp(self.h, C.byref(s)))
        return Stats.from_c(s)

    def solve_batch(self, lb, ub, cutoff=None, want_x=True, x0=None, y0=None, weight0=None, want_y=False,
                    branch_tol=None):
        """B node LPs over the loaded matrix.  x0 / y0 ([B][n] / [B][m]) / weight0 ([B]): start points;
        branch_tol ([B]): early MIPCU_BRANCH stop (mipcu_solve_lp_batch_warm).  Returns (stats, x) or, with want_y, (stats, x, y)."""
        lb = np.ascontiguousarray(lb, dtype=np.float64)
        ub = np.ascontiguousarray(ub, dtype=np.float64)
        B = lb.shape[0]
        cutoff = None if cutoff is None else _f64(cutoff, B)
        stats = (_Stats * B)()
        x = np.empty((B, self.n)) if want_x else None
        y = np.empty((B, self.m)) if want_y else None
        if x0 is None and y0 is None and weight0 is None and branch_tol is None and not want_y:
            self._check(self.lib.mipcu_solve_lp_batch(self.h, B, _ptr(lb), _ptr(ub), _ptr(cutoff), stats, _ptr(x)))
        else:
            x0 = None if x0 is None else _f64(x0, B * self.n)
            y0 = None if y0 is None else _f64(y0, B * self.m)
            weight0 = None if weight0 is None else _f64(weight0, B)
            branch_tol = None if branch_tol is None else _f64(branch_tol, B)
            self._check(self.lib.mipcu_solve_lp_batch_warm(self.h, B, _ptr(lb), _ptr(ub), _ptr(cutoff), _ptr(x0),
                                                           _ptr(y0), _ptr(weight0), _ptr(branch_tol), stats,
                                                           _ptr(x), _ptr(y)))
        out = [Stats.from_c(s) for s in stats]
        return (out, x, y) if want_y else (out, x)

    def x(self):
        out = np.empty(self.n)
        self._check(self.lib.mipcu_get_x(self.h, _ptr(out)))
        return out

    def y(self):
        out = np.empty(self.m)
        self._check(self.lib.mipcu_get_y(self.h, _ptr(out)))
        return out

    # ---- hooks --------------------------------------------------------------------------------
    def spmv(self, v, transpose=False):
        v = _f64(v, self.m if transpose else self.n)
        out = np.empty(self.n if transpose else self.m)
        self._check(self.lib.mipcu_spmv(self.h, int(transpose), _ptr(v), _ptr(out)))
        return out

    def prepare(self):
        self._check(self.lib.mipcu_prepare(self.h))

    def scaling(self):
        d1, d2 = np.empty(self.m), np.empty(self.n)
        self._check(self.lib.mipcu_get_scaling(self.h, _ptr(d1), _ptr(d2)))
        return d1, d2

    def debug_iterate(self, iters):
        xbar, y = np.empty(self.n), np.empty(self.m)
        self._check(self.lib.mipcu_debug_iterate(self.h, iters, _ptr(xbar), _ptr(y)))
        return xbar, y

    def time_kernel(self, which, reps=50, flush_l2=False):
        us, nbytes = C.c_double(), C.c_int64()
        self._check(self.lib.mipcu_time_kernel(self.h, which, reps, int(flush_l2), C.byref(us),
                                               C.byref(nbytes)))
        return us.value, nbytes.value
