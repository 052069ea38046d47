"""BASELINE configs 4 / 5 as MIPs at FULL size under a time limit (nobody here can prove their optima:
SURVEY.md section 7.3-7): the Cuda backend's branch and bound through the plugin C API against SciPy-HiGHS
`milp` (the SCIP / lp_solve stand-in) with the same limit on the same instance.  Checked, size-independently:
the incumbent is integral and feasible for every row and bound (recomputed here in FP64), its objective is
the reported one, the reported dual bound is on the right side of BOTH solvers' incumbents and not below
the root LP bound, and the two solvers' [bound, incumbent] intervals intersect.
usage: python scripts/mip_fullsize.py [--config 4] [--seconds 60] [--gpus 1] [--no-highs]"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from miplib_b200 import build
from oracle import gen

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--no-highs", action="store_true")
args = ap.parse_args()

L = gen.config(args.config)
A = L.A.tocsr(); A.sort_indices()
m, n = A.shape
lib = C.CDLL(str(build.FRONT_LIB))
lib.miplib_get_last_error.restype = C.c_char_p
lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
lib.miplib_cuda_set_time_limit.argtypes = [C.c_void_p, C.c_double]
lib.miplib_cuda_set_nr_threads.argtypes = [C.c_void_p, C.c_int64]
vp = lambda a: a.ctypes.data_as(C.c_void_p)
ptr = A.indptr.astype(np.int64); idx = A.indices.astype(np.int32); val = A.data.astype(np.float64)
integer = L.integer if L.integer is not None else np.ones(n, bool)
binary = integer & (L.lb == 0) & (L.ub == 1)
vtype = np.where(binary, 1, np.where(integer, 2, 0)).astype(np.int32)
lo = np.maximum(L.row_lo, -1e30); hi = np.minimum(L.row_hi, 1e30)
c = np.ascontiguousarray(L.c, dtype=np.float64); lb = np.ascontiguousarray(L.lb, dtype=np.float64)
ub = np.ascontiguousarray(L.ub, dtype=np.float64)


def ck(rc):
    if rc:
        raise RuntimeError(lib.miplib_get_last_error().decode())


s = C.c_void_p()
ck(lib.miplib_create_solver(C.byref(s), 5))
ck(lib.miplib_cuda_set_verbose(s, 0))
ck(lib.miplib_cuda_load_csr(s, m, n, vp(ptr), vp(idx), vp(val), vp(c), vp(lb), vp(ub), vp(vtype), vp(lo), vp(hi),
                            int(L.maximize)))
ck(lib.miplib_cuda_set_time_limit(s, args.seconds))
ck(lib.miplib_cuda_set_nr_threads(s, args.gpus))
res, has = C.c_int(), C.c_int()
t0 = time.perf_counter()
ck(lib.miplib_cuda_solve(s, C.byref(res), C.byref(has)))
secs = time.perf_counter() - t0
it, launches, nodes, insecs = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
ck(lib.miplib_cuda_get_info(s, C.byref(it), C.byref(launches), C.byref(nodes), C.byref(insecs)))
bound, proven = C.c_double(), C.c_int()
ck(lib.miplib_cuda_get_bound(s, C.byref(bound), C.byref(proven)))
out = {"config": args.config, "m": m, "n": n, "nnz": int(A.nnz), "time_limit_s": args.seconds, "gpus": args.gpus,
       "cuda": {"result": res.value, "has_incumbent": bool(has.value), "seconds": secs, "nodes": nodes.value,
                "lp_iterations": it.value, "kernel_launches": launches.value, "dual_bound": bound.value,
                "proven_optimal": bool(proven.value)}}
sgn = -1.0 if L.maximize else 1.0
if has.value:
    x = np.empty(n); obj = C.c_double()
    ck(lib.miplib_cuda_get_values(s, vp(x), C.byref(obj)))
    ax = A @ x
    out["cuda"].update(
        incumbent=obj.value,
        objective_recomputed=float(c @ x),
        max_integrality_violation=float(np.abs(x[integer] - np.round(x[integer])).max()),
        max_bound_violation=float(max((lb - x).max(), (x - ub).max(), 0.0)),
        max_row_violation=float(max((L.row_lo - ax)[np.isfinite(L.row_lo)].max(initial=0.0),
                                    (ax - L.row_hi)[np.isfinite(L.row_hi)].max(initial=0.0), 0.0)),
        relative_gap=abs(obj.value - bound.value) / max(1e-9, abs(obj.value)))
lib.miplib_destroy_solver(s)
print(json.dumps(out), flush=True)

if not args.no_highs:
    from scipy.optimize import milp, LinearConstraint, Bounds
    t0 = time.perf_counter()
    r = milp(sgn * L.c, constraints=LinearConstraint(L.A, L.row_lo, L.row_hi), bounds=Bounds(L.lb, L.ub),
             integrality=integer.astype(int), options={"time_limit": args.seconds, "disp": False})
    hs = {"label": "SciPy-HiGHS milp, stand-in for the reference's SCIP / lp_solve backends (not installable offline)",
          "status": int(r.status), "message": str(r.message)[:80], "seconds": time.perf_counter() - t0,
          "incumbent": None if r.x is None else float(sgn * r.fun),
          "dual_bound": None if getattr(r, "mip_dual_bound", None) is None else float(sgn * r.mip_dual_bound),
          "nodes": int(getattr(r, "mip_node_count", -1) or -1)}
    out["highs"] = hs
    if has.value and hs["incumbent"] is not None and hs["dual_bound"] is not None:
        lo_c, hi_c = sorted((out["cuda"]["dual_bound"], out["cuda"]["incumbent"]))
        lo_h, hi_h = sorted((hs["dual_bound"], hs["incumbent"]))
        tol = 1e-6 * (1 + abs(hi_c))
        out["intervals_intersect"] = bool(lo_c <= hi_h + tol and lo_h <= hi_c + tol)
    print(json.dumps(out), flush=True)
