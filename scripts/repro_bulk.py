import sys, ctypes as C, faulthandler
sys.path.insert(0, ".")
faulthandler.enable()
import numpy as np
from oracle import gen
from miplib_b200 import build
lib = C.CDLL(str(build.FRONT_LIB))
lib.miplib_get_last_error.restype = C.c_char_p
which = sys.argv[1] if len(sys.argv) > 1 else "lp"
L = gen.config(1) if which == "lp" else gen.set_cover(200, 500, 8, 20264)
types = None if which == "lp" else np.ones(L.n, np.int32)
solver = C.c_void_p()
print("create", lib.miplib_create_solver(C.byref(solver), 5), flush=True)
A = L.A.tocsr(); A.sort_indices()
arrs = [np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
        np.ascontiguousarray(A.data, dtype=np.float64), np.ascontiguousarray(L.c, dtype=np.float64),
        np.ascontiguousarray(L.lb, dtype=np.float64), np.ascontiguousarray(L.ub, dtype=np.float64),
        None if types is None else np.ascontiguousarray(types, dtype=np.int32),
        np.ascontiguousarray(np.maximum(L.row_lo, -1e30), dtype=np.float64),
        np.ascontiguousarray(np.minimum(L.row_hi, 1e30), dtype=np.float64)]
ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
print("load", lib.miplib_cuda_load_csr(solver, L.m, L.n, *ptrs, int(L.maximize)), flush=True)
result, has = C.c_int(-1), C.c_int(-1)
lib.miplib_cuda_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
print("solve", lib.miplib_cuda_solve(solver, C.byref(result), C.byref(has)), result.value, has.value, flush=True)
obj = C.c_double()
lib.miplib_cuda_get_values.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
print("values", lib.miplib_cuda_get_values(solver, None, C.byref(obj)), obj.value, flush=True)
lib.miplib_destroy_solver.argtypes = [C.c_void_p]
print("destroy", lib.miplib_destroy_solver(solver), flush=True)
