"""What a load-time binning of rows / columns is worth on the GPU (config 3): the matrix is permuted on the host
the way a load-time pass would do it, loaded, and the two fused kernels are timed (mipcu_time_kernel) beside the
host-computed requests per nonzero (scripts/locality_lab.py).   usage: python scripts/locality_gpu.py [--config 3]"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from miplib_b200 import Engine
from oracle import gen
from scripts.locality_lab import requests_per_nnz

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
args = ap.parse_args()
L = gen.config(args.config)
A = L.A.tocsr(); A.sort_indices()


def first_entry(M):
    lens = np.diff(M.indptr)
    return np.where(lens > 0, M.indices[np.minimum(M.indptr[:-1], M.nnz - 1)], 0), lens


def variant(name, rows=None, cols=None):
    M = A
    lo, hi, c, lb, ub = L.row_lo, L.row_hi, L.c, L.lb, L.ub
    if rows is not None:
        M = M[rows]; lo = lo[rows]; hi = hi[rows]
    if cols is not None:
        M = M[:, cols]; c = c[cols]; lb = lb[cols]; ub = ub[cols]
    M = M.tocsr(); M.sort_indices()
    MT = M.T.tocsr(); MT.sort_indices()
    rec = {"variant": name, "row_requests_per_nnz": requests_per_nnz(M), "col_requests_per_nnz": requests_per_nnz(MT)}
    Lv = gen.LP(M, c, lb, ub, lo, hi, L.maximize)
    with Engine(0) as e:
        e.load_lp(Lv)
        e.prepare()
        for _ in range(2):
            rec["row_step_us"], _ = e.time_kernel(2, 200, False)
            rec["col_step_us"], _ = e.time_kernel(3, 200, False)
        st = e.solve()
        rec.update(status=st.status, iterations=st.iterations, objective=st.primal_obj,
                   in_solve_row_us=st.row_kernel_us, in_solve_col_us=st.col_kernel_us, solve_seconds=st.solve_seconds)
    print(json.dumps(rec), flush=True)
    return M


variant("as generated")
f, lens = first_entry(A)
rows_len = np.argsort(-lens, kind="stable")               # what the engine does today
rows_bin = np.lexsort((f, -lens))                          # (length, first column)
variant("rows by (length, first column)", rows=rows_bin)
A1 = A[rows_len].tocsr()
T1 = A1.T.tocsr(); T1.sort_indices()
fT, lensT = first_entry(T1)
cols_bin = np.lexsort((fT, -lensT))                        # columns by (length, first internal row)
variant("columns by (length, first row)", rows=rows_len, cols=cols_bin)
A2 = A1[:, cols_bin].tocsr(); A2.sort_indices()
f2, lens2 = first_entry(A2)
# rows re-binned on the new column numbering WITHOUT changing their order across lengths would renumber
# the rows the columns were binned on; measure that too
variant("columns binned, then rows by (length, first column)", rows=rows_len[np.lexsort((f2, -lens2))], cols=cols_bin)
