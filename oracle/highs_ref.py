"""Exact-solver stand-in for the reference's SCIP / lp_solve backends (TEST INFRASTRUCTURE).

The reference's own solve is one opaque call into an un-vendored third-party library
(``SCIPsolve`` reference miplib/scip/solver.cpp:370; ``::solve(lprec*)``
miplib/lpsolve/solver.cpp:156).  SCIP 8.0.3 / lp_solve 5.5 (docker/Dockerfile:45,64-66) are
not installable offline, so the independent optimum comes from SciPy's bundled HiGHS.
The LP optimal objective is solver independent, which is what parity is anchored on.
HiGHS is NOT SCIP/lp_solve: every report says "HiGHS stand-in".
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog, milp, LinearConstraint, Bounds

from .gen import LP


def _split_rows(lpm: LP):
    """row_lo <= A x <= row_hi  ->  (A_ub, b_ub, A_eq, b_eq) as linprog wants them."""
    A = lpm.A.tocsr()
    eq = lpm.row_lo == lpm.row_hi
    up = np.isfinite(lpm.row_hi) & ~eq
    lo = np.isfinite(lpm.row_lo) & ~eq
    blocks, rhs = [], []
    if up.any():
        blocks.append(A[up]); rhs.append(lpm.row_hi[up])
    if lo.any():
        blocks.append(-A[lo]); rhs.append(-lpm.row_lo[lo])
    A_ub = sp.vstack(blocks).tocsr() if blocks else None
    b_ub = np.concatenate(rhs) if rhs else None
    A_eq = A[eq] if eq.any() else None
    b_eq = lpm.row_hi[eq] if eq.any() else None
    return A_ub, b_ub, A_eq, b_eq


def solve_lp(lpm: LP, method: str = "highs-ds", time_limit: float | None = None):
    """LP relaxation optimum.  Returns dict(status, obj, x, nit); obj is in the model's own sense."""
    sgn = -1.0 if lpm.maximize else 1.0
    A_ub, b_ub, A_eq, b_eq = _split_rows(lpm)
    bounds = np.column_stack([np.where(np.isfinite(lpm.lb), lpm.lb, -np.inf),
                              np.where(np.isfinite(lpm.ub), lpm.ub, np.inf)])
    opts = {}
    if time_limit is not None:
        opts["time_limit"] = float(time_limit)
    r = linprog(sgn * lpm.c, A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq,
                bounds=bounds, method=method, options=opts)
    return {"status": int(r.status), "message": r.message,
            "obj": (sgn * r.fun) if r.fun is not None else None,
            "x": r.x, "nit": getattr(r, "nit", None)}


def solve_mip(lpm: LP, time_limit: float | None = None):
    """Proven-optimal MIP incumbent via scipy.optimize.milp (HiGHS branch and cut)."""
    sgn = -1.0 if lpm.maximize else 1.0
    integ = np.zeros(lpm.n) if lpm.integer is None else lpm.integer.astype(float)
    opts = {}
    if time_limit is not None:
        opts["time_limit"] = float(time_limit)
    r = milp(sgn * lpm.c, constraints=LinearConstraint(lpm.A, lpm.row_lo, lpm.row_hi),
             integrality=integ, bounds=Bounds(lpm.lb, lpm.ub), options=opts)
    return {"status": int(r.status), "message": r.message,
            "obj": (sgn * r.fun) if r.fun is not None else None, "x": r.x}


def solve_lp_pdlp(lpm: LP, eps: float = 1e-6, time_limit: float | None = None):
    """Second, independent first-order solve: HiGHS' own PDLP (cuPDLP-C, single CPU thread) through
    the highspy binding SciPy bundles.  Exact solvers do not finish configs 2 and 3 in this
    container; this one does (config 2: 9040 iterations, 121 s), and it shares no
    code, scaling, step-size rule or restart scheme with miplib_b200 - only the LP.
    Returns dict(status, obj, x, iters, secs, max_primal_infeas, max_dual_infeas)."""
    import time
    import scipy.optimize._highspy._core as core
    A = lpm.A.tocsc()
    A.sort_indices()
    sgn = -1.0 if lpm.maximize else 1.0
    inf = core.kHighsInf
    lp = core.HighsLp()
    lp.num_col_, lp.num_row_ = lpm.n, lpm.m
    lp.col_cost_ = (sgn * lpm.c).astype(np.float64)
    lp.col_lower_ = np.where(np.isfinite(lpm.lb), lpm.lb, -inf)
    lp.col_upper_ = np.where(np.isfinite(lpm.ub), lpm.ub, inf)
    lp.row_lower_ = np.where(np.isfinite(lpm.row_lo), lpm.row_lo, -inf)
    lp.row_upper_ = np.where(np.isfinite(lpm.row_hi), lpm.row_hi, inf)
    lp.a_matrix_.format_ = core.MatrixFormat.kColwise
    lp.a_matrix_.num_col_, lp.a_matrix_.num_row_ = lpm.n, lpm.m
    lp.a_matrix_.start_ = A.indptr.astype(np.int32)
    lp.a_matrix_.index_ = A.indices.astype(np.int32)
    lp.a_matrix_.value_ = A.data.astype(np.float64)
    h = core._Highs()
    h.setOptionValue("output_flag", False)
    h.setOptionValue("solver", "pdlp")
    h.setOptionValue("presolve", "off")
    if time_limit is not None:
        h.setOptionValue("time_limit", float(time_limit))
    # pdlp_optimality_tolerance (relative duality gap) keeps its default 1e-7: tighter than eps
    for key in ("primal_feasibility_tolerance", "dual_feasibility_tolerance"):
        h.setOptionValue(key, float(eps))
    h.passModel(lp)
    t0 = time.time()
    h.run()
    secs = time.time() - t0
    info = h.getInfo()
    return {"status": str(h.getModelStatus()).split(".")[-1], "obj": sgn * info.objective_function_value,
            "x": np.asarray(h.getSolution().col_value), "iters": int(getattr(info, "pdlp_iteration_count", -1)),
            "secs": secs, "max_primal_infeas": info.max_primal_infeasibility,
            "max_dual_infeas": info.max_dual_infeasibility}
