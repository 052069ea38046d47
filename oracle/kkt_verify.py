"""Independent FP64 optimality-certificate check of a returned (x, y) (TEST INFRASTRUCTURE).

Shares no code with the CUDA path or with pdhg_ref.py: it recomputes, from the ORIGINAL
unscaled model with SciPy-sparse, the relative primal residual, dual residual and duality gap.
By weak duality ``dual_obj <= optimum <= primal_obj`` up to the feasibility errors, so a verified
1e-6 triple is agreement with ANY exact solver's objective (the reference's SCIP / lp_solve
backends included - reference miplib/scip/solver.cpp:370, miplib/lpsolve/solver.cpp:156) to about
1e-6 (1 + |obj|), with no solver in the loop.  This is the parity instrument for C2 / C3, where
no factorisation-based CPU solver finishes.
"""
from __future__ import annotations

import numpy as np

from .gen import LP


def certificate(lpm: LP, x: np.ndarray, y: np.ndarray) -> dict:
    """y is the multiplier of the row constraints in the model's own sense (for a maximise
    model it is the multiplier of the equivalent minimisation, negated)."""
    sgn = -1.0 if lpm.maximize else 1.0
    A = lpm.A.tocsr()
    c = sgn * np.asarray(lpm.c, dtype=np.float64)
    y = sgn * np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    lo, hi, lb, ub = lpm.row_lo, lpm.row_hi, lpm.lb, lpm.ub

    ax = A @ x
    row_viol = np.maximum(lo - ax, 0.0) + np.maximum(ax - hi, 0.0)
    bound_viol = np.maximum(lb - x, 0.0) + np.maximum(x - ub, 0.0)
    b = np.maximum(np.where(np.isfinite(lo), np.abs(lo), 0.0), np.where(np.isfinite(hi), np.abs(hi), 0.0))
    rel_p = np.linalg.norm(row_viol) / (1.0 + np.linalg.norm(b))

    r = c - A.T @ y                       # reduced costs
    # a positive reduced cost needs a finite lower bound to be priced, a negative one a finite upper
    dual_viol = np.where(np.isfinite(lb), 0.0, np.maximum(r, 0.0)) + \
        np.where(np.isfinite(ub), 0.0, np.maximum(-r, 0.0))
    # a positive multiplier needs a finite row_lo, a negative one a finite row_hi
    y_viol = np.where(np.isfinite(lo), 0.0, np.maximum(y, 0.0)) + \
        np.where(np.isfinite(hi), 0.0, np.maximum(-y, 0.0))
    rel_d = np.sqrt(np.sum(dual_viol ** 2) + np.sum(y_viol ** 2)) / (1.0 + np.linalg.norm(c))

    with np.errstate(invalid="ignore"):
        dobj = (np.sum(np.where(y > 0, np.where(np.isfinite(lo), lo, 0.0) * y, 0.0))
                + np.sum(np.where(y < 0, np.where(np.isfinite(hi), hi, 0.0) * y, 0.0))
                + np.sum(np.where(r > 0, np.where(np.isfinite(lb), lb, 0.0) * r, 0.0))
                + np.sum(np.where(r < 0, np.where(np.isfinite(ub), ub, 0.0) * r, 0.0)))
    pobj = float(c @ x)
    rel_g = abs(pobj - dobj) / (1.0 + abs(pobj) + abs(dobj))
    return {"rel_primal_res": float(rel_p), "rel_dual_res": float(rel_d), "rel_gap": float(rel_g),
            "primal_obj": sgn * pobj, "dual_obj": sgn * float(dobj),
            "max_bound_violation": float(bound_viol.max()) if x.size else 0.0}
