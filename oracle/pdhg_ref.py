"""CPU restatement (NumPy / SciPy-sparse, FP64) of the algorithm the CUDA backend runs
(TEST INFRASTRUCTURE - never imported by the product path).

What this is the oracle *for*.  In the reference the whole solve is one opaque call into an
un-vendored third-party library (``SCIPsolve`` reference miplib/scip/solver.cpp:370;
``::solve(lprec*)`` miplib/lpsolve/solver.cpp:156; SCIP 8.0.3 / lp_solve 5.5 pinned in
docker/Dockerfile:45,64-66), so there is no reference arithmetic to restate.  The reference
only fixes (i) the lowering semantics - a row ``sum a_j x_j + k {<=,=} 0`` becomes
``row_hi = -k, row_lo = Equal ? -k : -inf`` (miplib/scip/solver.cpp:108-120,
miplib/lpsolve/solver.cpp:115-130), Binary => [0,1] (miplib/lpsolve/var.cpp:112-123), the
objective constant is dropped (miplib/lpsolve/solver.cpp:66-86) - and (ii) the result contract
``pair<Result, has_solution>`` (miplib/lpsolve/solver.cpp:171-199).  This module restates the
*published* first-order method the new backend uses for the LP relaxation: diagonally
preconditioned (Ruiz + Pock-Chambolle), restarted, reflected-Halpern PDHG (PDLP: Applegate
et al. 2021; r2HPDHG: Lu & Yang 2024).  It is written operation-for-operation like the CUDA
kernels in ``miplib_b200/csrc`` so that kernel-level and iterate-level parity can be checked.

PARITY PINNING.  The reference holds no golden vector beyond 3x3 toy models
(test/unit/expr.cpp:134-149, :201-216, test/unit/solver.cpp:29-65); those known-answer tests
are re-hosted assertion for assertion in tests/cpp/frontend_tests.cpp and tests/cpp/solver_tests.cpp
(cases "KAT-A" .. "KAT-D"), run by tests/test_frontend.py.  At BASELINE.json sizes this oracle is pinned against SciPy-HiGHS
optimal objectives (``highs_ref.py``, golden values in tests/golden/objectives.json) - a
stand-in for SCIP/lp_solve, which cannot be installed offline: "parity unpinned" against the
reference's own backends, pinned against an independent exact solver.

Canonical problem (after the sign flip for maximise)::

    min c'x  s.t.  lo <= A x <= hi,  l <= x <= u
"""
from __future__ import annotations

import dataclasses
import time
import numpy as np
import scipy.sparse as sp

from .gen import LP

# ---- constants shared with miplib_b200/csrc/params.h (same names) ---------------------------
RUIZ_PASSES = 10
STEP_SAFETY = 0.998
POWER_BATCH = 10
POWER_MAX = 400
POWER_TOL = 1e-6
BETA_SUFFICIENT = 0.2
BETA_NECESSARY = 0.8
BETA_ARTIFICIAL = 0.36
WEIGHT_SMOOTHING = 0.5
WEIGHT_EPS = 1e-10
RAY_TOL = 1e-8


def hash_unit_vector(n: int) -> np.ndarray:
    """Deterministic start vector for the power iteration: v_j = frac(j * 2654435761 / 2^32) + 0.5
    (Knuth multiplicative hash), bit-identical to ``k_hash_vector`` in csrc/setup.cu."""
    j = np.arange(n, dtype=np.uint64)
    h = (j * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    return h.astype(np.float64) / 4294967296.0 + 0.5


def _row_col_scale(K: sp.csr_matrix, r: np.ndarray, c: np.ndarray) -> sp.csr_matrix:
    """K_ij <- (K_ij * r_i) * c_j, in exactly that order (the CUDA kernel does the same on both
    the CSR and the CSC copy, so the two device copies stay bit-identical)."""
    K = K.copy()
    rows = np.repeat(np.arange(K.shape[0]), np.diff(K.indptr))
    K.data = (K.data * r[rows]) * c[K.indices]
    return K


def _segment_reduce(ufunc, values: np.ndarray, ptr: np.ndarray) -> np.ndarray:
    """ufunc.reduceat over CSR segments, returning 0 for empty segments."""
    nseg = ptr.size - 1
    out = np.zeros(nseg)
    nonempty = ptr[1:] > ptr[:-1]
    if values.size and nonempty.any():
        starts = ptr[:-1][nonempty]
        out[nonempty] = ufunc.reduceat(values, starts)
    return out


def ruiz_pc_scale(A: sp.csr_matrix, ruiz_passes: int = RUIZ_PASSES, pock_chambolle: bool = True):
    """Ruiz (inf-norm, ``ruiz_passes`` sweeps) then Pock-Chambolle alpha=1 equilibration.
    Returns (K, d1, d2) with K = diag(d1) A diag(d2).  Row norms are reduced over the CSR
    storage order, column norms over the CSC storage order (rows increasing inside a column),
    like the two device copies."""
    K = A.tocsr().copy().astype(np.float64)
    K.sort_indices()
    m, n = K.shape
    d1 = np.ones(m)
    d2 = np.ones(n)
    rows = np.repeat(np.arange(m), np.diff(K.indptr))
    perm = np.argsort(K.indices, kind="stable")            # CSR position -> CSC order
    colptr = np.concatenate([[0], np.cumsum(np.bincount(K.indices, minlength=n))])
    for sweep in range(ruiz_passes + (1 if pock_chambolle else 0)):
        a = np.abs(K.data)
        uf = np.maximum if sweep < ruiz_passes else np.add
        rn = _segment_reduce(uf, a, K.indptr)
        cn = _segment_reduce(uf, a[perm], colptr)
        r = np.where(rn > 0, 1.0 / np.sqrt(np.where(rn > 0, rn, 1.0)), 1.0)
        c = np.where(cn > 0, 1.0 / np.sqrt(np.where(cn > 0, cn, 1.0)), 1.0)
        K.data = (K.data * r[rows]) * c[K.indices]
        d1 *= r
        d2 *= c
    return K, d1, d2


def spectral_norm(K: sp.csr_matrix, KT: sp.csr_matrix):
    """Power iteration on K'K from the hash vector, in batches of POWER_BATCH, until the estimate
    moves by < POWER_TOL relative between batches (or POWER_MAX iterations)."""
    v = hash_unit_vector(K.shape[1])
    nv = np.linalg.norm(v)
    if nv == 0 or K.nnz == 0:
        return 0.0, 0
    v /= nv
    prev = 0.0
    s = 0.0
    it = 0
    while it < POWER_MAX:
        for _ in range(POWER_BATCH):
            u = K @ v
            v = KT @ u
            s2 = np.linalg.norm(v)
            if s2 == 0:
                return 0.0, it
            v /= s2
            it += 1
        s = np.sqrt(s2)
        if abs(s - prev) <= POWER_TOL * s:
            break
        prev = s
    return float(s), it


@dataclasses.dataclass
class Scaled:
    K: sp.csr_matrix
    KT: sp.csr_matrix
    d1: np.ndarray
    d2: np.ndarray
    c: np.ndarray
    l: np.ndarray
    u: np.ndarray
    lo: np.ndarray
    hi: np.ndarray
    bnorm: float     # unscaled ||b||_2, b_i = max(|lo_i|, |hi_i|) over finite sides
    cnorm: float     # unscaled ||c||_2
    w0: float        # initial primal weight
    eta: float       # constant step size
    sgn: float       # -1 if the model maximises


def finite_abs_max(lo, hi):
    b = np.where(np.isfinite(lo), np.abs(lo), 0.0)
    return np.maximum(b, np.where(np.isfinite(hi), np.abs(hi), 0.0))


def prepare(lpm: LP, step_size: float | None = None, scale: bool = True) -> Scaled:
    sgn = -1.0 if lpm.maximize else 1.0
    c0 = sgn * np.asarray(lpm.c, dtype=np.float64)
    if scale:
        K, d1, d2 = ruiz_pc_scale(lpm.A)
    else:
        K = lpm.A.tocsr().astype(np.float64); d1 = np.ones(lpm.m); d2 = np.ones(lpm.n)
    KT = K.T.tocsr()
    KT.sort_indices()
    c = c0 * d2
    l = lpm.lb / d2
    u = lpm.ub / d2
    lo = lpm.row_lo * d1
    hi = lpm.row_hi * d1
    bnorm = float(np.linalg.norm(finite_abs_max(lpm.row_lo, lpm.row_hi)))
    cnorm = float(np.linalg.norm(c0))
    bs = float(np.linalg.norm(finite_abs_max(lo, hi)))
    cs = float(np.linalg.norm(c))
    w0 = cs / bs if (cs > 0 and bs > 0) else 1.0
    if step_size is None:
        s, _ = spectral_norm(K, KT)
        eta = STEP_SAFETY / s if s > 0 else 1.0
    else:
        eta = float(step_size)
    return Scaled(K, KT, d1, d2, c, l, u, lo, hi, bnorm, cnorm, w0, eta, sgn)


# ---- the two fused per-iteration kernels, restated -------------------------------------------
def row_step(S: Scaled, xbar, y, y0, sigma, a):
    """k_row_step: kx = K xbar ; yhat = sigma (clamp(t, lo, hi) - t), t = kx - y/sigma ;
    y <- a (2 yhat - y) + (1 - a) y0."""
    kx = S.K @ xbar
    t = kx - y / sigma
    yhat = sigma * (np.clip(t, S.lo, S.hi) - t)
    ynew = a * (2.0 * yhat - y) + (1.0 - a) * y0
    return yhat, ynew, kx


def col_step(S: Scaled, yhat, xbar, x0, aty, aty0, tau, a):
    """k_col_step: atyh = K' yhat ; x <- a xbar + (1-a) x0 ; aty <- a (2 atyh - aty) + (1-a) aty0 ;
    xhat = clamp(x - tau (c - aty), l, u) ; xbar <- 2 xhat - x."""
    atyh = S.KT @ yhat
    x = a * xbar + (1.0 - a) * x0
    aty_new = a * (2.0 * atyh - aty) + (1.0 - a) * aty0
    xhat = np.clip(x - tau * (S.c - aty_new), S.l, S.u)
    xbar_new = 2.0 * xhat - x
    return atyh, aty_new, xhat, xbar_new


def kkt_scaled(S: Scaled, xhat, yhat, atyh, kxhat):
    """Unscaled relative KKT triple of (D2 xhat, D1 yhat) computed from scaled quantities, as
    k_kkt_row / k_col_step<CHECK> do: A x = D1^-1 K xhat ; c - A'y = D2^-1 (c~ - K'yhat)."""
    pres = np.linalg.norm((kxhat - np.clip(kxhat, S.lo, S.hi)) / S.d1)
    r = S.c - atyh
    rp = np.maximum(r, 0.0)
    rm = np.maximum(-r, 0.0)
    viol = np.where(np.isfinite(S.l), 0.0, rp) + np.where(np.isfinite(S.u), 0.0, rm)
    dres = np.linalg.norm(viol / S.d2)
    yp = np.maximum(yhat, 0.0)
    ym = np.maximum(-yhat, 0.0)
    dobj = (np.where(np.isfinite(S.lo), S.lo, 0.0) @ yp - np.where(np.isfinite(S.hi), S.hi, 0.0) @ ym
            + np.where(np.isfinite(S.l), S.l, 0.0) @ rp - np.where(np.isfinite(S.u), S.u, 0.0) @ rm)
    pobj = S.c @ xhat
    rel_p = pres / (1.0 + S.bnorm)
    rel_d = dres / (1.0 + S.cnorm)
    rel_g = abs(pobj - dobj) / (1.0 + abs(pobj) + abs(dobj))
    return rel_p, rel_d, rel_g, pobj, dobj


def ray_tests(S: Scaled, dx, dy, daty, kdx):
    """(dual ray dy proves primal infeasibility, primal ray dx proves dual infeasibility) - the two
    tests of k_ctrl_check on the displacement T(z) - z."""
    fin = np.isfinite
    nd, nx = np.linalg.norm(dy), np.linalg.norm(dx)
    dual_ok = primal_ok = False
    if nd > 1e-12:
        dp, dm = np.maximum(dy, 0), np.maximum(-dy, 0)
        qp, qm = np.maximum(-daty, 0), np.maximum(daty, 0)
        val = (np.where(fin(S.lo), S.lo, 0) @ dp - np.where(fin(S.hi), S.hi, 0) @ dm
               + np.where(fin(S.l), S.l, 0) @ qp - np.where(fin(S.u), S.u, 0) @ qm) / nd
        viol2 = np.sum((np.where(fin(S.lo), 0, dp) + np.where(fin(S.hi), 0, dm)) ** 2) \
            + np.sum((np.where(fin(S.l), 0, qp) + np.where(fin(S.u), 0, qm)) ** 2)
        dual_ok = val > 1e-9 and np.sqrt(viol2) / nd <= RAY_TOL * val
    if nx > 1e-12:
        val = (S.c @ dx) / nx
        viol2 = np.sum((np.where(fin(S.l), np.maximum(-dx, 0), 0) + np.where(fin(S.u), np.maximum(dx, 0), 0)) ** 2) \
            + np.sum((np.where(fin(S.lo), np.maximum(-kdx, 0), 0) + np.where(fin(S.hi), np.maximum(kdx, 0), 0)) ** 2)
        primal_ok = val < -1e-9 and np.sqrt(viol2) / nx <= RAY_TOL * -val
    return dual_ok, primal_ok


def solve(lpm: LP, eps: float = 1e-6, max_iter: int = 200000, period: int = 64,
          step_size: float | None = None, x_start=None, y_start=None, scale: bool = True,
          trace_at: tuple = (), verbose: bool = False):
    """Restarted reflected-Halpern PDHG.  Returns a dict with status, x, y (unscaled), obj
    (in the model's own sense), iters, restarts, rel_pres, rel_dres, rel_gap and ``trace``: the
    scaled (xbar, y) pairs after the iteration counts listed in ``trace_at``."""
    t_start = time.time()
    S = prepare(lpm, step_size, scale)
    n, m = lpm.n, lpm.m
    w = S.w0
    eta = S.eta
    x0 = np.clip(np.zeros(n) if x_start is None else np.asarray(x_start, float) / S.d2, S.l, S.u)
    y0 = np.zeros(m) if y_start is None else np.asarray(y_start, float) / S.d1
    aty0 = S.KT @ y0

    def restart_apply(x0, aty0, w):
        tau = eta / w
        xhat = np.clip(x0 - tau * (S.c - aty0), S.l, S.u)
        return xhat, 2.0 * xhat - x0

    xhat_first, xbar = restart_apply(x0, aty0, w)
    y = y0.copy()
    aty = aty0.copy()
    k = 0            # iterations since the last restart
    total = 0
    restarts = 0
    fp0 = fp_prev = None
    ray_hits = [0, 0]
    trace = {}
    hist = []
    res = dict(status="iteration_limit")
    xhat_chk = xhat_first
    while total < max_iter:
        tau = eta / w
        sigma = eta * w
        for i in range(period):
            a = (k + 1.0) / (k + 2.0)
            first = (i == 0)
            last = (i == period - 1)
            y_k = y
            yhat, y, kx_k = row_step(S, xbar, y, y0, sigma, a)
            if first or last:
                xh_k = xhat_first if first else xhat_chk
                dx = xbar - xh_k                      # = xhat_k - x_k
                dy = yhat - y_k
                if last:
                    kxhat = S.K @ xh_k
            atyh, aty_new, xhat_next, xbar_new = col_step(S, yhat, xbar, x0, aty, aty0, tau, a)
            if first or last:
                daty = atyh - aty
                fp2 = (w / eta) * (dx @ dx) - 2.0 * (dx @ daty) + (dy @ dy) / (eta * w)
                fp = np.sqrt(max(fp2, 0.0))
            aty = aty_new
            xbar = xbar_new
            if i == period - 2:
                xhat_chk = xhat_next
            if last:
                xhat_first = xhat_next
            k += 1
            total += 1
            if total in trace_at:
                trace[total] = (xbar.copy(), y.copy())
            if first and k == 1:
                fp0 = fp
                fp_prev = fp
        # ---- end of block: KKT at (xhat_chk, yhat), restart decision -------------------------
        rel_p, rel_d, rel_g, pobj, dobj = kkt_scaled(S, xh_k, yhat, atyh, kxhat)
        hist.append((total, k, rel_p, rel_d, rel_g, S.sgn * pobj, S.sgn * dobj, w, fp))
        if verbose:
            print(f"it {total:7d} k {k:6d} pres {rel_p:.2e} dres {rel_d:.2e} gap {rel_g:.2e} "
                  f"pobj {S.sgn*pobj:.9f} w {w:.3e} fp {fp:.2e}")
        if max(rel_p, rel_d, rel_g) <= eps:
            res = dict(status="optimal")
            break
        # certificates of infeasibility from the displacement (k_ctrl_check, same tests)
        verdict = ray_tests(S, dx, dy, daty, kx_k - kxhat)
        ray_hits = [ray_hits[0] + 1 if verdict[0] else 0, ray_hits[1] + 1 if verdict[1] else 0]
        if ray_hits[0] >= 2:
            res = dict(status="infeasible")
            break
        if ray_hits[1] >= 2:
            res = dict(status="unbounded" if rel_p <= 1e-4 else "infeasible_or_unbounded")
            break
        do_restart = (fp <= BETA_SUFFICIENT * fp0
                      or (fp <= BETA_NECESSARY * fp0 and fp > fp_prev)
                      or k >= BETA_ARTIFICIAL * total)
        fp_prev = fp
        if do_restart:
            restarts += 1
            dxn = np.linalg.norm(xh_k - x0)
            dyn = np.linalg.norm(yhat - y0)
            if dxn > WEIGHT_EPS and dyn > WEIGHT_EPS:
                w = np.exp(WEIGHT_SMOOTHING * np.log(dyn / dxn) + (1.0 - WEIGHT_SMOOTHING) * np.log(w))
            x0 = xh_k.copy()
            y0 = yhat.copy()
            aty0 = atyh.copy()
            y = y0.copy()
            aty = aty0.copy()
            xhat_first, xbar = restart_apply(x0, aty0, w)
            k = 0
    res.update(x=xh_k * S.d2, y=yhat * S.d1 * S.sgn, obj=S.sgn * pobj, dual_obj=S.sgn * dobj,
               iters=total, restarts=restarts, rel_pres=rel_p, rel_dres=rel_d, rel_gap=rel_g,
               trace=trace, hist=hist, eta=eta, w=w, secs=time.time() - t_start)
    return res
