"""Row-sharded restatement of oracle/pdhg_ref.py (TEST INFRASTRUCTURE).

Each rank holds a row block of A; the only data-path exchange is the sum over ranks of the partial
products K_g' yhat_g, plus a few scalars on check iterations and the column norms during the
equilibration - exactly what miplib_b200/csrc/dist.cu does with NCCL.  Here the collectives go
through a tiny `comm` object (torch.distributed over gloo in tests/test_sharded_cpu.py), so the
N > 1 host logic is covered on CPU with world_size 2."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import pdhg_ref as R


class LocalComm:
    """world_size 1: every collective is the identity."""
    def sum(self, a): return a
    def max(self, a): return a


def _segment(uf, values, ptr):
    return R._segment_reduce(uf, values, ptr)


def scale_sharded(Ag: sp.csr_matrix, comm):
    """Ruiz + Pock-Chambolle on a row block; column norms are combined across ranks."""
    K = Ag.tocsr().copy().astype(np.float64)
    K.sort_indices()
    mg, n = K.shape
    d1 = np.ones(mg)
    d2 = np.ones(n)
    rows = np.repeat(np.arange(mg), np.diff(K.indptr))
    perm = np.argsort(K.indices, kind="stable")
    colptr = np.concatenate([[0], np.cumsum(np.bincount(K.indices, minlength=n))])
    for sweep in range(R.RUIZ_PASSES + 1):
        a = np.abs(K.data)
        ruiz = sweep < R.RUIZ_PASSES
        uf = np.maximum if ruiz else np.add
        rn = _segment(uf, a, K.indptr)
        cn = _segment(uf, a[perm], colptr)
        cn = comm.max(cn) if ruiz else comm.sum(cn)
        r = np.where(rn > 0, 1.0 / np.sqrt(np.where(rn > 0, rn, 1.0)), 1.0)
        c = np.where(cn > 0, 1.0 / np.sqrt(np.where(cn > 0, cn, 1.0)), 1.0)
        K.data = (K.data * r[rows]) * c[K.indices]
        d1 *= r
        d2 *= c
    return K, d1, d2


def solve_sharded(Ag, c, lb, ub, lo_g, hi_g, comm, maximize=False, eps=1e-6, max_iter=200000, period=64):
    """Same algorithm and schedule as pdhg_ref.solve, on one rank's row block."""
    sgn = -1.0 if maximize else 1.0
    c0 = sgn * np.asarray(c, float)
    K, d1, d2 = scale_sharded(Ag, comm)
    KT = K.T.tocsr(); KT.sort_indices()
    cs = c0 * d2; l = lb / d2; u = ub / d2; lo = lo_g * d1; hi = hi_g * d1
    bnorm = float(np.sqrt(comm.sum(np.array([np.sum(R.finite_abs_max(lo_g, hi_g) ** 2)]))[0]))
    cnorm = float(np.linalg.norm(c0))
    bs = float(np.sqrt(comm.sum(np.array([np.sum(R.finite_abs_max(lo, hi) ** 2)]))[0]))
    csn = float(np.linalg.norm(cs))
    w = csn / bs if (csn > 0 and bs > 0) else 1.0
    # power iteration with the summed K'K v
    v = R.hash_unit_vector(K.shape[1]); v /= np.linalg.norm(v)
    prev = s = 0.0; it = 0
    while it < R.POWER_MAX:
        for _ in range(R.POWER_BATCH):
            v = comm.sum(KT @ (K @ v)); s2 = np.linalg.norm(v); v /= s2; it += 1
        s = np.sqrt(s2)
        if abs(s - prev) <= R.POWER_TOL * s: break
        prev = s
    eta = R.STEP_SAFETY / s
    n, mg = K.shape[1], K.shape[0]
    x0 = np.clip(np.zeros(n), l, u); y0 = np.zeros(mg); aty0 = comm.sum(KT @ y0)

    def seed(x0, aty0, w):
        xh = np.clip(x0 - (eta / w) * (cs - aty0), l, u)
        return xh, 2.0 * xh - x0
    xhat_first, xbar = seed(x0, aty0, w)
    y = y0.copy(); aty = aty0.copy(); k = total = restarts = 0
    fp0 = fp_prev = None; xhat_chk = xhat_first
    status = "iteration_limit"
    while total < max_iter:
        tau, sigma = eta / w, eta * w
        for i in range(period):
            a = (k + 1.0) / (k + 2.0)
            first, last = i == 0, i == period - 1
            y_k = y
            kx = K @ xbar
            t = kx - y / sigma
            yhat = sigma * (np.clip(t, lo, hi) - t)
            y = a * (2.0 * yhat - y) + (1.0 - a) * y0
            if first or last:
                xh_k = xhat_first if first else xhat_chk
                dx = xbar - xh_k
                dy2 = comm.sum(np.array([np.sum((yhat - y_k) ** 2)]))[0]
                if last:
                    kxh = K @ xh_k
            atyh = comm.sum(KT @ yhat)                      # the one data-path collective
            x = a * xbar + (1.0 - a) * x0
            aty_new = a * (2.0 * atyh - aty) + (1.0 - a) * aty0
            xhat_next = np.clip(x - tau * (cs - aty_new), l, u)
            if first or last:
                fp = np.sqrt(max((w / eta) * (dx @ dx) - 2.0 * (dx @ (atyh - aty)) + dy2 / (eta * w), 0.0))
            aty = aty_new; xbar = 2.0 * xhat_next - x
            if i == period - 2: xhat_chk = xhat_next
            if last: xhat_first = xhat_next
            k += 1; total += 1
            if first and k == 1: fp0 = fp_prev = fp
        pres2, dy0, dobj_row = comm.sum(np.array([
            np.sum(((kxh - np.clip(kxh, lo, hi)) / d1) ** 2), np.sum((yhat - y0) ** 2),
            np.where(np.isfinite(lo), lo, 0.0) @ np.maximum(yhat, 0) - np.where(np.isfinite(hi), hi, 0.0) @ np.maximum(-yhat, 0)]))
        r = cs - atyh; rp = np.maximum(r, 0); rm = np.maximum(-r, 0)
        viol = (np.where(np.isfinite(l), 0, rp) + np.where(np.isfinite(u), 0, rm)) / d2
        dobj = dobj_row + np.where(np.isfinite(l), l, 0) @ rp - np.where(np.isfinite(u), u, 0) @ rm
        pobj = cs @ xh_k
        rel_p = np.sqrt(pres2) / (1 + bnorm); rel_d = np.linalg.norm(viol) / (1 + cnorm)
        rel_g = abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj))
        if max(rel_p, rel_d, rel_g) <= eps:
            status = "optimal"; break
        restart = fp <= R.BETA_SUFFICIENT * fp0 or (fp <= R.BETA_NECESSARY * fp0 and fp > fp_prev) \
            or k >= R.BETA_ARTIFICIAL * total
        fp_prev = fp
        if restart:
            restarts += 1
            dxn = np.linalg.norm(xh_k - x0); dyn = np.sqrt(dy0)
            if dxn > R.WEIGHT_EPS and dyn > R.WEIGHT_EPS:
                w = np.exp(R.WEIGHT_SMOOTHING * np.log(dyn / dxn) + (1 - R.WEIGHT_SMOOTHING) * np.log(w))
            x0 = xh_k.copy(); y0 = yhat.copy(); aty0 = atyh.copy(); y = y0.copy(); aty = aty0.copy()
            xhat_first, xbar = seed(x0, aty0, w); k = 0
    return dict(status=status, x=xh_k * d2, y_local=sgn * yhat * d1, obj=sgn * pobj, iters=total,
                restarts=restarts, rel_pres=rel_p, rel_dres=rel_d, rel_gap=rel_g)
