"""Row-sharded restatement of oracle/pdhg_ref.py (TEST INFRASTRUCTURE).

Each rank holds a row block of A; the only data-path exchange is the sum over ranks of the partial
products K_g' yhat_g, plus a few scalars on check iterations and the column norms during the
equilibration - exactly what miplib_b200/csrc/dist.cu does with NCCL.  Here the collectives go
through a tiny `comm` object (torch.distributed over gloo in tests/test_sharded_cpu.py), so the
N > 1 host logic is covered on CPU with world_size 2.  `solve_block_owner` restates the default
exchange of miplib_b200/csrc/p2p.cu (row block of K + column block of K' per rank, two all-gathers
per iteration) the same way."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import pdhg_ref as R


class LocalComm:
    """world_size 1: every collective is the identity."""
    def sum(self, a): return a
    def max(self, a): return a
    def gather(self, a): return [a]


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


def solve_block_owner(Ag, c, lb, ub, lo_g, hi_g, comm, rank, world, maximize=False, eps=1e-6,
                      max_iter=200000, period=64):
    """The block-owner exchange of miplib_b200/csrc/p2p.cu (MIPCU_PARAM_SHARD_EXCHANGE = 2): rank g
    owns row block g of K (row step on its rows, yhat slices all-gathered) AND column block g of K'
    (whole columns cut out of all ranks' blocks at load; primal step on its column slice, xbar
    slices all-gathered).  No partial-sum vector, no reduction on the data path: every column sum
    runs over the whole column in global row order, as in the single-process oracle.
    `comm.gather(a)` returns the list of every rank's array, in rank order."""
    sgn = -1.0 if maximize else 1.0
    c0 = sgn * np.asarray(c, float)
    K, d1, d2 = scale_sharded(Ag, comm)
    KT = K.T.tocsr(); KT.sort_indices()
    n, mg = K.shape[1], K.shape[0]
    j0, j1 = (n * rank) // world, (n * (rank + 1)) // world
    # column block of the WHOLE scaled matrix: all ranks' blocks stacked, columns [j0, j1)
    blocks = comm.gather((K.indptr, K.indices, K.data, mg))
    Kfull = sp.vstack([sp.csr_matrix((d, i, p), shape=(r, n)) for p, i, d, r in blocks]).tocsr()
    KcT = Kfull[:, j0:j1].T.tocsr(); KcT.sort_indices()           # (j1 - j0) x m_glob
    cs = c0 * d2; l = lb / d2; u = ub / d2; lo = lo_g * d1; hi = hi_g * d1
    bnorm = float(np.sqrt(comm.sum(np.array([np.sum(R.finite_abs_max(lo_g, hi_g) ** 2)]))[0]))
    cnorm = float(np.linalg.norm(c0))
    bs = float(np.sqrt(comm.sum(np.array([np.sum(R.finite_abs_max(lo, hi) ** 2)]))[0]))
    csn = float(np.linalg.norm(cs))
    w = csn / bs if (csn > 0 and bs > 0) else 1.0
    v = R.hash_unit_vector(n); v /= np.linalg.norm(v)
    prev = s = 0.0; it = 0
    while it < R.POWER_MAX:
        for _ in range(R.POWER_BATCH):
            v = comm.sum(KT @ (K @ v)); s2 = np.linalg.norm(v); v /= s2; it += 1
        s = np.sqrt(s2)
        if abs(s - prev) <= R.POWER_TOL * s: break
        prev = s
    eta = R.STEP_SAFETY / s
    sl = slice(j0, j1)                                   # column-side state lives on the slice only
    cs_s, l_s, u_s, d2_s = cs[sl], l[sl], u[sl], d2[sl]
    x0 = np.clip(np.zeros(j1 - j0), l_s, u_s); y0 = np.zeros(mg)
    aty0 = KcT @ np.concatenate(comm.gather(y0))

    def seed(x0, aty0, w):
        xh = np.clip(x0 - (eta / w) * (cs_s - aty0), l_s, u_s)
        return xh, np.concatenate(comm.gather(2.0 * xh - x0))        # xbar is kept whole
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
            kx = K @ xbar                                             # row step: this rank's rows
            t = kx - y / sigma
            yhat = sigma * (np.clip(t, lo, hi) - t)
            y = a * (2.0 * yhat - y) + (1.0 - a) * y0
            yhat_full = np.concatenate(comm.gather(yhat))             # exchange 1: yhat slices
            if first or last:
                xh_k = xhat_first if first else xhat_chk              # slice
                dx = xbar[sl] - xh_k
                if last:
                    kxh = K @ np.concatenate(comm.gather(xh_k))       # KKT pass gathers the whole candidate
            atyh = KcT @ yhat_full                                    # col step: this rank's columns
            x = a * xbar[sl] + (1.0 - a) * x0
            aty_new = a * (2.0 * atyh - aty) + (1.0 - a) * aty0
            xhat_next = np.clip(x - tau * (cs_s - aty_new), l_s, u_s)
            if first or last:
                dy2, dx2, dxd = comm.sum(np.array([np.sum((yhat - y_k) ** 2), dx @ dx, dx @ (atyh - aty)]))
                fp = np.sqrt(max((w / eta) * dx2 - 2.0 * dxd + dy2 / (eta * w), 0.0))
            aty = aty_new
            xbar = np.concatenate(comm.gather(2.0 * xhat_next - x))   # exchange 2: xbar slices
            if i == period - 2: xhat_chk = xhat_next
            if last: xhat_first = xhat_next
            k += 1; total += 1
            if first and k == 1: fp0 = fp_prev = fp
        r = cs_s - atyh; rp = np.maximum(r, 0); rm = np.maximum(-r, 0)
        viol = (np.where(np.isfinite(l_s), 0, rp) + np.where(np.isfinite(u_s), 0, rm)) / d2_s
        pres2, dy0, dobj, dres2, pobj, dx0 = comm.sum(np.array([
            np.sum(((kxh - np.clip(kxh, lo, hi)) / d1) ** 2), np.sum((yhat - y0) ** 2),
            np.where(np.isfinite(lo), lo, 0.0) @ np.maximum(yhat, 0) - np.where(np.isfinite(hi), hi, 0.0) @ np.maximum(-yhat, 0)
            + np.where(np.isfinite(l_s), l_s, 0) @ rp - np.where(np.isfinite(u_s), u_s, 0) @ rm,
            viol @ viol, cs_s @ xh_k, np.sum((xh_k - x0) ** 2)]))
        rel_p = np.sqrt(pres2) / (1 + bnorm); rel_d = np.sqrt(dres2) / (1 + cnorm)
        rel_g = abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj))
        if max(rel_p, rel_d, rel_g) <= eps:
            status = "optimal"; break
        restart = fp <= R.BETA_SUFFICIENT * fp0 or (fp <= R.BETA_NECESSARY * fp0 and fp > fp_prev) \
            or k >= R.BETA_ARTIFICIAL * total
        fp_prev = fp
        if restart:
            restarts += 1
            dxn = np.sqrt(dx0); dyn = np.sqrt(dy0)
            if dxn > R.WEIGHT_EPS and dyn > R.WEIGHT_EPS:
                w = np.exp(R.WEIGHT_SMOOTHING * np.log(dyn / dxn) + (1 - R.WEIGHT_SMOOTHING) * np.log(w))
            x0 = xh_k.copy(); y0 = yhat.copy(); aty0 = atyh.copy(); y = y0.copy(); aty = aty0.copy()
            xhat_first, xbar = seed(x0, aty0, w); k = 0
    x_full = np.concatenate(comm.gather(xh_k * d2_s))
    return dict(status=status, x=x_full, y_local=sgn * yhat * d1, obj=sgn * pobj, iters=total,
                restarts=restarts, rel_pres=rel_p, rel_dres=rel_d, rel_gap=rel_g)


# ---- 2-D (R x C) block partition: the next multi-GPU layout ("next") --------
class GridComm:
    """Collectives of an R x C process grid, rank = r * C + c, on top of ONE primitive
    `gather_world(obj) -> list over all ranks` (torch.distributed.all_gather_object in the tests).
    Sums run in rank order, so every member of a group computes bit-identical results."""

    def __init__(self, rank, R, C, gather_world):
        self.rank, self.R, self.C = rank, R, C
        self.r, self.c = divmod(rank, C)
        self._gather = gather_world
        self.received_bytes = 0          # data-path ingress of this rank (vectors only)

    def row_group(self, a, count=True):
        """what the C ranks of this rank's ROW group hold, ordered by c"""
        allv = self._gather(a)
        out = [allv[self.r * self.C + c] for c in range(self.C)]
        if count:
            self.received_bytes += sum(np.asarray(v).nbytes for i, v in enumerate(out) if i != self.c)
        return out

    def col_group(self, a, count=True):
        """what the R ranks of this rank's COLUMN group hold, ordered by r"""
        allv = self._gather(a)
        out = [allv[r * self.C + self.c] for r in range(self.R)]
        if count:
            self.received_bytes += sum(np.asarray(v).nbytes for i, v in enumerate(out) if i != self.r)
        return out

    def world_sum(self, a):
        return sum(self._gather(np.asarray(a, dtype=np.float64)))


def _sub(n, k, parts):
    return (n * k) // parts, (n * (k + 1)) // parts


def solve_grid(Ablk, c_c, lb_c, ub_c, lo_r, hi_r, comm: GridComm, maximize=False, eps=1e-6,
               max_iter=200000, period=64):
    """Rank (r, c) of an R x C grid holds the block K[rows of group r, columns of group c].
    Row step: partial product with the rank's n/C slice of xbar, reduce-scatter over the C ranks of
    the row group (each rank sums its sub-slice of the group's rows, in rank order), dual step on
    the sub-slice, all-gather of yhat inside the row group.  Column step: the mirror image over the
    R ranks of the column group.  Ingress per rank and iteration: 16 ((C-1) m + (R-1) n) / (R C)
    bytes against 8 (m + n) (RC - 1) / (RC) for the block-owner exchange (10 MB against 21 MB for
    config 3 on a 2 x 4 grid).  Check scalars are summed over the whole grid.

    Inputs are the rank's shares: Ablk (m_r x n_c), c / lb / ub of column group c, lo / hi of row
    group r.  Returns the rank's view (x of its column group, y of its row group)."""
    GR, GC, r, c = comm.R, comm.C, comm.r, comm.c
    sgn = -1.0 if maximize else 1.0
    K = Ablk.tocsr().copy().astype(np.float64)
    K.sort_indices()
    mr, nc = K.shape
    # -- Ruiz + Pock-Chambolle: a row's norm spans its row group, a column's its column group
    d1 = np.ones(mr); d2 = np.ones(nc)
    rows = np.repeat(np.arange(mr), np.diff(K.indptr))
    perm = np.argsort(K.indices, kind="stable")
    colptr = np.concatenate([[0], np.cumsum(np.bincount(K.indices, minlength=nc))])
    for sweep in range(R.RUIZ_PASSES + 1):
        a = np.abs(K.data)
        ruiz = sweep < R.RUIZ_PASSES
        uf = np.maximum if ruiz else np.add
        rn = _segment(uf, a, K.indptr)
        cn = _segment(uf, a[perm], colptr)
        rn = np.maximum.reduce(comm.row_group(rn, False)) if ruiz else sum(comm.row_group(rn, False))
        cn = np.maximum.reduce(comm.col_group(cn, False)) if ruiz else sum(comm.col_group(cn, False))
        rf = np.where(rn > 0, 1.0 / np.sqrt(np.where(rn > 0, rn, 1.0)), 1.0)
        cf = np.where(cn > 0, 1.0 / np.sqrt(np.where(cn > 0, cn, 1.0)), 1.0)
        K.data = (K.data * rf[rows]) * cf[K.indices]
        d1 *= rf
        d2 *= cf
    KT = K.T.tocsr(); KT.sort_indices()
    c0 = sgn * np.asarray(c_c, float)
    cs = c0 * d2; l = lb_c / d2; u = ub_c / d2; lo = lo_r * d1; hi = hi_r * d1
    # every row group / column group is counted once in the global norms
    once_row = 1.0 if c == 0 else 0.0
    once_col = 1.0 if r == 0 else 0.0
    bn2, bs2, cn2, cs2 = comm.world_sum(np.array([
        once_row * np.sum(R.finite_abs_max(lo_r, hi_r) ** 2), once_row * np.sum(R.finite_abs_max(lo, hi) ** 2),
        once_col * (c0 @ c0), once_col * (cs @ cs)]))
    bnorm, bs, cnorm, csn = np.sqrt(bn2), np.sqrt(bs2), np.sqrt(cn2), np.sqrt(cs2)
    w = csn / bs if (csn > 0 and bs > 0) else 1.0
    # -- step size: power iteration on K'K with both products reduced inside the groups
    col_sizes = comm.row_group(nc, False)                 # sizes of the column groups, ordered by c
    offset = int(sum(col_sizes[:c]))
    n_total = int(sum(col_sizes))
    v = R.hash_unit_vector(n_total)[offset:offset + nc].copy()
    v /= np.sqrt(comm.world_sum(np.array([once_col * (v @ v)]))[0])
    prev = s = 0.0; it = 0
    while it < R.POWER_MAX:
        for _ in range(R.POWER_BATCH):
            uvec = sum(comm.row_group(K @ v, False))
            v = sum(comm.col_group(KT @ uvec, False))
            s2 = np.sqrt(comm.world_sum(np.array([once_col * (v @ v)]))[0]); v /= s2; it += 1
        s = np.sqrt(s2)
        if abs(s - prev) <= R.POWER_TOL * s: break
        prev = s
    eta = R.STEP_SAFETY / s
    # -- ownership inside the groups: the rank's sub-slice of its row group's rows / column group's columns
    i0, i1 = _sub(mr, c, GC)           # rows of group r this rank steps
    j0, j1 = _sub(nc, r, GR)           # columns of group c this rank steps
    rs, csl = slice(i0, i1), slice(j0, j1)
    cs_s, l_s, u_s, d2_s = cs[csl], l[csl], u[csl], d2[csl]
    lo_s, hi_s, d1_s = lo[rs], hi[rs], d1[rs]

    # reduce-scatter: only the rank's sub-slice of each peer's partial vector has to travel (the
    # in-process transport of this restatement hands over whole vectors; the byte count is the
    # protocol's)
    def reduce_rows(partial):          # over the row group: my sub-slice of the group's rows
        comm.received_bytes += (GC - 1) * 8 * (i1 - i0)
        return sum(p[rs] for p in comm.row_group(partial, False))
    def reduce_cols(partial):
        comm.received_bytes += (GR - 1) * 8 * (j1 - j0)
        return sum(p[csl] for p in comm.col_group(partial, False))
    def gather_rows(sub):              # all-gather inside the row group: the whole y-like vector of group r
        return np.concatenate(comm.row_group(sub))
    def gather_cols(sub):
        return np.concatenate(comm.col_group(sub))

    x0 = np.clip(np.zeros(j1 - j0), l_s, u_s); y0 = np.zeros(i1 - i0)
    aty0 = reduce_cols(KT @ gather_rows(y0))

    def seed(x0, aty0, w):
        xh = np.clip(x0 - (eta / w) * (cs_s - aty0), l_s, u_s)
        return xh, gather_cols(2.0 * xh - x0)             # xbar of the column group, whole
    xhat_first, xbar = seed(x0, aty0, w)
    y = y0.copy(); aty = aty0.copy(); k = total = restarts = 0
    fp0 = fp_prev = None; xhat_chk = xhat_first
    status = "iteration_limit"
    comm.received_bytes = 0
    iters_counted = 0
    while total < max_iter:
        tau, sigma = eta / w, eta * w
        for i in range(period):
            a = (k + 1.0) / (k + 2.0)
            first, last = i == 0, i == period - 1
            y_k = y
            kx = reduce_rows(K @ xbar)                                  # row step: partial, reduce-scatter
            t = kx - y / sigma
            yhat = sigma * (np.clip(t, lo_s, hi_s) - t)
            y = a * (2.0 * yhat - y) + (1.0 - a) * y0
            yhat_grp = gather_rows(yhat)                                # all-gather inside the row group
            if first or last:
                xh_k = xhat_first if first else xhat_chk
                dx = xbar[csl] - xh_k
                if last:
                    kxh = reduce_rows(K @ gather_cols(xh_k))
            atyh = reduce_cols(KT @ yhat_grp)                           # column step: partial, reduce-scatter
            x = a * xbar[csl] + (1.0 - a) * x0
            aty_new = a * (2.0 * atyh - aty) + (1.0 - a) * aty0
            xhat_next = np.clip(x - tau * (cs_s - aty_new), l_s, u_s)
            if first or last:
                dy2, dx2, dxd = comm.world_sum(np.array([np.sum((yhat - y_k) ** 2), dx @ dx, dx @ (atyh - aty)]))
                fp = np.sqrt(max((w / eta) * dx2 - 2.0 * dxd + dy2 / (eta * w), 0.0))
            aty = aty_new
            xbar = gather_cols(2.0 * xhat_next - x)                     # all-gather inside the column group
            if i == period - 2: xhat_chk = xhat_next
            if last: xhat_first = xhat_next
            k += 1; total += 1; iters_counted += 1
            if first and k == 1: fp0 = fp_prev = fp
        rr = cs_s - atyh; rp = np.maximum(rr, 0); rm = np.maximum(-rr, 0)
        viol = (np.where(np.isfinite(l_s), 0, rp) + np.where(np.isfinite(u_s), 0, rm)) / d2_s
        pres2, dy0, dobj, dres2, pobj, dx0 = comm.world_sum(np.array([
            np.sum(((kxh - np.clip(kxh, lo_s, hi_s)) / d1_s) ** 2), np.sum((yhat - y0) ** 2),
            np.where(np.isfinite(lo_s), lo_s, 0.0) @ np.maximum(yhat, 0) - np.where(np.isfinite(hi_s), hi_s, 0.0) @ np.maximum(-yhat, 0)
            + np.where(np.isfinite(l_s), l_s, 0) @ rp - np.where(np.isfinite(u_s), u_s, 0) @ rm,
            viol @ viol, cs_s @ xh_k, np.sum((xh_k - x0) ** 2)]))
        rel_p = np.sqrt(pres2) / (1 + bnorm); rel_d = np.sqrt(dres2) / (1 + cnorm)
        rel_g = abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj))
        if max(rel_p, rel_d, rel_g) <= eps:
            status = "optimal"; break
        restart = fp <= R.BETA_SUFFICIENT * fp0 or (fp <= R.BETA_NECESSARY * fp0 and fp > fp_prev) \
            or k >= R.BETA_ARTIFICIAL * total
        fp_prev = fp
        if restart:
            restarts += 1
            dxn = np.sqrt(dx0); dyn = np.sqrt(dy0)
            if dxn > R.WEIGHT_EPS and dyn > R.WEIGHT_EPS:
                w = np.exp(R.WEIGHT_SMOOTHING * np.log(dyn / dxn) + (1 - R.WEIGHT_SMOOTHING) * np.log(w))
            x0 = xh_k.copy(); y0 = yhat.copy(); aty0 = atyh.copy(); y = y0.copy(); aty = aty0.copy()
            xhat_first, xbar = seed(x0, aty0, w); k = 0
    ingress = comm.received_bytes / max(iters_counted, 1)
    x_grp = gather_cols(xh_k * d2_s)
    y_grp = gather_rows(sgn * yhat * d1_s)
    return dict(status=status, x_group=x_grp, y_group=y_grp, obj=sgn * pobj, iters=total, restarts=restarts,
                rel_pres=rel_p, rel_dres=rel_d, rel_gap=rel_g, ingress_bytes_per_iter=ingress)
