"""Scratch prototype to compare PDHG variants (not shipped; see pdhg_ref.py for the oracle)."""
import sys, time
import numpy as np
import scipy.sparse as sp
from oracle import gen, highs_ref

INF = np.inf


def scale_problem(A, ruiz_iters=10, pc=True):
    m, n = A.shape
    K = A.copy().tocsr()
    d1 = np.ones(m); d2 = np.ones(n)
    for _ in range(ruiz_iters):
        absK = abs(K)
        rmax = np.asarray(absK.max(axis=1).todense()).ravel()
        cmax = np.asarray(absK.max(axis=0).todense()).ravel()
        r = np.where(rmax > 0, 1 / np.sqrt(rmax), 1.0)
        c = np.where(cmax > 0, 1 / np.sqrt(cmax), 1.0)
        K = sp.diags(r) @ K @ sp.diags(c)
        d1 *= r; d2 *= c
    if pc:
        absK = abs(K)
        rs = np.asarray(absK.sum(axis=1)).ravel()
        cs = np.asarray(absK.sum(axis=0)).ravel()
        r = np.where(rs > 0, 1 / np.sqrt(rs), 1.0)
        c = np.where(cs > 0, 1 / np.sqrt(cs), 1.0)
        K = sp.diags(r) @ K @ sp.diags(c)
        d1 *= r; d2 *= c
    return K.tocsr(), d1, d2


def kkt(A, AT, c, lb, ub, lo, hi, x, y, bnorm, cnorm):
    ax = A @ x
    pres = np.linalg.norm(ax - np.clip(ax, lo, hi))
    r = c - AT @ y
    rp = np.maximum(r, 0); rm = np.maximum(-r, 0)
    viol = np.where(np.isfinite(lb), 0, rp) + np.where(np.isfinite(ub), 0, rm)
    dres = np.linalg.norm(viol)
    yp = np.maximum(y, 0); ym = np.maximum(-y, 0)
    dobj = (np.where(np.isfinite(lo), lo, 0) @ yp - np.where(np.isfinite(hi), hi, 0) @ ym
            + np.where(np.isfinite(lb), lb, 0) @ rp - np.where(np.isfinite(ub), ub, 0) @ rm)
    pobj = c @ x
    gap = abs(pobj - dobj)
    return pres / (1 + bnorm), dres / (1 + cnorm), gap / (1 + abs(pobj) + abs(dobj)), pobj, dobj


def power_iter(K, KT, iters=50, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(K.shape[1]); v /= np.linalg.norm(v)
    s = 0
    for _ in range(iters):
        u = K @ v
        v = KT @ u
        s = np.linalg.norm(v)
        v /= s
    return np.sqrt(s)


def solve(L, variant="halpern", eps=1e-6, max_iter=200000, check=64, restart_to="hat", verbose=True,
          reflect=1.0, eval_at="hat"):
    A = L.A.tocsr(); AT = A.T.tocsr()
    sgn = -1.0 if L.maximize else 1.0
    c0 = sgn * L.c
    K, d1, d2 = scale_problem(A)
    KT = K.T.tocsr()
    c = c0 * d2; lb = L.lb / d2; ub = L.ub / d2; lo = L.row_lo * d1; hi = L.row_hi * d1
    m, n = K.shape
    bvec = np.where(np.isfinite(L.row_lo), np.abs(L.row_lo), 0)
    bvec = np.maximum(bvec, np.where(np.isfinite(L.row_hi), np.abs(L.row_hi), 0))
    bnorm = np.linalg.norm(bvec); cnorm = np.linalg.norm(c0)
    bs = np.where(np.isfinite(lo), np.abs(lo), 0)
    bs = np.maximum(bs, np.where(np.isfinite(hi), np.abs(hi), 0))
    bsn = np.linalg.norm(bs); csn = np.linalg.norm(c)
    w = csn / bsn if (csn > 0 and bsn > 0) else 1.0
    x = np.clip(np.zeros(n), lb, ub); y = np.zeros(m)

    def dualprox(y, kxbar, sigma):
        t = kxbar - y / sigma
        return sigma * (np.clip(t, lo, hi) - t)

    def check_term(xs, ys):
        return kkt(A, AT, c0, L.lb, L.ub, L.row_lo, L.row_hi, xs * d2, ys * d1, bnorm, cnorm)

    t0 = time.time()
    total = 0
    nrestart = 0
    if variant == "halpern":
        eta = 0.998 / power_iter(K, KT)
        x0 = x.copy(); y0 = y.copy()
        aty = KT @ y
        k = 0
        r0 = None; rprev = None
        while total < max_iter:
            tau = eta / w; sigma = eta * w
            xh = np.clip(x - tau * (c - aty), lb, ub)
            kxb = K @ (2 * xh - x)
            yh = dualprox(y, kxb, sigma)
            atyh = KT @ yh
            total += 1
            do_check = (total % check == 0) or k == 0
            if do_check:
                dx = xh - x; dy = yh - y
                fp = np.sqrt(max(w / eta * (dx @ dx) - 2 * (dx @ (atyh - aty)) + (dy @ dy) / (eta * w), 0))
                if k == 0:
                    r0 = fp; rprev = fp
            # halpern update
            a = (k + 1) / (k + 2)
            xn = a * ((1 + reflect) * xh - reflect * x) + (1 - a) * x0
            yn = a * ((1 + reflect) * yh - reflect * y) + (1 - a) * y0
            x, y = xn, yn
            aty = KT @ y   # (prototype: recompute; product code combines linearly)
            k += 1
            if total % check == 0:
                xe, ye = (xh, yh) if eval_at == "hat" else (x, y)
                rp, rd, rg, pobj, dobj = check_term(xe, ye)
                if verbose and (total % (check * 16) == 0):
                    print(f"it {total:7d} k {k:6d} pres {rp:.2e} dres {rd:.2e} gap {rg:.2e} pobj {sgn*pobj:.8f} w {w:.3e} fp {fp:.2e}")
                if max(rp, rd, rg) <= eps:
                    return dict(iters=total, restarts=nrestart, obj=sgn * pobj, dobj=sgn * dobj, rp=rp, rd=rd, rg=rg,
                                secs=time.time() - t0, x=xe * d2, y=ye * d1)
                restart = False
                if fp <= 0.2 * r0: restart = True
                elif fp <= 0.8 * r0 and fp > rprev: restart = True
                elif k >= 0.36 * total: restart = True
                rprev = fp
                if restart:
                    nrestart += 1
                    if restart_to == "hat":
                        xr, yr = xh, yh
                    else:
                        xr, yr = x, y
                    dxn = np.linalg.norm(xr - x0); dyn = np.linalg.norm(yr - y0)
                    if dxn > 1e-10 and dyn > 1e-10:
                        w = np.exp(0.5 * np.log(dyn / dxn) + 0.5 * np.log(w))
                    x, y = xr.copy(), yr.copy()
                    x0 = x.copy(); y0 = y.copy()
                    aty = KT @ y
                    k = 0
    else:  # PDLP adaptive, averaged
        eta = 1.0 / abs(K).max()
        aty = KT @ y
        xs = np.zeros(n); ys = np.zeros(m); ws = 0.0
        xr0 = x.copy(); yr0 = y.copy()
        k = 0

        def kkt_err_scaled(xx, yy, w):
            ax = K @ xx
            pr = ax - np.clip(ax, lo, hi)
            r = c - KT @ yy
            rp_ = np.maximum(r, 0); rm_ = np.maximum(-r, 0)
            viol = np.where(np.isfinite(lb), 0, rp_) + np.where(np.isfinite(ub), 0, rm_)
            yp = np.maximum(yy, 0); ym = np.maximum(-yy, 0)
            dobj = (np.where(np.isfinite(lo), lo, 0) @ yp - np.where(np.isfinite(hi), hi, 0) @ ym
                    + np.where(np.isfinite(lb), lb, 0) @ rp_ - np.where(np.isfinite(ub), ub, 0) @ rm_)
            gap = c @ xx - dobj
            return np.sqrt(w * (pr @ pr) + (viol @ viol) / w + gap * gap)
        err_last_restart = None; err_prev_cand = None
        while total < max_iter:
            while True:
                tau = eta / w; sigma = eta * w
                xn = np.clip(x - tau * (c - aty), lb, ub)
                kxb = K @ (2 * xn - x)
                yn = dualprox(y, kxb, sigma)
                atyn = KT @ yn
                total += 1
                dx = xn - x; dy = yn - y
                mov = 0.5 * w * (dx @ dx) + 0.5 * (dy @ dy) / w
                inter = abs(dx @ (atyn - aty))
                lim = mov / inter if inter > 0 else np.inf
                eta_used = eta
                eta = min((1 - (total + 1) ** -0.3) * lim, (1 + (total + 1) ** -0.6) * eta)
                if eta_used <= lim:
                    break
            x, y, aty = xn, yn, atyn
            xs += eta_used * x; ys += eta_used * y; ws += eta_used
            k += 1
            if total % check < 1 or (total % check) < 0:
                pass
            if k % check == 0:
                xa = xs / ws; ya = ys / ws
                rp, rd, rg, pobj, dobj = check_term(x, y)
                rpa, rda, rga, pobja, dobja = check_term(xa, ya)
                if verbose and (k % (check * 16) == 0):
                    print(f"it {total:7d} k {k:6d} cur {rp:.1e} {rd:.1e} {rg:.1e} avg {rpa:.1e} {rda:.1e} {rga:.1e} pobj {sgn*pobja:.8f} w {w:.3e} eta {eta:.2e}")
                if max(rp, rd, rg) <= eps:
                    return dict(iters=total, restarts=nrestart, obj=sgn * pobj, rp=rp, rd=rd, rg=rg, secs=time.time() - t0, x=x*d2, y=y*d1)
                if max(rpa, rda, rga) <= eps:
                    return dict(iters=total, restarts=nrestart, obj=sgn * pobja, rp=rpa, rd=rda, rg=rga, secs=time.time() - t0, x=xa*d2, y=ya*d1)
                ec = kkt_err_scaled(x, y, w); ea = kkt_err_scaled(xa, ya, w)
                if ec < ea: cand, ecand = (x, y), ec
                else: cand, ecand = (xa, ya), ea
                if err_last_restart is None:
                    err_last_restart = kkt_err_scaled(xr0, yr0, w)
                restart = False
                if ecand <= 0.2 * err_last_restart: restart = True
                elif ecand <= 0.8 * err_last_restart and err_prev_cand is not None and ecand > err_prev_cand: restart = True
                elif k >= 0.36 * total: restart = True
                err_prev_cand = ecand
                if restart:
                    nrestart += 1
                    xr, yr = cand
                    dxn = np.linalg.norm(xr - xr0); dyn = np.linalg.norm(yr - yr0)
                    if dxn > 1e-10 and dyn > 1e-10:
                        w = np.exp(0.5 * np.log(dyn / dxn) + 0.5 * np.log(w))
                    x = xr.copy(); y = yr.copy(); aty = KT @ y
                    xr0 = x.copy(); yr0 = y.copy()
                    xs[:] = 0; ys[:] = 0; ws = 0.0; k = 0
                    err_last_restart = kkt_err_scaled(x, y, w); err_prev_cand = None
    return dict(iters=total, restarts=nrestart, obj=None, secs=time.time() - t0)


if __name__ == "__main__":
    which = sys.argv[1]
    variant = sys.argv[2]
    kw = {}
    for a in sys.argv[3:]:
        k_, v_ = a.split("=")
        try: v_ = float(v_)
        except ValueError: pass
        kw[k_] = v_
    if which == "c1": L = gen.config(1)
    elif which == "c2": L = gen.config(2)
    elif which == "mid": L = gen.lp(10000, 20000, 8, 1)
    elif which == "sc": L = gen.set_cover(400, 1000, 8, 20264); L.integer = None
    elif which == "mkp": L = gen.mkp(10, 100, 4, 5, 20265); L.integer = None
    kw.setdefault("verbose", True); r = solve(L, variant, **kw)
    r.pop("x", None); r.pop("y", None)
    print(r)
