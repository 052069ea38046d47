"""Scratch: iteration counts of restart / primal-weight variants on the CPU oracle (not shipped)."""
import sys, time, itertools
import numpy as np
from multiprocessing import Pool
from oracle import gen, pdhg_ref as P


def solve(lpm, eps=1e-6, max_iter=60000, period=64, rule="pdlp", theta=0.5, kp=0.99, ki=0.96, kd=0.0,
          bs=0.2, bn=0.8, ba=0.36, S=None, reflect=1.0):
    S = S or P.prepare(lpm)
    n, m = lpm.n, lpm.m
    w, eta = S.w0, S.eta
    x0 = np.clip(np.zeros(n), S.l, S.u); y0 = np.zeros(m); aty0 = S.KT @ y0
    def restart_apply(x0, aty0, w):
        tau = eta / w
        xhat = np.clip(x0 - tau * (S.c - aty0), S.l, S.u)
        return xhat, 2.0 * xhat - x0
    xhat_first, xbar = restart_apply(x0, aty0, w)
    y = y0.copy(); aty = aty0.copy()
    k = total = restarts = 0
    fp0 = fp_prev = None
    esum = 0.0; eprev = 0.0
    xhat_chk = xhat_first
    while total < max_iter:
        tau, sigma = eta / w, eta * w
        for i in range(period):
            a = (k + 1.0) / (k + 2.0)
            first, last = i == 0, i == period - 1
            y_k = y
            yhat, y, kx_k = P.row_step(S, xbar, y, y0, sigma, a)
            if first or last:
                xh_k = xhat_first if first else xhat_chk
                dx = xbar - xh_k; dy = yhat - y_k
                if last: kxhat = S.K @ xh_k
            atyh, aty_new, xhat_next, xbar_new = P.col_step(S, yhat, xbar, x0, aty, aty0, tau, a)
            if first or last:
                daty = atyh - aty
                fp = np.sqrt(max((w / eta) * (dx @ dx) - 2.0 * (dx @ daty) + (dy @ dy) / (eta * w), 0.0))
            aty, xbar = aty_new, xbar_new
            if i == period - 2: xhat_chk = xhat_next
            if last: xhat_first = xhat_next
            k += 1; total += 1
            if first and k == 1: fp0 = fp_prev = fp
        rel_p, rel_d, rel_g, pobj, dobj = P.kkt_scaled(S, xh_k, yhat, atyh, kxhat)
        if max(rel_p, rel_d, rel_g) <= eps:
            return dict(status="optimal", iters=total, restarts=restarts, obj=S.sgn * pobj)
        do_restart = (fp <= bs * fp0 or (fp <= bn * fp0 and fp > fp_prev) or k >= ba * total)
        fp_prev = fp
        if do_restart:
            restarts += 1
            dxn = np.linalg.norm(xh_k - x0); dyn = np.linalg.norm(yhat - y0)
            if dxn > 1e-10 and dyn > 1e-10:
                if rule == "pdlp":
                    w = np.exp(theta * np.log(dyn / dxn) + (1 - theta) * np.log(w))
                else:   # PID on e = log(w dx / dy)
                    e = np.log(w * dxn / dyn)
                    esum = ki * esum + e if rule == "pid_leaky" else esum + e
                    w = np.exp(np.log(w) - (kp * e + (ki * esum if rule == "pid" else 0.0) + kd * (e - eprev)))
                    eprev = e
            x0, y0, aty0 = xh_k.copy(), yhat.copy(), atyh.copy()
            y, aty = y0.copy(), aty0.copy()
            xhat_first, xbar = restart_apply(x0, aty0, w)
            k = 0
    return dict(status="limit", iters=total, restarts=restarts, obj=S.sgn * pobj)


PROBLEMS = {
    "C1": lambda: gen.config(1),
    "lp300": lambda: gen.lp(300, 500, 6, 7),
    "mid": lambda: gen.lp_fast(10000, 20000, 8, 1),
    "sc400": lambda: gen.set_cover(400, 1000, 8, 20264),
    "sc2k": lambda: gen.set_cover(2000, 5000, 8, 20264),
    "C2": lambda: gen.config(2),
}

def run(job):
    name, kw = job
    L = PROBLEMS[name]()
    t = time.time()
    r = solve(L, **kw)
    return name, kw, r["status"], r["iters"], r["restarts"], round(time.time() - t, 1)

if __name__ == "__main__":
    names = sys.argv[1].split(",")
    variants = [
        dict(),
        dict(period=32), dict(period=16),
        dict(theta=0.8), dict(theta=1.0), dict(theta=0.3),
        dict(rule="p", kp=0.99), dict(rule="pid", kp=0.99, ki=0.96),
        dict(rule="pid", kp=0.5, ki=0.2), dict(rule="pid", kp=0.8, ki=0.1),
        dict(bs=0.1), dict(bs=0.3, bn=0.9), dict(ba=0.5), dict(ba=0.25),
    ]
    jobs = [(nm, v) for nm in names for v in variants]
    with Pool(8) as pool:
        for res in pool.imap(run, jobs):
            print(*res, flush=True)
