"""Scratch: node counts of a best-first B&B with / without implied-bound (x_j <= z_g) rows, using
HiGHS as the node LP solver (CPU).  Decides whether strengthening is worth building."""
import sys, heapq, time
import numpy as np, scipy.sparse as sp
from scipy.optimize import linprog
sys.path.insert(0, ".")
from oracle import gen

def bnb(L, extra=None, node_limit=20000, rcfix=False, start_best=np.inf):
    A = L.A.tocsr(); lo, hi = L.row_lo.copy(), L.row_hi.copy()
    if extra is not None:
        A = sp.vstack([A, extra[0]]).tocsr(); lo = np.concatenate([lo, extra[1]]); hi = np.concatenate([hi, extra[2]])
    sgn = -1.0 if L.maximize else 1.0
    c = sgn * L.c
    n = L.n
    fin_hi, fin_lo = np.isfinite(hi) & (hi < 1e19), np.isfinite(lo) & (lo > -1e19)
    Aub = sp.vstack([A[fin_hi], -A[fin_lo]]).tocsr(); bub = np.concatenate([hi[fin_hi], -lo[fin_lo]])
    best, nodes, lps = start_best, 0, 0
    heap = [(-np.inf, 0, L.lb.copy(), L.ub.copy())]
    cnt = 1
    while heap and nodes < node_limit:
        bound, _, lb, ub = heapq.heappop(heap)
        if bound >= best - 1 + 1e-6: continue
        nodes += 1
        r = linprog(c, A_ub=Aub, b_ub=bub, bounds=np.stack([lb, ub], 1), method="highs-ds")
        lps += 1
        if r.status != 0: continue
        nb = np.ceil(r.fun - 1e-6)
        if nb >= best - 1 + 1e-6: continue
        x = r.x
        for cand in (np.round(x), np.floor(x + 1e-6)):
            cand = np.clip(cand, lb, ub)
            ax = A @ cand
            if np.all(ax <= hi + 1e-9) and np.all(ax >= lo - 1e-9): best = min(best, c @ cand)
        if rcfix and np.isfinite(best):
            rc = r.lower.marginals + r.upper.marginals      # reduced costs (>= 0 at lb, <= 0 at ub)
            gap = best - 1 + 1e-6 - r.fun
            fix0 = (rc > gap + 1e-9) & (x <= lb + 1e-9) & (ub > lb)
            fix1 = (-rc > gap + 1e-9) & (x >= ub - 1e-9) & (ub > lb)
            ub = np.where(fix0, lb, ub); lb = np.where(fix1, ub, lb)
        frac = np.abs(x - np.round(x))
        j = int(np.argmax(frac))
        if frac[j] < 1e-6:
            best = min(best, c @ np.round(x)); continue
        for side in (0, 1):
            l2, u2 = lb.copy(), ub.copy()
            if side == 0: u2[j] = np.floor(x[j])
            else: l2[j] = np.floor(x[j]) + 1
            heapq.heappush(heap, (nb, cnt, l2, u2)); cnt += 1
    return sgn * best, nodes, lps, len(heap)

def vub_rows(L, m_knap, n, G):
    grp = np.arange(n) % G
    rows = np.concatenate([np.arange(n), np.arange(n)]); cols = np.concatenate([np.arange(n), n + grp])
    vals = np.concatenate([np.ones(n), -np.ones(n)])
    E = sp.csr_matrix((vals, (rows, cols)), shape=(n, n + G))
    return E, np.full(n, -np.inf), np.zeros(n)

for (m, n, k, G, want) in ((8, 80, 4, 4, 1085), (10, 100, 4, 5, 1526)):
    L = gen.mkp(m, n, k, G, 20265)
    sg = -1.0
    for tag, kw in (("vub+rc", dict(rcfix=True)), ("vub+rc+opt", dict(rcfix=True, start_best=sg*want)), ("vub+opt", dict(start_best=sg*want))):
        t = time.time()
        obj, nodes, lps, left = bnb(L, vub_rows(L, m, n, G), node_limit=40000, **kw)
        print(f"mkp {m}x{n}: {tag:10s} obj {obj} (want {want}) nodes {nodes} open {left} {time.time()-t:.1f}s", flush=True)
