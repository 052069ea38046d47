"""Seeded synthetic instance generators for BASELINE.json's configs (TEST INFRASTRUCTURE).

This file is part of ``oracle/``: it is test infrastructure, not product code.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` may import it (it makes the
synthetic inputs; it computes nothing the product ships).

Draw order follows BASELINE.md section 5 exactly, so the golden objectives recorded there
(C1 -> -698.9207296656, set cover 200x500 -> 382, MKP 5x60 -> 875 ...) reproduce.

Canonical LP form used everywhere in this repo (and by ``include/mipcu.h``)::

    min / max  c'x   s.t.   row_lo <= A x <= row_hi,   lb <= x <= ub

which is what the reference's adapters lower a miplib row ``sum a_j x_j + k {<=,=} 0``
into: ``row_hi = -k``, ``row_lo = (Equal ? -k : -inf)`` (reference
miplib/scip/solver.cpp:108-120, miplib/lpsolve/solver.cpp:115-130).
"""
from __future__ import annotations

import dataclasses
import numpy as np
import scipy.sparse as sp

INF = float("inf")


@dataclasses.dataclass
class LP:
    """A model in canonical row-range form; ``integer`` marks columns that must be integral."""
    A: sp.csr_matrix            # m x n, sorted column indices inside each row
    c: np.ndarray               # n
    lb: np.ndarray              # n
    ub: np.ndarray              # n
    row_lo: np.ndarray          # m  (-inf allowed)
    row_hi: np.ndarray          # m  (+inf allowed)
    maximize: bool = False
    integer: np.ndarray | None = None   # bool[n] or None (pure LP)
    name: str = ""

    @property
    def m(self) -> int:
        return self.A.shape[0]

    @property
    def n(self) -> int:
        return self.A.shape[1]

    @property
    def nnz(self) -> int:
        return int(self.A.nnz)


def _random_pattern(rng: np.random.Generator, m: int, n: int, k: int):
    """Step 1 of every recipe: for each column j, k distinct rows uniform in [0, m)."""
    rows = np.empty(n * k, dtype=np.int64)
    for j in range(n):
        rows[j * k:(j + 1) * k] = rng.choice(m, k, replace=False)
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    return rows, cols


def _canonical_csr(vals, rows, cols, m, n) -> sp.csr_matrix:
    A = sp.csr_matrix((vals, (rows, cols)), shape=(m, n))
    A.sum_duplicates()
    A.sort_indices()
    return A


def lp(m: int, n: int, k: int, seed: int, name: str = "") -> LP:
    """LP(m, n, k, seed) of BASELINE.md section 5 (configs C1, C2, C3): packing-type, feasible
    (x0 is a witness), bounded by the box 0 <= x <= 1; first m//4 rows are equalities."""
    rng = np.random.default_rng(seed)
    rows, cols = _random_pattern(rng, m, n, k)
    vals = rng.uniform(0.1, 1.0, n * k)
    A = _canonical_csr(vals, rows, cols, m, n)
    x0 = rng.uniform(0, 1, n)
    ax = A @ x0
    meq = m // 4
    s = rng.uniform(0.05, 0.5, m - meq)
    c = -rng.uniform(0.1, 1, n)
    row_hi = ax.copy()
    row_hi[meq:] += s
    row_lo = np.full(m, -INF)
    row_lo[:meq] = ax[:meq]
    return LP(A, c, np.zeros(n), np.ones(n), row_lo, row_hi, False, None,
              name or f"lp_{m}x{n}_k{k}_s{seed}")


def set_cover(m: int, n: int, k: int, seed: int, name: str = "") -> LP:
    """SetCover(m, n, k, seed) (config C4): min c'x, A x >= 1, x binary, integer costs."""
    rng = np.random.default_rng(seed)
    rows, cols = _random_pattern(rng, m, n, k)
    A = _canonical_csr(np.ones(n * k), rows, cols, m, n).tolil()
    counts = np.diff(A.tocsr().indptr)
    for i in np.flatnonzero(counts == 0):
        A[i, rng.integers(n)] = 1.0
    A = A.tocsr()
    A.sort_indices()
    c = rng.integers(1, 101, n).astype(np.float64)
    return LP(A, c, np.zeros(n), np.ones(n), np.ones(m), np.full(m, INF), False,
              np.ones(n, dtype=bool), name or f"setcover_{m}x{n}_k{k}_s{seed}")


def mkp(m: int, n: int, k: int, G: int, seed: int, name: str = "") -> LP:
    """MKP(m, n, k, G, seed) (config C5): multi-dimensional knapsack with G group switches
    z_g added through indicator constraints ``(z_g == 0) >> (sum_{j in g} x_j <= 0)``, which the
    reference reformulates (miplib/constr.cpp:225-271) to ``sum_{j in g} x_j - |g| z_g <= 0``.
    Columns are [x_0..x_{n-1}, z_0..z_{G-1}]; maximise p'x - f'z."""
    k = min(k, m)
    rng = np.random.default_rng(seed)
    rows, cols = _random_pattern(rng, m, n, k)
    w = rng.integers(1, 101, n * k).astype(np.float64)
    W = _canonical_csr(w, rows, cols, m, n)
    cap = np.floor(0.25 * np.asarray(W.sum(axis=1)).ravel())
    p = rng.integers(1, 101, n).astype(np.float64)
    f = rng.integers(50, 201, G).astype(np.float64)
    grp = np.arange(n) % G
    gsize = np.bincount(grp, minlength=G).astype(np.float64)
    # indicator rows after big-M reformulation: sum_{j in g} x_j - |g| z_g <= 0
    ind_rows = np.concatenate([grp, np.arange(G)])
    ind_cols = np.concatenate([np.arange(n), n + np.arange(G)])
    ind_vals = np.concatenate([np.ones(n), -gsize])
    top = sp.hstack([W, sp.csr_matrix((m, G))]).tocsr()
    bot = sp.csr_matrix((ind_vals, (ind_rows, ind_cols)), shape=(G, n + G))
    A = sp.vstack([top, bot]).tocsr()
    A.sort_indices()
    c = np.concatenate([p, -f])
    N = n + G
    return LP(A, c, np.zeros(N), np.ones(N), np.full(m + G, -INF),
              np.concatenate([cap, np.zeros(G)]), True, np.ones(N, dtype=bool),
              name or f"mkp_{m}x{n}_k{k}_G{G}_s{seed}")


# BASELINE.json configs, by number (1-based as in SURVEY.md section 8d)
CONFIGS = {
    1: lambda: lp(1000, 2000, 8, 20261, "C1"),
    2: lambda: lp(100000, 200000, 8, 20262, "C2"),
    3: lambda: lp(1000000, 2000000, 16, 20263, "C3"),
    4: lambda: set_cover(20000, 50000, 8, 20264, "C4"),
    5: lambda: mkp(500, 100000, 8, 100, 20265, "C5"),
}


def config(i: int) -> LP:
    return CONFIGS[i]()


def lp_fast(m: int, n: int, k: int, seed: int, name: str = "") -> LP:
    """Same family as :func:`lp` but with a vectorised pattern draw (rows drawn with replacement
    then de-duplicated), for throw-away large instances in property tests.  NOT the BASELINE
    recipe: no golden value is attached to it."""
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, m, n * k)
    cols = np.repeat(np.arange(n, dtype=np.int64), k)
    vals = rng.uniform(0.1, 1.0, n * k)
    A = sp.csr_matrix((vals, (rows, cols)), shape=(m, n))
    A.sum_duplicates()
    A.sort_indices()
    x0 = rng.uniform(0, 1, n)
    ax = A @ x0
    meq = m // 4
    s = rng.uniform(0.05, 0.5, m - meq)
    c = -rng.uniform(0.1, 1, n)
    row_hi = ax.copy()
    row_hi[meq:] += s
    row_lo = np.full(m, -INF)
    row_lo[:meq] = ax[:meq]
    return LP(A, c, np.zeros(n), np.ones(n), row_lo, row_hi, False, None,
              name or f"lpfast_{m}x{n}_k{k}_s{seed}")
