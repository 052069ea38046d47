"""Host-side partitioning for the row-sharded solve: rank g of G owns the
contiguous row block [g*m/G, (g+1)*m/G) of A (its CSR rows and the matching slices of row_lo /
row_hi); the column vectors c, lb, ub are replicated.  Used by bench.py and by the tests."""
from __future__ import annotations

import numpy as np


def row_block(m: int, rank: int, world: int) -> tuple[int, int]:
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (m * rank) // world, (m * (rank + 1)) // world


def shard_csr(A, row_lo, row_hi, rank: int, world: int):
    """(indptr, indices, data, row_lo, row_hi) of this rank's row block; indptr re-based to 0."""
    A = A.tocsr()
    r0, r1 = row_block(A.shape[0], rank, world)
    s, e = A.indptr[r0], A.indptr[r1]
    indptr = (A.indptr[r0:r1 + 1] - s).astype(np.int64)
    return (indptr, A.indices[s:e].astype(np.int32), np.asarray(A.data[s:e], dtype=np.float64),
            np.asarray(row_lo[r0:r1], dtype=np.float64), np.asarray(row_hi[r0:r1], dtype=np.float64))
