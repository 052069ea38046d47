"""Requests per nonzero of the lane-per-row sliced-ELL gather, computed exactly on the host.

The fused SpMV kernels (miplib_b200/csrc/kernels.cuh, TPR = 0) give lane l of a warp row l of a 32-row slice; at
step k the warp gathers the k-th entry of its 32 rows, and the L1 sends one request per DISTINCT 128-byte line
(16 doubles) among them - the unit that binds these kernels (l1tex__m_l1tex2xbar_req_cycles_active 85 - 87 %).
This script rebuilds the slices the way setup.cu does (rows sorted by decreasing length, consecutive groups of
32) and counts distinct lines per (slice, k), for the matrix as generated and after two load-time orderings:
  * "bin": rows sorted by (length, smallest column), columns untouched - brings the first entries of a slice's
    rows into neighbouring lines;
  * "rcm": reverse Cuthill-McKee on the bipartite graph (scipy), rows and columns both permuted - the classic
    bandwidth-reducing ordering (only run below --rcm-max nonzeros).
requests / nonzero = 1.0 means every gathered double costs its own request (no locality at all).
usage: python scripts/locality_lab.py [--configs 1,2,3,4,5]"""
import argparse
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import scipy.sparse as sp

from oracle import gen


def requests_per_nnz(A: sp.csr_matrix, row_order=None) -> float:
    """A: CSR, columns sorted inside a row.  row_order: permutation applied before the length sort."""
    A = A.tocsr()
    if row_order is not None:
        A = A[row_order]
    lens = np.diff(A.indptr)
    order = np.argsort(-lens, kind="stable")            # setup.cu: sort_rows_by_length (stable, decreasing)
    m = A.shape[0]
    total_req = 0
    nnz = int(A.nnz)
    ptr, idx = A.indptr, A.indices
    step = 32 * 8192                                      # slices per pass (bounded memory)
    for s0 in range(0, m, step):
        rows = order[s0:s0 + step]
        pad = (-len(rows)) % 32
        L = lens[rows]
        maxlen = int(L.max()) if len(L) else 0
        if maxlen == 0:
            continue
        # [rows, maxlen] line ids, -1 where the row has ended
        k = np.arange(maxlen)[None, :]
        pos = ptr[rows][:, None] + k
        valid = k < L[:, None]
        lines = np.where(valid, idx[np.minimum(pos, nnz - 1)] >> 4, -1).astype(np.int64)
        if pad:
            lines = np.vstack([lines, -np.ones((pad, maxlen), dtype=np.int64)])
        lines = lines.reshape(-1, 32, maxlen)
        srt = np.sort(lines, axis=1)
        distinct = (srt[:, 1:, :] != srt[:, :-1, :]).sum(axis=1) + 1      # distinct values incl. -1
        has_pad = (srt[:, 0, :] == -1)
        all_pad = (srt[:, -1, :] == -1)
        total_req += int((distinct - has_pad.astype(np.int64)).sum()) + 0 * int(all_pad.sum())
    return total_req / max(nnz, 1)


def bin_order(A: sp.csr_matrix):
    lens = np.diff(A.indptr)
    first = np.where(lens > 0, A.indices[np.minimum(A.indptr[:-1], A.nnz - 1)], 0)
    return np.lexsort((first, -lens))


def rcm_orders(A: sp.csr_matrix):
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    m, n = A.shape
    P = (A != 0).astype(np.int8)
    G = sp.bmat([[None, P], [P.T, None]], format="csr")
    perm = reverse_cuthill_mckee(G, symmetric_mode=True)
    rows = perm[perm < m]
    cols = perm[perm >= m] - m
    return rows, cols


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--rcm-max", type=int, default=2_000_000)
    args = ap.parse_args()
    for cfg in [int(c) for c in args.configs.split(",")]:
        L = gen.config(cfg)
        A = L.A.tocsr(); A.sort_indices()
        AT = A.T.tocsr(); AT.sort_indices()
        rec = {"config": cfg, "m": A.shape[0], "n": A.shape[1], "nnz": int(A.nnz)}
        t = time.time()
        rec["row_step_as_generated"] = requests_per_nnz(A)          # K xbar: rows gather columns
        rec["col_step_as_generated"] = requests_per_nnz(AT)         # K' yhat: columns gather rows
        rec["row_step_binned"] = requests_per_nnz(A, bin_order(A))
        rec["col_step_binned"] = requests_per_nnz(AT, bin_order(AT))
        if A.nnz <= args.rcm_max:
            r, c = rcm_orders(A)
            B = A[r][:, c].tocsr(); B.sort_indices()
            BT = B.T.tocsr(); BT.sort_indices()
            rec["row_step_rcm"] = requests_per_nnz(B)
            rec["col_step_rcm"] = requests_per_nnz(BT)
        rec["seconds"] = round(time.time() - t, 1)
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
