"""Where the plugin path (miplib_create_solver(Cuda) + miplib_cuda_load_csr + Solver::solve() + values)
spends its host time on config 3: runs bench.py's plugin_e2e step with MIPLIB_CUDA_TIMING laps on stderr.
usage: MIPLIB_CUDA_TIMING=1 python scripts/plugin_e2e.py [--config 3] [--steps 3]"""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

import bench
from oracle import gen

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
L = gen.config(args.config)
A = L.A.tocsr()
A.sort_indices()
host = dict(indptr=bench.pinned_like(A.indptr.astype(np.int64)), indices=bench.pinned_like(A.indices.astype(np.int32)),
            data=bench.pinned_like(A.data), c=bench.pinned_like(L.c), lb=bench.pinned_like(L.lb),
            ub=bench.pinned_like(L.ub), row_lo=bench.pinned_like(L.row_lo), row_hi=bench.pinned_like(L.row_hi))
print(json.dumps(bench.plugin_e2e(args, L, A, host)))
