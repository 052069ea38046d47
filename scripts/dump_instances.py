"""Write BASELINE.json's LP configs as free-format MPS through the backend's own writer
(miplib_cuda_dump -> Solver::dump, the reference's hand-off format: miplib/lpsolve/solver.cpp:262-274)
and record their SHA-256, so that an external SCIP 8.0.3 / lp_solve 5.5 / any exact solver can be run
on byte-identical instances and finally pin parity against the reference's own backends
(VERDICT round 1 item 5).  No GPU needed: dumping never touches the device.

    python scripts/dump_instances.py [--configs 1 2] [--out /tmp/miplib_instances] [--check]

`--check` compares against tests/golden/instances.json instead of rewriting it.
With SCIP:      scip -c "read C1.mps optimize display solution quit"
With lp_solve:  lp_solve -fmps C1.mps -S4          (expected objective values: tests/golden/objectives.json)
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from miplib_b200 import build          # noqa: E402
from oracle import gen                 # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "instances.json"


def dump(cfg: int, path: Path) -> dict:
    lib = C.CDLL(str(build.FRONT_LIB))
    lib.miplib_get_last_error.restype = C.c_char_p
    lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
    L = gen.config(cfg)
    A = L.A.tocsr()
    A.sort_indices()
    arrs = [np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64), np.ascontiguousarray(L.c, dtype=np.float64),
            np.ascontiguousarray(L.lb, dtype=np.float64), np.ascontiguousarray(L.ub, dtype=np.float64),
            None if L.integer is None else np.ascontiguousarray(np.where(L.integer, 1, 0), dtype=np.int32),
            np.ascontiguousarray(np.maximum(L.row_lo, -1e30)), np.ascontiguousarray(np.minimum(L.row_hi, 1e30))]
    solver = C.c_void_p()
    if lib.miplib_create_solver(C.byref(solver), 5):
        raise RuntimeError(lib.miplib_get_last_error().decode())
    ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
    if lib.miplib_cuda_load_csr(solver, L.m, L.n, *ptrs, int(L.maximize)):
        raise RuntimeError(lib.miplib_get_last_error().decode())
    if lib.miplib_cuda_dump(solver, str(path).encode()):
        raise RuntimeError(lib.miplib_get_last_error().decode())
    lib.miplib_destroy_solver(solver)
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return {"recipe": f"oracle.gen.config({cfg})", "rows": L.m, "columns": L.n, "nonzeros": L.nnz,
            "bytes": path.stat().st_size, "sha256": h.hexdigest()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", type=int, nargs="+", default=[1, 2])
    ap.add_argument("--out", default="/tmp/miplib_instances")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    build.build_front()
    got = {}
    for cfg in args.configs:
        got[f"C{cfg}.mps"] = dump(cfg, out / f"C{cfg}.mps")
        print(f"C{cfg}.mps", got[f"C{cfg}.mps"], flush=True)
    if args.check:
        want = json.loads(GOLDEN.read_text())
        bad = [k for k, v in got.items() if want.get(k, {}).get("sha256") != v["sha256"]]
        if bad:
            sys.exit(f"checksum mismatch: {bad}")
        print("checksums match tests/golden/instances.json")
    else:
        old = json.loads(GOLDEN.read_text()) if GOLDEN.exists() else {}
        old.update(got)
        GOLDEN.write_text(json.dumps(old, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
