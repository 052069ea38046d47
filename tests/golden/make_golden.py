"""Regenerates tests/golden/objectives.json and tests/golden/pdhg_trace_c1.npz.

Run in the build container (needs SciPy's HiGHS; minutes, because of the 10k x 20k IPM solve):
    python tests/golden/make_golden.py [--with-mid] [--with-c2] [--with-c3]
--with-c2 / --with-c3: configs 2 / 3 by HiGHS' own PDLP (2 minutes / hours on one core), the only
independent solver that finishes them here.
The values are the independent exact-solver optima the parity tests compare the CUDA backend
against (HiGHS stand-in for the reference's SCIP / lp_solve backends).
"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import gen, highs_ref, pdhg_ref  # noqa: E402

HERE = Path(__file__).resolve().parent
out = {}


def put(key, recipe, kind, value, extra=None):
    out[key] = {"recipe": recipe, "kind": kind, "objective": value, **(extra or {})}
    print(key, value, flush=True)


L = gen.config(1)
put("C1_lp", "lp(1000,2000,8,20261)", "highs-ds", highs_ref.solve_lp(L, "highs-ds")["obj"])
for m, n in [(200, 500), (400, 1000)]:
    S = gen.set_cover(m, n, 8, 20264)
    put(f"setcover_{m}x{n}_mip", f"set_cover({m},{n},8,20264)", "highs-milp", highs_ref.solve_mip(S)["obj"])
    put(f"setcover_{m}x{n}_lp", f"set_cover({m},{n},8,20264) relaxation", "highs-ds", highs_ref.solve_lp(S)["obj"])
for m, n, G in [(5, 60, 4), (8, 80, 4), (10, 100, 5)]:
    S = gen.mkp(m, n, 4, G, 20265)
    put(f"mkp_{m}x{n}_mip", f"mkp({m},{n},4,{G},20265)", "highs-milp", highs_ref.solve_mip(S)["obj"])
    put(f"mkp_{m}x{n}_lp", f"mkp({m},{n},4,{G},20265) relaxation", "highs-ds", highs_ref.solve_lp(S)["obj"])
for seed in (7, 8, 9):
    S = gen.lp(300, 500, 6, seed)
    put(f"lp_300x500_s{seed}", f"lp(300,500,6,{seed})", "highs-ds", highs_ref.solve_lp(S)["obj"])
if "--with-mid" in sys.argv:
    S = gen.lp(10000, 20000, 8, 1)
    put("mid_lp", "lp(10000,20000,8,1)", "highs-ipm", highs_ref.solve_lp(S, "highs-ipm")["obj"])
else:
    old = json.loads((HERE / "objectives.json").read_text()) if (HERE / "objectives.json").exists() else {}
    if "mid_lp" in old:
        out["mid_lp"] = old["mid_lp"]
old = json.loads((HERE / "objectives.json").read_text()) if (HERE / "objectives.json").exists() else {}
for cfg in (2, 3):
    key = f"C{cfg}_lp"
    if f"--with-c{cfg}" in sys.argv:
        r = highs_ref.solve_lp_pdlp(gen.config(cfg), 1e-6)
        assert r["status"] == "kOptimal", r["status"]
        put(key, f"config({cfg})", "highs-pdlp 1e-6", r["obj"],
            {"iterations": r["iters"], "seconds": round(r["secs"], 1),
             "max_primal_infeasibility": r["max_primal_infeas"], "max_dual_infeasibility": r["max_dual_infeas"]})
    elif key in old:
        out[key] = old[key]
(HERE / "objectives.json").write_text(json.dumps(out, indent=1) + "\n")

# iterate-level golden vectors of the oracle on C1: scaled (xbar, y) after 10 and 64 iterations,
# scaling vectors, step size; small enough to commit (float64, compressed)
r = pdhg_ref.solve(L, period=64, max_iter=64, trace_at=(10, 64))
S = pdhg_ref.prepare(L)
np.savez_compressed(HERE / "pdhg_trace_c1.npz", d1=S.d1, d2=S.d2, eta=S.eta, w0=S.w0,
                    xbar10=r["trace"][10][0], y10=r["trace"][10][1],
                    xbar64=r["trace"][64][0], y64=r["trace"][64][1])
print("wrote", HERE / "objectives.json", HERE / "pdhg_trace_c1.npz")
