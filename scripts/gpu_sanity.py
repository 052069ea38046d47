"""First-contact script for a B200 box: SpMV parity, a full C1 solve against the oracle, and
kernel timings on C2 / C3-shaped instances.  `python scripts/gpu_sanity.py [big]`."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from oracle import gen, pdhg_ref, kkt_verify
from miplib_b200 import Engine

big = "big" in sys.argv
out = {}
L = gen.config(1)
with Engine(0) as e:
    e.load_lp(L)
    rng = np.random.default_rng(0)
    v = rng.standard_normal(L.n); w = rng.standard_normal(L.m)
    r1 = e.spmv(v); r2 = e.spmv(w, transpose=True)
    print("spmv rel err", np.abs(r1 - L.A @ v).max() / np.abs(L.A @ v).max(),
          np.abs(r2 - L.A.T @ w).max() / np.abs(L.A.T @ w).max())
    e.set_param("verbose", 1 if "-v" in sys.argv else 0)
    t = time.time(); s = e.solve(); print("C1 solve", time.time() - t, s)
    x, y = e.x(), e.y()
    print("cert", kkt_verify.certificate(L, x, y))
    r1 = e.spmv(v)
    print("spmv after scaling rel err", np.abs(r1 - L.A @ v).max() / np.abs(L.A @ v).max())
    t = time.time(); s = e.solve(); print("C1 re-solve", time.time() - t, s.iterations, s.solve_seconds, s.iter_seconds)
    ref = pdhg_ref.solve(L)
    print("oracle iters", ref["iters"], ref["restarts"], ref["obj"], "cuda", s.iterations, s.restarts, s.primal_obj)
    xb, yy = e.debug_iterate(10)
    tr = pdhg_ref.solve(L, period=1000, max_iter=1000, trace_at=(10,))["trace"][10]
    print("iterate parity @10", np.abs(xb - tr[0]).max(), np.abs(yy - tr[1]).max())

def timing(L, name):
    with Engine(0) as e:
        t = time.time(); e.load_lp(L); tl = time.time() - t
        t = time.time(); e.prepare(); tp = time.time() - t
        res = {"load_s": tl, "prepare_s": tp}
        for tpr in (4, 8, 16, 32):
            e.set_param("threads_per_row", tpr)
            for which, nm in ((0, "Kx"), (1, "KTy"), (2, "row_step"), (3, "col_step")):
                for fl in (0, 1):
                    us, nb = e.time_kernel(which, 20, fl)
                    res[f"tpr{tpr}_{nm}_{'flush' if fl else 'hot'}"] = (round(us, 1), round(nb / us / 1e3, 1))
        e.set_param("threads_per_row", 0)
        t = time.time(); s = e.solve(); res["solve_wall"] = time.time() - t
        res["stats"] = s.__dict__
        x, y = e.x(), e.y()
        res["cert"] = kkt_verify.certificate(L, x, y)
    print(name, json.dumps(res, indent=1, default=str))
    return res

t = time.time(); L2 = gen.config(2); print("gen C2", time.time() - t)
out["C2"] = timing(L2, "C2")
if big:
    t = time.time(); L3 = gen.config(3); print("gen C3", time.time() - t)
    out["C3"] = timing(L3, "C3")
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sanity.json", "w"), indent=1, default=str)
