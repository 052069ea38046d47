"""Small end-to-end case for memory / race checking.
compute-sanitizer answers "closed on this pool" (gpurun_out/r02_memcheck_single.out), so the checks are
the library's own: MIPCU_PARAM_VALIDATE = 1 makes device kernels verify every index structure the
iteration kernels dereference (setup.cu:validate_structures), results are compared with SciPy, and the
sharded mode repeats the two-rank solve and demands bit-identical vectors on both ranks every time
(a race in the cross-GPU protocol shows up as a difference).  What the case exercises:
config 1 through every kernel family of the single-GPU path - SELL lane-per-row kernels (direct
launches and the CUDA-graph block), CSR lanes-per-row kernels, the TMA and CSR-stream variants, the
batched node-LP kernels (short-row and CTA-per-row), warm start, bounds update, SpMV hooks - with
iteration limits that keep a 20-50x slower instrumented run short.

    compute-sanitizer --tool memcheck  python scripts/sanitize_case.py
    compute-sanitizer --tool racecheck python scripts/sanitize_case.py
    python scripts/sanitize_case.py sharded        (under torchrun-less mp.spawn: 2 ranks, 2 GPUs)
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from oracle import gen            # noqa: E402
from miplib_b200 import Engine    # noqa: E402


def single():
    L = gen.config(1)
    with Engine(0) as e:
        e.set_param("validate", 1)
        e.load_lp(L)
        e.set_param("iter_limit", 512)
        v = np.random.default_rng(0).standard_normal(L.n)
        assert np.abs(e.spmv(v) - L.A @ v).max() <= 1e-12
        w = np.random.default_rng(1).standard_normal(L.m)
        assert np.abs(e.spmv(w, transpose=True) - L.A.T @ w).max() <= 1e-12
        for graph in (1, 0):                       # graph replay, then direct launches
            e.set_param("use_graph", graph)
            st = e.solve()
            print("sell graph", graph, st.status, st.iterations, flush=True)
        for tpr in (2, 8, 32):                     # CSR lanes-per-row kernels
            e.set_param("threads_per_row", tpr)
            st = e.solve()
            print("csr tpr", tpr, st.status, st.iterations, flush=True)
        e.set_param("threads_per_row", 0)
        for tma in (1, 2):                         # TMA-streamed tiles, CSR-stream
            e.set_param("use_tma", tma)
            st = e.solve()
            print("variant", tma, st.status, st.iterations, flush=True)
        e.set_param("use_tma", -1)
        e.warm_start(e.x(), e.y())
        e.set_bounds(np.arange(4), np.zeros(4), np.full(4, 0.5))
        st = e.solve()
        print("warm", st.status, st.iterations, flush=True)
        e.time_kernel(2, 3, True)
        e.time_kernel(5, 3, False)
    for name, Lb in (("setcover", gen.set_cover(200, 500, 8, 20264)), ("mkp long rows", gen.mkp(5, 2000, 4, 10, 20265))):
        with Engine(0) as e:
            e.set_param("validate", 1)
            e.load_lp(Lb)
            e.set_param("iter_limit", 256)
            B = 5
            lb = np.tile(Lb.lb, (B, 1)); ub = np.tile(Lb.ub, (B, 1))
            ub[1, :3] = 0.0
            root = e.solve()
            stats, x, y = e.solve_batch(lb, ub, x0=np.tile(e.x(), (B, 1)), y0=np.tile(e.y(), (B, 1)),
                                        weight0=np.full(B, root.primal_weight), branch_tol=np.full(B, 1e-4), want_y=True)
            print("batch", name, [s.status for s in stats], flush=True)
            stats, x = e.solve_batch(lb, ub)
            print("batch cold", name, [s.status for s in stats], flush=True)


def _rank(rank, world, uid, multicast, out):
    from miplib_b200 import sharding
    L = gen.lp(3000, 5000, 8, 21)
    ptr, idx, val, lo, hi = sharding.shard_csr(L.A, L.row_lo, L.row_hi, rank, world)
    res = []
    with Engine(rank, rank, world, uid) as e:
        e.set_param("multicast", multicast)
        e.set_param("validate", 1)
        for exchange in (2, 1, 0):
            e.set_param("shard_exchange", exchange)
            e.load(ptr, idx, val, L.c, L.lb, L.ub, lo, hi, L.maximize, n=L.n)
            for rep in range(6):
                if exchange == 2 and rep == 3:        # re-load under kept mappings in the middle of the series
                    e.load(ptr, idx, val, L.c, L.lb, L.ub, lo, hi, L.maximize, n=L.n)
                st = e.solve()
                res.append((exchange, st.status, st.iterations, st.primal_obj, e.x().tobytes(), e.y().tobytes()))
            if rank == 0:
                print("sharded exchange", exchange, "multicast", int(e.get_param("multicast")), st.status, st.iterations, flush=True)
    out[rank] = res


def sharded():
    import torch.multiprocessing as mp
    for multicast in (1, 0):
        uid = Engine.unique_id()
        with mp.Manager() as mgr:
            out = mgr.dict()
            mp.spawn(_rank, args=(2, uid, multicast, out), nprocs=2, join=True)
            r0, r1 = out[0], out[1]
        for a, b in zip(r0, r1):
            assert a[:5] == b[:5], "ranks disagree on x / status / iterations"       # x is replicated bit for bit
        for ex in (2, 1, 0):
            runs = [a for a in r0 if a[0] == ex]
            assert all(r[1] == "optimal" for r in runs), runs[0][:4]
            assert all(r[2:] == runs[0][2:] for r in runs), f"exchange {ex}: repeated solves differ (race?)"
        print("multicast", multicast, ": 18 two-rank solves, bit-identical across repeats and ranks", flush=True)


if __name__ == "__main__":
    sharded() if len(sys.argv) > 1 and sys.argv[1] == "sharded" else single()
    print("sanitize_case done", flush=True)
