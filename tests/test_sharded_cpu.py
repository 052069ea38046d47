"""N > 1 host logic on CPU: world_size-2 gloo run of the row-sharded oracle (the same exchange
pattern dist.cu performs with NCCL) against the single-process oracle, plus the partitioning
helper bench.py uses."""
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from miplib_b200 import sharding
from oracle import gen, pdhg_dist_ref, pdhg_ref


def test_row_blocks_partition_the_rows():
    for m, w in [(10, 3), (1000000, 8), (7, 8), (5, 1)]:
        blocks = [sharding.row_block(m, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == m
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        assert max(e - s for s, e in blocks) - min(e - s for s, e in blocks) <= 1
    with pytest.raises(ValueError):
        sharding.row_block(10, 3, 3)
    L = gen.lp(50, 80, 4, 3)
    parts = [sharding.shard_csr(L.A, L.row_lo, L.row_hi, r, 3) for r in range(3)]
    assert sum(p[0].size - 1 for p in parts) == L.m
    assert np.array_equal(np.concatenate([p[2] for p in parts]), L.A.tocsr().data)
    assert all(p[0][0] == 0 for p in parts)


class GlooComm:
    def _all(self, a, op):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).copy())
        dist.all_reduce(t, op=op)
        return t.numpy()
    def sum(self, a): return self._all(a, dist.ReduceOp.SUM)
    def max(self, a): return self._all(a, dist.ReduceOp.MAX)
    def gather(self, a):
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, a)
        return out


def _worker(rank, world, port, out, exchange="allreduce"):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        L = gen.lp(300, 500, 6, 7)
        ptr, idx, val, lo, hi = sharding.shard_csr(L.A, L.row_lo, L.row_hi, rank, world)
        Ag = sp.csr_matrix((val, idx, ptr), shape=(ptr.size - 1, L.n))
        if exchange == "owner":
            r = pdhg_dist_ref.solve_block_owner(Ag, L.c, L.lb, L.ub, lo, hi, GlooComm(), rank, world)
        else:
            r = pdhg_dist_ref.solve_sharded(Ag, L.c, L.lb, L.ub, lo, hi, GlooComm())
        out[rank] = (r["status"], r["obj"], r["iters"], r["restarts"], r["x"].copy(), r["y_local"].copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allreduce", "owner"])
def test_sharded_oracle_world2_gloo_matches_single_process(exchange):
    """allreduce: partial K_g' yhat_g summed over the ranks (dist.cu); owner: the block-owner exchange
    of p2p.cu - row block + column block per rank, two all-gathers per iteration."""
    world = 2
    port = 29500 + np.random.default_rng().integers(0, 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, int(port), out, exchange), nprocs=world, join=True)
        res = dict(out)
    ref = pdhg_ref.solve(gen.lp(300, 500, 6, 7))
    L = gen.lp(300, 500, 6, 7)
    assert res[0][0] == res[1][0] == "optimal"
    # both ranks reach the same decision at the same iteration and hold the same x
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]
    assert np.array_equal(res[0][4], res[1][4])
    # and it is the single-process answer (summation order differs by one partial sum)
    assert abs(res[0][1] - ref["obj"]) <= 1e-6 * (1 + abs(ref["obj"]))
    assert abs(res[0][2] - ref["iters"]) <= 2 * 64
    y = np.concatenate([res[0][5], res[1][5]])
    from oracle import kkt_verify
    cert = kkt_verify.certificate(L, res[0][4], y)
    assert max(cert["rel_primal_res"], cert["rel_dual_res"], cert["rel_gap"]) <= 1.05e-6


def test_block_owner_oracle_world1_is_the_plain_oracle():
    L = gen.lp(300, 500, 6, 8)
    r = pdhg_dist_ref.solve_block_owner(L.A, L.c, L.lb, L.ub, L.row_lo, L.row_hi, pdhg_dist_ref.LocalComm(), 0, 1)
    ref = pdhg_ref.solve(L)
    assert r["iters"] == ref["iters"] and abs(r["obj"] - ref["obj"]) <= 1e-12 * (1 + abs(ref["obj"]))
    assert np.abs(r["x"] - ref["x"]).max() <= 1e-12


def test_sharded_oracle_world1_is_the_plain_oracle():
    L = gen.lp(300, 500, 6, 8)
    r = pdhg_dist_ref.solve_sharded(L.A, L.c, L.lb, L.ub, L.row_lo, L.row_hi, pdhg_dist_ref.LocalComm())
    ref = pdhg_ref.solve(L)
    assert r["iters"] == ref["iters"] and abs(r["obj"] - ref["obj"]) <= 1e-12 * (1 + abs(ref["obj"]))


# ---- 2-D (R x C) block partition: the layout DESIGN.md section 5.1 names next for 8 GPUs ---------
class _ThreadWorld:
    """In-process stand-in for all_gather_object: W threads meet at a barrier."""
    def __init__(self, W):
        import threading
        self.slots = [None] * W
        self.barrier = threading.Barrier(W)

    def gather(self, rank, obj):
        self.slots[rank] = obj
        self.barrier.wait()
        out = list(self.slots)
        self.barrier.wait()
        return out


def _grid_shares(L, R, C, rank):
    A = L.A.tocsr()
    m, n = A.shape
    r, c = divmod(rank, C)
    r0, r1 = (m * r) // R, (m * (r + 1)) // R
    c0, c1 = (n * c) // C, (n * (c + 1)) // C
    return A[r0:r1, c0:c1], L.c[c0:c1], L.lb[c0:c1], L.ub[c0:c1], L.row_lo[r0:r1], L.row_hi[r0:r1]


@pytest.mark.parametrize("R,C", [(2, 2), (2, 4)])
def test_grid_partition_oracle_matches_single_process(R, C):
    """Row step = partial product + reduce-scatter inside the row group + all-gather of yhat; column
    step = the mirror image inside the column group.  Same decisions as the single-process oracle,
    and the ingress per rank is the formula DESIGN.md quotes (half the block-owner exchange's at
    2 x 4)."""
    import threading
    from oracle import kkt_verify
    L = gen.lp(300, 500, 6, 7)
    W = R * C
    world = _ThreadWorld(W)
    res = [None] * W

    def work(rank):
        comm = pdhg_dist_ref.GridComm(rank, R, C, lambda obj: world.gather(rank, obj))
        try:
            res[rank] = pdhg_dist_ref.solve_grid(*_grid_shares(L, R, C, rank), comm)
        except BaseException as e:          # let the other threads out of the barrier
            res[rank] = e
            world.barrier.abort()

    threads = [threading.Thread(target=work, args=(k,)) for k in range(W)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert all(isinstance(v, dict) for v in res), res
    ref = pdhg_ref.solve(L)
    assert {v["status"] for v in res} == {"optimal"}
    assert {v["iters"] for v in res} == {ref["iters"]} and {v["restarts"] for v in res} == {ref["restarts"]}
    assert abs(res[0]["obj"] - ref["obj"]) <= 1e-12 * (1 + abs(ref["obj"]))
    for c in range(C):                       # the R ranks of a column group hold the same x, bit for bit
        assert all(np.array_equal(res[c]["x_group"], res[r * C + c]["x_group"]) for r in range(R))
    x = np.concatenate([res[c]["x_group"] for c in range(C)])
    y = np.concatenate([res[r * C]["y_group"] for r in range(R)])
    cert = kkt_verify.certificate(L, x, y)
    assert max(cert["rel_primal_res"], cert["rel_dual_res"], cert["rel_gap"]) <= 1.05e-6
    formula = 16 * ((C - 1) * L.m + (R - 1) * L.n) / (R * C)
    owner = 8 * (L.m + L.n) * (W - 1) / W
    assert formula <= res[0]["ingress_bytes_per_iter"] <= 1.03 * formula       # + the KKT pass, 1 in 64
    if (R, C) == (2, 4):
        assert res[0]["ingress_bytes_per_iter"] < 0.52 * owner


def _grid_worker(rank, world, port, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        L = gen.lp(300, 500, 6, 7)

        def gather_world(obj):
            got = [None] * world
            dist.all_gather_object(got, obj)
            return got
        comm = pdhg_dist_ref.GridComm(rank, 1, world, gather_world)          # 1 x 2: column groups
        r = pdhg_dist_ref.solve_grid(*_grid_shares(L, 1, world, rank), comm)
        out[rank] = (r["status"], r["obj"], r["iters"], r["x_group"].copy(), r["y_group"].copy())
    finally:
        dist.destroy_process_group()


def test_grid_partition_oracle_world2_gloo():
    world = 2
    port = 29500 + np.random.default_rng().integers(0, 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_grid_worker, args=(world, int(port), out), nprocs=world, join=True)
        res = dict(out)
    ref = pdhg_ref.solve(gen.lp(300, 500, 6, 7))
    assert res[0][0] == res[1][0] == "optimal" and res[0][2] == res[1][2] == ref["iters"]
    assert abs(res[0][1] - ref["obj"]) <= 1e-12 * (1 + abs(ref["obj"]))
    assert np.array_equal(res[0][4], res[1][4])                   # one row group: both hold the whole y
    x = np.concatenate([res[0][3], res[1][3]])
    assert np.abs(x - ref["x"]).max() <= 1e-9
