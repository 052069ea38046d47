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
