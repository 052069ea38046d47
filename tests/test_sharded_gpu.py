"""Row-sharded solve on >= 2 GPUs of one box (one process per GPU, NCCL all-reduce of the K'y
partial sums) against the single-GPU solve and the independent certificate.  Skipped on a
one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`."""
import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import gen, kkt_verify

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid, which, exchange, out):
    from miplib_b200 import Engine, sharding
    L = gen.lp(3000, 5000, 8, 21) if which == "small" else gen.config(2)
    ptr, idx, val, lo, hi = sharding.shard_csr(L.A, L.row_lo, L.row_hi, rank, world)
    with Engine(rank, rank, world, uid) as e:
        e.set_param("shard_exchange", exchange)       # 2: block-owner, 1: fused P2P kernel, 0: ncclAllReduce
        e.load(ptr, idx, val, L.c, L.lb, L.ub, lo, hi, L.maximize, n=L.n)
        assert e.get_param("shard_exchange") == exchange      # no silent fall-back to NCCL
        if which == "small":
            # a solve, then a re-load of the same shape: the peer mappings (and the multicast window)
            # are kept, only the matrix-dependent parts are rebuilt; the solve below must not notice
            e.solve()
            e.load(ptr, idx, val, L.c, L.lb, L.ub, lo, hi, L.maximize, n=L.n)
            assert e.get_param("shard_exchange") == exchange
        st = e.solve()
        out[rank] = (st.status, st.primal_obj, st.iterations, st.restarts, e.x(), e.y(),
                     st.rel_primal_res, st.rel_gap)


@pytest.mark.parametrize("exchange", [2, 1, 0])
@pytest.mark.parametrize("which", ["small", "C2"])
def test_two_rank_solve_matches_single_gpu(built, which, exchange):
    from miplib_b200 import Engine
    if Engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    uid = Engine.unique_id()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, uid, which, exchange, out), nprocs=world, join=True)
        res = dict(out)
    L = gen.lp(3000, 5000, 8, 21) if which == "small" else gen.config(2)
    with Engine(0) as e:
        e.load_lp(L)
        one = e.solve()
    assert res[0][0] == res[1][0] == "optimal"
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]        # same decisions on every rank
    assert np.array_equal(res[0][4], res[1][4])                      # x is replicated bit for bit
    assert abs(res[0][1] - one.primal_obj) <= 1e-6 * (1 + abs(one.primal_obj))
    assert abs(res[0][2] - one.iterations) <= 4 * 64
    y = np.concatenate([res[0][5], res[1][5]])
    cert = kkt_verify.certificate(L, res[0][4], y)
    assert max(cert["rel_primal_res"], cert["rel_dual_res"], cert["rel_gap"]) <= 1.05e-6, cert


def test_branch_and_bound_deals_node_rounds_to_two_gpus(built, golden, tmp_path):
    """set_nr_threads(2): every round of node LPs is split between two contexts (one per GPU) by
    host threads; the proven optimum must not change."""
    from miplib_b200 import Engine
    if Engine.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.test_frontend import run_unit_test, write_models
    models = [("setcover_400x1000_mip", gen.set_cover(400, 1000, 8, 20264), golden["setcover_400x1000_mip"]["objective"], 1e-9),
              ("mkp_5x60_mip", gen.mkp(5, 60, 4, 4, 20265), golden["mkp_5x60_mip"]["objective"], 1e-9)]
    path = tmp_path / "models.txt"
    write_models(path, models)
    rc, out = run_unit_test(built, "--filter", "Generated", env={"MIPLIB_TEST_MODELS": str(path), "MIPLIB_TEST_GPUS": "2"})
    print(out)
    assert rc == 0 and "0 failed" in out and "0 skipped" in out, out
    assert "objective 867 expected 867" in out and "objective 875 expected 875" in out, out
