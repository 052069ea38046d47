"""Infeasible / unbounded LPs: the Result contract of the reference's adapters
(miplib/lpsolve/solver.cpp:171-199: status 2 -> Infeasible, 3 -> Unbounded) served by certificate
tests on the PDHG displacement.  Oracle on CPU; the CUDA path against it on the GPU."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import gen, highs_ref, pdhg_ref

INF = np.inf


def cases():
    A = sp.csr_matrix(np.array([[1.0, 1.0], [1.0, 1.0]]))
    yield "tiny_infeasible", gen.LP(A, np.array([1.0, 0.0]), np.zeros(2), np.ones(2),
                                    np.array([-INF, 2.0]), np.array([1.0, INF])), "infeasible"
    A = sp.csr_matrix(np.array([[1.0, -1.0]]))
    yield "tiny_unbounded", gen.LP(A, np.array([-1.0, 0.0]), np.zeros(2), np.full(2, INF),
                                   np.array([-INF]), np.array([1.0])), "unbounded"
    L = gen.lp(300, 500, 6, 8)                       # a contradictory copy of row 5
    A = sp.vstack([L.A, L.A[5]]).tocsr()
    yield "lp_infeasible", gen.LP(A, L.c, L.lb, L.ub, np.append(L.row_lo, L.row_hi[5] + 1.0),
                                  np.append(L.row_hi, INF)), "infeasible"
    L = gen.lp(300, 500, 6, 9)                       # column 0: unbounded above, only relaxes a <= row
    A = L.A.tolil(); A[:, 0] = 0; A[250, 0] = -1.0; A = A.tocsr()
    ub = L.ub.copy(); ub[0] = INF
    yield "lp_unbounded", gen.LP(A, L.c, L.lb, ub, L.row_lo, L.row_hi), "unbounded"


def test_oracle_certificates_agree_with_exact_solver():
    for name, L, kind in cases():
        ref = highs_ref.solve_lp(L)
        assert ref["status"] in ((2,) if kind == "infeasible" else (3, 4)), (name, ref["status"], ref["message"])
        r = pdhg_ref.solve(L, max_iter=100000)
        if kind == "infeasible":
            assert r["status"] == "infeasible", (name, r["status"])
        else:
            assert r["status"] in ("unbounded", "infeasible_or_unbounded"), (name, r["status"])
        assert r["iters"] <= 5000
    # and no false alarm on feasible models
    for L in (gen.lp(300, 500, 6, 7), gen.set_cover(200, 500, 8, 20264), gen.mkp(5, 60, 4, 4, 20265)):
        assert pdhg_ref.solve(L)["status"] == "optimal"


@pytest.mark.gpu
def test_cuda_certificates_match_oracle(built):
    from miplib_b200 import Engine
    with Engine(0) as e:
        for name, L, kind in cases():
            ref = pdhg_ref.solve(L, max_iter=100000)
            e.load_lp(L)
            st = e.solve()
            if kind == "infeasible":
                assert st.status == "infeasible", (name, st)
            else:
                assert st.status in ("unbounded", "infeasible_or_unbounded"), (name, st)
            assert abs(st.iterations - ref["iters"]) <= 4 * 64, (name, st.iterations, ref["iters"])
        for L in (gen.lp(300, 500, 6, 7), gen.config(1)):
            e.load_lp(L)
            assert e.solve().status == "optimal"
