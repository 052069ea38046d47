"""The C++ front-end (Solver / Var / Expr / Constr / capi) and the Backend::Cuda adapter, driven
through the re-hosted reference test cases in tests/cpp (binary miplib_b200/miplib_unit_test).

CPU part: every front-end known answer of the reference's test/unit/*.cpp (strings, bounds,
reformulations, policies, error conventions) plus KAT-A / KAT-B, which bound propagation decides
without a device.  GPU part: the solve-level KATs and generated models replayed through the
expression API against the exact-solver goldens."""
import os
import subprocess

import numpy as np
import pytest

from oracle import gen


def run_unit_test(built, *args, env=None, timeout=600):
    exe = built.UNIT_TEST
    assert exe.exists()
    p = subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=timeout,
                       env={**os.environ, **(env or {})})
    return p.returncode, p.stdout + p.stderr


def test_frontend_known_answers_on_cpu(built):
    rc, out = run_unit_test(built, "--cpu-only")
    assert rc == 0, out
    last = out.strip().splitlines()[-1]
    assert "0 failed" in last, out
    assert int(last.split()[0]) >= 16, out
    assert "PASS KAT-A" in out and "PASS KAT-B" in out


def test_libmiplib_exports_reference_capi(built):
    out = subprocess.run(["nm", "-D", "--defined-only", str(built.FRONT_LIB)], capture_output=True,
                         text=True, check=True).stdout
    for sym in ["miplib_get_last_error", "miplib_create_solver", "miplib_destroy_solver",
                "miplib_shallow_copy_solver", "miplib_create_var", "miplib_destroy_var",
                "miplib_shallow_copy_var"]:                     # reference miplib/capi.hpp:16-23
        assert f" T {sym}" in out, sym
    # and every entry point the Cuda backend adds (miplib_b200/miplib/cuda/capi.hpp) is exported
    import re
    header = (built.ROOT / "miplib_b200" / "miplib" / "cuda" / "capi.hpp").read_text()
    declared = re.findall(r"^int\s+(miplib_cuda_\w+)\s*\(", header, flags=re.M)
    assert len(declared) >= 11 and "miplib_cuda_get_bound" in declared
    for sym in declared:
        assert f" T {sym}" in out, sym


def write_models(path, models, indicator_rows=None):
    """Text fixture read by tests/cpp/solver_tests.cpp ('Generated models ...').
    indicator_rows: {label: G} - the LAST G rows of that model are big-M switch rows
    `sum_{j in g} x_j - |g| z_g <= 0` (oracle/gen.py: mkp); they are written as indicator
    constraints `(z_g == 0) >> (sum_{j in g} x_j <= 0)` instead, so that the C++ side posts them
    through Solver::add(IndicatorConstr) and the adapter's reformulation has to recreate the rows."""
    def f(v):
        return "inf" if v == np.inf else "-inf" if v == -np.inf else repr(float(v))
    with open(path, "w") as fh:
        fh.write(f"{len(models)}\n")
        for label, L, expected, tol in models:
            A = L.A.tocsr()
            integer = L.integer if L.integer is not None else np.zeros(L.n, bool)
            G = (indicator_rows or {}).get(label, 0)
            switches = []
            for i in range(L.m - G, L.m):
                s, e = A.indptr[i], A.indptr[i + 1]
                cols, vals = A.indices[s:e], A.data[s:e]
                z = cols[vals < 0]
                assert z.size == 1 and np.all(vals[vals > 0] == 1.0) and L.row_hi[i] == 0.0
                assert -vals[vals < 0][0] == (vals > 0).sum()            # big-M = |group|
                switches.append((int(z[0]), [int(c) for c in cols[vals > 0]]))
            m_lin = L.m - G
            fh.write(f"{label} {L.n} {m_lin} {G}\n")
            for j in range(L.n):
                binary = integer[j] and L.lb[j] == 0 and L.ub[j] == 1
                t = 1 if binary else 2 if integer[j] else 0
                fh.write(f"{t} {f(L.lb[j])} {f(L.ub[j])} {f(L.c[j])}\n")
            for i in range(m_lin):
                s, e = A.indptr[i], A.indptr[i + 1]
                terms = " ".join(f"{c} {f(v)}" for c, v in zip(A.indices[s:e], A.data[s:e]))
                fh.write(f"{f(L.row_lo[i])} {f(L.row_hi[i])} {e - s} {terms}\n")
            for z, members in switches:
                fh.write(f"{z} {len(members)} " + " ".join(map(str, members)) + "\n")
            fh.write(f"{int(L.maximize)} {f(expected)} {tol}\n")


def _read_mps_with_highs(path):
    """The dumped model as HiGHS (an independent MPS reader) sees it."""
    import scipy.optimize._highspy._core as core
    import scipy.sparse as sp
    h = core._Highs()
    h.setOptionValue("output_flag", False)
    assert h.readModel(str(path)) == core.HighsStatus.kOk
    lp = h.getLp()
    a = lp.a_matrix_
    shape = (lp.num_row_, lp.num_col_)
    if a.format_ == core.MatrixFormat.kColwise:
        A = sp.csc_matrix((np.array(a.value_), np.array(a.index_), np.array(a.start_)), shape=shape).tocsr()
    else:
        A = sp.csr_matrix((np.array(a.value_), np.array(a.index_), np.array(a.start_)), shape=shape)
    A.sort_indices()
    inf = core.kHighsInf
    clean = lambda v: np.where(np.array(v) >= inf, np.inf, np.where(np.array(v) <= -inf, -np.inf, np.array(v)))
    integ = np.array([int(t) != 0 for t in lp.integrality_], bool) if len(lp.integrality_) else np.zeros(shape[1], bool)
    # LP files order the columns by first appearance: put them back in creation order x0, x1, ...
    names = list(lp.col_names_)
    order = np.argsort([int(nm[1:]) for nm in names]) if names and all(nm[:1] == "x" and nm[1:].isdigit() for nm in names) \
        else np.arange(shape[1])
    A = A[:, order].tocsr()
    A.sort_indices()
    return dict(h=h, A=A, c=np.array(lp.col_cost_)[order], lb=clean(lp.col_lower_)[order], ub=clean(lp.col_upper_)[order],
                lo=clean(lp.row_lower_), hi=clean(lp.row_upper_), integer=integ[order],
                maximize=lp.sense_ == core.ObjSense.kMaximize)


def test_lowering_through_the_expression_api_equals_the_generator(built, golden, tmp_path):
    """Rows a5 - a10 of SURVEY.md section 8 without a device: the fixture models are built through
    Var / Expr / Constr, lowered by the Cuda adapter and written with Solver::dump(); HiGHS reads
    the MPS files back.  The lowered model must be the generator's model entry for entry (a `>=`
    row may come back as the negated `<=` row, the reference's own convention,
    miplib/constr.cpp:117-143), and HiGHS must find the golden optimum on it."""
    models = [
        ("C1_lp", gen.config(1), golden["C1_lp"]["objective"], 1e-6),
        ("lp_300x500_s7", gen.lp(300, 500, 6, 7), golden["lp_300x500_s7"]["objective"], 1e-6),
        ("setcover_200x500_mip", gen.set_cover(200, 500, 8, 20264), golden["setcover_200x500_mip"]["objective"], 1e-9),
        ("mkp_5x60_mip", gen.mkp(5, 60, 4, 4, 20265), golden["mkp_5x60_mip"]["objective"], 1e-9),
        # the same recipe with its 4 switch rows posted as indicator constraints
        # (z == 0) >> (sum x <= 0): IndicatorConstr::reformulation (reference miplib/constr.cpp:225-271)
        # must give back exactly the generator's big-M rows
        ("mkp_5x60_indicator", gen.mkp(5, 60, 4, 4, 20265), golden["mkp_5x60_mip"]["objective"], 1e-9),
        ("mkp_8x80_indicator", gen.mkp(8, 80, 4, 4, 20265), golden["mkp_8x80_mip"]["objective"], 1e-9),
    ]
    path = tmp_path / "models.txt"
    write_models(path, models, indicator_rows={"mkp_5x60_indicator": 4, "mkp_8x80_indicator": 4})
    rc, out = run_unit_test(built, "--cpu-only", "--filter", "dumped as MPS",
                            env={"MIPLIB_TEST_MODELS": str(path), "MIPLIB_TEST_DUMP_DIR": str(tmp_path)})
    assert rc == 0 and "1 passed" in out, out

    def upper_form(A, lo, hi):
        """every one-sided row as `a x <= b` (a `>=` row negated), two-sided rows untouched"""
        A = A.tocsr().copy()
        flip = np.isfinite(lo) & ~np.isfinite(hi)
        sign = np.where(flip, -1.0, 1.0)
        A = A.multiply(sign[:, None]).tocsr()
        A.sort_indices()
        return A, np.where(flip, -hi, lo), np.where(flip, -lo, hi)

    for label, L, expected, tol in models:
        M = _read_mps_with_highs(tmp_path / f"{label}.mps")
        assert M["A"].shape == (L.m, L.n), label
        A0, lo0, hi0 = upper_form(L.A, L.row_lo, L.row_hi)
        A1, lo1, hi1 = upper_form(M["A"], M["lo"], M["hi"])
        assert np.array_equal(A0.indptr, A1.indptr) and np.array_equal(A0.indices, A1.indices), label
        assert np.array_equal(A0.data, A1.data), label                  # 17 significant digits: exact
        assert np.array_equal(lo0, lo1) and np.array_equal(hi0, hi1), label
        assert np.array_equal(M["c"], L.c) and M["maximize"] == bool(L.maximize), label
        assert np.array_equal(M["lb"], L.lb) and np.array_equal(M["ub"], L.ub), label
        integer = L.integer if L.integer is not None else np.zeros(L.n, bool)
        assert np.array_equal(M["integer"], integer), label
        # and an exact solver finds the golden optimum on the dumped file
        h = M["h"]
        h.run()
        obj = h.getInfo().objective_function_value
        assert abs(obj - expected) <= max(tol, 1e-9) * (1 + abs(expected)), (label, obj, expected)
        # the LP-format writer describes the same model
        import scipy.optimize._highspy._core as core
        h2 = core._Highs()
        h2.setOptionValue("output_flag", False)
        assert h2.readModel(str(tmp_path / f"{label}.lp")) == core.HighsStatus.kOk
        h2.run()
        obj2 = h2.getInfo().objective_function_value
        assert abs(obj2 - expected) <= max(tol, 1e-9) * (1 + abs(expected)), (label, obj2, expected)


def test_dump_writers_keep_every_kind_of_bound_and_row(built, tmp_path):
    """A small model with every column kind the API can express (free, one-sided, boxed, fixed,
    negative bounds, integer, binary) and every row sense, lowered and dumped; both files, read by
    HiGHS, must give back the model - and the same optimum from either file."""
    import scipy.sparse as sp
    inf = np.inf
    lb = np.array([-inf, -inf, -3.0, 2.0, 0.0, 1.5, -4.0, 0.0])
    ub = np.array([inf, 5.0, inf, 7.0, 1.0, 1.5, -1.0, 10.0])
    integer = np.array([False, False, False, True, True, False, True, False])
    c = np.array([1.0, -2.0, 0.5, 3.0, -1.0, 0.0, 2.0, -0.25])
    A = sp.csr_matrix(np.array([
        [1.0, 1.0, 0.0, 0.0, 2.0, 0.0, 0.0, 1.0],      # <=
        [1.0, -1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0],     # >=
        [0.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0, 0.0],      # ==
        [-1.0, 0.0, 0.0, 0.5, 0.0, 0.0, 0.0, 2.0],     # <= with a negative right-hand side
        [0.0, 2.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0],      # >= 0: no RHS line in the MPS file
    ]))
    row_lo = np.array([-inf, -2.0, 3.5, -inf, 0.0])
    row_hi = np.array([8.0, inf, 3.5, -1.25, inf])
    L = gen.LP(A, c, lb, ub, row_lo, row_hi, False, integer, "mixed")
    path = tmp_path / "models.txt"
    write_models(path, [("mixed", L, 0.0, 1e-9)])
    rc, out = run_unit_test(built, "--cpu-only", "--filter", "dumped as MPS",
                            env={"MIPLIB_TEST_MODELS": str(path), "MIPLIB_TEST_DUMP_DIR": str(tmp_path)})
    assert rc == 0 and "1 passed" in out, out
    objs = []
    for ext in ("mps", "lp"):
        M = _read_mps_with_highs(tmp_path / f"mixed.{ext}")
        assert M["A"].shape == (L.m, L.n), ext
        flip = np.isfinite(M["lo"]) & ~np.isfinite(M["hi"])            # `>=` rows may come back negated
        dense = M["A"].toarray() * np.where(flip, -1.0, 1.0)[:, None]
        lo = np.where(flip, -M["hi"], M["lo"]); hi = np.where(flip, -M["lo"], M["hi"])
        flip0 = np.isfinite(row_lo) & ~np.isfinite(row_hi)
        dense0 = A.toarray() * np.where(flip0, -1.0, 1.0)[:, None]
        lo0 = np.where(flip0, -row_hi, row_lo); hi0 = np.where(flip0, -row_lo, row_hi)
        assert np.array_equal(dense, dense0), ext
        assert np.array_equal(lo, lo0) and np.array_equal(hi, hi0), ext
        assert np.array_equal(M["c"], c) and np.array_equal(M["lb"], lb) and np.array_equal(M["ub"], ub), ext
        assert np.array_equal(M["integer"], integer), ext
        M["h"].run()
        objs.append(M["h"].getInfo().objective_function_value)
    assert abs(objs[0] - objs[1]) <= 1e-9 * (1 + abs(objs[0])), objs


@pytest.mark.gpu
def test_kats_and_generated_models_on_gpu(built, golden, tmp_path):
    models = [
        ("C1_lp", gen.config(1), golden["C1_lp"]["objective"], 1e-6),
        ("lp_300x500_s7", gen.lp(300, 500, 6, 7), golden["lp_300x500_s7"]["objective"], 1e-6),
        ("setcover_200x500_mip", gen.set_cover(200, 500, 8, 20264), golden["setcover_200x500_mip"]["objective"], 1e-9),
        ("mkp_5x60_mip", gen.mkp(5, 60, 4, 4, 20265), golden["mkp_5x60_mip"]["objective"], 1e-9),
        ("setcover_400x1000_mip", gen.set_cover(400, 1000, 8, 20264), golden["setcover_400x1000_mip"]["objective"], 1e-9),
        # 7.5 k nodes: only tractable with the implied-bound rows of the big-M switch rows (bnb.hpp)
        ("mkp_8x80_mip", gen.mkp(8, 80, 4, 4, 20265), golden["mkp_8x80_mip"]["objective"], 1e-9),
    ]
    path = tmp_path / "models.txt"
    write_models(path, models)
    rc, out = run_unit_test(built, env={"MIPLIB_TEST_MODELS": str(path)})
    print(out)
    assert rc == 0, out
    assert "0 failed" in out and "0 skipped" in out, out


@pytest.mark.gpu
def test_mkp_10x100_mip_on_gpu(built, golden, tmp_path):
    """The last golden of the MIP parity tier (10 x 100 knapsack with 5 indicator
    switches -> 1526, HiGHS `milp` proves it with 5 242 nodes): identical incumbent objective
    through Solver / Var / Constr / IndicatorConstr on Backend::Cuda, proven by the tree."""
    models = [("mkp_10x100_mip", gen.mkp(10, 100, 4, 5, 20265), golden["mkp_10x100_mip"]["objective"], 1e-9)]
    path = tmp_path / "models.txt"
    write_models(path, models)
    rc, out = run_unit_test(built, "--filter", "Generated models",
                            env={"MIPLIB_TEST_MODELS": str(path), "MIPLIB_TEST_TIME_LIMIT": "900"}, timeout=1200)
    print(out)
    assert rc == 0, out
    assert "0 failed" in out and "0 skipped" in out, out


# ---- bulk ingress through the C entry points (miplib/cuda/capi.hpp) ---------------------------
def _front(built):
    import ctypes as C
    lib = C.CDLL(str(built.FRONT_LIB))
    lib.miplib_get_last_error.restype = C.c_char_p
    return lib, C


def _load_csr(lib, C, solver, L, types=None):
    A = L.A.tocsr(); A.sort_indices()
    arrs = [np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64), np.ascontiguousarray(L.c, dtype=np.float64),
            np.ascontiguousarray(L.lb, dtype=np.float64), np.ascontiguousarray(L.ub, dtype=np.float64),
            None if types is None else np.ascontiguousarray(types, dtype=np.int32),
            np.ascontiguousarray(np.maximum(L.row_lo, -1e30), dtype=np.float64),
            np.ascontiguousarray(np.minimum(L.row_hi, 1e30), dtype=np.float64)]
    ptrs = [None if a is None else a.ctypes.data_as(C.c_void_p) for a in arrs]
    lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
    return lib.miplib_cuda_load_csr(solver, L.m, L.n, *ptrs, int(L.maximize)), arrs


def test_bulk_ingress_without_a_device_stays_host_only(built, monkeypatch):
    """The eager device load of miplib_cuda_load_csr is best effort: with no usable GPU (this container) the call
    must still build the row store and succeed, even when the size threshold says "load the device now"."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    monkeypatch.setenv("MIPLIB_CUDA_EAGER_NNZ", "0")
    lib, C = _front(built)
    solver = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(solver), 5) == 0
    L = gen.config(1)
    rc, _ = _load_csr(lib, C, solver, L)
    assert rc == 0, lib.miplib_get_last_error()
    n, m = C.c_int64(), C.c_int64()
    assert lib.miplib_cuda_get_shape(solver, C.byref(n), C.byref(m)) == 0
    assert (n.value, m.value) == (L.n, L.m)
    result, has = C.c_int(-1), C.c_int(-1)
    assert lib.miplib_cuda_solve(solver, C.byref(result), C.byref(has)) == 1       # no device: an error, not a fallback
    assert b"no CUDA device" in lib.miplib_get_last_error() or b"Cuda backend" in lib.miplib_get_last_error()
    assert lib.miplib_destroy_solver(solver) == 0


def test_bulk_ingress_builds_the_same_row_store(built):
    lib, C = _front(built)
    solver = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(solver), 5) == 0            # miplib_SolverBackend::Cuda
    L = gen.config(1)
    rc, _ = _load_csr(lib, C, solver, L)
    assert rc == 0, lib.miplib_get_last_error()
    n, m = C.c_int64(), C.c_int64()
    assert lib.miplib_cuda_get_shape(solver, C.byref(n), C.byref(m)) == 0
    assert (n.value, m.value) == (L.n, L.m)
    obj = C.c_double()
    assert lib.miplib_cuda_get_values(solver, None, C.byref(obj)) == 1   # no solution yet
    assert b"before a solution" in lib.miplib_get_last_error()
    bad = gen.LP(L.A[:, ::-1], L.c, L.lb, L.ub, L.row_lo, L.row_hi)
    A = bad.A.tocsr(); A.indices[:2] = A.indices[:2][::-1].copy()        # a row with decreasing columns
    bad.A = A
    s2 = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(s2), 5) == 0
    arrs = [np.array([0, 2], dtype=np.int64), np.array([1, 0], dtype=np.int32), np.array([1.0, 1.0])]
    lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
    z = np.zeros(2); o = np.ones(2); r = np.array([1.0])
    rc = lib.miplib_cuda_load_csr(s2, 1, 2, *[a.ctypes.data_as(C.c_void_p) for a in arrs],
                                  z.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                                  o.ctypes.data_as(C.c_void_p), None, r.ctypes.data_as(C.c_void_p),
                                  r.ctypes.data_as(C.c_void_p), 0)
    assert rc == 1 and b"increasing" in lib.miplib_get_last_error()
    assert lib.miplib_destroy_solver(solver) == 0 and lib.miplib_destroy_solver(s2) == 0


HAND_MPS = """NAME hand
OBJSENSE
    MAX
ROWS
 N cost
 N ignored
 L lim1
 G lim2
 E eq1
 L rng1
 G rng2
 E rng3
 E rng4
COLUMNS
 a cost 1.5 lim1 1
 a lim2 2 eq1 1
 a ignored 7
 b cost -2 lim1 3
 b rng1 1 rng2 -1
 MARKER 'MARKER' 'INTORG'
 i1 cost 1 eq1 2
 i1 rng3 1
 i2 lim2 1 rng4 4
 i3 cost 2 lim1 1
 MARKER 'MARKER' 'INTEND'
 c lim1 -1 rng1 2.5
 d cost 0.25 rng2 1
 e rng3 -3 rng4 1
 f lim2 1
 g eq1 1 rng1 1
 h cost -1 lim1 1
RHS
 rhs lim1 10 lim2 -4
 rhs eq1 3 cost 99
 rhs rng1 8 rng2 1
 rhs rng3 2 rng4 -2
RANGES
 rng rng1 3 rng2 2.5
 rng rng3 4 rng4 -1.5
BOUNDS
 UP bnd a 4
 MI bnd b
 UP bnd b -1
 LO bnd c -2.5
 UP bnd c 6
 FX bnd d 1.25
 FR bnd e 0
 MI bnd f
 UP bnd f 9
 PL g 1e30
 BV bnd h 1
 LI bnd i1 -3
 UI bnd i1 5
 UP bnd i2 12
ENDATA
"""


def test_mps_reader_agrees_with_highs_reader(built, golden, tmp_path):
    """miplib_cuda_read_mps (miplib/cuda/mps.cpp) against HiGHS' reader on a file with every
    section and bound type, then a byte-for-byte dump -> read -> dump round trip of generated
    models.  No device needed: reader, row store and writers are host code."""
    lib, C = _front(built)
    lib.miplib_cuda_read_mps.argtypes = [C.c_void_p, C.c_char_p]
    lib.miplib_cuda_dump.argtypes = [C.c_void_p, C.c_char_p]
    src = tmp_path / "hand.mps"
    src.write_text(HAND_MPS)
    solver = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(solver), 5) == 0
    assert lib.miplib_cuda_read_mps(solver, str(src).encode()) == 0, lib.miplib_get_last_error()
    n, m = C.c_int64(), C.c_int64()
    assert lib.miplib_cuda_get_shape(solver, C.byref(n), C.byref(m)) == 0
    assert (n.value, m.value) == (11, 7)
    out = tmp_path / "hand_out.mps"
    assert lib.miplib_cuda_dump(solver, str(out).encode()) == 0, lib.miplib_get_last_error()
    assert lib.miplib_destroy_solver(solver) == 0
    want, got = _read_mps_with_highs(src), _read_mps_with_highs(out)
    assert got["maximize"] and want["maximize"]
    assert np.array_equal(got["A"].toarray(), want["A"].toarray())
    assert np.array_equal(got["c"], want["c"])
    assert np.array_equal(got["lb"], want["lb"]) and np.array_equal(got["ub"], want["ub"])
    assert np.array_equal(got["integer"], want["integer"])
    assert np.array_equal(np.isfinite(got["lo"]), np.isfinite(want["lo"]))
    assert np.array_equal(np.isfinite(got["hi"]), np.isfinite(want["hi"]))
    fin = np.isfinite(want["lo"])
    assert np.abs(got["lo"][fin] - want["lo"][fin]).max() <= 1e-12          # ranges are written as hi - lo
    assert np.array_equal(got["hi"][np.isfinite(want["hi"])], want["hi"][np.isfinite(want["hi"])])
    # what the file says, spelled out for the cases readers disagree on most often
    i3 = 4                                                            # integer by MARKER only, no BOUNDS line: binary
    assert want["integer"][i3] and (want["lb"][i3], want["ub"][i3]) == (0.0, 1.0)
    assert (got["lb"][i3], got["ub"][i3]) == (0.0, 1.0)
    assert want["lb"][1] == -np.inf and want["ub"][1] == -1
    assert (want["lo"][3], want["hi"][3]) == (5.0, 8.0)               # L row with a range
    assert (want["lo"][4], want["hi"][4]) == (1.0, 3.5)               # G row with a range
    assert (want["lo"][5], want["hi"][5]) == (2.0, 6.0)               # E row, positive range
    assert (want["lo"][6], want["hi"][6]) == (-3.5, -2.0)             # E row, negative range
    for M in (want, got):                                             # and both describe the same optimum
        M["h"].run()
    # ... up to the objective constant (RHS of the objective row, -99 here), which the backend drops
    # like the constant of set_objective (reference miplib/lpsolve/solver.cpp:70-73)
    assert abs(want["h"].getInfo().objective_function_value + 99.0 - got["h"].getInfo().objective_function_value) <= 1e-9

    # A negative UP bound on a column whose lower bound is still the default 0: lp_solve, SCIP,
    # CPLEX and Gurobi (the reference's backends) move the lower bound to -inf; HiGHS keeps [0, -1]
    # and warns.  The reader follows the reference's backends.
    quirk = tmp_path / "quirk.mps"
    quirk.write_text("NAME q\nROWS\n N obj\n L r0\nCOLUMNS\n x obj 1 r0 1\n y obj 1 r0 1\nRHS\n rhs r0 5\n"
                     "BOUNDS\n UP bnd x -2\n LO bnd y 1\n UP bnd y 3\nENDATA\n")
    s3 = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(s3), 5) == 0
    assert lib.miplib_cuda_read_mps(s3, str(quirk).encode()) == 0, lib.miplib_get_last_error()
    assert lib.miplib_cuda_dump(s3, str(tmp_path / "quirk_out.mps").encode()) == 0
    assert lib.miplib_destroy_solver(s3) == 0
    Q = _read_mps_with_highs(tmp_path / "quirk_out.mps")
    assert list(Q["lb"]) == [-np.inf, 1.0] and list(Q["ub"]) == [-2.0, 3.0]
    # malformed input is an error with the file position, not a crash
    broken = tmp_path / "broken.mps"
    broken.write_text("NAME b\nROWS\n N obj\n L r0\nCOLUMNS\n x obj 1 nosuchrow 1\nENDATA\n")
    s4 = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(s4), 5) == 0
    assert lib.miplib_cuda_read_mps(s4, str(broken).encode()) == 1
    assert b"broken.mps:6" in lib.miplib_get_last_error() and b"nosuchrow" in lib.miplib_get_last_error()
    assert lib.miplib_cuda_read_mps(s4, str(tmp_path / "missing.mps").encode()) == 1
    assert lib.miplib_destroy_solver(s4) == 0

    # generated models: expression API -> dump -> read_mps -> dump must reproduce the file
    models = [("C1_lp", gen.config(1), 0.0, 1e-6),
              ("setcover_200x500_mip", gen.set_cover(200, 500, 8, 20264), 0.0, 1e-9),
              ("mkp_5x60_mip", gen.mkp(5, 60, 4, 4, 20265), 0.0, 1e-9)]
    path = tmp_path / "models.txt"
    write_models(path, models)
    rc, text = run_unit_test(built, "--cpu-only", "--filter", "dumped as MPS",
                             env={"MIPLIB_TEST_MODELS": str(path), "MIPLIB_TEST_DUMP_DIR": str(tmp_path)})
    assert rc == 0 and "1 passed" in text, text
    for label, L, _, _ in models:
        s2 = C.c_void_p()
        assert lib.miplib_create_solver(C.byref(s2), 5) == 0
        assert lib.miplib_cuda_read_mps(s2, str(tmp_path / f"{label}.mps").encode()) == 0, lib.miplib_get_last_error()
        again = tmp_path / f"{label}_again.mps"
        assert lib.miplib_cuda_dump(s2, str(again).encode()) == 0
        assert lib.miplib_destroy_solver(s2) == 0
        assert again.read_text() == (tmp_path / f"{label}.mps").read_text(), label


@pytest.mark.gpu
def test_bulk_ingress_solves_lp_and_mip(built, golden):
    lib, C = _front(built)
    for key, L, types in [("C1_lp", gen.config(1), None),
                          ("setcover_200x500_mip", gen.set_cover(200, 500, 8, 20264), np.ones(500, np.int32))]:
        solver = C.c_void_p()
        assert lib.miplib_create_solver(C.byref(solver), 5) == 0
        rc, keep = _load_csr(lib, C, solver, L, types)
        assert rc == 0, lib.miplib_get_last_error()
        result, has = C.c_int(-1), C.c_int(-1)
        assert lib.miplib_cuda_solve(solver, C.byref(result), C.byref(has)) == 0, lib.miplib_get_last_error()
        assert (result.value, has.value) == (0, 1)                     # Solver::Result::Optimal
        x = np.empty(L.n); obj = C.c_double()
        assert lib.miplib_cuda_get_values(solver, x.ctypes.data_as(C.c_void_p), C.byref(obj)) == 0
        exp = golden[key]["objective"]
        assert abs(obj.value - exp) <= 1e-6 * (1 + abs(exp))
        assert abs(float(L.c @ x) - obj.value) <= 1e-9 * (1 + abs(exp))
        assert lib.miplib_destroy_solver(solver) == 0


def test_dumped_instances_are_pinned_by_checksum_and_solve_to_the_golden(built, golden, tmp_path):
    """scripts/dump_instances.py: the config-1 / config-2 MPS files an external SCIP / lp_solve would
    be run on.  Their SHA-256 is pinned in tests/golden/instances.json, the C1 file read back by an
    independent reader (HiGHS) is the generator's model entry for entry, and an exact solve of the
    FILE gives the golden objective - so a parity run elsewhere is a run on this instance."""
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    p = subprocess.run([sys.executable, str(root / "scripts" / "dump_instances.py"), "--configs", "1", "2",
                        "--out", str(tmp_path), "--check"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    M = _read_mps_with_highs(tmp_path / "C1.mps")
    L = gen.config(1)
    A = L.A.tocsr(); A.sort_indices()
    assert (M["A"] != A).nnz == 0
    assert np.array_equal(M["c"], L.c) and np.array_equal(M["lb"], L.lb) and np.array_equal(M["ub"], L.ub)
    assert np.array_equal(M["lo"], L.row_lo) and np.array_equal(M["hi"], L.row_hi)
    M["h"].run()
    obj = M["h"].getInfo().objective_function_value
    assert abs(obj - golden["C1_lp"]["objective"]) <= 1e-9 * (1 + abs(obj))


@pytest.mark.gpu
def test_full_size_mip_configs_under_a_time_limit(built):
    """BASELINE configs 4 / 5 as MIPs at FULL size (20k x 50k set cover, 600 x 100 100 knapsack with switch rows):
    nobody here can prove their optima (HiGHS has no bound for config 4 and no incumbent for config 5 after a
    minute), so the branch and bound runs under a time limit and what it returns is checked size-independently:
    Interrupted with an incumbent that is integral and feasible for every row and bound (recomputed in FP64), the
    reported objective is c'x, the dual bound is on the right side of the incumbent and not below the root LP
    bound, and the gap is what rounding + 1-opt + a few thousand nodes give (measured: 6.3 % / 0.24 %)."""
    lib, C = _front(built)
    lib.miplib_cuda_set_time_limit.argtypes = [C.c_void_p, C.c_double]
    for cfg, root_bound, max_gap in ((4, 40211.9658, 0.10), (5, 2305927.33, 0.01)):
        L = gen.config(cfg)
        solver = C.c_void_p()
        assert lib.miplib_create_solver(C.byref(solver), 5) == 0
        lib.miplib_cuda_set_verbose(solver, 0)
        rc, keep = _load_csr(lib, C, solver, L, np.ones(L.n, np.int32))       # every column Binary
        assert rc == 0, lib.miplib_get_last_error()
        assert lib.miplib_cuda_set_time_limit(solver, 12.0) == 0
        result, has = C.c_int(-1), C.c_int(-1)
        assert lib.miplib_cuda_solve(solver, C.byref(result), C.byref(has)) == 0, lib.miplib_get_last_error()
        assert (result.value, has.value) == (4, 1)                            # Solver::Result::Interrupted, incumbent
        x = np.empty(L.n); obj = C.c_double(); bound = C.c_double(); proven = C.c_int(-1)
        assert lib.miplib_cuda_get_values(solver, x.ctypes.data_as(C.c_void_p), C.byref(obj)) == 0
        assert lib.miplib_cuda_get_bound(solver, C.byref(bound), C.byref(proven)) == 0
        assert proven.value == 0
        assert np.abs(x - np.round(x)).max() == 0.0 and x.min() >= 0.0 and x.max() <= 1.0
        ax = L.A @ x
        assert np.all(ax <= L.row_hi + 1e-9) and np.all(ax >= L.row_lo - 1e-9)
        assert abs(float(L.c @ x) - obj.value) <= 1e-9 * (1 + abs(obj.value))
        sgn = -1.0 if L.maximize else 1.0
        assert sgn * bound.value <= sgn * obj.value                            # bound on the right side
        assert sgn * bound.value >= sgn * root_bound - 2e-6 * (1 + abs(root_bound))   # no worse than the root LP
        assert abs(obj.value - bound.value) <= max_gap * abs(obj.value), (cfg, obj.value, bound.value)
        assert lib.miplib_destroy_solver(solver) == 0


@pytest.mark.gpu
def test_bulk_ingress_eager_device_load_then_append(built, golden, monkeypatch):
    """miplib_cuda_load_csr on an empty solver of a big model loads the device at once, out of the caller's arrays,
    while host threads fill the row store (threshold lowered here so that config 1 takes that path).  The solve
    must then find the model on the device (same optimum), and a later structural change - a second block of
    columns and rows appended by another load_csr - must reach the device too: the optimum of the block-diagonal
    model is the sum of the two."""
    monkeypatch.setenv("MIPLIB_CUDA_EAGER_NNZ", "0")
    lib, C = _front(built)
    L1, L2 = gen.config(1), gen.lp(300, 500, 6, 7)
    solver = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(solver), 5) == 0
    lib.miplib_cuda_set_verbose(solver, 0)
    rc, keep1 = _load_csr(lib, C, solver, L1)
    assert rc == 0, lib.miplib_get_last_error()
    result, has, obj = C.c_int(-1), C.c_int(-1), C.c_double()
    assert lib.miplib_cuda_solve(solver, C.byref(result), C.byref(has)) == 0, lib.miplib_get_last_error()
    assert (result.value, has.value) == (0, 1)
    x = np.empty(L1.n)
    assert lib.miplib_cuda_get_values(solver, x.ctypes.data_as(C.c_void_p), C.byref(obj)) == 0
    e1 = golden["C1_lp"]["objective"]
    assert abs(obj.value - e1) <= 1e-6 * (1 + abs(e1))
    assert abs(float(L1.c @ x) - obj.value) <= 1e-9 * (1 + abs(e1))
    # append a second, independent block: load_csr REPLACES the objective, so the first block's columns cost 0 now
    rc, keep2 = _load_csr(lib, C, solver, L2)
    assert rc == 0, lib.miplib_get_last_error()
    n, m = C.c_int64(), C.c_int64()
    assert lib.miplib_cuda_get_shape(solver, C.byref(n), C.byref(m)) == 0
    assert (n.value, m.value) == (L1.n + L2.n, L1.m + L2.m)
    assert lib.miplib_cuda_solve(solver, C.byref(result), C.byref(has)) == 0, lib.miplib_get_last_error()
    assert (result.value, has.value) == (0, 1)
    xx = np.empty(L1.n + L2.n)
    assert lib.miplib_cuda_get_values(solver, xx.ctypes.data_as(C.c_void_p), C.byref(obj)) == 0
    e2 = golden["lp_300x500_s7"]["objective"]
    assert abs(obj.value - e2) <= 2e-6 * (1 + abs(e2))
    # the first block is still constrained by its own rows
    ax = L1.A @ xx[:L1.n]
    assert np.all(ax <= L1.row_hi + 1e-4 * (1 + np.abs(L1.row_hi))) and np.all(ax >= L1.row_lo - 1e-4 * (1 + np.abs(L1.row_lo)))
    assert abs(float(L2.c @ xx[L1.n:]) - obj.value) <= 1e-9 * (1 + abs(e2))
    assert lib.miplib_destroy_solver(solver) == 0


@pytest.mark.parametrize("case", ["setcover", "mkp"])
def test_branch_and_bound_local_machinery_on_cpu(built, case):
    """Host side of the tree, no GPU involved (miplib_cuda_selftest_local_search): (i) propagating a child's box
    from the rows of its branched column only gives the same verdict and the same box as propagating from every
    row; (ii) an LP point (HiGHS on the CPU), threshold-rounded, comes back from the greedy repair and the 1-opt
    pass integral and feasible for every row, cheaper than plain rounding-up / rounding-down, and 1-opt: no single
    column can be moved to its cheaper side without breaking a row."""
    from oracle import highs_ref
    lib, C = _front(built)
    L = gen.set_cover(300, 700, 8, 20264) if case == "setcover" else gen.mkp(12, 300, 4, 6, 20265)
    lp = highs_ref.solve_lp(L)
    assert lp["status"] == 0
    x_lp = np.ascontiguousarray(lp["x"], dtype=np.float64)
    solver = C.c_void_p()
    assert lib.miplib_create_solver(C.byref(solver), 5) == 0
    rc, keep = _load_csr(lib, C, solver, L, np.ones(L.n, np.int32))          # every column Binary
    assert rc == 0, lib.miplib_get_last_error()
    x = np.empty(L.n); feas = C.c_int(-1); mism = C.c_int(-1)
    lib.miplib_cuda_selftest_local_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.miplib_cuda_selftest_local_search(solver, x_lp.ctypes.data_as(C.c_void_p), 200,
                                                 x.ctypes.data_as(C.c_void_p), C.byref(feas), C.byref(mism)) == 0, \
        lib.miplib_get_last_error()
    assert mism.value == 0
    assert feas.value == 1
    assert np.array_equal(x, np.round(x)) and x.min() >= 0 and x.max() <= 1
    A = L.A.tocsc()
    ax = L.A @ x
    assert np.all(ax <= L.row_hi + 1e-9) and np.all(ax >= L.row_lo - 1e-9)
    sgn = -1.0 if L.maximize else 1.0
    cost = sgn * L.c
    # no worse than the plain rounding of this model class when that is feasible (covering: up), and than the
    # trivial point otherwise (the switch rows of the knapsack make rounding down infeasible; all-zero is feasible)
    plain = np.ceil(x_lp - 1e-9) if case == "setcover" else np.floor(x_lp + 1e-9)
    ap = L.A @ plain
    if not (np.all(ap <= L.row_hi + 1e-9) and np.all(ap >= L.row_lo - 1e-9)):
        plain = np.zeros(L.n)
        assert np.all(L.row_lo <= 1e-9) and np.all(L.row_hi >= -1e-9)
    assert cost @ x <= cost @ plain + 1e-9
    # 1-opt: no column can take one step towards its cheaper side
    for j in range(L.n):
        if cost[j] == 0:
            continue
        d = -1.0 if cost[j] > 0 else 1.0
        if not (0 <= x[j] + d <= 1):
            continue
        col = A[:, j]
        a2 = ax.copy(); a2[col.indices] += col.data * d
        assert np.any(a2 > L.row_hi + 1e-9) or np.any(a2 < L.row_lo - 1e-9), j
    # within reach of the LP bound (measured: a few per cent on these sizes)
    assert abs(cost @ x - sgn * lp["obj"]) <= 0.15 * abs(lp["obj"])
    assert lib.miplib_destroy_solver(solver) == 0
