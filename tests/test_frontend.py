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


def write_models(path, models):
    """Text fixture read by tests/cpp/solver_tests.cpp ('Generated models ...')."""
    def f(v):
        return "inf" if v == np.inf else "-inf" if v == -np.inf else repr(float(v))
    with open(path, "w") as fh:
        fh.write(f"{len(models)}\n")
        for label, L, expected, tol in models:
            A = L.A.tocsr()
            integer = L.integer if L.integer is not None else np.zeros(L.n, bool)
            fh.write(f"{label} {L.n} {L.m}\n")
            for j in range(L.n):
                binary = integer[j] and L.lb[j] == 0 and L.ub[j] == 1
                t = 1 if binary else 2 if integer[j] else 0
                fh.write(f"{t} {f(L.lb[j])} {f(L.ub[j])} {f(L.c[j])}\n")
            for i in range(L.m):
                s, e = A.indptr[i], A.indptr[i + 1]
                terms = " ".join(f"{c} {f(v)}" for c, v in zip(A.indices[s:e], A.data[s:e]))
                fh.write(f"{f(L.row_lo[i])} {f(L.row_hi[i])} {e - s} {terms}\n")
            fh.write(f"{int(L.maximize)} {f(expected)} {tol}\n")


@pytest.mark.gpu
def test_kats_and_generated_models_on_gpu(built, golden, tmp_path):
    models = [
        ("C1_lp", gen.config(1), golden["C1_lp"]["objective"], 1e-6),
        ("lp_300x500_s7", gen.lp(300, 500, 6, 7), golden["lp_300x500_s7"]["objective"], 1e-6),
        ("setcover_200x500_mip", gen.set_cover(200, 500, 8, 20264), golden["setcover_200x500_mip"]["objective"], 1e-9),
        ("mkp_5x60_mip", gen.mkp(5, 60, 4, 4, 20265), golden["mkp_5x60_mip"]["objective"], 1e-9),
        ("setcover_400x1000_mip", gen.set_cover(400, 1000, 8, 20264), golden["setcover_400x1000_mip"]["objective"], 1e-9),
    ]
    path = tmp_path / "models.txt"
    write_models(path, models)
    rc, out = run_unit_test(built, env={"MIPLIB_TEST_MODELS": str(path)})
    print(out)
    assert rc == 0, out
    assert "0 failed" in out and "0 skipped" in out, out
