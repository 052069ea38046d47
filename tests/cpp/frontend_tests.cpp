// tests/cpp/frontend_tests.cpp - the assertions of the reference's front-end test cases
// (test/unit/expr.cpp:11-114, :220-426; test/unit/solver.cpp:69-152), re-hosted on the Cuda
// backend.  Only the ASSERTIONS are taken over (they are the known-answer set); the harness and
// the library under test are new.  None of these needs a GPU: a CudaSolver only touches the
// device inside solve().
#include <miplib/capi.hpp>
#include <miplib/solver.hpp>

#include "harness.hpp"

using namespace miplib;
static constexpr Solver::Backend B = Solver::Backend::Cuda;

CPU_TEST("Expressions: printing, cancellation, degree limits (ref test/unit/expr.cpp:33-56)")
{
  Solver solver(B);
  Var v1(solver, Var::Type::Binary, "v1");
  Var v2(solver, Var::Type::Binary, "v2");
  Var v3(solver, Var::Type::Continuous, "v3");

  REQUIRE_EQ_STR(Expr(3), "3");
  REQUIRE_EQ_STR(-v1, "-v1");
  REQUIRE_EQ_STR(v1 - v2, "v1 - v2");
  REQUIRE_EQ_STR(2 * v1 - v2, "2 v1 - v2");
  REQUIRE_EQ_STR(2 * v1 - 2 * v1, "0");
  REQUIRE_EQ_STR(2 * (v1 - v2), "2 v1 - 2 v2");
  REQUIRE_EQ_STR(2 * (v1 - v2) + v1, "3 v1 - 2 v2");
  REQUIRE_EQ_STR(2 * (v1 - v2) - 2 * v1, "-2 v2");
  REQUIRE_EQ_STR(2 * (v1 - v2) / 2, "v1 - v2");
  REQUIRE_EQ_STR(v1 * v2, "v1 v2");
  REQUIRE_EQ_STR((v1 + v2) * (v2 + v3), "v1 v2 + v1 v3 + v2 v2 + v2 v3");
  REQUIRE_EQ_STR((v1 + v2) * (v1 - v2), "v1 v1 - v2 v2");
  REQUIRE_EQ_STR((1 * v1 + 2 * v2) * (3 * v2 + 4 * v3), "3 v1 v2 + 4 v1 v3 + 6 v2 v2 + 8 v2 v3");

  REQUIRE_THROWS_AS(v1 * v2 * v3, std::logic_error);               // cubic
  REQUIRE_THROWS_AS((v1 + 2) * (v2 + 2) * v3, std::logic_error);
  REQUIRE_THROWS_AS(v1 * v1 * v1 * v1, std::logic_error);          // quartic
}

CPU_TEST("Expressions: in-place building and handle semantics")
{
  Solver solver(B);
  Var a(solver, Var::Type::Continuous, 0, 1, "a");
  Var b(solver, Var::Type::Continuous, 0, 1, "b");
  Expr e;
  for (int i = 0; i < 1000; ++i) { e += a; e -= b; }
  e += 2.5;
  REQUIRE_EQ_STR(e, "1000 a - 1000 b + 2.5");
  e -= 1000 * Expr(a);
  REQUIRE_EQ_STR(e, "-1000 b + 2.5");
  REQUIRE(e.linear_vars().size() == 1 && e.linear_coeffs()[0] == -1000);
  Expr shared = e;              // a handle, like the reference's Expr (miplib/expr.hpp:47)
  shared += b;
  REQUIRE_EQ_STR(e, "-999 b + 2.5");
  Expr deep = e.copy();
  deep += b;
  REQUIRE_EQ_STR(e, "-999 b + 2.5");
  REQUIRE_EQ_STR(deep, "-998 b + 2.5");
  REQUIRE((a * b).is_quadratic() && !(a * b).is_linear() && Expr(1).is_constant());
  REQUIRE_EQ_STR((a * b + a).div_by_nonzero(a), "b + 1");
  REQUIRE_THROWS_AS((a * b + 1).div_by_nonzero(a), std::logic_error);
  REQUIRE_THROWS_AS(Expr(1).solver(), std::logic_error);
  REQUIRE((2 * a + b).arity() == 2);
}

CPU_TEST("Constraints: canonical rows, negation, indicator printing (ref test/unit/expr.cpp:83-113)")
{
  Solver solver(B);
  Var v1(solver, Var::Type::Binary, "v1");
  Var v2(solver, Var::Type::Binary, "v2");
  Var v3(solver, Var::Type::Integer, "v3");

  REQUIRE_EQ_STR(v1 == v2, "v1 - v2 = 0");
  REQUIRE_EQ_STR(v1 >= v2, "-v1 + v2 <= 0");
  REQUIRE_EQ_STR(v1 >= 1, "-v1 + 1 <= 0");
  REQUIRE_EQ_STR(v1 * v2 >= 1, "-v1 v2 + 1 <= 0");
  REQUIRE_EQ_STR((1 * v1 + 2 * v2) * (3 * v2 + 4 * v3) == 2, "3 v1 v2 + 4 v1 v3 + 6 v2 v2 + 8 v2 v3 - 2 = 0");

  REQUIRE_THROWS_AS(!v3, std::logic_error);
  REQUIRE_THROWS_AS(!(v1 + 1), std::logic_error);
  REQUIRE_THROWS_AS(!(2 * v1), std::logic_error);
  REQUIRE_EQ_STR(!v1, "v1 = 0");
  REQUIRE_EQ_STR(!(1 - v1), "-v1 + 1 = 0");

  REQUIRE_EQ_STR((v1 == 0) >> (v2 >= 1), "v1 = 0 -> -v2 + 1 <= 0");
  REQUIRE_EQ_STR(!v1 >> (v2 >= 1), "v1 = 0 -> -v2 + 1 <= 0");
  REQUIRE_EQ_STR((2 * v1 == 0) >> (v2 >= 1), "2 v1 = 0 -> -v2 + 1 <= 0");
  REQUIRE_EQ_STR((v1 == 0) << (v2 == 1), "v2 - 1 = 0 -> v1 = 0");
  REQUIRE_EQ_STR(!v1 << v2, "v2 - 1 = 0 -> v1 = 0");
  REQUIRE_EQ_STR((v1 == 1) >> (2 * v2 + v3 <= 1), "v1 - 1 = 0 -> 2 v2 + v3 - 1 <= 0");
  REQUIRE_THROWS_AS(Expr(1) <= Expr(2), std::logic_error);
}

CPU_TEST("Lower/upper bounds: defaults and interval arithmetic (ref test/unit/expr.cpp:238-297)")
{
  Solver solver(B);
  Var v1(solver, Var::Type::Binary);
  REQUIRE(v1.lb() == 0);
  REQUIRE(v1.ub() == 1);
  Var v2(solver, Var::Type::Integer);
  REQUIRE(v2.lb() == -solver.infinity());
  REQUIRE(v2.ub() == solver.infinity());
  Var v3(solver, Var::Type::Continuous);
  REQUIRE(v3.lb() == -solver.infinity());
  REQUIRE(v3.ub() == solver.infinity());
  Var v4(solver, Var::Type::Continuous, 1, 3);
  REQUIRE(v4.lb() == 1);
  REQUIRE(v4.ub() == 3);
  Var v5(solver, Var::Type::Continuous, -1, 3);
  REQUIRE(v5.lb() == -1);
  REQUIRE(v5.ub() == 3);
  Var v6(solver, Var::Type::Integer, -1);
  REQUIRE(v6.lb() == -1);
  REQUIRE(v6.ub() == solver.infinity());

  REQUIRE((2 * v1).lb() == 0);
  REQUIRE((2 * v1).ub() == 2);
  REQUIRE((2 * v1 + 3).lb() == 3);
  REQUIRE((2 * v1 + 3).ub() == 5);
  REQUIRE((v1 + v2).lb() == -solver.infinity());
  REQUIRE((v1 + v2).ub() == solver.infinity());
  REQUIRE((v2 + v3).lb() == -solver.infinity());
  REQUIRE((v2 + v3).ub() == solver.infinity());
  REQUIRE((2 * v4 - v5).lb() == -1);
  REQUIRE((2 * v4 - v5).ub() == 7);
  REQUIRE((v1 * v1).lb() == 0);
  REQUIRE((v1 * v1).ub() == 1);
  REQUIRE((v1 * v2).lb() == -solver.infinity());
  REQUIRE((v1 * v2).ub() == solver.infinity());
  REQUIRE((v2 * v3).lb() == -solver.infinity());
  REQUIRE((v2 * v3).ub() == solver.infinity());
  REQUIRE((v4 * v5).lb() == -3);
  REQUIRE((v4 * v5).ub() == 9);
  REQUIRE((v5 * v5).lb() == 0);
  REQUIRE((v5 * v5).ub() == 9);
  REQUIRE((v5 * v6).lb() == -solver.infinity());
  REQUIRE((v5 * v6).ub() == solver.infinity());
  v4.set_ub(2);
  REQUIRE((v4 * v5).lb() == -2);
  REQUIRE((v4 * v5).ub() == 6);
  auto const b = (2 * v4 - v5).bounds();
  REQUIRE(b.first == -1 && b.second == 5);
  REQUIRE_THROWS_AS((v4 * v5).bounds(), std::logic_error);
}

// ---- "Indicator constraint reformulation" (ref test/unit/expr.cpp:319-425), one case per SECTION
CPU_TEST("Reformulation: none when the implicand is unbounded")
{
  Solver solver(B);
  Var z(solver, Var::Type::Binary, "z");
  {
    Var x(solver, Var::Type::Integer, "x");
    REQUIRE(!(z >> (x <= 0)).has_reformulation());
  }
  {
    Var x(solver, Var::Type::Integer, -1, std::nullopt, "x");
    REQUIRE(!(z >> (x <= 0)).has_reformulation());
    REQUIRE_THROWS_AS((z >> (x <= 0)).reformulation(), std::logic_error);
  }
  {
    Var x(solver, Var::Type::Integer, std::nullopt, 1, "x");
    REQUIRE(!(z >> (x == 0)).has_reformulation());
  }
}

CPU_TEST("Reformulation: inequalities (big-M strings)")
{
  Solver solver(B);
  Var z(solver, Var::Type::Binary, "z");
  Var x(solver, Var::Type::Integer, std::nullopt, 2, "x");
  {
    REQUIRE((z >> (x <= 0)).has_reformulation());
    auto const r = (z >> (x <= 0)).reformulation();
    REQUIRE(r.size() == 1);
    REQUIRE_EQ_STR(r[0], "x + 2 z - 2 <= 0");
  }
  {
    auto const r = ((z == 1) >> (x <= 0)).reformulation();
    REQUIRE(r.size() == 1);
    REQUIRE_EQ_STR(r[0], "x + 2 z - 2 <= 0");
  }
  {
    REQUIRE((!z >> (x <= 0)).has_reformulation());
    auto const r = (!z >> (x <= 0)).reformulation();
    REQUIRE(r.size() == 1);
    REQUIRE_EQ_STR(r[0], "x - 2 z <= 0");
  }
}

CPU_TEST("Reformulation: equalities give up to two rows")
{
  Solver solver(B);
  Var z(solver, Var::Type::Binary, "z");
  Var x(solver, Var::Type::Integer, 2, 4, "x");
  {
    REQUIRE((z >> (x == 3)).has_reformulation());
    auto const r = (z >> (x == 3)).reformulation();
    REQUIRE(r.size() == 2);
    REQUIRE_EQ_STR(r[0], "x + z - 4 <= 0");
    REQUIRE_EQ_STR(r[1], "-x + z + 2 <= 0");
  }
  {
    auto const r = (z >> (x == 2)).reformulation();
    REQUIRE(r.size() == 1);
    REQUIRE_EQ_STR(r[0], "x + 2 z - 4 <= 0");
  }
}

CPU_TEST("Reformulation: non-unary implicants")
{
  Solver solver(B);
  Var z1(solver, Var::Type::Binary, "z1");
  Var z2(solver, Var::Type::Binary, "z2");
  {
    Var x(solver, Var::Type::Integer, std::nullopt, 1, "x");
    REQUIRE(!((z1 - z2) >> (x == 0)).has_reformulation());     // implicant is not a half space
  }
  {
    Var x(solver, Var::Type::Integer, std::nullopt, 2, "x");
    auto const r = ((z1 + z2 == 2) >> (x <= 0)).reformulation();
    REQUIRE(r.size() == 1);
    REQUIRE_EQ_STR(r[0], "x + 2 z1 + 2 z2 - 4 <= 0");
    auto const r2 = ((-z1 - z2 == -2) >> (x <= 0)).reformulation();
    REQUIRE(r2.size() == 1);
    REQUIRE_EQ_STR(r2[0], "x + 2 z1 + 2 z2 - 4 <= 0");
  }
  {
    Var x(solver, Var::Type::Integer, 2, 4, "x");
    auto const r = ((z1 + z2 == 2) >> (x == 3)).reformulation();
    REQUIRE(r.size() == 2);
    REQUIRE_EQ_STR(r[0], "x + z1 + z2 - 5 <= 0");
    REQUIRE_EQ_STR(r[1], "-x + z1 + z2 + 1 <= 0");
  }
}

// ---- "Indicator constraint automatic reformulation" (ref test/unit/solver.cpp:85-151) -----------
CPU_TEST("Indicator policy: PassThrough throws on every form the backend cannot take")
{
  Solver solver(B);
  solver.set_indicator_constraint_policy(Solver::IndicatorConstraintPolicy::PassThrough);
  Var z1(solver, Var::Type::Binary, "z1");
  Var z2(solver, Var::Type::Binary, "z2");
  Var zi(solver, Var::Type::Integer, "zi");
  REQUIRE_THROWS_AS(solver.add((z1 == 0) << (zi >= 1)), std::logic_error);
  REQUIRE_THROWS_AS(solver.add((z1 + z2 == 2) >> (z2 >= 1)), std::logic_error);
  REQUIRE_THROWS_AS(solver.add((z1 == 0) >> (z2 * z1 >= 1)), std::logic_error);
  REQUIRE_THROWS_AS(solver.add((z1 * z2 == 0) >> (z2 >= 1)), std::logic_error);
}

CPU_TEST("Indicator policy: ReformulateIfUnsupported linearises half-space implicants")
{
  Solver solver(B);
  solver.set_indicator_constraint_policy(Solver::IndicatorConstraintPolicy::ReformulateIfUnsupported);
  Var z1(solver, Var::Type::Binary, "z1");
  Var z2(solver, Var::Type::Binary, "z2");
  Var zi(solver, Var::Type::Integer, "zi");
  REQUIRE_THROWS_AS(solver.add((z1 == 0) << (zi >= 1)), std::logic_error);   // inequality implicant
  REQUIRE_NOTHROW(solver.add((z1 + z2 == 2) >> (z2 >= 1)));
  // like Lpsolve, no quadratic rows: the reference skips those two sections for it
  REQUIRE(!solver.supports_quadratic_constraints() && !solver.supports_quadratic_objective());
  REQUIRE_THROWS_AS(solver.add(z1 * z2 <= 1), std::logic_error);
}

CPU_TEST("Solver facade: error conventions of the plugin seam")
{
  Solver solver(B, false);
  Var v1(solver, Var::Type::Integer, 1, 3, "v1");
  Var v2(solver, Var::Type::Continuous, 0, 1, "v2");
  REQUIRE_THROWS_AS(v1.value(), std::logic_error);                       // no solution yet
  REQUIRE_THROWS_AS(solver.get_objective_value(), std::logic_error);
  Constr c = v1 + v2 <= 3;
  solver.add(c);
  REQUIRE_THROWS_AS(solver.add(c), std::logic_error);                    // posted twice
  REQUIRE_THROWS_AS(solver.add(v1 >= 4), std::logic_error);              // trivially unsat
  REQUIRE_THROWS_AS(solver.maximize(v1 * v2), std::logic_error);         // quadratic objective
  REQUIRE_THROWS_AS(Solver(Solver::Backend::Scip), std::logic_error);    // not compiled
  REQUIRE(Solver::backend_is_compiled(B) && !Solver::backend_is_compiled(Solver::Backend::Gurobi));
  REQUIRE(solver.infinity() == 1e20);
  REQUIRE_EQ_STR(B, "Cuda");
  REQUIRE_EQ_STR(Solver::Backend::Lpsolve, "Lpsolve");
  solver.set_feasibility_tolerance(1e-7);
  REQUIRE(solver.get_feasibility_tolerance() == 1e-7);
  Solver other(B, false);
  Var w(other, Var::Type::Continuous, 0, 1, "w");
  REQUIRE_THROWS_AS(solver.add(w + v2 <= 1), std::logic_error);          // variable of another solver
  solver.remove(c);
  REQUIRE_NOTHROW(solver.add(c));                                        // can be posted again
  REQUIRE(Solver::backend_info().count(B) == 1);
}

CPU_TEST("C API: create / copy / destroy and the appended Cuda enumerator (ref capi.hpp:6-25)")
{
  REQUIRE(static_cast<int>(Cuda) == 5 && static_cast<int>(BestAtRunTime) == 4);
  Solver* s = nullptr;
  REQUIRE(miplib_create_solver(&s, Cuda) == 0);
  REQUIRE(s != nullptr);
  Solver* s2 = nullptr;
  REQUIRE(miplib_shallow_copy_solver(&s2, s) == 0);
  Var* v = nullptr;
  REQUIRE(miplib_create_var(&v, s, miplib_VarType::Binary) == 0);
  REQUIRE(v->ub() == 1);
  Var* v2 = nullptr;
  REQUIRE(miplib_shallow_copy_var(&v2, v) == 0);
  REQUIRE(v2->is_same(*v));
  REQUIRE(miplib_destroy_var(v2) == 0);
  REQUIRE(miplib_destroy_var(v) == 0);
  REQUIRE(miplib_destroy_solver(s2) == 0);
  REQUIRE(miplib_destroy_solver(s) == 0);
  Solver* bad = nullptr;
  REQUIRE(miplib_create_solver(&bad, Scip) == 1);
  REQUIRE(std::string(miplib_get_last_error()).find("not compiled") != std::string::npos);
}

CPU_TEST("Row scaling helper (ref miplib/util/scale.cpp:52-112)")
{
  Solver solver(B);
  Var a(solver, Var::Type::Continuous, 0, 1, "a");
  Var b(solver, Var::Type::Continuous, 0, 1, "b");
  Constr c = 1024 * 1024 * a + 4096 * b <= 2048;
  Constr s = c.scale();
  // range [2048, 1048576] -> divisor sqrt(1048576) = 1024 -> factor 2^-10
  REQUIRE_EQ_STR(s, "1024 a + 4 b - 2 <= 0");
  Constr ok = a + b <= 1;
  REQUIRE_EQ_STR(ok.scale(), "a + b - 1 <= 0");                          // already in [1e-4, 1e4]
  Var u(solver, Var::Type::Continuous, "u");
  REQUIRE_THROWS_AS((1e6 * u <= 1).scale(), std::logic_error);           // unbounded domain
}

CPU_TEST("dump() writes MPS and LP text of the lowered model")
{
  Solver solver(B, false);
  Var x(solver, Var::Type::Integer, 0, 4, "x");
  Var y(solver, Var::Type::Continuous, "y");
  solver.add(2 * x + y <= 7);
  solver.add(x - y == 1);
  solver.set_objective(Solver::Sense::Maximize, 3 * x + y);
  solver.dump("/tmp/miplib_cuda_dump_test.mps");
  solver.dump("/tmp/miplib_cuda_dump_test.lp");
  REQUIRE_THROWS_AS(solver.dump("/tmp/model.txt"), std::logic_error);
}
