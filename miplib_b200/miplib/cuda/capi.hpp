// miplib/cuda/capi.hpp - C entry points NEW with the Cuda backend: bulk
// model ingress that bypasses the expression tree, plus solve / read-back, in the style of the
// reference's miplib/capi.hpp (rc 0 / 1, message via miplib_get_last_error()).
//
// Why: the reference's C API stops at creating solvers and variables (miplib/capi.hpp:16-23), and
// building a 32 M-nonzero model through Var / Expr objects costs tens of seconds and gigabytes on
// the host before solve() is ever called (reference miplib/expr.cpp:53-119, :599-633).  These
// calls hand CSR arrays straight to the backend's row store.
#pragma once

#include <cstdint>

#include "../capi.hpp"

extern "C" {

// Appends n columns and m rows to a Cuda-backend solver.  col_idx counts from the first NEW column
// and must be increasing inside a row.  var_type: miplib_VarType values as int32, or NULL for all
// Continuous.  Bounds / row sides beyond +-1e20 mean "none".  Replaces the objective.
int miplib_cuda_load_csr(miplib::Solver* p_solver, std::int64_t m, std::int64_t n,
                         std::int64_t const* row_ptr, std::int32_t const* col_idx, double const* val,
                         double const* c, double const* lb, double const* ub,
                         std::int32_t const* var_type, double const* row_lo, double const* row_hi,
                         int maximize);

// Reads a free-format MPS file into a Cuda-backend solver (appends its columns and rows, replaces
// the objective): N / L / G / E rows, integer markers, RHS, RANGES, every BOUNDS type, OBJSENSE.
int miplib_cuda_read_mps(miplib::Solver* p_solver, char const* path);

// Solver::dump(): the lowered model as free-format MPS (*.mps) or CPLEX-LP text (*.lp).
int miplib_cuda_dump(miplib::Solver* p_solver, char const* path);

// Solver::solve(); *result receives Solver::Result as int, *has_solution 0 / 1.
int miplib_cuda_solve(miplib::Solver* p_solver, int* result, int* has_solution);

// Values of ALL columns in creation order (x may be NULL) and the objective value.
int miplib_cuda_get_values(miplib::Solver* p_solver, double* x, double* objective);

// Statistics of the last solve: PDHG iterations over all LP solves, kernels launched, branch-and-
// bound nodes, wall seconds inside Solver::solve().  Any pointer may be NULL.
int miplib_cuda_get_info(miplib::Solver* p_solver, std::int64_t* lp_iterations, std::int64_t* kernel_launches,
                         std::int64_t* nodes, double* seconds);

// The reference's C API always creates verbose solvers (miplib/capi.cpp:58-63 -> Solver's default
// second argument, miplib/solver.hpp:34); this turns the backend's iteration log on / off.
int miplib_cuda_set_verbose(miplib::Solver* p_solver, int verbose);

// Solver::set_time_limit / set_nr_threads (= GPUs the branch and bound deals its node LPs to) for hosts
// that only hold the C handle, and the dual side of the last solve: *best_bound is the proven bound on the
// optimum (the incumbent's objective once the tree is finished; the smallest bound among the open nodes
// of an interrupted tree; the dual objective of a continuous model), *proven_optimal 0 / 1.
int miplib_cuda_set_time_limit(miplib::Solver* p_solver, double seconds);
int miplib_cuda_set_nr_threads(miplib::Solver* p_solver, std::int64_t gpus);
int miplib_cuda_get_bound(miplib::Solver* p_solver, double* best_bound, int* proven_optimal);

// Host-only self test of the branch and bound's local machinery on the loaded (pure-integer) model - no device
// involved: propagation from the rows of a branched column only must equal propagation from every row (`trials`
// random branchings; *propagate_mismatches), and x_lp, threshold-rounded, must come back from the greedy repair and
// the 1-opt pass as a point (x_out, n doubles) that satisfies every row (*feasible).
int miplib_cuda_selftest_local_search(miplib::Solver* p_solver, double const* x_lp, int trials, double* x_out,
                                      int* feasible, int* propagate_mismatches);

// Number of columns / alive rows of the lowered model.
int miplib_cuda_get_shape(miplib::Solver* p_solver, std::int64_t* n, std::int64_t* m);

}
