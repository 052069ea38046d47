// miplib/cuda/bnb.hpp - branch and bound around the device LP engine, for models with Binary /
// Integer columns.  In the reference this lives inside the vendor call (SCIPsolve,
// miplib/scip/solver.cpp:370; lp_solve's B&B behind ::solve, miplib/lpsolve/solver.cpp:156).
//
// Host side: best-first tree, activity-based bound propagation (decides most small models with
// no LP at all and catches the structurally infeasible nodes PDHG cannot certify), rounding
// heuristics, pruning by the DUAL objective of the node LP (a valid bound, unlike the primal value
// of an inexact first-order solve).  Device side: one PDHG solve per node, warm started from the
// parent, stopped early by MIPCU_PARAM_OBJ_CUTOFF as soon as its dual bound crosses the incumbent -
// or, the other way round, as soon as a 1e-4-accurate iterate shows that the node can NOT be pruned
// by bound (MIPCU_BRANCH): such a node is branched on at once; only nodes whose LP value is close to
// the incumbent, and nodes whose LP point looks integral, are solved to the full tolerance.
//
// Root strengthening: implied-bound rows.  A row in which fixing one binary z to a value forces
// another column to a tighter bound (the big-M rows IndicatorConstr::reformulation() emits,
// reference miplib/constr.cpp:225-271: sum_{j in g} x_j - |g| z <= 0) yields the valid rows
// x_j <= u0 + (u - u0) z, here x_j - z <= 0.  SCIP and lp_solve separate the same family inside
// their solve; without them the aggregated big-M row leaves a 20-30 % root gap on config 5's
// family and the tree is two orders of magnitude larger (scripts/bnb_proto.py).
#pragma once

#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

#include "solver.hpp"

namespace miplib {

struct CudaBranchAndBound
{
  explicit CudaBranchAndBound(CudaSolver& solver);
  std::pair<Solver::Result, bool> run();

  // Host-only self test of the tree's local machinery (no device needed; miplib_cuda_selftest_local_search):
  // the root box is propagated, `trials` random single-column branchings are propagated twice - from every
  // row and from the rows of the branched column only - and must give the same verdict and the same box;
  // then x_lp is threshold-rounded (0.5), repaired and polished by one_opt.  Returns the number of
  // propagation mismatches; *feasible says whether the returned point satisfies every row and bound.
  int selftest_local_search(std::vector<double> const& x_lp, int trials, std::vector<double>& x_out, bool* feasible);

  private:
  struct Change { std::int32_t col; double lb, ub; };
  struct Node
  {
    std::vector<Change> changes;   // bounds that differ from the root box
    double bound;                  // lower bound on the (minimisation-sense) objective
    int depth;
    // parent LP solution (shared by both children) and the primal weight its solve ended with:
    // the start point of this node's LP (mipcu_solve_lp_batch_warm)
    std::shared_ptr<std::vector<double> const> x, y;
    double weight = 0;
    bool exact = false;            // solve this node's LP to full accuracy (no early MIPCU_BRANCH stop)
    // The parent's box was propagated to a fixpoint before it was split, so only the rows of the branched
    // column can tighten anything in the child: propagate() starts from them (-1: from every row).  Valid
    // while the row set is the one the parent saw (rows_epoch).
    std::int32_t branched = -1;
    std::int64_t epoch = -1;
  };

  bool propagate(std::vector<double>& lb, std::vector<double>& ub, std::int32_t seed_col = -1) const;
  std::size_t add_implied_bound_rows(std::vector<double> const& lb, std::vector<double> const& ub);
  void rebuild_columns();
  bool row_feasible(std::vector<double> const& x, double tol_scale) const;
  double objective(std::vector<double> const& x) const;
  bool try_incumbent(std::vector<double> const& x);   // false: x violates a row or a lazy handler
  bool absorb_lazy_rows();         // rows posted by lazy handlers -> this tree's row set; true if any
  void try_roundings(std::vector<double> const& x, std::vector<double> const& lb,
                     std::vector<double> const& ub);
  bool one_opt(std::vector<double>& x) const;   // true: an integer column was moved to a cheaper value
  bool repair(std::vector<double>& x) const;    // greedy moves until every row holds; false: stuck
  void try_threshold_roundings(std::vector<double> const& x);   // x_j >= theta -> up, then repair + one_opt
  double cutoff() const;           // nodes whose bound reaches this cannot improve the incumbent

  CudaSolver& S;
  std::size_t n = 0, m = 0;
  double sgn = 1;                  // -1 when the model maximises: the tree minimises sgn * c'x
  std::vector<std::int64_t> rptr, cptr;
  std::vector<std::int32_t> ridx, cidx;
  std::vector<double> rval, cval, rlo, rhi, cost;
  std::vector<char> is_int;
  bool pure_integer = true, objective_integral = true;
  double best = 0;
  bool have_incumbent = false;
  bool lazy_rejected = false, lazy_added = false;   // outcome of the last try_incumbent
  std::vector<double> best_x;
  std::vector<double> rounding_scratch;   // try_roundings' candidate (kept between calls)
  std::vector<double> root_lb, root_ub;   // the root box after propagation (what an incumbent must respect)
  std::vector<std::int32_t> opt_order;    // integer columns with a cost, most expensive first (one_opt)
  std::int64_t rows_epoch = 0;            // bumped whenever rows join the tree (lazy handlers, implied bounds)
  std::int64_t polish_candidates = 0;     // feasible roundings seen (one_opt runs on the first 32, then every 8th)
  std::int64_t threshold_calls = 0;       // nodes whose LP point went through try_threshold_roundings
  double cost_norm = 0;            // ||c||_2: turns the engine's relative dual residual back into an absolute one
};

}  // namespace miplib
