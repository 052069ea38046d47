// miplib/cuda/bnb.cpp - see bnb.hpp.
#include "bnb.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <deque>
#include <iostream>
#include <limits>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>

#include <mipcu.h>

namespace miplib {

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

void check(mipcu_ctx* ctx, int rc)
{
  if (rc) throw std::logic_error(std::string("Cuda backend: ") + mipcu_last_error(ctx));
}

bool integral(double v) { return std::isfinite(v) && std::floor(v) == v; }

}  // namespace

CudaBranchAndBound::CudaBranchAndBound(CudaSolver& solver) : S(solver)
{
  n = S.m_columns.size();
  sgn = S.m_sense == Solver::Sense::Maximize ? -1.0 : 1.0;
  double const inf = S.infinity();
  // alive rows, CSR
  rptr.push_back(0);
  for (std::size_t r = 0; r < S.m_row_alive.size(); ++r)
  {
    if (!S.m_row_alive[r]) continue;
    for (std::int64_t k = S.m_row_start[r]; k < S.m_row_start[r + 1]; ++k)
    {
      ridx.push_back(S.m_row_idx[k]);
      rval.push_back(S.m_row_val[k]);
    }
    rptr.push_back(static_cast<std::int64_t>(ridx.size()));
    rlo.push_back(S.m_row_lo[r] <= -inf ? -kInf : S.m_row_lo[r]);
    rhi.push_back(S.m_row_hi[r] >= inf ? kInf : S.m_row_hi[r]);
  }
  m = rlo.size();
  rebuild_columns();
  cost.resize(n);
  is_int.resize(n);
  for (std::size_t j = 0; j < n; ++j)
  {
    cost[j] = sgn * (j < S.m_objective.size() ? S.m_objective[j] : 0.0);
    is_int[j] = S.m_columns[j].type != Var::Type::Continuous;
    if (!is_int[j]) pure_integer = false;
    if (cost[j] != 0 && (!is_int[j] || !integral(cost[j]))) objective_integral = false;
    cost_norm += cost[j] * cost[j];
  }
  cost_norm = std::sqrt(cost_norm);
}

// CSC of the rows (which rows to revisit when a column's box shrinks)
void CudaBranchAndBound::rebuild_columns()
{
  m = rlo.size();
  cptr.assign(n + 1, 0);
  for (std::int32_t j : ridx) cptr[j + 1]++;
  for (std::size_t j = 0; j < n; ++j) cptr[j + 1] += cptr[j];
  cidx.resize(ridx.size());
  cval.resize(ridx.size());
  std::vector<std::int64_t> fill(cptr.begin(), cptr.end() - 1);
  for (std::size_t r = 0; r < m; ++r)
    for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k)
    {
      cidx[fill[ridx[k]]] = static_cast<std::int32_t>(r);
      cval[fill[ridx[k]]++] = rval[k];
    }
}

// Implied-bound rows from the root box (see bnb.hpp).  Each side of a row is normalised to
// sum alpha_k x_k <= b.  Fixing a binary z at the value that RAISES the row's minimum activity
// (0 if alpha_z < 0, 1 if alpha_z > 0) by delta = |alpha_z| leaves column j the room
//   alpha_j x_j <= b - (minact + delta) + min(alpha_j l_j, alpha_j u_j);
// if that is tighter than j's own bound, "bound(z fixed) ... bound(z free)" interpolated in z is
// a valid row with two nonzeros.  Only pairs with |alpha_z| >= |alpha_j| are generated (the other
// orientation yields the same row), and a z is skipped in O(1) unless delta + (largest range of
// another term) exceeds the row's slack b - minact, which keeps knapsack rows linear.
std::size_t CudaBranchAndBound::add_implied_bound_rows(std::vector<double> const& lb, std::vector<double> const& ub)
{
  std::size_t const m0 = m, cap = 4 * n + 1024;
  std::size_t added = 0;
  auto is_binary = [&](std::int32_t j) { return is_int[j] && lb[j] == 0.0 && ub[j] == 1.0; };
  auto emit = [&](std::int32_t j, double aj, std::int32_t z, double az, double lo, double hi) {
    if (j < z) { ridx.push_back(j); rval.push_back(aj); ridx.push_back(z); rval.push_back(az); }
    else { ridx.push_back(z); rval.push_back(az); ridx.push_back(j); rval.push_back(aj); }
    rptr.push_back(static_cast<std::int64_t>(ridx.size()));
    rlo.push_back(lo);
    rhi.push_back(hi);
    ++added;
  };
  for (std::size_t r = 0; r < m0 && added < cap; ++r)
  {
    std::int64_t const k0 = rptr[r], k1 = rptr[r + 1];
    if (k1 - k0 < 2) continue;
    for (int side = 0; side < 2 && added < cap; ++side)
    {
      double const s = side == 0 ? 1.0 : -1.0;
      double const b = side == 0 ? rhi[r] : -rlo[r];
      if (std::isinf(b)) continue;
      double minact = 0, range1 = 0, range2 = 0;   // two largest term ranges
      bool bounded = true;
      for (std::int64_t k = k0; k < k1 && bounded; ++k)
      {
        double const a = s * rval[k], l = lb[ridx[k]], u = ub[ridx[k]];
        if (std::isinf(l) || std::isinf(u)) { bounded = false; break; }
        minact += std::min(a * l, a * u);
        double const range = std::abs(a) * (u - l);
        if (range > range1) { range2 = range1; range1 = range; } else if (range > range2) range2 = range;
      }
      if (!bounded) continue;
      double const slack = b - minact;
      if (slack < 0) continue;
      for (std::int64_t kz = k0; kz < k1 && added < cap; ++kz)
      {
        std::int32_t const z = ridx[kz];
        double const az = s * rval[kz];
        if (az == 0 || !is_binary(z)) continue;
        double const delta = std::abs(az);
        double const other = (delta == range1) ? range2 : range1;   // binary: range == |alpha|
        if (delta + other <= slack * (1 + 1e-12) + 1e-9) continue;
        bool const z_at_zero = az < 0;
        for (std::int64_t k = k0; k < k1 && added < cap; ++k)
        {
          if (k == kz) continue;
          std::int32_t const j = ridx[k];
          double const aj = s * rval[k], l = lb[j], u = ub[j];
          if (aj == 0 || l == u) continue;
          if (std::abs(aj) > delta || (std::abs(aj) == delta && j < z)) continue;
          double const room = slack - delta;   // what |alpha_j| (x_j - its min-activity side) may still use
          if (room >= std::abs(aj) * (u - l) - 1e-9) continue;          // no tightening
          if (room < -1e-9) continue;                                   // z cannot take that value: propagation's job
          if (aj > 0)
          {
            double nu = l + std::max(room, 0.0) / aj;
            if (is_int[j]) nu = std::floor(nu + 1e-6);
            if (nu > u - 1e-9) continue;
            // z = 0 forces x_j <= nu:  x_j - (u - nu) z <= nu ;  z = 1 forces it:  x_j + (u - nu) z <= u
            if (z_at_zero) emit(j, 1.0, z, -(u - nu), -kInf, nu);
            else emit(j, 1.0, z, (u - nu), -kInf, u);
          }
          else
          {
            double nl = u - std::max(room, 0.0) / -aj;
            if (is_int[j]) nl = std::ceil(nl - 1e-6);
            if (nl < l + 1e-9) continue;
            // z = 0 forces x_j >= nl:  x_j + (nl - l) z >= nl ;  z = 1 forces it:  x_j - (nl - l) z >= l
            if (z_at_zero) emit(j, 1.0, z, (nl - l), nl, kInf);
            else emit(j, 1.0, z, -(nl - l), l, kInf);
          }
        }
      }
    }
  }
  if (added) { rebuild_columns(); ++rows_epoch; }
  return added;
}

double CudaBranchAndBound::objective(std::vector<double> const& x) const
{
  double v = 0;
  for (std::size_t j = 0; j < n; ++j) v += cost[j] * x[j];
  return v;
}

double CudaBranchAndBound::cutoff() const
{
  if (!have_incumbent) return kInf;
  // an improving solution of an all-integer objective is better by at least 1
  return objective_integral ? best - 1.0 + 1e-6 : best - 1e-9 * (1.0 + std::abs(best));
}

// Activity-based bound tightening to a fixpoint (bounded work).  Returns false when some row can
// no longer be satisfied inside the box.
bool CudaBranchAndBound::propagate(std::vector<double>& lb, std::vector<double>& ub, std::int32_t seed_col) const
{
  std::deque<std::int32_t> work;
  std::vector<char> queued(m, seed_col < 0 ? 1 : 0);
  if (seed_col < 0)
    for (std::size_t r = 0; r < m; ++r) work.push_back(static_cast<std::int32_t>(r));
  else      // the box was at a fixpoint before seed_col's bounds changed: only its rows can tighten anything
    for (std::int64_t k = cptr[seed_col]; k < cptr[seed_col + 1]; ++k)
      if (!queued[cidx[k]]) { queued[cidx[k]] = 1; work.push_back(cidx[k]); }
  std::int64_t budget = 30 * static_cast<std::int64_t>(m) + 2000;
  auto touch = [&](std::int32_t col) {
    for (std::int64_t k = cptr[col]; k < cptr[col + 1]; ++k)
      if (!queued[cidx[k]]) { queued[cidx[k]] = 1; work.push_back(cidx[k]); }
  };
  while (!work.empty() && budget-- > 0)
  {
    std::int32_t const r = work.front();
    work.pop_front();
    queued[r] = 0;
    double amin = 0, amax = 0;
    int min_inf = 0, max_inf = 0;
    for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k)
    {
      double const a = rval[k], l = lb[ridx[k]], u = ub[ridx[k]];
      double const lo_side = a > 0 ? l : u, hi_side = a > 0 ? u : l;
      if (std::isinf(lo_side)) ++min_inf; else amin += a * lo_side;
      if (std::isinf(hi_side)) ++max_inf; else amax += a * hi_side;
    }
    double const tol_hi = 1e-7 * (1.0 + std::abs(std::isinf(rhi[r]) ? 0.0 : rhi[r]));
    double const tol_lo = 1e-7 * (1.0 + std::abs(std::isinf(rlo[r]) ? 0.0 : rlo[r]));
    if (min_inf == 0 && amin > rhi[r] + tol_hi) return false;
    if (max_inf == 0 && amax < rlo[r] - tol_lo) return false;
    for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k)
    {
      std::int32_t const j = ridx[k];
      double const a = rval[k], l = lb[j], u = ub[j];
      double new_l = l, new_u = u;
      // a x_j <= hi - (min activity of the other terms)
      if (!std::isinf(rhi[r]))
      {
        double const mine = a > 0 ? l : u;
        int const others_inf = min_inf - (std::isinf(mine) ? 1 : 0);
        if (others_inf == 0)
        {
          double const rest = amin - (std::isinf(mine) ? 0.0 : a * mine);
          double const bnd = (rhi[r] - rest) / a;
          if (a > 0) new_u = std::min(new_u, bnd); else new_l = std::max(new_l, bnd);
        }
      }
      // a x_j >= lo - (max activity of the other terms)
      if (!std::isinf(rlo[r]))
      {
        double const mine = a > 0 ? u : l;
        int const others_inf = max_inf - (std::isinf(mine) ? 1 : 0);
        if (others_inf == 0)
        {
          double const rest = amax - (std::isinf(mine) ? 0.0 : a * mine);
          double const bnd = (rlo[r] - rest) / a;
          if (a > 0) new_l = std::max(new_l, bnd); else new_u = std::min(new_u, bnd);
        }
      }
      if (is_int[j])
      {
        new_u = std::floor(new_u + 1e-6);
        new_l = std::ceil(new_l - 1e-6);
      }
      double const span = std::max(1.0, std::max(std::abs(l), std::abs(u)));
      bool changed = false;
      // continuous boxes only move on a real improvement, so the loop cannot creep forever
      double const min_step = is_int[j] ? 0.5 : 1e-4 * (std::isinf(u - l) ? span : std::max(u - l, 1e-9));
      if (new_u < u - min_step || (std::isinf(u) && !std::isinf(new_u))) { ub[j] = new_u; changed = true; }
      if (new_l > l + min_step || (std::isinf(l) && !std::isinf(new_l))) { lb[j] = new_l; changed = true; }
      if (lb[j] > ub[j])
      {
        if (lb[j] - ub[j] > 1e-7 * span) return false;
        lb[j] = ub[j] = 0.5 * (lb[j] + ub[j]);   // met within rounding: the column is fixed
        if (is_int[j]) lb[j] = ub[j] = std::round(lb[j]);
      }
      if (changed) touch(j);
    }
  }
  return true;
}

bool CudaBranchAndBound::row_feasible(std::vector<double> const& x, double tol_scale) const
{
  for (std::size_t r = 0; r < m; ++r)
  {
    double act = 0;
    for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k) act += rval[k] * x[ridx[k]];
    if (act > rhi[r] + tol_scale * (1.0 + std::abs(std::isinf(rhi[r]) ? 0.0 : rhi[r]))) return false;
    if (act < rlo[r] - tol_scale * (1.0 + std::abs(std::isinf(rlo[r]) ? 0.0 : rlo[r]))) return false;
  }
  return true;
}

bool CudaBranchAndBound::try_incumbent(std::vector<double> const& x)
{
  double const tol = pure_integer ? 1e-9 : std::max(10 * S.m_feas_tol, 1e-6);
  lazy_rejected = lazy_added = false;
  if (!row_feasible(x, tol)) return false;
  double const v = objective(x);
  if (!have_incumbent || v < best - 1e-12 * (1 + std::abs(best)))
  {
    // lazy constraints (reference miplib/lazy.hpp:19-41): the handlers see every point that is
    // about to become the incumbent; a rejected point is dropped and the rows they posted join
    // the tree before the next round of node LPs
    if (!S.lazy_accepts(x, &lazy_added))
    {
      lazy_rejected = true;
      return false;
    }
    have_incumbent = true;
    best = v;
    best_x = x;
  }
  return true;
}

bool CudaBranchAndBound::absorb_lazy_rows()
{
  if (S.m_lazy_new_rows.empty()) return false;
  double const inf = S.infinity();
  for (std::int64_t r : S.m_lazy_new_rows)
  {
    for (std::int64_t k = S.m_row_start[r]; k < S.m_row_start[r + 1]; ++k)
    {
      ridx.push_back(S.m_row_idx[k]);
      rval.push_back(S.m_row_val[k]);
    }
    rptr.push_back(static_cast<std::int64_t>(ridx.size()));
    rlo.push_back(S.m_row_lo[r] <= -inf ? -kInf : S.m_row_lo[r]);
    rhi.push_back(S.m_row_hi[r] >= inf ? kInf : S.m_row_hi[r]);
  }
  S.m_info.lazy_rows += static_cast<std::int64_t>(S.m_lazy_new_rows.size());
  S.m_lazy_new_rows.clear();
  rebuild_columns();
  ++rows_epoch;
  return true;
}

// 1-opt on a feasible point (SCIP's "oneopt" idea): every integer column with a cost is moved as far
// towards its cheaper side as the root box and the slack of its rows allow, the most expensive columns
// first, row activities kept up to date; repeated while something moves.  A rounded-up LP point of a
// covering model carries many redundant columns (config 4: 87 585 after rounding, 40 212 is the LP bound);
// this is what removes them.  x must be row-feasible on entry and stays so.
bool CudaBranchAndBound::one_opt(std::vector<double>& x) const
{
  if (opt_order.empty() || cptr.size() != n + 1) return false;
  std::vector<double> act(m, 0.0);
  for (std::size_t r = 0; r < m; ++r)
  {
    double a = 0;
    for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k) a += rval[k] * x[ridx[k]];
    act[r] = a;
  }
  bool moved_any = false;
  for (int pass = 0; pass < 4; ++pass)
  {
    bool moved = false;
    for (std::int32_t j : opt_order)
    {
      double const dir = cost[j] > 0 ? -1.0 : 1.0;                 // the direction that lowers the objective
      double room = dir < 0 ? x[j] - root_lb[j] : root_ub[j] - x[j];
      if (!(room >= 1.0 - 1e-9)) continue;
      for (std::int64_t k = cptr[j]; k < cptr[j + 1] && room >= 1.0 - 1e-9; ++k)
      {
        std::int32_t const r = cidx[k];
        double const d = cval[k] * dir;                            // change of the row's activity per unit step
        if (d > 0 && !std::isinf(rhi[r])) room = std::min(room, (rhi[r] - act[r]) / d);
        else if (d < 0 && !std::isinf(rlo[r])) room = std::min(room, (act[r] - rlo[r]) / -d);
      }
      double const step = std::floor(room + 1e-9);
      if (!(step >= 1.0) || std::isinf(step)) continue;
      x[j] += dir * step;
      for (std::int64_t k = cptr[j]; k < cptr[j + 1]; ++k) act[cidx[k]] += cval[k] * dir * step;
      moved = true;
    }
    if (!moved) break;
    moved_any = true;
  }
  return moved_any;
}

// Greedy repair of an integral point inside the root box: while a row is violated, move the integer column of that
// row that buys the most violation per unit of cost, provided the move breaks no row that holds (rows are
// re-checked through the column copy).  Covering and packing rows with non-negative coefficients - configs 4
// and 5 - always find such a move; anything else may get stuck, which is reported.
bool CudaBranchAndBound::repair(std::vector<double>& x) const
{
  if (cptr.size() != n + 1) return false;
  std::vector<double> act(m, 0.0);
  for (std::size_t r = 0; r < m; ++r)
  {
    double a = 0;
    for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k) a += rval[k] * x[ridx[k]];
    act[r] = a;
  }
  auto violated = [&](std::size_t r) {
    if (act[r] < rlo[r] - 1e-9 * (1.0 + std::abs(rlo[r]))) return -1;       // too small
    if (act[r] > rhi[r] + 1e-9 * (1.0 + std::abs(rhi[r]))) return 1;        // too large
    return 0;
  };
  for (int pass = 0; pass < 3; ++pass)
  {
    bool any = false;
    for (std::size_t r = 0; r < m; ++r)
    {
      int guard = 0;
      for (int v = violated(r); v != 0 && guard < 64; v = violated(r), ++guard)
      {
        any = true;
        double const need = v < 0 ? rlo[r] - act[r] : act[r] - rhi[r];
        std::int32_t pick = -1;
        double pick_dir = 0, pick_score = kInf;
        for (std::int64_t k = rptr[r]; k < rptr[r + 1]; ++k)
        {
          std::int32_t const j = ridx[k];
          if (!is_int[j] || rval[k] == 0) continue;
          double const dir = (rval[k] > 0) == (v < 0) ? 1.0 : -1.0;         // the move that helps this row
          if (dir > 0 ? x[j] + 1.0 > root_ub[j] + 1e-9 : x[j] - 1.0 < root_lb[j] - 1e-9) continue;
          bool ok = true;                                                     // must not break a row that holds
          for (std::int64_t q = cptr[j]; q < cptr[j + 1] && ok; ++q)
          {
            std::int32_t const i = cidx[q];
            if (static_cast<std::size_t>(i) == r) continue;
            double const a2 = act[i] + cval[q] * dir;
            if (a2 > rhi[i] + 1e-9 * (1.0 + std::abs(rhi[i])) && cval[q] * dir > 0) ok = false;
            if (a2 < rlo[i] - 1e-9 * (1.0 + std::abs(rlo[i])) && cval[q] * dir < 0) ok = false;
          }
          if (!ok) continue;
          // cost of the move per unit of this row's violation it removes (other rows it helps count too)
          double gain = std::min(std::abs(rval[k]), need);
          for (std::int64_t q = cptr[j]; q < cptr[j + 1]; ++q)
          {
            std::int32_t const i = cidx[q];
            if (static_cast<std::size_t>(i) == r) continue;
            double const d = cval[q] * dir;
            if (d > 0 && act[i] < rlo[i]) gain += std::min(d, rlo[i] - act[i]);
            if (d < 0 && act[i] > rhi[i]) gain += std::min(-d, act[i] - rhi[i]);
          }
          double const score = (cost[j] * dir) / std::max(gain, 1e-12);
          if (score < pick_score) { pick_score = score; pick = j; pick_dir = dir; }
        }
        if (pick < 0) return false;
        x[pick] += pick_dir;
        for (std::int64_t q = cptr[pick]; q < cptr[pick + 1]; ++q) act[cidx[q]] += cval[q] * pick_dir;
      }
      if (violated(r) != 0) return false;
    }
    if (!any) return true;
  }
  for (std::size_t r = 0; r < m; ++r)
    if (violated(r) != 0) return false;
  return true;
}

// LP-guided start points for the two local searches: integer columns at or above a threshold go up, the rest
// down; the point is repaired greedily, polished by one_opt and offered as an incumbent.  Pure-integer models
// only, at the first nodes of the tree and then at every 64th (each call is a few passes over the matrix on the
// host).  On config 4 at full size this takes the incumbent from 42 920 (rounding up + one_opt) to within a few
// per cent of the LP bound.
void CudaBranchAndBound::try_threshold_roundings(std::vector<double> const& x)
{
  if (!pure_integer || cptr.size() != n + 1) return;
  ++threshold_calls;
  if (threshold_calls > 8 && threshold_calls % 64 != 0) return;
  std::vector<double> cand(n);
  for (double theta : {0.9, 0.6, 0.35, 0.15})
  {
    for (std::size_t j = 0; j < n; ++j)
    {
      double const f = x[j] - std::floor(x[j] + 1e-9);
      double v = f >= theta ? std::ceil(x[j] - 1e-9) : std::floor(x[j] + 1e-9);
      cand[j] = std::min(std::max(v, root_lb[j]), root_ub[j]);
    }
    if (!repair(cand)) continue;
    one_opt(cand);
    try_incumbent(cand);
  }
}

// nearest / up / down rounding of the integer columns of an LP point, clipped to the node box; a rounding
// that is feasible and not far above the incumbent is also handed to one_opt
void CudaBranchAndBound::try_roundings(std::vector<double> const& x, std::vector<double> const& lb,
                                       std::vector<double> const& ub)
{
  auto rounded = [&](std::size_t j, int mode) {
    double v = x[j];
    if (is_int[j])
    {
      v = mode == 0 ? std::round(v) : mode == 1 ? std::ceil(v - 1e-6) : std::floor(v + 1e-6);
      v = std::min(std::max(v, std::ceil(lb[j] - 1e-9)), std::floor(ub[j] + 1e-9));
    }
    return v;
  };
  // the three objectives in one pass over the columns, before anything proportional to the matrix: a point
  // more than 25 % above the incumbent is neither an incumbent nor worth polishing
  double value[3] = {0, 0, 0};
  for (std::size_t j = 0; j < n; ++j)
  {
    if (cost[j] == 0) continue;
    for (int mode = 0; mode < 3; ++mode) value[mode] += cost[j] * rounded(j, mode);
  }
  std::vector<double>& cand = rounding_scratch;
  for (int mode = 0; mode < 3; ++mode)
  {
    bool const had = have_incumbent;
    double const before = best;
    if (had && value[mode] > best + 0.25 * std::abs(best)) continue;
    cand.resize(n);
    for (std::size_t j = 0; j < n; ++j) cand[j] = rounded(j, mode);
    if (!try_incumbent(cand)) continue;                     // infeasible, or rejected by a lazy handler
    // worth polishing: the first feasible point, a new incumbent, and - every few nodes, it is a pass or
    // four over the matrix - a point within 25 % of the incumbent
    bool const improved = !had || best < before;
    bool const turn = ++polish_candidates <= 32 || polish_candidates % 8 == 0;
    if (pure_integer && (improved || turn) && one_opt(cand)) try_incumbent(cand);
  }
}

int CudaBranchAndBound::selftest_local_search(std::vector<double> const& x_lp, int trials, std::vector<double>& x_out,
                                              bool* feasible)
{
  double const inf = S.infinity();
  root_lb.assign(n, 0.0);
  root_ub.assign(n, 0.0);
  for (std::size_t j = 0; j < n; ++j)
  {
    root_lb[j] = S.m_columns[j].lb <= -inf ? -kInf : S.m_columns[j].lb;
    root_ub[j] = S.m_columns[j].ub >= inf ? kInf : S.m_columns[j].ub;
    if (is_int[j]) { root_lb[j] = std::ceil(root_lb[j] - 1e-9); root_ub[j] = std::floor(root_ub[j] + 1e-9); }
  }
  if (feasible) *feasible = false;
  if (!propagate(root_lb, root_ub)) return -1;
  opt_order.clear();
  for (std::size_t j = 0; j < n; ++j)
    if (is_int[j] && cost[j] != 0) opt_order.push_back(static_cast<std::int32_t>(j));
  std::stable_sort(opt_order.begin(), opt_order.end(),
                   [&](std::int32_t a, std::int32_t b) { return std::abs(cost[a]) > std::abs(cost[b]); });
  // incremental against full propagation of single-column branchings (a deterministic walk over the columns)
  int mismatches = 0;
  std::uint64_t state = 88172645463325252ull;
  for (int t = 0; t < trials && n > 0; ++t)
  {
    state ^= state << 13; state ^= state >> 7; state ^= state << 17;
    std::size_t const j = static_cast<std::size_t>(state % n);
    if (!is_int[j] || root_lb[j] >= root_ub[j]) continue;
    for (int side = 0; side < 2; ++side)
    {
      std::vector<double> la(root_lb), ua(root_ub), lb2(root_lb), ub2(root_ub);
      if (side == 0) { ua[j] = root_lb[j]; ub2[j] = root_lb[j]; }
      else { la[j] = root_lb[j] + 1; lb2[j] = root_lb[j] + 1; }
      bool const full = propagate(la, ua), seeded = propagate(lb2, ub2, static_cast<std::int32_t>(j));
      if (full != seeded || (full && (la != lb2 || ua != ub2))) ++mismatches;
    }
  }
  x_out.assign(n, 0.0);
  if (x_lp.size() == n && pure_integer)
  {
    for (std::size_t j = 0; j < n; ++j)
    {
      double const f = x_lp[j] - std::floor(x_lp[j] + 1e-9);
      double const v = f >= 0.5 ? std::ceil(x_lp[j] - 1e-9) : std::floor(x_lp[j] + 1e-9);
      x_out[j] = std::min(std::max(v, root_lb[j]), root_ub[j]);
    }
    if (repair(x_out))
    {
      one_opt(x_out);
      if (feasible) *feasible = row_feasible(x_out, 1e-9);
    }
  }
  return mismatches;
}

std::pair<Solver::Result, bool> CudaBranchAndBound::run()
{
  using R = Solver::Result;
  auto const t0 = std::chrono::steady_clock::now();
  auto elapsed = [&]() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };
  double const inf = S.infinity();
  root_lb.assign(n, 0.0);
  root_ub.assign(n, 0.0);
  for (std::size_t j = 0; j < n; ++j)
  {
    root_lb[j] = S.m_columns[j].lb <= -inf ? -kInf : S.m_columns[j].lb;
    root_ub[j] = S.m_columns[j].ub >= inf ? kInf : S.m_columns[j].ub;
    if (is_int[j])
    {
      root_lb[j] = std::ceil(root_lb[j] - 1e-9);
      root_ub[j] = std::floor(root_ub[j] + 1e-9);
    }
  }
  if (!propagate(root_lb, root_ub)) return {R::Infeasible, false};
  opt_order.clear();
  for (std::size_t j = 0; j < n; ++j)
    if (is_int[j] && cost[j] != 0) opt_order.push_back(static_cast<std::int32_t>(j));
  std::stable_sort(opt_order.begin(), opt_order.end(),
                   [&](std::int32_t a, std::int32_t b) { return std::abs(cost[a]) > std::abs(cost[b]); });

  auto finish = [&](bool complete) -> std::pair<R, bool> {
    S.m_info.proven_optimal = complete && have_incumbent;
    if (have_incumbent)
    {
      S.m_solution = best_x;
      S.m_has_solution = true;
      double v = 0;
      for (std::size_t j = 0; j < n; ++j) v += (j < S.m_objective.size() ? S.m_objective[j] : 0.0) * best_x[j];
      S.m_objective_value = v;
    }
    if (complete) return have_incumbent ? std::make_pair(R::Optimal, true) : std::make_pair(R::Infeasible, false);
    return {R::Interrupted, have_incumbent};
  };

  // a user start point that happens to be feasible is a free incumbent
  if (S.m_start_x.size() == n)
  {
    std::vector<double> cand(S.m_start_x);
    bool ok = true;
    for (std::size_t j = 0; j < n && ok; ++j)
    {
      if (is_int[j]) cand[j] = std::round(cand[j]);
      ok = cand[j] >= root_lb[j] - 1e-9 && cand[j] <= root_ub[j] + 1e-9;
    }
    if (ok) try_incumbent(cand);
  }

  auto all_fixed = [&](std::vector<double> const& lb, std::vector<double> const& ub) {
    for (std::size_t j = 0; j < n; ++j)
      if (lb[j] != ub[j]) return false;
    return true;
  };
  if (all_fixed(root_lb, root_ub))
  {
    try_incumbent(root_lb);
    return finish(true);
  }

  // device model(s): the primary context plus one per extra GPU asked for with set_nr_threads
  bool const user_verbose = S.m_verbose;
  bool custom_rows = absorb_lazy_rows();      // a start point may already have met the handlers
  S.m_info.cuts = static_cast<std::int64_t>(add_implied_bound_rows(root_lb, root_ub));
  if (S.m_info.cuts) custom_rows = true;
  if (custom_rows && !propagate(root_lb, root_ub)) return finish(true);
  if (user_verbose && S.m_info.cuts)
    std::fprintf(stderr, "[miplib cuda b&b] %lld implied-bound rows added at the root\n",
                 static_cast<long long>(S.m_info.cuts));
  std::vector<mipcu_ctx*> gpus;
  auto upload = [&]() {
    if (custom_rows)
    {
      std::vector<double> lo(rlo), hi(rhi);
      for (double& v : lo) if (std::isinf(v)) v = -inf;
      for (double& v : hi) if (std::isinf(v)) v = inf;
      S.upload_rows(rptr, ridx, rval, lo, hi);
    }
    else
      S.upload_model();
    gpus = S.node_contexts();
    for (mipcu_ctx* ctx : gpus)
    {
      check(ctx, mipcu_set_param(ctx, MIPCU_PARAM_EPS_REL, S.m_feas_tol));
      check(ctx, mipcu_set_param(ctx, MIPCU_PARAM_VERBOSE, 0));
      check(ctx, mipcu_set_param(ctx, MIPCU_PARAM_ITER_LIMIT, 100000));
      check(ctx, mipcu_set_param(ctx, MIPCU_PARAM_OBJ_CUTOFF, std::nan("")));
    }
  };
  upload();

  // no feasible point of the box can cost more than this: a dual bound above it proves the
  // node LP infeasible (PDHG's dual iterate diverges on infeasible problems)
  auto box_objective_max = [&](std::vector<double> const& lb, std::vector<double> const& ub) {
    double v = 0;
    for (std::size_t j = 0; j < n; ++j)
    {
      if (cost[j] == 0) continue;
      double const side = cost[j] > 0 ? ub[j] : lb[j];
      if (std::isinf(side)) return kInf;
      v += cost[j] * side;
    }
    return v;
  };

  auto by_bound = [](Node const& a, Node const& b) {
    return a.bound != b.bound ? a.bound > b.bound : a.depth < b.depth;
  };
  std::priority_queue<Node, std::vector<Node>, decltype(by_bound)> open(by_bound);
  open.push(Node{{}, -kInf, 0, nullptr, nullptr, 0.0, false});
  bool complete = true;
  double lost_bound = kInf;      // smallest bound among the nodes given up undecided (they leave no open node)
  bool unbounded_relaxation = false;
  std::int64_t const node_limit = 4000000;
  std::size_t const per_gpu = 128;                       // node LPs per GPU per round
  std::size_t const round_cap = per_gpu * gpus.size();

  // finite_box: every column of the node's box has two finite bounds.  Only then is the engine's dual
  // objective a Lagrangian bound valid for ANY multiplier; with a free or one-sided column the reduced
  // costs on the infinite side are counted into the dual residual instead (solve.cu: Q_DRES2), so the
  // dual objective may overshoot the node optimum by about that residual times |x|.
  struct Pending { Node node; std::vector<double> lb, ub; double cut; bool finite_box; };
  // Everything one round owns.  Two of them: while the device solves the node LPs of one round, the host
  // decides the previous one (incumbents, branching) and assembles the next (see the loop below).
  struct RoundData
  {
    std::vector<Pending> round;
    std::vector<double> box_lb, box_ub, cuts, xs, ys, x0s, y0s, w0s, btol;
    std::vector<mipcu_stats> stats;
    std::string error;
  };
  RoundData bufs[2];
  // accuracy at which a node that cannot be pruned by bound is handed back for branching
  double const branch_tol = std::max(1e-4, S.m_feas_tol);
  // parents' solutions ride on the open nodes only while that stays cheap on the host
  std::size_t const warm_budget_bytes = std::size_t(2) << 30;
  // MIPLIB_CUDA_TIMING: where the wall time of the tree goes (printed once at the end)
  bool const timing = std::getenv("MIPLIB_CUDA_TIMING") != nullptr;
  double t_assemble = 0, t_pack = 0, t_gpu = 0, t_post = 0;
  std::int64_t rounds = 0;
  auto lap_from = [&](double since) { return elapsed() - since; };
  // ---- a round: the best open nodes, tightened on the host, packed for the device; false: nothing to solve ----
  auto assemble = [&](RoundData& R) -> bool {
    auto& round = R.round; auto& box_lb = R.box_lb; auto& box_ub = R.box_ub; auto& cuts = R.cuts; auto& xs = R.xs;
    auto& ys = R.ys; auto& x0s = R.x0s; auto& y0s = R.y0s; auto& w0s = R.w0s; auto& btol = R.btol; auto& stats = R.stats;
    R.error.clear();
    // ---- a round: the best open nodes, tightened on the host, the survivors solved together ----
    double t_mark = elapsed();
    round.clear();
    while (!open.empty() && round.size() < round_cap)
    {
      if (!S.m_lazy_new_rows.empty()) break;   // a handler just posted rows: take them in first
      Node node = open.top();
      open.pop();
      if (node.bound >= cutoff()) continue;            // best-first: cannot improve the incumbent
      S.m_info.nodes += 1;
      Pending p{std::move(node), root_lb, root_ub, kInf, true};
      for (Change const& c : p.node.changes) { p.lb[c.col] = c.lb; p.ub[c.col] = c.ub; }
      if (!propagate(p.lb, p.ub, p.node.epoch == rows_epoch ? p.node.branched : -1)) continue;
      if (all_fixed(p.lb, p.ub))
      {
        try_incumbent(p.lb);
        continue;
      }
      round.push_back(std::move(p));
    }
    t_assemble += lap_from(t_mark);
    t_mark = elapsed();
    if (round.empty()) return false;
    ++rounds;
    std::size_t const B = round.size();
    box_lb.resize(B * n);
    box_ub.resize(B * n);
    cuts.resize(B);
    stats.assign(B, mipcu_stats{});
    xs.resize(B * n);
    ys.resize(B * m);
    x0s.assign(B * n, 0.0);
    y0s.assign(B * m, 0.0);
    w0s.assign(B, 0.0);
    btol.assign(B, 0.0);
    for (std::size_t b = 0; b < B; ++b)
    {
      Pending& p = round[b];
      for (std::size_t j = 0; j < n && p.finite_box; ++j)
        p.finite_box = !std::isinf(p.lb[j]) && !std::isinf(p.ub[j]);
      if (p.finite_box)
      {
        // a dual bound above the largest objective value the box allows proves the node LP infeasible
        double const umax = box_objective_max(p.lb, p.ub);
        double const cut = std::min(cutoff(), std::isinf(umax) ? kInf : umax + 1e-6 * (1.0 + std::abs(umax)));
        cuts[b] = std::isinf(cut) ? std::nan("") : sgn * cut;
        if (!p.node.exact && S.m_lazy_handlers.empty()) btol[b] = branch_tol;
      }
      else
        cuts[b] = std::nan("");      // no early stop on a bound that may overshoot: decided after the solve
      for (std::size_t j = 0; j < n; ++j)
      {
        box_lb[b * n + j] = std::isinf(p.lb[j]) ? -inf : p.lb[j];
        box_ub[b * n + j] = std::isinf(p.ub[j]) ? inf : p.ub[j];
      }
      // start from the parent's LP solution (rows may have joined the model since: then only x)
      if (p.node.x && p.node.x->size() == n)
      {
        std::copy(p.node.x->begin(), p.node.x->end(), x0s.begin() + b * n);
        if (p.node.y && p.node.y->size() == m) std::copy(p.node.y->begin(), p.node.y->end(), y0s.begin() + b * m);
        w0s[b] = p.node.weight;
      }
    }
    t_pack += lap_from(t_mark);
    return true;
  };

  // ---- deal the round to the GPUs: nodes are independent, no device-side exchange (runs on its own thread) ----
  auto run_gpu = [&](RoundData& R) {
    auto& box_lb = R.box_lb; auto& box_ub = R.box_ub; auto& cuts = R.cuts; auto& xs = R.xs; auto& ys = R.ys;
    auto& x0s = R.x0s; auto& y0s = R.y0s; auto& w0s = R.w0s; auto& btol = R.btol; auto& stats = R.stats;
    std::size_t const B = R.round.size();
    {
      std::size_t const G = std::min(gpus.size(), B);
      std::vector<std::thread> workers;
      std::vector<std::string> errors(G);
      for (std::size_t g = 0; g < G; ++g)
      {
        std::size_t const b0 = B * g / G, b1 = B * (g + 1) / G;
        auto job = [&, g, b0, b1]() {
          mipcu_ctx* ctx = gpus[g];
          if (S.m_time_limit > 0)
            mipcu_set_param(ctx, MIPCU_PARAM_TIME_LIMIT, std::max(0.01, S.m_time_limit - elapsed()));
          if (mipcu_solve_lp_batch_warm(ctx, static_cast<std::int32_t>(b1 - b0), box_lb.data() + b0 * n,
                                        box_ub.data() + b0 * n, cuts.data() + b0, x0s.data() + b0 * n,
                                        y0s.data() + b0 * m, w0s.data() + b0, btol.data() + b0,
                                        stats.data() + b0, xs.data() + b0 * n, ys.data() + b0 * m))
            errors[g] = mipcu_last_error(ctx);
        };
        if (G == 1) job(); else workers.emplace_back(job);
      }
      for (std::thread& t : workers) t.join();
      for (std::string const& e : errors)
        if (!e.empty()) R.error = e;
    }
  };

  // ---- decide the nodes of a solved round: incumbents, pruning, branching ----
  auto decide = [&](RoundData& R) {
    auto& round = R.round; auto& xs = R.xs; auto& ys = R.ys; auto& stats = R.stats;
    std::size_t const B = round.size();
    double const t_mark = elapsed();
    S.m_info.lp_solves += static_cast<std::int64_t>(B);
    S.m_info.kernel_launches += stats[0].kernel_launches;
    if (user_verbose)
      std::fprintf(stderr, "[miplib cuda b&b] round of %zu node LPs, %lld nodes so far, %zu open, incumbent %g\n",
                   B, static_cast<long long>(S.m_info.nodes), open.size(),
                   have_incumbent ? sgn * best : std::nan(""));

    std::vector<double> x;                 // one buffer for the whole round (a fresh 8 n-byte vector per node
    std::vector<Change> inherited;         // goes through mmap / munmap on big models)
    for (std::size_t b = 0; b < B; ++b)
    {
      Node const& node = round[b].node;
      std::vector<double> const& lb = round[b].lb;
      std::vector<double> const& ub = round[b].ub;
      mipcu_stats const& lp = stats[b];
      S.m_info.lp_iterations += lp.iterations;
      x.assign(xs.begin() + b * n, xs.begin() + (b + 1) * n);
      if (lp.status == MIPCU_CUTOFF) continue;       // bound (or infeasibility) proven by the dual
      if (lp.status == MIPCU_INFEASIBLE) continue;   // Farkas certificate from the engine
      if (lp.status == MIPCU_UNBOUNDED || lp.status == MIPCU_INFEASIBLE_OR_UNBOUNDED)
      {
        unbounded_relaxation = true;                 // an unbounded node LP: no finite optimum to prove
        complete = false;
        lost_bound = -kInf;
        continue;
      }
      if (lp.status == MIPCU_NUMERICAL_ERROR || lp.status == MIPCU_TIME_LIMIT)
      {
        complete = false;
        lost_bound = std::min(lost_bound, node.bound);
        continue;
      }
      double node_bound = node.bound;
      bool const early = lp.status == MIPCU_BRANCH;      // cannot be pruned by bound: branch on the 1e-4 point
      if (early)
      {
        // the dual objective of ANY multiplier bounds a finite box (only finite boxes stop early)
        double const d = sgn * lp.dual_obj;
        node_bound = std::max(node_bound, d - std::max(1e-6, S.m_feas_tol) * (1.0 + std::abs(d)));
        try_roundings(x, lb, ub);
        try_threshold_roundings(x);
        if (node_bound >= cutoff()) continue;
      }
      if (lp.status == MIPCU_OPTIMAL)
      {
        double const d = sgn * lp.dual_obj, p = sgn * lp.primal_obj;
        double const tol = std::max(1e-6, S.m_feas_tol);
        if (round[b].finite_box)
        {
          node_bound = std::max(node_bound, d - tol * (1.0 + std::abs(d)));
          if (objective_integral) node_bound = std::ceil(node_bound - 1e-9);
        }
        else
        {
          // the dual objective is exact only up to the reduced-cost violations on infinite sides:
          // (absolute dual residual) x (norm of the columns that have such a side), plus the gap tolerance
          double xfree = 0;
          for (std::size_t j = 0; j < n; ++j)
            if (std::isinf(lb[j]) || std::isinf(ub[j])) xfree += x[j] * x[j];
          double const slack = 2.0 * tol * (1.0 + std::abs(p) + std::abs(d)) +
                               2.0 * lp.rel_dual_res * (1.0 + cost_norm) * std::sqrt(xfree);
          node_bound = std::max(node_bound, d - slack);
          if (objective_integral) node_bound = std::ceil(node_bound - 1e-9);
        }
        try_roundings(x, lb, ub);
        try_threshold_roundings(x);
        if (node_bound >= cutoff()) continue;
      }
      // branching column: the most fractional integer column of the LP point
      std::int32_t pick = -1;
      double worst = S.m_int_tol;
      for (std::size_t j = 0; j < n; ++j)
        if (is_int[j] && lb[j] < ub[j])
        {
          double const f = std::abs(x[j] - std::round(x[j]));
          if (f > worst) { worst = f; pick = static_cast<std::int32_t>(j); }
        }
      if (early && worst < 0.02)
      {
        // a 1e-4 point that looks integral decides nothing: the same box again, to full accuracy
        Node again = node;
        again.exact = true;
        again.bound = node_bound;
        again.x = std::make_shared<std::vector<double> const>(x);
        again.y = std::make_shared<std::vector<double> const>(ys.begin() + b * m, ys.begin() + (b + 1) * m);
        again.weight = lp.primal_weight;
        S.m_info.nodes -= 1;                             // counted again when it is popped
        open.push(std::move(again));
        continue;
      }
      if (pick < 0)
      {
        if (lp.status == MIPCU_OPTIMAL)
        {
          // integral LP optimum: round the integer columns exactly; continuous ones keep their LP value
          std::vector<double> cand(x);
          for (std::size_t j = 0; j < n; ++j)
            if (is_int[j]) cand[j] = std::min(std::max(std::round(cand[j]), lb[j]), ub[j]);
          if (!try_incumbent(cand))
          {
            if (lazy_rejected && lazy_added)
            {
              // the handlers cut this point off: solve the same box again under the new rows
              Node again;
              again.x = std::make_shared<std::vector<double> const>(x);   // rows changed: the dual starts cold
              again.weight = lp.primal_weight;
              for (std::size_t j = 0; j < n; ++j)
                if (lb[j] != root_lb[j] || ub[j] != root_ub[j])
                  again.changes.push_back({static_cast<std::int32_t>(j), lb[j], ub[j]});
              again.bound = node_bound;
              again.depth = node.depth;
              open.push(std::move(again));
            }
            else
            {
              complete = false;                        // an integral LP point that rounding broke
              lost_bound = std::min(lost_bound, node_bound);
            }
          }
        }
        else
        {
          complete = false;                            // undecided node with nothing to branch on
          lost_bound = std::min(lost_bound, node_bound);
        }
        continue;
      }
      double const split = std::floor(x[pick]);
      std::shared_ptr<std::vector<double> const> px, py;
      if ((lp.status == MIPCU_OPTIMAL || early) && (open.size() + 2) * (n + m) * sizeof(double) <= warm_budget_bytes)
      {
        px = std::make_shared<std::vector<double> const>(x);
        py = std::make_shared<std::vector<double> const>(ys.begin() + b * m, ys.begin() + (b + 1) * m);
      }
      // carry every bound propagation found, so the children start from the tightened box
      inherited.clear();
      for (std::size_t j = 0; j < n; ++j)
        if (lb[j] != root_lb[j] || ub[j] != root_ub[j])
          inherited.push_back({static_cast<std::int32_t>(j), lb[j], ub[j]});
      for (int side = 0; side < 2; ++side)
      {
        Node child;
        child.x = px;
        child.y = py;
        child.weight = px ? lp.primal_weight : 0.0;
        child.changes = inherited;
        double cl = lb[pick], cu = ub[pick];
        if (side == 0) cu = split; else cl = split + 1;
        if (cl > cu) continue;
        bool found = false;
        for (Change& c : child.changes)
          if (c.col == pick) { c.lb = cl; c.ub = cu; found = true; }
        if (!found) child.changes.push_back({pick, cl, cu});
        child.bound = node_bound;
        child.depth = node.depth + 1;
        child.branched = pick;
        child.epoch = rows_epoch;
        open.push(std::move(child));
      }
    }
    t_post += lap_from(t_mark);
  };

  // ---- the loop.  With no lazy-constraint handler attached the rounds are PIPELINED: the host assembles
  // round k + 1 from the open list as it stands, hands it to the device as soon as round k comes back, and
  // decides round k (roundings, local searches, branching: half of the wall time of a big tree) while the
  // device works.  Round k + 1 therefore does not see the children of round k - they join the list one round
  // later, which a best-first tree with thousands of open nodes does not notice; while the tree is small the
  // open list is empty at that point and the loop simply waits (no overlap, same order as before).  Handlers
  // may post rows while a round is decided, and rows mean a new device model: then one round at a time.
  bool const pipelined = S.m_lazy_handlers.empty() && !(std::getenv("MIPLIB_CUDA_BNB_SYNC") != nullptr);
  auto limits_hit = [&]() {
    return (S.m_time_limit > 0 && elapsed() > S.m_time_limit) || S.m_info.nodes >= node_limit;
  };
  int cur = 0;
  std::thread device;
  RoundData* in_flight = nullptr;
  auto wait_device = [&]() -> RoundData* {
    if (!in_flight) return nullptr;
    double const t_mark = elapsed();
    device.join();
    t_gpu += lap_from(t_mark);
    RoundData* done = in_flight;
    in_flight = nullptr;
    if (!done->error.empty()) throw std::logic_error("Cuda backend: " + done->error);
    return done;
  };
  try
  {
    while (true)
    {
      if (limits_hit())
      {
        if (!open.empty()) complete = false;
        if (RoundData* done = wait_device()) decide(*done);
        if (!open.empty()) complete = false;
        break;
      }
      // rows posted by lazy handlers during the last round join the tree and the device model
      if (!in_flight && absorb_lazy_rows())
      {
        custom_rows = true;
        if (!propagate(root_lb, root_ub)) break;      // nothing left inside the root box
        upload();
      }
      RoundData& next = bufs[cur];
      bool const have_next = assemble(next);
      RoundData* done = wait_device();
      if (have_next)
      {
        in_flight = &next;
        device = std::thread(run_gpu, std::ref(next));
        cur ^= 1;
        if (!pipelined) done = wait_device();          // one round at a time
      }
      if (done) decide(*done);
      if (!in_flight && !have_next && !done && open.empty()) break;
    }
  }
  catch (...)
  {
    if (device.joinable()) device.join();
    throw;
  }
  if (device.joinable()) device.join();
  if (timing)
    std::fprintf(stderr, "[miplib cuda b&b] %lld rounds, %lld nodes: assemble (pop + propagate) %.2f s, pack boxes %.2f s, "
                 "waiting for the device %.2f s, decide / round / branch %.2f s\n",
                 static_cast<long long>(rounds), static_cast<long long>(S.m_info.nodes), t_assemble, t_pack, t_gpu, t_post);
  // Global dual bound.  A finished tree proves the incumbent; an interrupted one (time or node limit) is
  // bounded by its best undecided node: rounds are processed whole, so every such node is either in
  // `open` here - ordered by bound, its top carries the smallest one - or was given up undecided
  // (numerical error, time limit inside its LP, nothing left to branch on), which lost_bound remembers.
  if (complete)
    S.m_info.best_bound = have_incumbent ? sgn * best : 0;
  else
  {
    double lower = std::min(lost_bound, open.empty() ? kInf : open.top().bound);
    if (have_incumbent) lower = std::min(lower, best);
    S.m_info.best_bound = sgn * lower;             // +-1e300-like values mean "nothing proven"
    S.m_info.bound_is_open_node = true;
  }
  if (unbounded_relaxation && !have_incumbent) return {R::InfeasibleOrUnbounded, false};
  return finish(complete);
}

}  // namespace miplib
