// miplib/solver.hpp - the Solver facade and the backend plugin interface detail::ISolver.
//
// Same interface as the reference's miplib/solver.hpp (Solver :18-121, detail::ISolver :125-188)
// plus ONE new enumerator, Backend::Cuda, appended after the existing ones so their values do not
// move.  Every backend is an adapter behind ISolver; the only one compiled in this tree is the
// B200 backend (miplib/cuda/solver.hpp), selected under WITH_CUDA exactly like the reference
// selects its backends under WITH_GUROBI / WITH_SCIP / WITH_LPSOLVE (miplib/solver.cpp:17-79).
#pragma once

#include <map>
#include <memory>
#include <optional>
#include <string>
#include <utility>

#include "constr.hpp"
#include "lazy.hpp"
#include "var.hpp"

namespace miplib {

namespace detail {
struct ISolver;
}

struct Solver
{
  enum class Backend { Gurobi, Scip, Lpsolve, BestAtCompileTime, BestAtRunTime, Cuda };
  enum class NonConvexPolicy { Error, Linearize, Branch };
  enum class IndicatorConstraintPolicy { PassThrough, Reformulate, ReformulateIfUnsupported };
  enum class Sense { Maximize, Minimize };
  enum class Result { Optimal, Infeasible, InfeasibleOrUnbounded, Unbounded, Interrupted, Error, Other };

  Solver(Backend backend, bool verbose = true);

  Backend const& backend() const { return m_backend; }

  void set_objective(Sense const& sense, Expr const& e);
  double get_objective_value() const;
  Sense get_objective_sense() const;

  // rejects rows that no point of the variable boxes can satisfy; scale = true first rescales
  // the row by a power of two (an indicator constraint is then reformulated first)
  void add(Constr const& constr, bool scale = false);
  void add(IndicatorConstr const& constr, bool scale = false);
  void remove(Constr const& constr);

  void add_lazy_constr_handler(LazyConstrHandler const& constr_handler, bool at_integral_only);

  void set_non_convex_policy(NonConvexPolicy policy);
  void set_indicator_constraint_policy(IndicatorConstraintPolicy policy);
  void set_constraint_autoscale(bool autoscale);

  void set_int_feasibility_tolerance(double value);
  void set_feasibility_tolerance(double value);
  void set_epsilon(double value);
  void set_nr_threads(std::size_t);
  double get_int_feasibility_tolerance() const;
  double get_feasibility_tolerance() const;
  double get_epsilon() const;

  // (result, whether a solution is available)
  std::pair<Result, bool> solve();
  std::pair<Result, bool> maximize(Expr const& e);
  std::pair<Result, bool> minimize(Expr const& e);

  bool supports_indicator_constraint(IndicatorConstr const& constr) const;
  bool supports_quadratic_constraints() const;
  bool supports_quadratic_objective() const;

  static bool backend_is_compiled(Backend const&);
  // compiled AND usable right now (for Cuda: an sm_100 device is visible)
  static bool backend_is_available(Backend const&);
  static std::map<Backend, std::string> backend_info();

  double infinity() const;
  void dump(std::string const& filename) const;
  void set_time_limit(double secs);
  void set_warm_start(PartialSolution const& partial_solution);
  void set_reoptimizing(bool);
  void setup_reoptimization();

  private:
  std::shared_ptr<detail::ISolver> p_impl;
  Backend m_backend;
  bool m_constraint_autoscale;
  friend struct Var;
  friend struct Constr;
  friend struct IndicatorConstr;
  friend struct CudaSolver;   // the one edit a new backend needs here
  friend struct CudaVar;
};

namespace detail {

struct ISolver
{
  virtual ~ISolver() = default;

  virtual std::shared_ptr<IVar> create_var(Solver const& solver, Var::Type const& type,
                                           std::optional<double> const& lb,
                                           std::optional<double> const& ub,
                                           std::optional<std::string> const& name) = 0;
  virtual std::shared_ptr<IConstr> create_constr(Constr::Type const& type, Expr const& e,
                                                 std::optional<std::string> const& name) = 0;
  virtual std::shared_ptr<IIndicatorConstr> create_indicator_constr(
    Constr const& implicant, Constr const& implicand, std::optional<std::string> const& name) = 0;

  virtual void set_objective(Solver::Sense const& sense, Expr const& e) = 0;
  virtual double get_objective_value() const = 0;
  virtual Solver::Sense get_objective_sense() const = 0;

  virtual void add(Constr const& constr) = 0;
  virtual void add(IndicatorConstr const& constr) = 0;
  virtual void remove(Constr const& constr) = 0;
  virtual void add_lazy_constr_handler(LazyConstrHandler const& constr, bool at_integral_only) = 0;

  virtual std::pair<Solver::Result, bool> solve() = 0;

  virtual void set_non_convex_policy(Solver::NonConvexPolicy policy) = 0;
  virtual void set_indicator_constraint_policy(Solver::IndicatorConstraintPolicy policy);

  virtual void set_int_feasibility_tolerance(double value) = 0;
  virtual void set_feasibility_tolerance(double value) = 0;
  virtual void set_epsilon(double value) = 0;
  virtual void set_nr_threads(std::size_t) = 0;
  virtual double get_int_feasibility_tolerance() const = 0;
  virtual double get_feasibility_tolerance() const = 0;
  virtual double get_epsilon() const = 0;

  virtual bool supports_indicator_constraint(IndicatorConstr const& constr) const = 0;
  virtual bool supports_quadratic_constraints() const = 0;
  virtual bool supports_quadratic_objective() const = 0;

  virtual double infinity() const = 0;
  virtual void set_time_limit(double secs) = 0;
  virtual void dump(std::string const& filename) const = 0;
  virtual void set_warm_start(PartialSolution const& partial_solution) = 0;
  virtual void set_reoptimizing(bool) = 0;
  virtual void setup_reoptimization() = 0;

  Solver::IndicatorConstraintPolicy m_indicator_constraint_policy =
    Solver::IndicatorConstraintPolicy::ReformulateIfUnsupported;
};

}  // namespace detail

std::ostream& operator<<(std::ostream& os, Solver::Backend const& solver_backend);

}  // namespace miplib
