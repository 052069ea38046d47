// miplib/solver.cpp - backend selection and one-line forwards (reference miplib/solver.cpp).
#include "solver.hpp"

#include <ostream>
#include <stdexcept>

#ifdef WITH_CUDA
#include "cuda/solver.hpp"
#endif

namespace miplib {

namespace {

std::shared_ptr<detail::ISolver> make_backend(Solver::Backend backend, bool verbose)
{
  using B = Solver::Backend;
  (void)verbose;
  switch (backend)
  {
    case B::Gurobi: throw std::logic_error("Request for Gurobi backend but it was not compiled.");
    case B::Scip: throw std::logic_error("Request for SCIP backend but it was not compiled.");
    case B::Lpsolve: throw std::logic_error("Request for Lpsolve backend but it was not compiled.");
    case B::Cuda:
    case B::BestAtCompileTime:
#ifdef WITH_CUDA
      return std::make_shared<CudaSolver>(verbose);
#else
      if (backend == B::Cuda) throw std::logic_error("Request for Cuda backend but it was not compiled.");
      throw std::logic_error("No MIP backends were compiled.");
#endif
    case B::BestAtRunTime:
#ifdef WITH_CUDA
      if (CudaSolver::is_available()) return std::make_shared<CudaSolver>(verbose);
#endif
      throw std::logic_error("No MIP backends are available.");
  }
  throw std::logic_error("Unknown backend.");
}

}  // namespace

Solver::Solver(Backend backend, bool verbose)
  : p_impl(make_backend(backend, verbose)), m_backend(backend), m_constraint_autoscale(false)
{}

void Solver::set_objective(Sense const& sense, Expr const& e) { p_impl->set_objective(sense, e); }
double Solver::get_objective_value() const { return p_impl->get_objective_value(); }
Solver::Sense Solver::get_objective_sense() const { return p_impl->get_objective_sense(); }

void Solver::add(Constr const& constr, bool scale)
{
  if (constr.must_be_violated())
    throw std::logic_error("Attempt to create a constraint that is trivially unsat.");
  if (scale || m_constraint_autoscale) p_impl->add(constr.scale());
  else p_impl->add(constr);
}

void Solver::add(IndicatorConstr const& constr, bool scale)
{
  using Policy = IndicatorConstraintPolicy;
  Policy const policy = p_impl->m_indicator_constraint_policy;
  bool const linearise = scale || policy == Policy::Reformulate ||
    (policy == Policy::ReformulateIfUnsupported && !supports_indicator_constraint(constr));
  if (!linearise)
  {
    p_impl->add(constr);
    return;
  }
  for (Constr const& row : constr.reformulation()) add(row, scale);
}

void Solver::remove(Constr const& constr) { p_impl->remove(constr); }

void Solver::add_lazy_constr_handler(LazyConstrHandler const& handler, bool at_integral_only)
{
  p_impl->add_lazy_constr_handler(handler, at_integral_only);
}

std::pair<Solver::Result, bool> Solver::solve() { return p_impl->solve(); }

std::pair<Solver::Result, bool> Solver::maximize(Expr const& e)
{
  set_objective(Sense::Maximize, e);
  return solve();
}

std::pair<Solver::Result, bool> Solver::minimize(Expr const& e)
{
  set_objective(Sense::Minimize, e);
  return solve();
}

void Solver::set_non_convex_policy(NonConvexPolicy policy) { p_impl->set_non_convex_policy(policy); }
void Solver::set_indicator_constraint_policy(IndicatorConstraintPolicy policy)
{
  p_impl->set_indicator_constraint_policy(policy);
}
void Solver::set_constraint_autoscale(bool autoscale) { m_constraint_autoscale = autoscale; }

void Solver::set_int_feasibility_tolerance(double v) { p_impl->set_int_feasibility_tolerance(v); }
void Solver::set_feasibility_tolerance(double v) { p_impl->set_feasibility_tolerance(v); }
void Solver::set_epsilon(double v) { p_impl->set_epsilon(v); }
void Solver::set_nr_threads(std::size_t n) { p_impl->set_nr_threads(n); }
double Solver::get_int_feasibility_tolerance() const { return p_impl->get_int_feasibility_tolerance(); }
double Solver::get_feasibility_tolerance() const { return p_impl->get_feasibility_tolerance(); }
double Solver::get_epsilon() const { return p_impl->get_epsilon(); }

bool Solver::supports_indicator_constraint(IndicatorConstr const& c) const
{
  return p_impl->supports_indicator_constraint(c);
}
bool Solver::supports_quadratic_constraints() const { return p_impl->supports_quadratic_constraints(); }
bool Solver::supports_quadratic_objective() const { return p_impl->supports_quadratic_objective(); }

double Solver::infinity() const { return p_impl->infinity(); }
void Solver::dump(std::string const& filename) const { p_impl->dump(filename); }
void Solver::set_time_limit(double secs) { p_impl->set_time_limit(secs); }
void Solver::set_warm_start(PartialSolution const& s) { p_impl->set_warm_start(s); }
void Solver::set_reoptimizing(bool v) { p_impl->set_reoptimizing(v); }
void Solver::setup_reoptimization() { p_impl->setup_reoptimization(); }

bool Solver::backend_is_compiled(Backend const& backend)
{
#ifdef WITH_CUDA
  return backend == Backend::Cuda;
#else
  (void)backend;
  return false;
#endif
}

bool Solver::backend_is_available(Backend const& backend)
{
  if (!backend_is_compiled(backend)) return false;
#ifdef WITH_CUDA
  return CudaSolver::is_available();
#else
  return false;
#endif
}

std::map<Solver::Backend, std::string> Solver::backend_info()
{
  std::map<Backend, std::string> info;
#ifdef WITH_CUDA
  info[Backend::Cuda] = CudaSolver::backend_info();
#endif
  return info;
}

namespace detail {
void ISolver::set_indicator_constraint_policy(Solver::IndicatorConstraintPolicy policy)
{
  m_indicator_constraint_policy = policy;
}
}  // namespace detail

std::ostream& operator<<(std::ostream& os, Solver::Backend const& backend)
{
  static char const* const names[] = {"Gurobi", "Scip", "Lpsolve", "BestAtCompileTime",
                                      "BestAtRunTime", "Cuda"};
  return os << names[static_cast<int>(backend)];
}

}  // namespace miplib
