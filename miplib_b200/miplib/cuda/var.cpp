#include "var.hpp"

#include <stdexcept>

#include "solver.hpp"

namespace miplib {

CudaSolver& CudaVar::backend() const { return static_cast<CudaSolver&>(*m_solver.p_impl); }

double CudaVar::value() const
{
  CudaSolver const& s = backend();
  // inside a lazy-constraint handler: the candidate under test (reference miplib/scip/var.cpp:69-74)
  if (s.m_callback_x && m_column < s.m_callback_x->size()) return (*s.m_callback_x)[m_column];
  if (!s.m_has_solution || m_column >= s.m_solution.size())
    throw std::logic_error("Attempt to access value of variable before a solution was found.");
  return s.m_solution[m_column];
}

Var::Type CudaVar::type() const { return backend().m_columns[m_column].type; }
std::optional<std::string> CudaVar::name() const { return backend().column_name(m_column); }
void CudaVar::set_name(std::string const& new_name) { backend().m_column_names[m_column] = new_name; }
double CudaVar::lb() const { return backend().m_columns[m_column].lb; }
double CudaVar::ub() const { return backend().m_columns[m_column].ub; }

void CudaVar::set_lb(double new_lb)
{
  backend().m_columns[m_column].lb = new_lb;
  backend().column_bounds_changed(m_column);
}

void CudaVar::set_ub(double new_ub)
{
  backend().m_columns[m_column].ub = new_ub;
  backend().column_bounds_changed(m_column);
}

void CudaVar::set_hint(double v) { backend().hint(m_column, v); }

}  // namespace miplib
