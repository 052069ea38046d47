#include "var.hpp"
#include "solver.hpp"

#include <ostream>
#include <sstream>

namespace miplib {

Var::Var(Solver const& solver, Type const& type, std::optional<double> const& lb,
         std::optional<double> const& ub, std::optional<std::string> const& name)
  : p_impl(solver.p_impl->create_var(solver, type, lb, ub, name))
{}

Var::Var(Solver const& solver, Type const& type, std::string const& name)
  : Var(solver, type, std::nullopt, std::nullopt, name)
{}

std::string Var::id() const
{
  if (auto nm = p_impl->name()) return *nm;
  std::ostringstream generated;
  generated << static_cast<void const*>(p_impl.get());
  return generated.str();
}

std::ostream& operator<<(std::ostream& os, Var const& v) { return os << v.id(); }

}  // namespace miplib
