// miplib/constr.cpp - behaviour follows the reference's miplib/constr.cpp: triviality tests use
// exact comparisons (:32-56, pinned by the test strings), reification (:62-98), operators
// (:117-143), big-M reformulation of indicator constraints (:205-271).
#include "constr.hpp"
#include "solver.hpp"
#include "util/scale.hpp"

#include <ostream>
#include <stdexcept>

namespace miplib {

Constr::Constr(Solver const& solver, Type const& type, Expr const& e,
               std::optional<std::string> const& name)
  : p_impl(solver.p_impl->create_constr(type, e, name))
{}

Expr Constr::expr() const { return p_impl->m_expr; }
Constr::Type Constr::type() const { return p_impl->m_type; }
std::optional<std::string> const& Constr::name() const { return p_impl->m_name; }

bool Constr::must_be_satisfied() const
{
  Expr const e = expr();
  if (e.ub() > 0) return false;                 // some point of the box violates "<= 0"
  if (type() == LessEqual) return true;
  return !(e.lb() < 0);                         // equality: the range must be exactly {0}
}

bool Constr::must_be_violated() const
{
  Expr const e = expr();
  if (e.lb() > 0) return true;
  return type() == Equal && e.ub() < 0;
}

bool Constr::is_reifiable() const
{
  if (type() != Equal) return false;
  Expr const e = expr();
  if (!e.must_be_integer()) return false;
  return !(e.lb() * e.ub() < 0);                // the range must not straddle zero
}

Expr Constr::reified() const
{
  if (!is_reifiable()) throw std::logic_error("Attempt to reify non-reifiable constraint.");
  if (must_be_violated())
    throw std::logic_error("Attempt to reify a constraint that is trivially violated.");
  if (must_be_satisfied())
    throw std::logic_error("Attempt to reify a constraint that is trivially satisfied.");
  Expr const e = expr();
  // integer valued and one-signed: |e| itself is 0 iff the equality holds
  return e.ub() > 0 ? e : -e;
}

std::ostream& operator<<(std::ostream& os, Constr const& c)
{
  return os << c.expr() << (c.type() == Constr::LessEqual ? " <= 0" : " = 0");
}

namespace {

Solver const& solver_of(Expr const& e1, Expr const& e2)
{
  if (!e1.is_constant()) return e1.solver();
  if (!e2.is_constant()) return e2.solver();
  throw std::logic_error("Attempt to create a constraint from constant expressions");
}

}  // namespace

Constr operator!(Expr const& e)
{
  if (!e.must_be_binary())
    throw std::logic_error("Attempt to negate a possibly non-binary expression");
  return e == 0;
}

Constr operator<=(Expr const& e1, Expr const& e2)
{
  return Constr(solver_of(e1, e2), Constr::LessEqual, e1 - e2);
}

Constr operator==(Expr const& e1, Expr const& e2)
{
  return Constr(solver_of(e1, e2), Constr::Equal, e1 - e2);
}

Constr Constr::scale(double skip_lb, double skip_ub, bool ignore_inf_var_bounds) const
{
  return detail::scale_gm(*this, skip_lb, skip_ub, ignore_inf_var_bounds);
}

// ---- indicator constraints --------------------------------------------------------------------

IndicatorConstr::IndicatorConstr(Solver const& solver, Constr const& implicant,
                                 Constr const& implicand, std::optional<std::string> const& name)
  : p_impl(solver.p_impl->create_indicator_constr(implicant, implicand, name))
{}

Constr const& IndicatorConstr::implicant() const { return p_impl->m_implicant; }
Constr const& IndicatorConstr::implicand() const { return p_impl->m_implicand; }
std::optional<std::string> const& IndicatorConstr::name() const { return p_impl->m_name; }

std::ostream& operator<<(std::ostream& os, IndicatorConstr const& ic)
{
  return os << ic.implicant() << " -> " << ic.implicand();
}

IndicatorConstr operator>>(Constr const& implicant, Constr const& implicand)
{
  return IndicatorConstr(implicand.expr().solver(), implicant, implicand);
}

IndicatorConstr operator<<(Constr const& implicand, Constr const& implicant)
{
  return IndicatorConstr(implicand.expr().solver(), implicant, implicand);
}

namespace detail {

std::shared_ptr<IIndicatorConstr> create_reformulatable_indicator_constr(
  Constr const& implicant, Constr const& implicand, std::optional<std::string> const& name)
{
  return std::make_shared<IIndicatorConstr>(implicant, implicand, name);
}

}  // namespace detail

bool IndicatorConstr::has_reformulation() const
{
  if (!implicant().is_reifiable()) return false;
  Expr const body = implicand().expr();
  double const inf = body.solver().infinity();
  if (body.ub() == inf) return false;
  if (implicand().type() == Constr::Equal && body.lb() == -inf) return false;
  return true;
}

// implicant  ->  body <= 0      becomes      body <= ub(body) * f
// implicant  ->  body == 0      adds         lb(body) * f <= body
// where f = implicant().reified() is 0 exactly when the implicant holds, so the big-M sides
// ub(body) / lb(body) (interval arithmetic over the variable boxes) switch the row off otherwise.
std::vector<Constr> IndicatorConstr::reformulation() const
{
  if (!implicant().is_reifiable())
    throw std::logic_error(
      "Attempt to reformulate indicator constraint with non reifiable implicant.");
  Expr const body = implicand().expr();
  double const inf = body.solver().infinity();

  double const big_up = body.ub();
  if (big_up == inf)
    throw std::logic_error(
      "Attempt to reformulate indicator constraint with unknown implicand upper bound."
      " Try bounding the domain of the involved variables.");

  std::vector<Constr> rows;
  if (big_up > 0) rows.push_back(body <= big_up * implicant().reified());   // else redundant
  if (implicand().type() == Constr::LessEqual) return rows;

  double const big_down = body.lb();
  if (big_down == -inf)
    throw std::logic_error(
      "Attempt to reformulate indicator constraint with unknown implicand lower bound."
      " Try bounding the domain of the involved variables.");
  if (big_down < 0) rows.push_back(big_down * implicant().reified() <= body);
  return rows;
}

std::vector<Constr> IndicatorConstr::scale(double skip_lb, double skip_ub) const
{
  std::vector<Constr> rows;
  for (Constr const& row : reformulation()) rows.push_back(row.scale(skip_lb, skip_ub));
  return rows;
}

}  // namespace miplib
