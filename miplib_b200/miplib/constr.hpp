// miplib/constr.hpp - linear/quadratic constraints and indicator constraints.
//
// Same interface as the reference's miplib/constr.hpp:22-141.  A Constr means `expr <= 0` or
// `expr == 0`: `e1 <= e2` is stored as `e1 - e2 <= 0` and `>=` is `<=` with the sides swapped
// (reference miplib/constr.cpp:117-143, miplib/constr.hpp:64-67) - this canonical row is what
// every backend lowers (row_hi = -constant, row_lo = -constant for Equal, else -inf).
#pragma once

#include <iosfwd>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "expr.hpp"
#include "util.hpp"

namespace miplib {

// rows whose term magnitudes all lie in [1e-4, 1e4] are left alone by Constr::scale()
static double constexpr MIN_MAX_ABS_SKIP_SCALE = 1e-4;
static double constexpr MAX_MAX_ABS_SKIP_SCALE = 1e4;

namespace detail {
struct IConstr;
struct IIndicatorConstr;
}  // namespace detail

struct Constr
{
  enum Type { LessEqual, Equal };

  Constr(Solver const& solver, Type const& type, Expr const& e,
         std::optional<std::string> const& name = std::nullopt);

  Expr expr() const;
  Type type() const;
  std::optional<std::string> const& name() const;

  // truth value expressible as a linear expression without extra variables?
  bool is_reifiable() const;
  // 0 when the constraint holds, positive when it does not
  Expr reified() const;

  bool must_be_satisfied() const;
  bool must_be_violated() const;

  Constr scale(double skip_lb = MIN_MAX_ABS_SKIP_SCALE, double skip_ub = MAX_MAX_ABS_SKIP_SCALE,
               bool ignore_inf_var_bounds = false) const;

  private:
  // backends hang their row handle on the impl and reach it by downcast; private behind a friend
  // list exactly as in the reference (miplib/constr.hpp:52-56) - the two Cuda lines are the edit a
  // maintainer makes there (tests/test_dropin.py applies it to a copy of
  // the reference's header and compiles the adapter against it)
  std::shared_ptr<detail::IConstr> p_impl;
  friend std::ostream& operator<<(std::ostream& os, Constr const& c);
  friend struct CudaSolver;
  friend struct CudaBranchAndBound;
};

std::ostream& operator<<(std::ostream& os, Constr const& c);

Constr operator!(Expr const& e);
Constr operator<=(Expr const& e1, Expr const& e2);
Constr operator==(Expr const& e1, Expr const& e2);
inline Constr operator>=(Expr const& e1, Expr const& e2) { return e2 <= e1; }

// implicant -> implicand
struct IndicatorConstr
{
  IndicatorConstr(Solver const& solver, Constr const& implicant, Constr const& implicand,
                  std::optional<std::string> const& name = std::nullopt);

  Constr const& implicant() const;
  Constr const& implicand() const;
  std::optional<std::string> const& name() const;

  // big-M linearisation: possible when the implicant is reifiable and the implicand is bounded
  bool has_reformulation() const;
  std::vector<Constr> reformulation() const;
  std::vector<Constr> scale(double skip_lb = MIN_MAX_ABS_SKIP_SCALE,
                            double skip_ub = MAX_MAX_ABS_SKIP_SCALE) const;

  std::shared_ptr<detail::IIndicatorConstr> p_impl;
};

std::ostream& operator<<(std::ostream& os, IndicatorConstr const& c);

IndicatorConstr operator>>(Constr const& implicant, Constr const& implicand);
IndicatorConstr operator<<(Constr const& implicand, Constr const& implicant);
inline IndicatorConstr operator>>(Expr const& implicant, Constr const& implicand)
{
  return (implicant == 1) >> implicand;
}
inline IndicatorConstr operator<<(Constr const& implicand, Expr const& implicant)
{
  return (implicant == 1) >> implicand;
}

namespace detail {

struct IConstr
{
  IConstr(Expr const& expr, Constr::Type const& type, std::optional<std::string> const& name)
    : m_expr(expr), m_type(type), m_name(name) {}
  virtual ~IConstr() = default;
  Expr m_expr;
  Constr::Type m_type;
  std::optional<std::string> m_name;
};

struct IIndicatorConstr
{
  IIndicatorConstr(Constr const& implicant, Constr const& implicand,
                   std::optional<std::string> const& name)
    : m_implicant(implicant), m_implicand(implicand), m_name(name) {}
  virtual ~IIndicatorConstr() = default;
  Constr m_implicant;
  Constr m_implicand;
  std::optional<std::string> m_name;
};

// for backends without native indicator rows (Lpsolve, Cuda): a plain holder that
// IndicatorConstr::reformulation() can work from
std::shared_ptr<IIndicatorConstr> create_reformulatable_indicator_constr(
  Constr const& implicant, Constr const& implicand, std::optional<std::string> const& name);

}  // namespace detail

}  // namespace miplib
