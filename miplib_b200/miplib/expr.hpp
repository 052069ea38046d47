// miplib/expr.hpp - (at most quadratic) polynomial expressions over Vars.
//
// Same public interface as the reference's miplib/expr.hpp:42-206 (constructors, arithmetic,
// linear_vars()/linear_coeffs() that backends lower from, lb()/ub() interval bounds, printing),
// different engine: the reference keeps two hash maps and copies them on every binary operator
// (reference miplib/expr.cpp:53-119, :599-633).  Here an expression is a flat term list that
// `+=` appends to in O(1) and that is put in canonical form (sorted by variable identity, equal
// variables merged, zero coefficients erased) only when somebody reads it.  Building a row of k
// terms with `+=` is therefore O(k) appends + one O(k log k) sort instead of k hash insertions,
// which is what makes 32 M-nonzero models loadable through the expression API.
//
// Like every handle in the library (Var, Constr, Solver) an Expr copy SHARES its terms with the
// original (reference miplib/expr.hpp:47); use copy() for an independent expression.
#pragma once

#include <iosfwd>
#include <memory>
#include <utility>
#include <vector>

#include "util.hpp"
#include "var.hpp"

namespace miplib {

namespace detail {

struct LinTerm
{
  Var var;
  double coef;
};

struct QuadTerm
{
  Var first, second;   // canonical form keeps first <= second in identity order
  double coef;
};

struct ExprImpl
{
  explicit ExprImpl(double c = 0) : constant(c) {}
  explicit ExprImpl(Var const& v) : constant(0) { linear.push_back({v, 1.0}); }

  void add_scaled(ExprImpl const& other, double factor);   // *this += factor * other
  void scale(double factor);
  void multiply(ExprImpl const& other);                    // throws on degree > 2
  void divide_by_var(Var const& v);
  void canonicalize() const;
  Solver const& solver() const;

  mutable std::vector<LinTerm> linear;
  mutable std::vector<QuadTerm> quad;
  double constant;
  mutable bool canonical = true;
};

std::ostream& operator<<(std::ostream& os, ExprImpl const& e);

}  // namespace detail

struct Expr
{
  Expr(double c = 0) : p_impl(std::make_shared<detail::ExprImpl>(c)) {}
  Expr(Var const& v) : p_impl(std::make_shared<detail::ExprImpl>(v)) {}
  Expr(Expr const& e) = default;            // shares the terms (handle semantics)
  Expr& operator=(Expr const&) = default;

  Expr copy() const;                        // deep copy

  bool is_constant() const;
  bool is_linear() const;
  bool is_quadratic() const;
  bool must_be_binary() const;
  bool must_be_integer() const;
  double constant() const { return p_impl->constant; }
  double is_zero() const { return is_constant() && constant() == 0; }

  // flat views of the canonical form; linear_vars()[i] goes with linear_coeffs()[i]
  std::vector<double> linear_coeffs() const;
  std::vector<Var> linear_vars() const;
  std::vector<double> quad_coeffs() const;
  std::vector<Var> quad_vars_1() const;
  std::vector<Var> quad_vars_2() const;

  std::pair<double, double> bounds() const;
  std::pair<double, double> numerical_range(bool ignore_inf_var_bounds) const;

  Solver const& solver() const { return p_impl->solver(); }

  Expr& operator+=(double c) { p_impl->constant += c; return *this; }
  Expr& operator+=(Var const& v);
  Expr& operator+=(Expr const& e);
  Expr& operator-=(double c) { p_impl->constant -= c; return *this; }
  Expr& operator-=(Var const& v);
  Expr& operator-=(Expr const& e);
  Expr& operator*=(double c) { p_impl->scale(c); return *this; }
  Expr& operator*=(Var const& v) { return *this *= Expr(v); }
  Expr& operator*=(Expr const& e);
  Expr& operator/=(double c) { p_impl->scale(1 / c); return *this; }

  Expr operator-() const;
  Expr operator+(double c) const;
  Expr operator+(Var const& v) const;
  Expr operator+(Expr const& e) const;
  Expr operator-(double c) const { return *this + (-c); }
  Expr operator-(Var const& v) const;
  Expr operator-(Expr const& e) const;
  Expr operator*(double c) const;
  Expr operator*(Var const& v) const;
  Expr operator*(Expr const& e) const;
  Expr operator/(double c) const { return *this * (1 / c); }
  Expr operator/(Expr const& e) const;      // e must be constant

  // divides by v assuming v != 0; every term must contain v
  Expr& div_by_nonzero(Var const& v) { p_impl->divide_by_var(v); return *this; }

  double lb() const;
  double ub() const;
  double value() const;

  std::vector<Var> vars() const;
  std::size_t arity() const { return vars().size(); }

  private:
  std::shared_ptr<detail::ExprImpl> p_impl;
  friend std::ostream& operator<<(std::ostream& os, Expr const& e);
};

std::ostream& operator<<(std::ostream& os, Expr const& e);

// mixed Var / double / Expr operators
inline Expr operator-(Var const& v) { return Expr(v) * -1.0; }
inline Expr operator+(Var const& a, double c) { return Expr(a) + c; }
inline Expr operator+(double c, Var const& a) { return Expr(a) + c; }
inline Expr operator+(double c, Expr const& e) { return e + c; }
inline Expr operator+(Var const& a, Var const& b) { return Expr(a) + b; }
inline Expr operator+(Var const& a, Expr const& e) { return e + a; }
inline Expr operator-(Var const& a, double c) { return Expr(a) + (-c); }
inline Expr operator-(double c, Var const& a) { return Expr(c) - a; }
inline Expr operator-(double c, Expr const& e) { return (-e) + c; }
inline Expr operator-(Var const& a, Var const& b) { return Expr(a) - b; }
inline Expr operator-(Var const& a, Expr const& e) { return Expr(a) - e; }
inline Expr operator*(Var const& a, double c) { return Expr(a) * c; }
inline Expr operator*(double c, Var const& a) { return Expr(a) * c; }
inline Expr operator*(double c, Expr const& e) { return e * c; }
inline Expr operator*(Var const& a, Var const& b) { return Expr(a) * b; }
inline Expr operator*(Var const& a, Expr const& e) { return Expr(a) * e; }
inline Expr operator/(Var const& a, double c) { return Expr(a) / c; }
inline Expr operator/(Var const& a, Expr const& e) { return Expr(a) / e; }

}  // namespace miplib
