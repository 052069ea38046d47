// miplib/var.hpp - variable handle of the modelling front-end.
//
// API-compatible with the reference's miplib/var.hpp (Var at :17-70, detail::IVar at :86-104,
// hash/equal_to/less at :111-155) so user code, tests and backend plugins compile unchanged.
// Written fresh: no Boost (the reference only uses boost::hash to hash a pair of Vars),
// <cmath>/<string> included explicitly, IVar gets the virtual destructor the reference forgot.
#pragma once

#include <cmath>
#include <cstddef>
#include <functional>
#include <iosfwd>
#include <memory>
#include <optional>
#include <string>
#include <type_traits>

namespace miplib {

struct Solver;

namespace detail {
struct IVar;
}

struct Var
{
  enum class Type { Continuous, Binary, Integer };

  Var(Solver const& solver, Type const& type,
      std::optional<double> const& lb = std::nullopt,
      std::optional<double> const& ub = std::nullopt,
      std::optional<std::string> const& name = std::nullopt);
  Var(Solver const& solver, Type const& type, std::string const& name);

  // identity is the backend object, not the name
  bool is_same(Var const& other) const { return p_impl == other.p_impl; }
  bool is_lex_less(Var const& other) const { return p_impl < other.p_impl; }

  // the backend name when there is one, else a unique generated string
  std::string id() const;
  std::optional<std::string> name() const;
  void set_name(std::string const& new_name);

  double value() const;
  template<class T> T value_as() const
  {
    if constexpr (std::is_integral<T>::value) return T(std::llround(value()));
    else return T(value());
  }

  Type type() const;
  Solver const& solver() const;

  double lb() const;
  double ub() const;
  void set_lb(double new_lb);
  void set_ub(double new_ub);
  void set_hint(double v);

  // public on purpose: backends downcast it (reference miplib/var.hpp:57)
  std::shared_ptr<detail::IVar> p_impl;
};

std::ostream& operator<<(std::ostream& os, Var const& v);

namespace detail {

// What a backend provides per variable (one object per column).
struct IVar
{
  virtual ~IVar() = default;

  virtual double value() const = 0;
  virtual Var::Type type() const = 0;
  virtual std::optional<std::string> name() const = 0;
  virtual void set_name(std::string const& new_name) = 0;
  virtual Solver const& solver() const = 0;
  virtual double lb() const = 0;
  virtual double ub() const = 0;
  virtual void set_lb(double new_lb) = 0;
  virtual void set_ub(double new_ub) = 0;
  virtual void set_hint(double v) = 0;
};

}  // namespace detail

inline std::optional<std::string> Var::name() const { return p_impl->name(); }
inline void Var::set_name(std::string const& new_name) { p_impl->set_name(new_name); }
inline double Var::value() const { return p_impl->value(); }
inline Var::Type Var::type() const { return p_impl->type(); }
inline Solver const& Var::solver() const { return p_impl->solver(); }
inline double Var::lb() const { return p_impl->lb(); }
inline double Var::ub() const { return p_impl->ub(); }
inline void Var::set_lb(double new_lb) { p_impl->set_lb(new_lb); }
inline void Var::set_ub(double new_ub) { p_impl->set_ub(new_ub); }
inline void Var::set_hint(double v) { p_impl->set_hint(v); }

}  // namespace miplib

namespace std {

template<> struct hash<miplib::Var>
{
  size_t operator()(miplib::Var const& v) const noexcept
  {
    return hash<miplib::detail::IVar*>()(v.p_impl.get());
  }
};

template<> struct equal_to<miplib::Var>
{
  bool operator()(miplib::Var const& a, miplib::Var const& b) const noexcept { return a.is_same(b); }
};

template<> struct less<miplib::Var>
{
  bool operator()(miplib::Var const& a, miplib::Var const& b) const noexcept { return a.is_lex_less(b); }
};

}  // namespace std
