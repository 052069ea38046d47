// miplib/lazy.hpp - lazy-constraint callback interface (reference miplib/lazy.hpp:11-66).
// The Cuda backend consults the handlers for every candidate point its solve is about to accept
// (integral branch-and-bound candidates, the optimum of a continuous model): see
// cuda/solver.hpp "lazy constraints" and cuda/bnb.cpp.
#pragma once

#include <memory>
#include <vector>

#include "constr.hpp"
#include "var.hpp"

namespace miplib {

namespace detail {
struct ICurrentStateHandle
{
  virtual ~ICurrentStateHandle() = default;
  virtual double value(IVar const& var) const = 0;
  virtual void add_lazy(Constr const& constr) = 0;
  virtual bool is_active() const = 0;
};
}  // namespace detail

struct ILazyConstrHandler
{
  virtual ~ILazyConstrHandler() = default;
  virtual bool is_feasible() = 0;
  virtual bool add() = 0;                       // true if at least one constraint was added
  virtual std::vector<Var> depends() const = 0;
};

struct LazyConstrHandler
{
  LazyConstrHandler(std::shared_ptr<ILazyConstrHandler> const& impl) : p_impl(impl) {}

  std::vector<Var> depends() const { return p_impl->depends(); }
  bool is_feasible() { return p_impl->is_feasible(); }
  bool add() { return p_impl->add(); }

  private:
  std::shared_ptr<ILazyConstrHandler> p_impl;
};

// forwards to an object with lazy_constraints_feasible / _add / _depends members
template<class Handler>
struct DefaultLazyConstrHandler : ILazyConstrHandler
{
  explicit DefaultLazyConstrHandler(Handler& handler) : m_handler(handler) {}
  bool is_feasible() override { return m_handler.lazy_constraints_feasible(); }
  bool add() override { return m_handler.lazy_constraints_add(); }
  std::vector<Var> depends() const override { return m_handler.lazy_constraints_depends(); }
  Handler& m_handler;
};

}  // namespace miplib
