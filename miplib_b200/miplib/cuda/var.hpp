// miplib/cuda/var.hpp - per-column handle of the Cuda backend (cf. reference
// miplib/lpsolve/var.hpp:8-41, the closest behavioural template).
#pragma once

#include "../solver.hpp"
#include "../var.hpp"

namespace miplib {

struct CudaSolver;

struct CudaVar : detail::IVar
{
  CudaVar(Solver const& solver, std::size_t column) : m_solver(solver), m_column(column) {}

  double value() const override;
  Var::Type type() const override;
  std::optional<std::string> name() const override;
  void set_name(std::string const& new_name) override;
  Solver const& solver() const override { return m_solver; }
  double lb() const override;
  double ub() const override;
  void set_lb(double new_lb) override;
  void set_ub(double new_ub) override;
  void set_hint(double v) override;

  CudaSolver& backend() const;

  Solver m_solver;         // shallow copy: a live Var keeps the solver (and its device memory) alive
  std::size_t m_column;    // column index in the lowered model; never changes
};

}  // namespace miplib
