// miplib/capi.cpp - C shim over Solver / Var (reference miplib/capi.cpp:7-109).
#include "capi.hpp"

#include <exception>
#include <stdexcept>
#include <string>

namespace {

std::string g_last_error;   // plain static like the reference's (capi.cpp:7): not thread safe

template<class Body>
int guarded(Body&& body) noexcept
{
  try
  {
    body();
    return 0;
  }
  catch (std::exception const& e)
  {
    g_last_error = e.what();
  }
  catch (...)
  {
    g_last_error = "unknown error";
  }
  return 1;
}

miplib::Solver::Backend to_backend(miplib_SolverBackend backend)
{
  using B = miplib::Solver::Backend;
  switch (backend)
  {
    case Gurobi: return B::Gurobi;
    case Scip: return B::Scip;
    case Lpsolve: return B::Lpsolve;
    case BestAtCompileTime: return B::BestAtCompileTime;
    case BestAtRunTime: return B::BestAtRunTime;
    case Cuda: return B::Cuda;
  }
  throw std::logic_error("Unsupported backend.");
}

miplib::Var::Type to_var_type(miplib_VarType type)
{
  using T = miplib::Var::Type;
  switch (type)
  {
    case miplib_VarType::Continuous: return T::Continuous;
    case miplib_VarType::Binary: return T::Binary;
    case miplib_VarType::Integer: return T::Integer;
  }
  throw std::logic_error("Unsupported var type.");
}

}  // namespace

// shared with the backend-specific C entry points (miplib/cuda/capi.cpp)
void miplib_store_error(char const* what) { g_last_error = what; }

extern "C" {

char const* miplib_get_last_error() { return g_last_error.c_str(); }

int miplib_create_solver(miplib::Solver** rp_solver, miplib_SolverBackend backend)
{
  return guarded([&] { *rp_solver = new miplib::Solver(to_backend(backend)); });
}

int miplib_destroy_solver(miplib::Solver* p_solver)
{
  return guarded([&] { delete p_solver; });
}

int miplib_shallow_copy_solver(miplib::Solver** rp_solver, miplib::Solver* p_solver)
{
  return guarded([&] { *rp_solver = new miplib::Solver(*p_solver); });
}

int miplib_create_var(miplib::Var** rp_var, miplib::Solver* p_solver, miplib_VarType type)
{
  return guarded([&] { *rp_var = new miplib::Var(*p_solver, to_var_type(type)); });
}

int miplib_destroy_var(miplib::Var* p_var)
{
  return guarded([&] { delete p_var; });
}

int miplib_shallow_copy_var(miplib::Var** rp_var, miplib::Var* p_var)
{
  return guarded([&] { *rp_var = new miplib::Var(*p_var); });
}

}
