#include <cstring>
#include "capi.hpp"

#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

#include "bnb.hpp"
#include "solver.hpp"

extern void miplib_store_error(char const* what);   // capi.cpp

namespace {

template<class Body>
int guarded(Body&& body) noexcept
{
  try
  {
    body();
    return 0;
  }
  catch (std::exception const& e) { miplib_store_error(e.what()); }
  catch (...) { miplib_store_error("unknown error"); }
  return 1;
}

}  // namespace

extern "C" {

int miplib_cuda_load_csr(miplib::Solver* p_solver, std::int64_t m, std::int64_t n,
                         std::int64_t const* row_ptr, std::int32_t const* col_idx, double const* val,
                         double const* c, double const* lb, double const* ub,
                         std::int32_t const* var_type, double const* row_lo, double const* row_hi,
                         int maximize)
{
  return guarded([&] {
    miplib::CudaSolver::of(*p_solver).load_csr(*p_solver, m, n, row_ptr, col_idx, val, c, lb, ub, var_type,
                                               row_lo, row_hi, maximize != 0);
  });
}

int miplib_cuda_read_mps(miplib::Solver* p_solver, char const* path)
{
  return guarded([&] {
    if (!path) throw std::logic_error("miplib_cuda_read_mps: null path");
    miplib::CudaSolver::of(*p_solver).read_mps(*p_solver, path);
  });
}

int miplib_cuda_dump(miplib::Solver* p_solver, char const* path)
{
  return guarded([&] {
    if (!path) throw std::logic_error("miplib_cuda_dump: null path");
    p_solver->dump(path);
  });
}

int miplib_cuda_solve(miplib::Solver* p_solver, int* result, int* has_solution)
{
  return guarded([&] {
    auto const [r, has] = p_solver->solve();
    if (result) *result = static_cast<int>(r);
    if (has_solution) *has_solution = has ? 1 : 0;
  });
}

int miplib_cuda_get_values(miplib::Solver* p_solver, double* x, double* objective)
{
  return guarded([&] {
    miplib::CudaSolver const& s = miplib::CudaSolver::of(*p_solver);
    if (!s.m_has_solution) throw std::logic_error("Attempt to access values before a solution was found.");
    if (x && !s.m_solution.empty()) std::memcpy(x, s.m_solution.data(), sizeof(double) * s.m_solution.size());
    if (objective) *objective = s.m_objective_value;
  });
}

int miplib_cuda_get_info(miplib::Solver* p_solver, std::int64_t* lp_iterations, std::int64_t* kernel_launches,
                         std::int64_t* nodes, double* seconds)
{
  return guarded([&] {
    miplib::CudaSolver const& s = miplib::CudaSolver::of(*p_solver);
    if (lp_iterations) *lp_iterations = s.m_info.lp_iterations;
    if (kernel_launches) *kernel_launches = s.m_info.kernel_launches;
    if (nodes) *nodes = s.m_info.nodes;
    if (seconds) *seconds = s.m_info.seconds;
  });
}

int miplib_cuda_set_verbose(miplib::Solver* p_solver, int verbose)
{
  return guarded([&] { miplib::CudaSolver::of(*p_solver).set_verbose(verbose != 0); });
}

int miplib_cuda_set_time_limit(miplib::Solver* p_solver, double seconds)
{
  return guarded([&] { p_solver->set_time_limit(seconds); });
}

int miplib_cuda_set_nr_threads(miplib::Solver* p_solver, std::int64_t gpus)
{
  return guarded([&] {
    if (gpus < 1) throw std::logic_error("miplib_cuda_set_nr_threads: at least one GPU");
    p_solver->set_nr_threads(static_cast<std::size_t>(gpus));
  });
}

int miplib_cuda_get_bound(miplib::Solver* p_solver, double* best_bound, int* proven_optimal)
{
  return guarded([&] {
    miplib::CudaSolver const& s = miplib::CudaSolver::of(*p_solver);
    if (best_bound) *best_bound = s.m_info.best_bound;
    if (proven_optimal) *proven_optimal = s.m_info.proven_optimal ? 1 : 0;
  });
}

int miplib_cuda_selftest_local_search(miplib::Solver* p_solver, double const* x_lp, int trials, double* x_out,
                                      int* feasible, int* propagate_mismatches)
{
  return guarded([&] {
    miplib::CudaSolver& s = miplib::CudaSolver::of(*p_solver);
    std::size_t const n = s.m_columns.size();
    s.m_objective.resize(n, 0.0);
    miplib::CudaBranchAndBound bnb(s);
    std::vector<double> in(x_lp ? std::vector<double>(x_lp, x_lp + n) : std::vector<double>()), out;
    bool ok = false;
    int const mism = bnb.selftest_local_search(in, trials, out, &ok);
    if (x_out && out.size() == n) std::memcpy(x_out, out.data(), sizeof(double) * n);
    if (feasible) *feasible = ok ? 1 : 0;
    if (propagate_mismatches) *propagate_mismatches = mism;
  });
}

int miplib_cuda_get_shape(miplib::Solver* p_solver, std::int64_t* n, std::int64_t* m)
{
  return guarded([&] {
    miplib::CudaSolver const& s = miplib::CudaSolver::of(*p_solver);
    if (n) *n = static_cast<std::int64_t>(s.m_columns.size());
    if (m)
    {
      std::int64_t alive = 0;
      for (char a : s.m_row_alive) alive += a ? 1 : 0;
      *m = alive;
    }
  });
}

}
