// miplib/cuda/solver.cpp - the Backend::Cuda adapter (see solver.hpp).
//
// Lowering rules, each pinned to what the reference's adapters do with the same front-end object:
//   row  `sum a_j x_j + k <= 0`  ->  row_hi = -k, row_lo = -inf        (miplib/lpsolve/solver.cpp:115-130:
//   row  `sum a_j x_j + k == 0`  ->  row_lo = row_hi = -k               type 1 / 3, rhs = -constant;
//                                                                       miplib/scip/solver.cpp:117-118)
//   posting the same Constr twice throws                                (miplib/lpsolve/solver.cpp:103-106)
//   quadratic rows / objectives throw                                   (miplib/lpsolve/solver.cpp:78-81,109-110)
//   Binary columns are [0,1] whatever bounds were passed                (miplib/lpsolve/var.cpp:112-123)
//   missing bounds are -+infinity()                                     (miplib/lpsolve/var.cpp:37-48)
//   the objective's constant term is dropped                            (miplib/lpsolve/solver.cpp:70-73,
//                                                                        miplib/scip/solver.cpp:148-157)
//   set_objective replaces the previous objective                       (lp_solve semantics, :73)
#include "solver.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <mutex>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <cstring>

#include <mipcu.h>

#include "bnb.hpp"
#include "constr.hpp"
#include "var.hpp"

namespace miplib {

namespace {

[[noreturn]] void fail(std::string const& what) { throw std::logic_error(what); }

void check(mipcu_ctx* ctx, int rc)
{
  if (rc) fail(std::string("Cuda backend: ") + mipcu_last_error(ctx));
}

// body(first, last) over [0, count) on a few host threads; small ranges run in place.  Used on the
// per-column / per-nonzero passes of bulk-loaded models, which sit on the end-to-end path of a solve.
template<class Body>
void parallel_ranges(std::int64_t count, Body&& body)
{
  unsigned const workers = count >= (std::int64_t(1) << 18)
                             ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1u;
  if (workers == 1) { body(std::int64_t(0), count); return; }
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < workers; ++w)
    pool.emplace_back([&, w] { body(count * w / workers, count * (w + 1) / workers); });
  body(std::int64_t(0), count / workers);
  for (std::thread& t : pool) t.join();
}

// wall-clock laps on stderr when MIPLIB_CUDA_TIMING is set (where a bulk-loaded solve spends its host time)
struct Laps
{
  char const* tag;
  bool on = std::getenv("MIPLIB_CUDA_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void operator()(char const* what) const
  {
    if (on)
      std::fprintf(stderr, "[%s] %s %.3f s\n", tag, what,
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
};

}  // namespace

CudaSolver::CudaSolver(bool verbose) : m_verbose(verbose) {}

// Engine contexts outlive the solver objects that used them: a context owns a CUDA stream and a
// device memory pool that keeps what a model freed, and building those anew costs far more than
// re-loading a model into a warm one (config 3: ~0.7 s against ~0.05 s).  A program that creates one
// Solver per model - the reference's usage, README.md:24-60 - therefore gets its next context out of
// this per-device cache, reset to default parameters with no model loaded.
namespace {
// heap-allocated and never destroyed: a Solver with static storage duration may be torn down after
// this translation unit's own statics
struct ContextCache
{
  std::mutex mutex;
  std::vector<std::pair<int, mipcu_ctx*>> free;          // (device, context)
};
ContextCache& context_cache()
{
  static ContextCache* cache = new ContextCache();
  return *cache;
}
constexpr std::size_t kCtxCacheMax = 16;

mipcu_ctx* acquire_context(int device)
{
  {
    ContextCache& cache = context_cache();
    std::lock_guard<std::mutex> lock(cache.mutex);
    for (std::size_t i = 0; i < cache.free.size(); ++i)
      if (cache.free[i].first == device)
      {
        mipcu_ctx* ctx = cache.free[i].second;
        cache.free.erase(cache.free.begin() + static_cast<std::ptrdiff_t>(i));
        return ctx;
      }
  }
  mipcu_ctx* ctx = nullptr;
  if (mipcu_create(&ctx, device)) fail(std::string("Cuda backend: ") + mipcu_last_error(nullptr));
  return ctx;
}

void release_context(int device, mipcu_ctx* ctx)
{
  if (!ctx) return;
  if (mipcu_reset(ctx) == 0)
  {
    ContextCache& cache = context_cache();
    std::lock_guard<std::mutex> lock(cache.mutex);
    if (cache.free.size() < kCtxCacheMax)
    {
      cache.free.emplace_back(device, ctx);
      return;
    }
  }
  mipcu_destroy(ctx);
}
}  // namespace

CudaSolver::~CudaSolver()
{
  for (std::size_t g = 0; g < m_extra_ctx.size(); ++g) release_context(static_cast<int>(g + 1), m_extra_ctx[g]);
  release_context(0, m_ctx);
}

bool CudaSolver::is_available() { return mipcu_device_count() > 0; }

std::string CudaSolver::backend_info()
{
  std::ostringstream s;
  s << "miplib Cuda backend (restarted Halpern PDHG, sm_100a), " << mipcu_device_count()
    << " usable device(s)";
  return s.str();
}

// ---- model building (host only; works without a GPU) --------------------------------------------

std::shared_ptr<detail::IVar> CudaSolver::create_var(
  Solver const& solver, Var::Type const& type, std::optional<double> const& lb,
  std::optional<double> const& ub, std::optional<std::string> const& name)
{
  Column col;
  col.type = type;
  if (name) m_column_names[m_columns.size()] = *name;
  if (type == Var::Type::Binary)
  {
    col.lb = 0;
    col.ub = 1;
  }
  else
  {
    col.lb = lb.value_or(-infinity());
    col.ub = ub.value_or(infinity());
  }
  m_columns.push_back(col);
  m_structure_dirty = true;
  return std::make_shared<CudaVar>(solver, m_columns.size() - 1);
}

std::shared_ptr<detail::IConstr> CudaSolver::create_constr(
  Constr::Type const& type, Expr const& e, std::optional<std::string> const& name)
{
  return std::make_shared<CudaConstr>(e, type, name);
}

std::shared_ptr<detail::IIndicatorConstr> CudaSolver::create_indicator_constr(
  Constr const& implicant, Constr const& implicand, std::optional<std::string> const& name)
{
  return detail::create_reformulatable_indicator_constr(implicant, implicand, name);
}

std::size_t CudaSolver::column_of(Var const& v) const
{
  auto const* impl = dynamic_cast<CudaVar const*>(v.p_impl.get());
  if (!impl || &impl->backend() != this)
    fail("Cuda backend: expression uses a variable that belongs to another solver.");
  return impl->m_column;
}

void CudaSolver::add(Constr const& constr)
{
  auto const& impl = static_cast<CudaConstr const&>(*constr.p_impl);
  if (impl.m_row >= 0) fail("Attempt to post the same constraint twice.");
  Expr const e = constr.expr();
  if (e.is_quadratic()) fail("Cuda backend does not support quadratic constraints.");

  std::vector<Var> const vars = e.linear_vars();
  std::vector<double> const coefs = e.linear_coeffs();
  // Expr hands the terms out in identity (pointer) order: sort by column so the CSR - and with
  // it the FP64 summation order on the device - is reproducible from run to run
  std::vector<std::pair<std::int32_t, double>> entries(vars.size());
  for (std::size_t i = 0; i < vars.size(); ++i)
    entries[i] = {static_cast<std::int32_t>(column_of(vars[i])), coefs[i]};
  std::sort(entries.begin(), entries.end());
  for (auto const& [col, coef] : entries)
  {
    m_row_idx.push_back(col);
    m_row_val.push_back(coef);
  }
  m_row_start.push_back(static_cast<std::int64_t>(m_row_idx.size()));
  double const rhs = -e.constant();
  m_row_hi.push_back(rhs);
  m_row_lo.push_back(constr.type() == Constr::Equal ? rhs : -infinity());
  m_row_alive.push_back(1);
  impl.m_row = static_cast<std::int64_t>(m_row_hi.size()) - 1;
  m_structure_dirty = true;
  // posted from inside a lazy-constraint handler (reference: ScipSolver::add forwards to
  // add_lazy while in a callback, miplib/scip/solver.cpp:197-203): the running solve takes it in
  if (is_in_callback()) m_lazy_new_rows.push_back(impl.m_row);
}

void CudaSolver::add(IndicatorConstr const&)
{
  fail("Cuda backend does not support indicator constraints (they are reformulated by default).");
}

void CudaSolver::remove(Constr const& constr)
{
  auto const& impl = static_cast<CudaConstr const&>(*constr.p_impl);
  if (impl.m_row < 0 || !m_row_alive[impl.m_row]) fail("Attempt to remove a constraint that was not posted.");
  m_row_alive[impl.m_row] = 0;
  impl.m_row = -1;
  m_structure_dirty = true;
}

// at_integral_only is accepted for interface parity; handlers are always run on the candidates the
// solve would otherwise accept (integral points / the LP optimum), never on fractional node LPs.
void CudaSolver::add_lazy_constr_handler(LazyConstrHandler const& handler, bool)
{
  // a handler built from a null pointer faults on first use, exactly as with the reference's
  // backends (miplib/lazy.hpp:28-41 offers no way to ask)
  m_lazy_handlers.push_back(handler);
}

bool CudaSolver::lazy_accepts(std::vector<double> const& x, bool* added)
{
  if (added) *added = false;
  if (m_lazy_handlers.empty()) return true;
  struct Scope
  {
    CudaSolver& s;
    std::vector<double> const* saved;
    ~Scope() { s.m_callback_x = saved; }
  } scope{*this, m_callback_x};
  m_callback_x = &x;
  bool ok = true;
  for (LazyConstrHandler& h : m_lazy_handlers)
  {
    if (h.is_feasible()) continue;
    ok = false;
    std::size_t const before = m_lazy_new_rows.size();
    h.add();
    if (added && m_lazy_new_rows.size() > before) *added = true;
  }
  return ok;
}

void CudaSolver::set_objective(Solver::Sense const& sense, Expr const& e)
{
  if (e.is_quadratic()) fail("Cuda backend does not support quadratic objective functions.");
  m_objective.assign(m_columns.size(), 0.0);
  std::vector<Var> const vars = e.linear_vars();
  std::vector<double> const coefs = e.linear_coeffs();
  for (std::size_t i = 0; i < vars.size(); ++i) m_objective[column_of(vars[i])] = coefs[i];
  m_sense = sense;
  m_objective_dirty = true;
}

double CudaSolver::get_objective_value() const
{
  if (!m_has_solution) fail("Attempt to access the objective value before a solution was found.");
  return m_objective_value;
}

void CudaSolver::set_nr_threads(std::size_t n)
{
  if (n == 0) fail("Cuda backend: the number of GPUs must be at least 1.");
  m_nr_gpus = n;
}

void CudaSolver::column_bounds_changed(std::size_t column)
{
  m_dirty_columns.push_back(static_cast<std::int64_t>(column));
}

void CudaSolver::hint(std::size_t column, double value)
{
  if (m_start_x.size() < m_columns.size()) m_start_x.resize(m_columns.size(), 0.0);
  m_start_x[column] = value;
}

void CudaSolver::set_warm_start(PartialSolution const& partial_solution)
{
  m_start_x.assign(m_columns.size(), 0.0);
  for (auto const& [var, value] : partial_solution) m_start_x[column_of(var)] = value;
}

// ---- bulk ingress -----------------------------------------------------------------------------------

CudaSolver& CudaSolver::of(Solver const& solver)
{
  auto* impl = dynamic_cast<CudaSolver*>(solver.p_impl.get());
  if (!impl) fail("This entry point needs a Solver created with Backend::Cuda.");
  return *impl;
}

void CudaSolver::load_csr(Solver const&, std::int64_t m, std::int64_t n, std::int64_t const* row_ptr,
                          std::int32_t const* col_idx, double const* val, double const* c,
                          double const* lb, double const* ub, std::int32_t const* var_type,
                          double const* row_lo, double const* row_hi, bool maximize)
{
  if (m < 0 || n < 0 || !row_ptr || (n > 0 && (!c || !lb || !ub)) || (m > 0 && (!row_lo || !row_hi)))
    fail("miplib_cuda_load_csr: bad arguments");
  if (row_ptr[0] != 0 || (row_ptr[m] > 0 && (!col_idx || !val))) fail("miplib_cuda_load_csr: bad row_ptr");
  std::size_t const first = m_columns.size();
  double const inf = infinity();
  // validate everything before the row store is touched: a bad array must not leave half a model behind
  // (big models: a few host threads - this copy is on the end-to-end path of every bulk-loaded solve)
  for (std::int64_t j = 0; j < n && var_type; ++j)
    if (var_type[j] < 0 || var_type[j] > 2) fail("miplib_cuda_load_csr: unknown variable type");
  std::int64_t const nnz_new = row_ptr[m];
  auto parallel = [](std::int64_t count, auto&& body) { parallel_ranges(count, body); };
  Laps const lap{"load_csr"};
  std::atomic<int> bad{0};
  parallel(m, [&](std::int64_t r0, std::int64_t r1) {
    for (std::int64_t i = r0; i < r1; ++i)
    {
      if (row_ptr[i + 1] < row_ptr[i] || row_ptr[i] < 0 || row_ptr[i + 1] > nnz_new) { bad = 1; return; }
      std::int32_t prev = -1;
      for (std::int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k)
      {
        if (col_idx[k] <= prev || col_idx[k] >= n) { bad = 2; return; }
        prev = col_idx[k];
      }
    }
  });
  if (bad == 1) fail("miplib_cuda_load_csr: row_ptr must be non-decreasing");
  if (bad == 2) fail("miplib_cuda_load_csr: column indices must be increasing inside a row and below n");
  lap("validated");
  // Every array below grows once (BigVec::resize default-initialises, bigvec.hpp) and its new tail is
  // filled by the worker threads: on config 3 (2 M columns, 32 M nonzeros) the single-threaded range
  // inserts this replaces were 0.11 s of a 2 s end-to-end solve.
  m_columns.resize(first + static_cast<std::size_t>(n));
  parallel(n, [&](std::int64_t j0, std::int64_t j1) {
    for (std::int64_t j = j0; j < j1; ++j)
    {
      Column& col = m_columns[first + static_cast<std::size_t>(j)];
      col.type = static_cast<Var::Type>(var_type ? var_type[j] : 0);
      col.lb = col.type == Var::Type::Binary ? 0.0 : std::max(lb[j], -inf);
      col.ub = col.type == Var::Type::Binary ? 1.0 : std::min(ub[j], inf);
    }
  });
  lap("columns");
  // A big model bulk-loaded into an EMPTY solver goes to the device right here, out of the caller's own
  // arrays (page-locked ones then move at PCIe speed; the row store below is pageable) and at the same
  // time as the host threads fill the row store: on config 3 the device load (30 - 50 ms) and the host
  // copy (15 - 35 ms) used to run one after the other, the upload reading the copy.  Best effort: without
  // a usable device the model simply stays dirty and solve() reports the missing GPU as before.
  // (MIPLIB_CUDA_EAGER_NNZ overrides the size threshold: the tests run this path on small models)
  char const* eager_env = std::getenv("MIPLIB_CUDA_EAGER_NNZ");
  std::int64_t const eager_from = eager_env ? std::atoll(eager_env) : (std::int64_t(1) << 22);
  bool const eager = first == 0 && m_row_lo.empty() && nnz_new >= eager_from && mipcu_device_count() > 0;
  std::thread device_load;
  struct JoinOnExit
  {
    std::thread& t;
    ~JoinOnExit() { if (t.joinable()) t.join(); }
  } join_on_exit{device_load};      // an exception below must not destroy a running thread
  std::string device_error;
  bool device_loaded = false;
  BigVec<double> eager_lb, eager_ub;
  if (eager)
  {
    eager_lb.resize(static_cast<std::size_t>(n));
    eager_ub.resize(static_cast<std::size_t>(n));
    parallel(n, [&](std::int64_t j0, std::int64_t j1) {
      for (std::int64_t j = j0; j < j1; ++j)
      {
        eager_lb[static_cast<std::size_t>(j)] = m_columns[static_cast<std::size_t>(j)].lb;
        eager_ub[static_cast<std::size_t>(j)] = m_columns[static_cast<std::size_t>(j)].ub;
      }
    });
    device_load = std::thread([&] {
      try
      {
        ensure_context();
        if (mipcu_load_lp(m_ctx, m, n, nnz_new, row_ptr, col_idx, val, c, eager_lb.data(), eager_ub.data(), row_lo,
                          row_hi, maximize))
          device_error = mipcu_last_error(m_ctx);
        else
          device_loaded = true;
      }
      catch (std::exception const& e)
      {
        device_error = e.what();
      }
    });
  }
  std::size_t const nnz_before = m_row_idx.size();
  m_row_idx.resize(nnz_before + static_cast<std::size_t>(nnz_new));
  m_row_val.resize(nnz_before + static_cast<std::size_t>(nnz_new));
  parallel(nnz_new, [&](std::int64_t k0, std::int64_t k1) {
    std::int32_t* const idx_out = m_row_idx.data() + nnz_before;
    if (first == 0)
      std::memcpy(idx_out + k0, col_idx + k0, sizeof(std::int32_t) * static_cast<std::size_t>(k1 - k0));
    else
      for (std::int64_t k = k0; k < k1; ++k) idx_out[k] = col_idx[k] + static_cast<std::int32_t>(first);
    std::memcpy(m_row_val.data() + nnz_before + k0, val + k0, sizeof(double) * static_cast<std::size_t>(k1 - k0));
  });
  lap("nonzeros copied");
  std::size_t const rows_before = m_row_lo.size();
  m_row_start.resize(rows_before + 1 + static_cast<std::size_t>(m));
  m_row_lo.resize(rows_before + static_cast<std::size_t>(m));
  m_row_hi.resize(rows_before + static_cast<std::size_t>(m));
  parallel(m, [&](std::int64_t i0, std::int64_t i1) {
    for (std::int64_t i = i0; i < i1; ++i)
    {
      m_row_start[rows_before + 1 + static_cast<std::size_t>(i)] = static_cast<std::int64_t>(nnz_before) + row_ptr[i + 1];
      m_row_lo[rows_before + static_cast<std::size_t>(i)] = std::max(row_lo[i], -inf);
      m_row_hi[rows_before + static_cast<std::size_t>(i)] = std::min(row_hi[i], inf);
    }
  });
  m_row_alive.insert(m_row_alive.end(), static_cast<std::size_t>(m), char(1));
  lap("row sides");
  if (first == 0)
    m_objective.assign(c, c + n);
  else
  {
    m_objective.assign(m_columns.size(), 0.0);
    std::copy(c, c + n, m_objective.begin() + static_cast<std::ptrdiff_t>(first));
  }
  lap("objective");
  m_sense = maximize ? Solver::Sense::Maximize : Solver::Sense::Minimize;
  m_structure_dirty = true;
  m_objective_dirty = true;
  if (device_load.joinable())
  {
    device_load.join();
    lap("device load joined");
    if (device_loaded) device_holds_row_store();
  }
}

// the device copy was made from exactly what the row store now holds (every row alive): the view the
// host-side passes read (polishing, cuts) is the row store itself
void CudaSolver::device_holds_row_store()
{
  m_device_row_of.resize(m_row_alive.size());
  std::iota(m_device_row_of.begin(), m_device_row_of.end(), std::int64_t(0));
  std::vector<std::int64_t>().swap(m_last_ptr);
  std::vector<std::int32_t>().swap(m_last_idx);
  std::vector<double>().swap(m_last_val);
  std::vector<double>().swap(m_last_lo);
  std::vector<double>().swap(m_last_hi);
  m_up = UploadedCsr{m_row_start.data(), m_row_idx.data(), m_row_val.data(), m_row_lo.data(), m_row_hi.data(),
                     m_row_lo.size(), m_row_idx.size()};
  m_structure_dirty = false;
  m_objective_dirty = false;
  m_dirty_columns.clear();
}

// ---- dump: free-format MPS or CPLEX-LP text of the lowered model ---------------------------------

void CudaSolver::dump(std::string const& filename) const
{
  auto ends_with = [&](char const* ext) {
    std::string e(ext);
    if (filename.size() < e.size()) return false;
    std::string tail = filename.substr(filename.size() - e.size());
    std::transform(tail.begin(), tail.end(), tail.begin(), ::tolower);
    return tail == e;
  };
  bool const as_mps = ends_with(".mps"), as_lp = ends_with(".lp");
  if (!as_mps && !as_lp) fail("Dumping Cuda backend models to this file type is not supported (use .mps or .lp).");
  // The text is assembled with snprintf and written with stdio: iostream NUMBER formatting
  // (std::num_put) crashes inside CPython hosts that have loaded an extension with its own copy
  // of the C++ runtime's locale tables (SciPy's does) - the reason bnb.cpp logs through stdio too.
  std::string out;
  auto num = [&](double v) {
    char buf[40];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    out += buf;
  };
  auto idx = [&](std::size_t v) { out += std::to_string(v); };
  std::size_t const n = m_columns.size();
  auto col_name = [&](std::size_t j) {
    auto const it = m_column_names.find(j);
    if (it != m_column_names.end() && it->second.find(' ') == std::string::npos) return it->second;
    return "x" + std::to_string(j);
  };
  std::vector<std::size_t> rows;
  for (std::size_t r = 0; r < m_row_alive.size(); ++r)
    if (m_row_alive[r]) rows.push_back(r);
  double const inf = infinity();
  auto obj = [&](std::size_t j) { return j < m_objective.size() ? m_objective[j] : 0.0; };

  if (as_mps)
  {
    out += "NAME miplib_cuda\n";
    if (m_sense == Solver::Sense::Maximize) out += "OBJSENSE\n    MAX\n";
    out += "ROWS\n N obj\n";
    for (std::size_t r : rows)
    {
      out += m_row_lo[r] == m_row_hi[r] ? " E r" : (m_row_hi[r] < inf ? " L r" : " G r");
      idx(r);
      out += "\n";
    }
    // column-major pass over the row store
    std::vector<std::vector<std::pair<std::size_t, double>>> by_col(n);
    for (std::size_t r : rows)
      for (std::int64_t k = m_row_start[r]; k < m_row_start[r + 1]; ++k)
        by_col[m_row_idx[k]].push_back({r, m_row_val[k]});
    out += "COLUMNS\n";
    bool in_int = false;
    for (std::size_t j = 0; j < n; ++j)
    {
      bool const is_int = m_columns[j].type != Var::Type::Continuous;
      if (is_int != in_int)
      {
        out += is_int ? " MARKER 'MARKER' 'INTORG'\n" : " MARKER 'MARKER' 'INTEND'\n";   // quoted: the MPS convention
        in_int = is_int;
      }
      std::string const name = col_name(j);
      out += " " + name + " obj ";
      num(obj(j));
      out += "\n";
      for (auto const& [r, v] : by_col[j])
      {
        out += " " + name + " r";
        idx(r);
        out += " ";
        num(v);
        out += "\n";
      }
    }
    if (in_int) out += " MARKER 'MARKER' 'INTEND'\n";
    out += "RHS\n";
    for (std::size_t r : rows)
    {
      double const rhs = m_row_hi[r] < inf ? m_row_hi[r] : m_row_lo[r];
      if (rhs != 0) { out += " rhs r"; idx(r); out += " "; num(rhs); out += "\n"; }
    }
    // two-sided rows (bulk ingress, read_mps): written as L rows with a range
    bool ranges = false;
    for (std::size_t r : rows)
      if (m_row_lo[r] != m_row_hi[r] && m_row_hi[r] < inf && m_row_lo[r] > -inf)
      {
        if (!ranges) { out += "RANGES\n"; ranges = true; }
        out += " rng r"; idx(r); out += " "; num(m_row_hi[r] - m_row_lo[r]); out += "\n";
      }
    out += "BOUNDS\n";
    for (std::size_t j = 0; j < n; ++j)
    {
      Column const& c = m_columns[j];
      std::string const name = col_name(j);
      bool const lo_inf = c.lb <= -inf, hi_inf = c.ub >= inf;
      if (lo_inf && hi_inf) out += " FR bnd " + name + "\n";
      else
      {
        if (lo_inf) out += " MI bnd " + name + "\n";
        else if (c.lb != 0 || c.type != Var::Type::Continuous) { out += " LO bnd " + name + " "; num(c.lb); out += "\n"; }
        if (!hi_inf) { out += " UP bnd " + name + " "; num(c.ub); out += "\n"; }
        else if (c.type != Var::Type::Continuous) out += " PL bnd " + name + "\n";
      }
    }
    out += "ENDATA\n";
  }
  else
  {
    auto term = [&](double v, std::size_t j) {
      out += v < 0 ? " - " : " + ";
      num(std::abs(v));
      out += " " + col_name(j);
    };
    out += m_sense == Solver::Sense::Maximize ? "Maximize\n obj:" : "Minimize\n obj:";
    bool any = false;
    for (std::size_t j = 0; j < n; ++j)
      if (obj(j) != 0) { term(obj(j), j); any = true; }
    if (!any && n) out += " 0 " + col_name(0);
    out += "\nSubject To\n";
    for (std::size_t r : rows)
    {
      auto body = [&]() {
        for (std::int64_t k = m_row_start[r]; k < m_row_start[r + 1]; ++k) term(m_row_val[k], m_row_idx[k]);
      };
      if (m_row_lo[r] == m_row_hi[r]) { out += " r"; idx(r); out += ":"; body(); out += " = "; num(m_row_hi[r]); out += "\n"; }
      else
      {
        if (m_row_hi[r] < inf) { out += " r"; idx(r); out += ":"; body(); out += " <= "; num(m_row_hi[r]); out += "\n"; }
        if (m_row_lo[r] > -inf) { out += " r"; idx(r); out += "_lo:"; body(); out += " >= "; num(m_row_lo[r]); out += "\n"; }
      }
    }
    out += "Bounds\n";
    for (std::size_t j = 0; j < n; ++j)
    {
      Column const& c = m_columns[j];
      if (c.lb <= -inf && c.ub >= inf) out += " " + col_name(j) + " free\n";
      else
      {
        out += " ";
        if (c.lb <= -inf) out += "-inf"; else num(c.lb);
        out += " <= " + col_name(j) + " <= ";
        if (c.ub >= inf) out += "+inf"; else num(c.ub);
        out += "\n";
      }
    }
    bool header = false;
    for (std::size_t j = 0; j < n; ++j)
      if (m_columns[j].type != Var::Type::Continuous)
      {
        if (!header) { out += "General\n"; header = true; }
        out += " " + col_name(j) + "\n";
      }
    out += "End\n";
  }
  std::FILE* file = std::fopen(filename.c_str(), "w");
  if (!file) fail("Cuda backend: cannot open " + filename + " for writing.");
  bool const ok = std::fwrite(out.data(), 1, out.size(), file) == out.size();
  if (std::fclose(file) != 0 || !ok) fail("Cuda backend: error writing " + filename + ".");
}

// ---- device side ---------------------------------------------------------------------------------

void CudaSolver::ensure_context()
{
  if (m_ctx) return;
  m_ctx = acquire_context(0);
}

void CudaSolver::upload_model()
{
  ensure_context();
  std::size_t const n = m_columns.size();
  m_objective.resize(n, 0.0);
  bool const maximize = m_sense == Solver::Sense::Maximize;
  if (m_structure_dirty)
  {
    Laps const lap{"upload"};
    BigVec<double> lb(n), ub(n);
    parallel_ranges(static_cast<std::int64_t>(n), [&](std::int64_t j0, std::int64_t j1) {
      for (std::int64_t j = j0; j < j1; ++j)
      {
        lb[static_cast<std::size_t>(j)] = m_columns[static_cast<std::size_t>(j)].lb;
        ub[static_cast<std::size_t>(j)] = m_columns[static_cast<std::size_t>(j)].ub;
      }
    });
    lap("bounds gathered");
    bool const all_alive = std::find(m_row_alive.begin(), m_row_alive.end(), char(0)) == m_row_alive.end();
    m_device_row_of.clear();
    if (all_alive)
    {
      // the row store IS the CSR the engine wants: hand it over as it is
      m_device_row_of.resize(m_row_alive.size());
      std::iota(m_device_row_of.begin(), m_device_row_of.end(), std::int64_t(0));
      std::vector<std::int64_t>().swap(m_last_ptr);
      std::vector<std::int32_t>().swap(m_last_idx);
      std::vector<double>().swap(m_last_val);
      std::vector<double>().swap(m_last_lo);
      std::vector<double>().swap(m_last_hi);
      m_up = UploadedCsr{m_row_start.data(), m_row_idx.data(), m_row_val.data(), m_row_lo.data(), m_row_hi.data(),
                         m_row_lo.size(), m_row_idx.size()};
    }
    else
    {
      std::vector<std::int64_t> ptr{0};
      std::vector<std::int32_t> idx;
      std::vector<double> val, lo, hi;
      idx.reserve(m_row_idx.size());
      val.reserve(m_row_val.size());
      for (std::size_t r = 0; r < m_row_alive.size(); ++r)
      {
        if (!m_row_alive[r]) continue;
        idx.insert(idx.end(), m_row_idx.begin() + m_row_start[r], m_row_idx.begin() + m_row_start[r + 1]);
        val.insert(val.end(), m_row_val.begin() + m_row_start[r], m_row_val.begin() + m_row_start[r + 1]);
        ptr.push_back(static_cast<std::int64_t>(idx.size()));
        lo.push_back(m_row_lo[r]);
        hi.push_back(m_row_hi[r]);
        m_device_row_of.push_back(static_cast<std::int64_t>(r));
      }
      m_last_ptr.swap(ptr);
      m_last_idx.swap(idx);
      m_last_val.swap(val);
      m_last_lo.swap(lo);
      m_last_hi.swap(hi);
      view_last();
    }
    check(m_ctx, mipcu_load_lp(m_ctx, static_cast<std::int64_t>(m_up.m), static_cast<std::int64_t>(n),
                               static_cast<std::int64_t>(m_up.nnz), m_up.ptr, m_up.idx, m_up.val,
                               m_objective.data(), lb.data(), ub.data(), m_up.lo, m_up.hi, maximize));
    lap("mipcu_load_lp");
    m_structure_dirty = false;
    m_objective_dirty = false;
    m_dirty_columns.clear();
    return;
  }
  if (m_objective_dirty)
  {
    check(m_ctx, mipcu_set_objective(m_ctx, m_objective.data(), maximize));
    m_objective_dirty = false;
  }
  if (!m_dirty_columns.empty())
  {
    std::sort(m_dirty_columns.begin(), m_dirty_columns.end());
    m_dirty_columns.erase(std::unique(m_dirty_columns.begin(), m_dirty_columns.end()), m_dirty_columns.end());
    std::vector<double> lb, ub;
    for (std::int64_t j : m_dirty_columns)
    {
      lb.push_back(m_columns[j].lb);
      ub.push_back(m_columns[j].ub);
    }
    check(m_ctx, mipcu_set_bounds(m_ctx, static_cast<std::int64_t>(m_dirty_columns.size()),
                                  m_dirty_columns.data(), lb.data(), ub.data()));
    m_dirty_columns.clear();
  }
}

// Contexts that take node LPs of a branch-and-bound round: the primary one plus a copy of the
// model on every extra GPU requested with set_nr_threads (nodes are
// independent, the host deals them out, nothing is exchanged on the device).
void CudaSolver::upload_rows(std::vector<std::int64_t> const& ptr, std::vector<std::int32_t> const& idx,
                             std::vector<double> const& val, std::vector<double> const& lo,
                             std::vector<double> const& hi)
{
  ensure_context();
  std::size_t const n = m_columns.size();
  m_objective.resize(n, 0.0);
  std::vector<double> lb(n), ub(n);
  for (std::size_t j = 0; j < n; ++j)
  {
    lb[j] = m_columns[j].lb;
    ub[j] = m_columns[j].ub;
  }
  check(m_ctx, mipcu_load_lp(m_ctx, static_cast<std::int64_t>(lo.size()), static_cast<std::int64_t>(n),
                             static_cast<std::int64_t>(idx.size()), ptr.data(), idx.data(), val.data(),
                             m_objective.data(), lb.data(), ub.data(), lo.data(), hi.data(),
                             m_sense == Solver::Sense::Maximize));
  m_last_ptr = ptr;
  m_last_idx = idx;
  m_last_val = val;
  m_last_lo = lo;
  m_last_hi = hi;
  view_last();
  m_structure_dirty = true;   // the device no longer holds the user's own rows
  m_objective_dirty = false;
  m_dirty_columns.clear();
}

void CudaSolver::view_last()
{
  m_up = UploadedCsr{m_last_ptr.data(), m_last_idx.data(), m_last_val.data(), m_last_lo.data(), m_last_hi.data(),
                     m_last_lo.size(), m_last_idx.size()};
}

std::vector<mipcu_ctx*> CudaSolver::node_contexts()
{
  std::vector<mipcu_ctx*> all{m_ctx};
  std::size_t const want = std::min<std::size_t>(m_nr_gpus, static_cast<std::size_t>(mipcu_device_count()));
  std::size_t const n = m_columns.size();
  std::vector<double> lb(n), ub(n);
  for (std::size_t j = 0; j < n; ++j)
  {
    lb[j] = m_columns[j].lb;
    ub[j] = m_columns[j].ub;
  }
  for (std::size_t g = 1; g < want; ++g)
  {
    if (m_extra_ctx.size() < g)
    {
      mipcu_ctx* ctx = nullptr;
      ctx = acquire_context(static_cast<int>(g));
      m_extra_ctx.push_back(ctx);
    }
    mipcu_ctx* ctx = m_extra_ctx[g - 1];
    check(ctx, mipcu_load_lp(ctx, static_cast<std::int64_t>(m_up.m), static_cast<std::int64_t>(n),
                             static_cast<std::int64_t>(m_up.nnz), m_up.ptr, m_up.idx, m_up.val,
                             m_objective.data(), lb.data(), ub.data(), m_up.lo, m_up.hi,
                             m_sense == Solver::Sense::Maximize));
    all.push_back(ctx);
  }
  return all;
}

CudaSolver::LpOutcome CudaSolver::solve_lp_on_device(std::vector<double>* x_out, std::vector<double>* y_out)
{
  mipcu_stats st{};
  check(m_ctx, mipcu_solve_lp(m_ctx, &st));
  m_info.lp_solves += 1;
  m_info.lp_iterations += st.iterations;
  m_info.kernel_launches += st.kernel_launches;
  bool const has_point = st.status != MIPCU_NUMERICAL_ERROR;
  if (x_out)
  {
    x_out->assign(m_columns.size(), 0.0);
    if (has_point) check(m_ctx, mipcu_get_x(m_ctx, x_out->data()));
  }
  if (y_out && has_point)
  {
    y_out->resize(m_device_row_of.size());
    check(m_ctx, mipcu_get_y(m_ctx, y_out->data()));
  }
  return {st.status, st.primal_obj, st.dual_obj, st.iterations, st.rel_primal_res};
}

std::pair<Solver::Result, bool> CudaSolver::solve()
{
  auto const t0 = std::chrono::steady_clock::now();
  m_info = SolveInfo();
  m_has_solution = false;
  m_solution.clear();
  m_lazy_new_rows.clear();
  m_objective.resize(m_columns.size(), 0.0);

  Laps const lap{"solve"};
  // one threaded pass over the columns (crossed boxes, integrality) and one over the rows (rows
  // without variables are decided here; an empty model is trivially optimal)
  std::atomic<bool> crossed{false}, integers{false}, empty_row_violated{false};
  parallel_ranges(static_cast<std::int64_t>(m_columns.size()), [&](std::int64_t j0, std::int64_t j1) {
    bool x = false, i = false;
    for (std::int64_t j = j0; j < j1; ++j)
    {
      Column const& c = m_columns[static_cast<std::size_t>(j)];
      x |= c.lb > c.ub;
      i |= c.type != Var::Type::Continuous;
    }
    if (x) crossed = true;
    if (i) integers = true;
  });
  if (crossed) return {Solver::Result::Infeasible, false};
  parallel_ranges(static_cast<std::int64_t>(m_row_alive.size()), [&](std::int64_t r0, std::int64_t r1) {
    for (std::int64_t q = r0; q < r1; ++q)
    {
      std::size_t const r = static_cast<std::size_t>(q);
      if (m_row_alive[r] && m_row_start[r] == m_row_start[r + 1] && (m_row_lo[r] > 0 || m_row_hi[r] < 0))
        empty_row_violated = true;
    }
  });
  if (empty_row_violated) return {Solver::Result::Infeasible, false};
  if (m_columns.empty())
  {
    m_has_solution = true;
    m_objective_value = 0;
    return {Solver::Result::Optimal, true};
  }
  bool const has_integers = integers;
  lap("model checks");
  auto result = has_integers ? solve_integer() : solve_continuous();
  lap("solved");
  m_info.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return result;
}

// Crossover-lite ("polishing").  The reference's backends return vertex
// solutions of a simplex method and its tests compare values with == (test/unit/solver.cpp:37);
// a first-order method stops at a point that satisfies the KKT conditions to 1e-6.  For small
// continuous models this guesses the optimal face from that point - columns sitting on a bound,
// rows that are tight - solves the remaining square (or consistent over-determined) system
// exactly by Gaussian elimination on the host, and accepts the result only if it satisfies every
// row and bound to 1e-9 and does not move the objective.  Anything else keeps the PDHG point.
void CudaSolver::polish_vertex(std::vector<double>& x)
{
  std::size_t const n = m_columns.size(), m = m_up.m;
  if (n == 0 || !m_up.ptr) return;
  // cheap way out before anything proportional to the model: far more columns than the limit
  if (m_polish_limit == 0) return;
  double const inf = infinity(), tol = 1e-5;
  {
    // a first-order optimum of a big model has far more columns off their bounds than the limit:
    // find that out before copying anything (stops at the first m_polish_limit + 1 free columns)
    std::size_t free_columns = 0;
    for (std::size_t j = 0; j < n && free_columns <= m_polish_limit; ++j)
    {
      double const l = m_columns[j].lb, u = m_columns[j].ub;
      bool const at_l = l > -inf && x[j] - l <= tol * (1 + std::abs(l));
      bool const at_u = u < inf && u - x[j] <= tol * (1 + std::abs(u));
      free_columns += !(at_l || at_u);
    }
    if (free_columns > m_polish_limit) return;
  }
  std::vector<double> snapped(x);
  std::vector<std::int32_t> free_of(n, -1);
  std::vector<std::int32_t> F;
  for (std::size_t j = 0; j < n; ++j)
  {
    double const l = m_columns[j].lb, u = m_columns[j].ub;
    if (l > -inf && x[j] - l <= tol * (1 + std::abs(l))) snapped[j] = l;
    else if (u < inf && u - x[j] <= tol * (1 + std::abs(u))) snapped[j] = u;
    else { free_of[j] = static_cast<std::int32_t>(F.size()); F.push_back(static_cast<std::int32_t>(j)); }
  }
  if (F.size() > m_polish_limit) return;
  // tight rows and their targets
  std::vector<std::int32_t> R;
  std::vector<double> target;
  for (std::size_t r = 0; r < m; ++r)
  {
    double act = 0;
    bool touches_free = false;
    for (std::int64_t k = m_up.ptr[r]; k < m_up.ptr[r + 1]; ++k)
    {
      act += m_up.val[k] * x[m_up.idx[k]];
      touches_free = touches_free || free_of[m_up.idx[k]] >= 0;
    }
    double const lo = m_up.lo[r], hi = m_up.hi[r];
    bool const at_hi = hi < inf && std::abs(act - hi) <= tol * (1 + std::abs(hi));
    bool const at_lo = lo > -inf && std::abs(act - lo) <= tol * (1 + std::abs(lo));
    if (!touches_free || !(at_hi || at_lo)) continue;
    R.push_back(static_cast<std::int32_t>(r));
    target.push_back(at_hi ? hi : lo);
  }
  std::size_t const nf = F.size(), nr = R.size();
  if (nr < nf || nr > 4 * m_polish_limit) return;          // face not pinned down by the tight rows
  std::vector<double> cand(snapped);
  if (nf > 0)
  {
    // dense [nr x (nf + 1)] system  M x_F = target - A_N x_N
    std::size_t const w = nf + 1;
    std::vector<double> M(nr * w, 0.0);
    for (std::size_t i = 0; i < nr; ++i)
    {
      std::int32_t const r = R[i];
      double rhs = target[i];
      for (std::int64_t k = m_up.ptr[r]; k < m_up.ptr[r + 1]; ++k)
      {
        std::int32_t const f = free_of[m_up.idx[k]];
        if (f >= 0) M[i * w + f] += m_up.val[k];
        else rhs -= m_up.val[k] * snapped[m_up.idx[k]];
      }
      M[i * w + nf] = rhs;
    }
    std::vector<std::size_t> pivot_row(nf);
    std::vector<char> used(nr, 0);
    for (std::size_t c = 0; c < nf; ++c)
    {
      std::size_t best = nr;
      double big = 1e-9;
      for (std::size_t i = 0; i < nr; ++i)
        if (!used[i] && std::abs(M[i * w + c]) > big) { big = std::abs(M[i * w + c]); best = i; }
      if (best == nr) return;                               // rank deficient: no unique vertex here
      used[best] = 1;
      pivot_row[c] = best;
      double const inv = 1.0 / M[best * w + c];
      for (std::size_t i = 0; i < nr; ++i)
      {
        if (i == best || M[i * w + c] == 0.0) continue;
        double const f = M[i * w + c] * inv;
        for (std::size_t k = c; k < w; ++k) M[i * w + k] -= f * M[best * w + k];
      }
    }
    for (std::size_t i = 0; i < nr; ++i)
      if (!used[i] && std::abs(M[i * w + nf]) > 1e-7) return;   // redundant tight rows disagree
    for (std::size_t c = 0; c < nf; ++c) cand[F[c]] = M[pivot_row[c] * w + nf] / M[pivot_row[c] * w + c];
  }
  // accept only a point that is feasible to 1e-9 everywhere and leaves the objective where it was
  double obj_old = 0, obj_new = 0;
  for (std::size_t j = 0; j < n; ++j)
  {
    double const l = m_columns[j].lb, u = m_columns[j].ub;
    if (cand[j] < l - 1e-9 * (1 + std::abs(l)) || cand[j] > u + 1e-9 * (1 + std::abs(u))) return;
    cand[j] = std::min(std::max(cand[j], l), u);
    obj_old += m_objective[j] * x[j];
    obj_new += m_objective[j] * cand[j];
  }
  if (std::abs(obj_new - obj_old) > 1e-5 * (1 + std::abs(obj_old))) return;
  for (std::size_t r = 0; r < m; ++r)
  {
    double act = 0;
    for (std::int64_t k = m_up.ptr[r]; k < m_up.ptr[r + 1]; ++k) act += m_up.val[k] * cand[m_up.idx[k]];
    double const lo = m_up.lo[r], hi = m_up.hi[r];
    if (hi < inf && act > hi + 1e-9 * (1 + std::abs(hi))) return;
    if (lo > -inf && act < lo - 1e-9 * (1 + std::abs(lo))) return;
  }
  x.swap(cand);
  m_info.polished = true;
}

std::pair<Solver::Result, bool> CudaSolver::solve_continuous()
{
  using R = Solver::Result;
  // lazy constraints on a continuous model: solve, let the handlers see the optimum, re-solve
  // with the rows they posted (warm started from the rejected point) until they accept
  std::vector<double> x, restart;
  for (int round = 0;; ++round)
  {
    upload_model();
    check(m_ctx, mipcu_set_param(m_ctx, MIPCU_PARAM_EPS_REL, m_feas_tol));
    check(m_ctx, mipcu_set_param(m_ctx, MIPCU_PARAM_TIME_LIMIT, m_time_limit));
    check(m_ctx, mipcu_set_param(m_ctx, MIPCU_PARAM_VERBOSE, m_verbose ? 1 : 0));
    check(m_ctx, mipcu_set_param(m_ctx, MIPCU_PARAM_OBJ_CUTOFF, std::nan("")));
    double const* start = !restart.empty() ? restart.data()
                          : m_start_x.size() == m_columns.size() ? m_start_x.data() : nullptr;
    check(m_ctx, mipcu_warm_start(m_ctx, start, nullptr));
    LpOutcome const lp = solve_lp_on_device(&x, nullptr);
    switch (lp.status)
    {
      case MIPCU_OPTIMAL:
      {
        bool added = false;
        m_lazy_new_rows.clear();
        if (!lazy_accepts(x, &added))
        {
          if (!added || round >= 1000) return {R::Other, false};   // rejected, nothing to learn from
          m_info.lazy_rows += static_cast<std::int64_t>(m_lazy_new_rows.size());
          restart = x;
          continue;
        }
        polish_vertex(x);
        m_solution = std::move(x);
        m_has_solution = true;
        m_objective_value = std::inner_product(m_solution.begin(), m_solution.end(), m_objective.begin(), 0.0);
        m_info.best_bound = lp.dual_obj;
        m_info.proven_optimal = true;
        return {R::Optimal, true};
      }
      case MIPCU_INFEASIBLE: return {R::Infeasible, false};
      case MIPCU_UNBOUNDED: return {R::Unbounded, false};
      case MIPCU_INFEASIBLE_OR_UNBOUNDED: return {R::InfeasibleOrUnbounded, false};
      case MIPCU_ITERATION_LIMIT:
      case MIPCU_TIME_LIMIT:
        // interrupted with a point in hand: like SCIP (has_solution = a primal solution exists,
        // reference miplib/scip/solver.cpp:379-410) the last iterate is reported as the incumbent when it
        // is feasible to ten times the tolerance; lp_solve's timeout row (miplib/lpsolve/solver.cpp:189-192)
        // has no solution, which is what a still-infeasible iterate maps to
        if (lp.rel_primal_res <= 10.0 * m_feas_tol && x.size() == m_columns.size())
        {
          m_solution = std::move(x);
          m_has_solution = true;
          m_objective_value = std::inner_product(m_solution.begin(), m_solution.end(), m_objective.begin(), 0.0);
          return {R::Interrupted, true};
        }
        return {R::Interrupted, false};
      case MIPCU_NUMERICAL_ERROR: return {R::Error, false};
      default: return {R::Other, false};
    }
  }
}

std::pair<Solver::Result, bool> CudaSolver::solve_integer()
{
  CudaBranchAndBound bnb(*this);
  return bnb.run();
}

}  // namespace miplib
