// miplib/cuda/solver.hpp - Solver::Backend::Cuda: the adapter between the modelling front-end
// and libmiplib_cuda.so (include/mipcu.h).
//
// It plays the role LpsolveSolver / ScipSolver play in the reference (miplib/lpsolve/solver.hpp,
// miplib/scip/solver.hpp): create_var / create_constr hand out handles, add() lowers a row
// immediately into host-side CSR triplets (the reference lowers into the vendor model,
// miplib/lpsolve/solver.cpp:99-131, miplib/scip/solver.cpp:69-143), set_objective() keeps a
// dense c, and solve() - the single vendor call in the reference (miplib/lpsolve/solver.cpp:156,
// miplib/scip/solver.cpp:370) - uploads the model through the C ABI and runs the PDHG LP engine,
// wrapped in a branch-and-bound when integer variables are present.
//
// Only the plugin seam is used (Expr::linear_vars/linear_coeffs/constant, Var::p_impl,
// Constr::p_impl), so the same file drops into the reference tree.
#pragma once

#include <cstdint>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

#include "../solver.hpp"
#include "bigvec.hpp"

struct mipcu_ctx;

namespace miplib {

struct CudaSolver : detail::ISolver
{
  explicit CudaSolver(bool verbose);
  ~CudaSolver() override;
  CudaSolver(CudaSolver const&) = delete;
  CudaSolver& operator=(CudaSolver const&) = delete;

  std::shared_ptr<detail::IVar> create_var(Solver const& solver, Var::Type const& type,
                                           std::optional<double> const& lb,
                                           std::optional<double> const& ub,
                                           std::optional<std::string> const& name) override;
  std::shared_ptr<detail::IConstr> create_constr(Constr::Type const& type, Expr const& e,
                                                 std::optional<std::string> const& name) override;
  std::shared_ptr<detail::IIndicatorConstr> create_indicator_constr(
    Constr const& implicant, Constr const& implicand, std::optional<std::string> const& name) override;

  void set_objective(Solver::Sense const& sense, Expr const& e) override;
  double get_objective_value() const override;
  Solver::Sense get_objective_sense() const override { return m_sense; }

  void add(Constr const& constr) override;
  void add(IndicatorConstr const& constr) override;
  void remove(Constr const& constr) override;
  void add_lazy_constr_handler(LazyConstrHandler const&, bool) override;

  std::pair<Solver::Result, bool> solve() override;

  void set_non_convex_policy(Solver::NonConvexPolicy) override {}
  void set_int_feasibility_tolerance(double value) override { m_int_tol = value; }
  void set_feasibility_tolerance(double value) override { m_feas_tol = value; }
  void set_epsilon(double value) override { m_epsilon = value; }
  void set_nr_threads(std::size_t n) override;       // = number of GPUs to use
  double get_int_feasibility_tolerance() const override { return m_int_tol; }
  double get_feasibility_tolerance() const override { return m_feas_tol; }
  double get_epsilon() const override { return m_epsilon; }

  bool supports_indicator_constraint(IndicatorConstr const&) const override { return false; }
  bool supports_quadratic_constraints() const override { return false; }
  bool supports_quadratic_objective() const override { return false; }

  double infinity() const override { return 1e20; }   // = MIPCU_INF = SCIP's sentinel
  void set_time_limit(double secs) override { m_time_limit = secs; }
  void dump(std::string const& filename) const override;
  void set_warm_start(PartialSolution const& partial_solution) override;
  void set_reoptimizing(bool) override {}
  void setup_reoptimization() override {}

  static std::string backend_info();
  static bool is_available();

  // the backend object behind a facade handle (throws if the Solver uses another backend)
  static CudaSolver& of(Solver const& solver);

  // Bulk ingress: n new columns and m new rows from CSR arrays, without
  // building Var / Expr / Constr objects.  col_idx counts from the first new column.
  void load_csr(Solver const& solver, std::int64_t m, std::int64_t n, std::int64_t const* row_ptr,
                std::int32_t const* col_idx, double const* val, double const* c, double const* lb,
                double const* ub, std::int32_t const* var_type, double const* row_lo,
                double const* row_hi, bool maximize);

  // Free-format MPS file -> n new columns and m new rows (mps.cpp); the counterpart of dump().
  // Column names are kept (dump() writes them back), row names are not.
  void read_mps(Solver const& solver, std::string const& path);

  // ---- model store (host) --------------------------------------------------------------------
  // 24 bytes and trivially copyable: loops over the columns of a multi-million-column model stream
  // 3 doubles per column, and the bulk ingress fills new columns from several threads.  Names are
  // rare on such models and live in a side map.
  struct Column
  {
    Var::Type type;
    double lb, ub;
  };
  BigVec<Column> m_columns;
  std::unordered_map<std::size_t, std::string> m_column_names;
  std::optional<std::string> column_name(std::size_t column) const
  {
    auto it = m_column_names.find(column);
    if (it == m_column_names.end()) return std::nullopt;
    return it->second;
  }
  // rows in posting order; row r owns entries [m_row_start[r], m_row_start[r+1]) sorted by column
  // (BigVec: the large buffers are recycled between solver objects, bigvec.hpp)
  BigVec<std::int64_t> m_row_start{0};
  BigVec<std::int32_t> m_row_idx;
  BigVec<double> m_row_val;
  std::vector<double> m_row_lo, m_row_hi;
  std::vector<char> m_row_alive;
  std::vector<double> m_objective;              // dense, minimise/maximise per m_sense
  Solver::Sense m_sense = Solver::Sense::Minimize;

  // ---- results -------------------------------------------------------------------------------
  std::vector<double> m_solution;
  bool m_has_solution = false;
  double m_objective_value = std::numeric_limits<double>::quiet_NaN();

  // ---- statistics of the last solve (tests and bench read them) --------------------------------
  struct SolveInfo
  {
    std::int64_t lp_solves = 0, lp_iterations = 0, nodes = 0, kernel_launches = 0, cuts = 0, lazy_rows = 0;
    bool polished = false;   // the continuous optimum was snapped to an exact vertex (polish_vertex)
    double seconds = 0, best_bound = 0;
    bool proven_optimal = false;
    bool bound_is_open_node = false;   // interrupted tree: best_bound is the smallest bound among the open nodes
  };
  SolveInfo m_info;

  // ---- lazy constraints (reference miplib/lazy.hpp:19-41, miplib/scip/solver.cpp:544-651) -----
  // Handlers are consulted for every candidate point the solve is about to accept (integral B&B
  // candidates, the LP optimum of a continuous model).  While one runs, Var::value() answers
  // with the candidate (ScipVar::value does the same, miplib/scip/var.cpp:69-74) and Solver::add
  // posts the row AND records it as "new during this solve" so the tree can take it in.
  std::vector<LazyConstrHandler> m_lazy_handlers;
  std::vector<double> const* m_callback_x = nullptr;   // non-null: inside a handler call
  std::vector<std::int64_t> m_lazy_new_rows;            // host rows posted from inside handlers
  bool is_in_callback() const { return m_callback_x != nullptr; }
  // true: every handler accepts x.  false: x was rejected; *added says whether a row was posted
  bool lazy_accepts(std::vector<double> const& x, bool* added);

  void column_bounds_changed(std::size_t column);
  void hint(std::size_t column, double value);   // IVar::set_hint: start value of one column
  void set_verbose(bool verbose) { m_verbose = verbose; }

  private:
  friend struct CudaBranchAndBound;
  void ensure_context();
  std::vector<mipcu_ctx*> node_contexts();   // primary + one loaded copy per extra GPU (B&B rounds)
  void upload_model();
  void device_holds_row_store();   // bookkeeping after a device load made straight from the row store's content
  // B&B: load the model with extra valid rows appended (cuts); the next plain solve re-uploads
  void upload_rows(std::vector<std::int64_t> const& ptr, std::vector<std::int32_t> const& idx,
                   std::vector<double> const& val, std::vector<double> const& lo,
                   std::vector<double> const& hi);
  struct LpOutcome
  {
    int status;
    double primal_obj, dual_obj;
    std::int64_t iterations;
    double rel_primal_res = 0;
  };
  LpOutcome solve_lp_on_device(std::vector<double>* x_out, std::vector<double>* y_out);
  std::pair<Solver::Result, bool> solve_continuous();
  void polish_vertex(std::vector<double>& x);   // crossover-lite, see solver.cpp
  std::pair<Solver::Result, bool> solve_integer();
  std::size_t column_of(Var const& v) const;

  mipcu_ctx* m_ctx = nullptr;
  std::vector<mipcu_ctx*> m_extra_ctx;           // devices 1.. for dealing node LPs (set_nr_threads)
  // the CSR last handed to the device (extra GPUs are loaded from it, polish_vertex reads it).  When
  // every row of the store is alive this is a VIEW of the row store itself - no second copy of a
  // 32 M-nonzero model on the host; otherwise it points into the compacted m_last_* vectors.
  struct UploadedCsr
  {
    std::int64_t const* ptr = nullptr;
    std::int32_t const* idx = nullptr;
    double const* val = nullptr;
    double const* lo = nullptr;
    double const* hi = nullptr;
    std::size_t m = 0, nnz = 0;
  };
  UploadedCsr m_up;
  void view_last();                              // m_up <- the m_last_* vectors
  std::vector<std::int64_t> m_last_ptr;
  std::vector<std::int32_t> m_last_idx;
  std::vector<double> m_last_val, m_last_lo, m_last_hi;
  bool m_verbose;
  bool m_structure_dirty = true, m_objective_dirty = true;
  std::vector<std::int64_t> m_dirty_columns;
  std::vector<std::int64_t> m_device_row_of;     // device row -> host row (dead rows are skipped)
  std::vector<double> m_start_x;                 // warm start (empty: none)
  double m_feas_tol = 1e-6, m_int_tol = 1e-6, m_epsilon = 1e-9;
  double m_time_limit = 0;
  std::size_t m_nr_gpus = 1;
  public:
  std::size_t m_polish_limit = 400;              // polish_vertex: most free columns it will factor (0: off)
};

}  // namespace miplib
