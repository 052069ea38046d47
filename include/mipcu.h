/* mipcu.h - thin C ABI of libmiplib_cuda.so, the sm_100a LP-relaxation engine behind
 * miplib's Solver::Backend::Cuda.
 *
 * This is the boundary a backend adapter binds ("B3").  In the reference,
 * each adapter forwards ISolver::solve() to one opaque third-party call:
 *     SCIPsolve(p_env)            reference miplib/scip/solver.cpp:370
 *     ::solve(p_lprec)            reference miplib/lpsolve/solver.cpp:156
 * after having lowered every row with
 *     SCIPcreateConsBasicLinear   reference miplib/scip/solver.cpp:108-120   (lhs/rhs form)
 *     add_constraintex            reference miplib/lpsolve/solver.cpp:115-130 (type 1/3, rhs=-k)
 *     set_obj_fnex / SCIPchgVarObj  miplib/lpsolve/solver.cpp:73, miplib/scip/solver.cpp:155
 * The functions below replace exactly those vendor entry points for the Cuda backend
 * (miplib_b200/miplib/cuda/solver.cpp is the adapter that calls them).
 *
 * Conventions (same as the reference's own C API, miplib/capi.cpp:7-28): plain pointers and
 * sizes, int return code 0 = ok / 1 = error, message via mipcu_last_error(); the caller owns
 * every host array (the context copies to the device); calls are synchronous on return; a
 * context is not thread-safe.  There is no CPU fallback: every entry point fails with rc=1 if
 * no sm_100 device is usable.
 *
 * Canonical problem:   min|max c'x   s.t.  row_lo <= A x <= row_hi,   lb <= x <= ub
 * with +-MIPCU_INF (or IEEE +-inf, or |v| >= 1e20) meaning "no bound".
 */
#ifndef MIPCU_H
#define MIPCU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIPCU_INF 1e20   /* same sentinel SCIP uses: reference miplib/scip/solver.cpp:463-466 */

typedef struct mipcu_ctx mipcu_ctx;

/* status codes in mipcu_stats.status; they map 1:1 onto Solver::Result
 * (reference miplib/solver.hpp:24-32, status switch miplib/lpsolve/solver.cpp:171-199) */
enum {
  MIPCU_OPTIMAL = 0,
  MIPCU_INFEASIBLE = 1,
  MIPCU_INFEASIBLE_OR_UNBOUNDED = 2,
  MIPCU_UNBOUNDED = 3,
  MIPCU_ITERATION_LIMIT = 4,   /* -> Result::Interrupted */
  MIPCU_TIME_LIMIT = 5,        /* -> Result::Interrupted */
  MIPCU_NUMERICAL_ERROR = 6,   /* -> Result::Error */
  MIPCU_CUTOFF = 7,            /* dual bound crossed MIPCU_PARAM_OBJ_CUTOFF (B&B pruning) */
  MIPCU_BRANCH = 8             /* node LP stopped early (mipcu_solve_lp_batch_warm, branch_tol): the iterate is
                                  feasible to branch_tol and its objective is below the cutoff by more than the
                                  error it can still have, so the node cannot be pruned by bound - branch on x.
                                  dual_obj is a valid (weaker) bound */
};

/* keys for mipcu_set_param / mipcu_get_param */
enum {
  MIPCU_PARAM_EPS_REL = 0,        /* relative KKT tolerance, default 1e-6 (set_feasibility_tolerance) */
  MIPCU_PARAM_ITER_LIMIT = 1,     /* default 2e6 */
  MIPCU_PARAM_TIME_LIMIT = 2,     /* seconds, default none (set_time_limit) */
  MIPCU_PARAM_CHECK_PERIOD = 3,   /* iterations per restart/KKT check block, default 64, >= 2 */
  MIPCU_PARAM_VERBOSE = 4,        /* 1: PDLP-style iteration log on stderr */
  MIPCU_PARAM_STEP_SIZE = 5,      /* >0 pins eta (skips the power iteration); 0 = automatic */
  MIPCU_PARAM_USE_GRAPH = 6,      /* -1 auto (graph for small problems), 0 never, 1 always */
  MIPCU_PARAM_SCALING = 7,        /* 1 (default): 10x Ruiz + Pock-Chambolle; 0: none */
  MIPCU_PARAM_THREADS_PER_ROW = 8,/* 0 auto; else 2,4,8,16,32 sub-warp width for both SpMVs */
  MIPCU_PARAM_OBJ_CUTOFF = 9,     /* stop with MIPCU_CUTOFF once the dual bound proves the optimum is no
                                     better than this objective value (model's own sense); NaN = off */
  MIPCU_PARAM_USE_TMA = 10,       /* kernel variant: -1 auto (default), 0: sub-warp per row, 1: TMA-streamed tiles,
                                     2: CSR-stream (products staged in shared memory) */
  MIPCU_PARAM_SHARD_EXCHANGE = 11,/* row-sharded contexts: 2 (default) block-owner exchange - every rank streams its row
                                     block of K and its column block of K', slices of yhat / xbar pushed into peer
                                     memory from the kernels' epilogues; 1: fused reduce-scatter + primal step +
                                     all-gather kernel over NVLink peer memory; 0: ncclAllReduce. Set before load */
  MIPCU_PARAM_USE_PDL = 12,       /* 1: the two iteration kernels are launched as programmatic dependents of each
                                     other (the head of the nonzero stream is fetched while the previous kernel drains);
                                     0: plain stream order.  Ignored when the block is replayed as a CUDA graph */
  MIPCU_PARAM_MULTICAST = 13,     /* block-owner exchange: 1 (default) push slices through an NVSwitch multicast window
                                     when the box supports it; 0: unicast stores into every peer.  Set before load;
                                     after load it reads back 1 only if the window is in use on every rank */
  MIPCU_PARAM_ROW_SORT = 14,      /* load-time ordering of the device copies, invisible through this ABI (every vector is
                                     permuted at the boundary).  4 (default): as 3, with a second (columns, rows) round.
                                     3: rows stored sorted by decreasing length (no
                                     padding in the sliced-ELL copy), columns binned by (length, first row), so that the
                                     32 columns of a slice start in neighbouring lines of the gathered vector, and the rows
                                     of equal length re-sorted by their first (new) column for the same reason (row-sharded
                                     contexts keep the caller's column numbering, i.e. behave as 1); 2: without the second
                                     row sort; 1: rows by length only; 0: the caller's order.  Set before load */
  MIPCU_PARAM_PEER_TIMEOUT = 15,  /* row-sharded contexts: seconds a kernel waits at a cross-GPU barrier for its peers
                                     (default 120; 0 = forever).  On expiry the solve returns rc = 1 with a message
                                     instead of hanging or trapping: sharded load / solve are collective and fail-stop */
  MIPCU_PARAM_VALIDATE = 16,      /* debug: 1 = after setup, device kernels check every index structure the iteration kernels
                                     dereference (CSR / sliced-ELL column indices in range, slice offsets monotone and
                                     consistent with the entry count, block-local permutations, the column block and the
                                     slice extents of row-sharded contexts); a violation fails the call with a message
                                     (2 = self-test: plants a bad index first, so the call must fail).
                                     The in-house stand-in for compute-sanitizer memcheck, which this pool does not offer */
  MIPCU_PARAM_COUNT_ = 17
};

typedef struct mipcu_stats {
  int32_t status;          /* MIPCU_* */
  int32_t restarts;
  int64_t iterations;
  int64_t kernel_launches; /* kernels of this library launched by the call (graph nodes count) */
  double  primal_obj;      /* in the model's own sense (max problems are not negated) */
  double  dual_obj;
  double  rel_primal_res;  /* ||Ax - proj(Ax)||_2 / (1 + ||b||_2)        (unscaled) */
  double  rel_dual_res;    /* ||reduced-cost violation||_2 / (1 + ||c||_2) (unscaled) */
  double  rel_gap;         /* |pobj - dobj| / (1 + |pobj| + |dobj|) */
  double  step_size;       /* eta */
  double  primal_weight;   /* omega at exit */
  double  solve_seconds;   /* wall time inside the call */
  double  iter_seconds;    /* CUDA-event time of the iteration loop only */
  double  setup_seconds;   /* scaling + step size estimate (first solve after a load) */
  double  row_kernel_us;   /* mean in-situ duration of the fused K.xbar / dual-step kernel */
  double  col_kernel_us;   /* mean in-situ duration of the fused K'.yhat / primal-step kernel */
  int64_t row_kernel_bytes;/* algorithmic bytes of one launch of each */
  int64_t col_kernel_bytes;
  double  col_partial_us;  /* row-sharded contexts: the local partial K_g'.yhat_g SpMV inside col_kernel_us */
  double  col_exchange_us; /* row-sharded contexts: exchange (+ primal step) inside col_kernel_us */
} mipcu_stats;

/* --- lifetime ------------------------------------------------------------------------------ */
int  mipcu_device_count(void);                         /* usable sm_100 devices; 0 if none */
int  mipcu_create(mipcu_ctx** out, int device);        /* one context = one GPU */
void mipcu_destroy(mipcu_ctx* ctx);
/* forget the loaded model and restore every parameter to its default; the device memory the model
 * held stays in the context's pool, so the next mipcu_load_lp of a similar size costs no allocation
 * (what lets an adapter recycle contexts between Solver objects).  Not for row-sharded contexts. */
int  mipcu_reset(mipcu_ctx* ctx);
const char* mipcu_last_error(const mipcu_ctx* ctx);    /* ctx may be NULL: last create() error */

/* --- row-sharded contexts (one process per GPU, one NCCL communicator) ---------------------- *
 * Rank 0 obtains an id with mipcu_dist_unique_id and ships it to the other ranks by any means
 * (torch.distributed broadcast in bench.py); every rank then calls mipcu_create_sharded.  Each
 * rank loads ITS row block [row_begin, row_end) of A with mipcu_load_lp (m = local rows, n and
 * the column vectors global) - SURVEY.md section 8e. */
#define MIPCU_UNIQUE_ID_BYTES 128
int  mipcu_dist_unique_id(void* id_out /* MIPCU_UNIQUE_ID_BYTES */);
int  mipcu_create_sharded(mipcu_ctx** out, int device, int rank, int world,
                          const void* unique_id /* MIPCU_UNIQUE_ID_BYTES */);

/* --- model ---------------------------------------------------------------------------------- */
/* CSR of A with column indices sorted inside each row and no duplicates (the adapter sorts:
 * the reference's Expr::linear_vars() order is pointer-hash order, miplib/expr.cpp:444-467). */
int  mipcu_load_lp(mipcu_ctx* ctx, int64_t m, int64_t n, int64_t nnz,
                   const int64_t* row_ptr /* m+1 */, const int32_t* col_idx /* nnz */,
                   const double* val /* nnz */,
                   const double* c /* n */, const double* lb /* n */, const double* ub /* n */,
                   const double* row_lo /* m */, const double* row_hi /* m */, int maximize);
/* change bounds of some columns (IVar::set_lb / set_ub, reference miplib/var.hpp:98-101; B&B) */
int  mipcu_set_bounds(mipcu_ctx* ctx, int64_t n_changed, const int64_t* cols,
                      const double* lb, const double* ub);
int  mipcu_set_objective(mipcu_ctx* ctx, const double* c /* n */, int maximize);
int  mipcu_set_param(mipcu_ctx* ctx, int key, double value);
int  mipcu_get_param(const mipcu_ctx* ctx, int key, double* value);
/* start point for the next solve (set_warm_start, reference miplib/solver.hpp:94); NULL = zeros */
int  mipcu_warm_start(mipcu_ctx* ctx, const double* x /* n or NULL */, const double* y /* m or NULL */);

/* --- solve ---------------------------------------------------------------------------------- */
int  mipcu_solve_lp(mipcu_ctx* ctx, mipcu_stats* out);
/* B node LPs that share A, c and the row sides but differ in column bounds; bounds are
 * [B][n] row-major on the host.  cutoff ([B], model's own sense, NaN = none) may be NULL: then
 * MIPCU_PARAM_OBJ_CUTOFF applies to every node.  x_out ([B][n]) may be NULL.  All nodes start
 * from x = clamp(0), y = 0 (mipcu_solve_lp_batch_warm takes start points) and run to
 * MIPCU_PARAM_EPS_REL / ITER_LIMIT / TIME_LIMIT. */
int  mipcu_solve_lp_batch(mipcu_ctx* ctx, int32_t B, const double* lb, const double* ub,
                          const double* cutoff, mipcu_stats* out /* [B] */, double* x_out);
/* The same with a start point per node - in a branch and bound, the parent's LP solution (what
 * lp_solve's and SCIP's tree do with the parent basis behind ::solve / SCIPsolve, reference
 * miplib/lpsolve/solver.cpp:156, miplib/scip/solver.cpp:370): x0 [B][n] and y0 [B][m] in the caller's
 * terms (x0 is clamped into the node's box; all-zero rows = cold start), primal_weight0 [B] the
 * mipcu_stats.primal_weight the parent's solve ended with (<= 0: default).  Any of the three may be
 * NULL.  branch_tol [B] (or NULL = all 0): a node with branch_tol > 0 stops with MIPCU_BRANCH as soon
 * as its relative primal / dual residual and gap are <= branch_tol and its primal objective is better
 * than the node's cutoff by more than 2 branch_tol (1 + |pobj| + |dobj|) - such a node cannot be pruned
 * by bound however accurately it is solved, so a branch and bound need not wait for 1e-6 (no cutoff:
 * any branch_tol-accurate point qualifies).  y_out ([B][m]) may be NULL.  stats[b].row/col_kernel_us
 * and _bytes describe the batched kernels with all B problems in flight.
 * A node's result (x, duals, iterations, status) does not depend on B or on which other nodes share
 * the call: the kernels pick their launch shape from the number of nodes still running, but every
 * node's sums are formed in one order and every rounding of its step is pinned (bit-identical). */
int  mipcu_solve_lp_batch_warm(mipcu_ctx* ctx, int32_t B, const double* lb, const double* ub,
                               const double* cutoff, const double* x0, const double* y0,
                               const double* primal_weight0, const double* branch_tol,
                               mipcu_stats* out /* [B] */, double* x_out, double* y_out);
int  mipcu_get_x(mipcu_ctx* ctx, double* x /* n */);
int  mipcu_get_y(mipcu_ctx* ctx, double* y /* m (local rows when sharded) */);

/* --- test / bench hooks ----------------------------------------------------------------------*/
/* out = A v (transpose = 0, v has n entries) or A' v (transpose = 1, v has m entries), on the
 * UNSCALED matrix as loaded, through the same SpMV kernel the solver uses. */
int  mipcu_spmv(mipcu_ctx* ctx, int transpose, const double* v, double* out);
/* diag scalings after the first solve / mipcu_prepare: d1 (m), d2 (n); K = D1 A D2 */
int  mipcu_prepare(mipcu_ctx* ctx);
int  mipcu_get_scaling(mipcu_ctx* ctx, double* d1 /* m */, double* d2 /* n */);
/* run `iters` PDHG iterations from the current start point with no restart / termination logic
 * and return the scaled iterates (xbar: n, y: m) - iterate-level parity against oracle/pdhg_ref */
int  mipcu_debug_iterate(mipcu_ctx* ctx, int64_t iters, double* xbar_out, double* y_out);
/* time `reps` back-to-back launches of one device-resident kernel with CUDA events on the
 * launching stream; which: 0 = plain K v, 1 = plain K' v, 2 = fused row step, 3 = fused col step,
 * 4 = one whole iteration as shipped (row step + col step), 5 = the same iteration with what an adaptive
 * step-size rule adds (per-iteration reductions in both epilogues + a single-CTA controller launch).
 * flush_l2 != 0 writes a 256 MB buffer before every launch (outside the timed events). */
int  mipcu_time_kernel(mipcu_ctx* ctx, int which, int reps, int flush_l2,
                       double* mean_us, int64_t* algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MIPCU_H */
