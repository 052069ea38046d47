// internal.h - context layout and helpers shared by the translation units of libmiplib_cuda.so.
// Nothing in here is part of the ABI (see include/mipcu.h for that).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached

#include "../../include/mipcu.h"
#include "params.h"

namespace mipcu {

// NVTX range over a host-side phase (load, setup, iteration block, KKT pass, node-LP round) so that
// Nsight Systems timelines of a host application show where the backend spends its time.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define MIPCU_CK(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess)                                                                     \
      throw ::mipcu::Error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" +        \
                           __FILE__ + ":" + std::to_string(__LINE__) + ")");                   \
  } while (0)

#define MIPCU_REQUIRE(cond, msg)                                                               \
  do {                                                                                         \
    if (!(cond)) throw ::mipcu::Error(std::string(msg));                                       \
  } while (0)

constexpr int kThreads = 256;        // CTA size of every kernel in this library
constexpr int kRowsPerCta = 256;     // rows (or columns) one CTA of a fused SpMV kernel owns

// One sparse matrix in CSR layout on the device.  The context keeps two: K by rows ("rows")
// and K' by rows, i.e. K by columns ("cols").  int32 indices: nnz < 2^31 is enforced at load.
struct Csr {
  int nrows = 0, ncols = 0;
  int64_t nnz = 0;
  int* ptr = nullptr;      // nrows + 1
  int* idx = nullptr;      // nnz (+ padding)
  double* val = nullptr;   // nnz (+ padding)
  int tpr = 8;             // threads per row chosen for this matrix (CSR kernels)
  // sliced-ELL copy of the SCALED matrix for the lane-per-row kernels (setup.cu:build_sell);
  // null when the matrix is not eligible (long or badly unbalanced rows)
  int* sptr = nullptr;              // 8 * nblk + 1 slice offsets (32 rows per slice)
  int* sidx = nullptr;              // padded entries, -1 = padding
  double* sval = nullptr;
  unsigned char* perm = nullptr;    // 256 * nblk: sorted position -> CTA-local row
  int64_t sell_entries = 0;         // padded entry count
  bool sell = false;                // SELL copy present and selected
  int tile_cap = 0;        // TMA kernels: entries one 64-row tile needs in shared memory (0: n/a)
  int stream_cap = 0;      // CSR-stream kernels: most nonzeros in one 256-row tile
};

// Quantities reduced per CTA on check iterations: partials[q * max_blocks + blockIdx.x].
enum Quantity {
  Q_DY2 = 0,     // sum (yhat - y)^2                      row step
  Q_DY0,         // sum (yhat - y0)^2                     row step
  Q_DOBJ_ROW,    // sum lo yhat+ - hi yhat-               row step
  Q_PRES2,       // sum ((K xhat - proj)/d1)^2            kkt row pass
  Q_DX2,         // sum (xhat - x)^2                      col step
  Q_DXDATY,      // sum (xhat - x)(K'yhat - K'y)          col step
  Q_DX0,         // sum (xhat - x0)^2                     col step
  Q_DRES2,       // sum (reduced-cost violation / d2)^2   col step
  Q_POBJ,        // sum c xhat                            col step
  Q_DOBJ_COL,    // sum l r+ - u r-                       col step
  // certificates of infeasibility from the displacement T(z) - z (dy = yhat - y, dx = xhat - x)
  Q_RAYD_ROW,    // sum lo dy+ - hi dy-                    row step   (Farkas value, row part)
  Q_RAYD_ROWV,   // sum (dy+ where lo = -inf, dy- where hi = +inf)^2
  Q_RAYD_COL,    // sum l r+ - u r-,  r = -K'dy            col step   (Farkas value, column part)
  Q_RAYD_COLV,   // sum (r+ where l = -inf, r- where u = +inf)^2
  Q_RAYP_C,      // sum c dx                               col step   (objective along the primal ray)
  Q_RAYP_COLV,   // sum (dx- where l finite, dx+ where u finite)^2
  Q_RAYP_ROWV,   // sum ((K dx)- where lo finite, (K dx)+ where hi finite)^2     kkt row pass
  Q_COUNT
};

// Device-resident control block.  The iteration kernels read it, the single-CTA controller
// kernels write it, the host only ever copies it out (no host round trip inside a block).
struct Scalars {
  double tau, sigma, eta, w;
  double eps;
  double cutoff;        // minimisation-sense objective cutoff for the dual bound (+inf: off)
  double branch_tol;    // batched node LPs: > 0 allows the early MIPCU_BRANCH stop at this tolerance (batch.cu)
  double bnorm, cnorm;
  double fp0, fp_prev, fp;
  double rel_p, rel_d, rel_g, pobj, dobj;
  long long k_base;     // iterations since the last restart, at the start of the current block
  long long total;      // iterations done, at the start of the current block
  int period;
  int done;             // 0 running; 1 + MIPCU_* status once finished
  int restart_flag;     // set by the controller, consumed by k_restart_apply
  int ray_hits_primal, ray_hits_dual;   // consecutive checks on which a certificate test passed
  int restarts;
  int nblk_row, nblk_col, max_blocks;
  int fault;            // set by a kernel whose cross-GPU barrier timed out (p2p.cu:own_barrier); the host reports it
  double ray_dual_value, ray_primal_value;   // last normalised Farkas value / c.dx (diagnostics)
};

// Row-sharded contexts: the peers' copies of the exchange buffers, mapped with CUDA IPC (p2p.cu)
constexpr int kMaxPeers = 16;
// flag block of a rank (u64 each): [0, W) ready, [W, 2W) done, [2W] CTA counter, [2W+1] epoch of the
// fused-exchange kernels; [kOwnFlagBase, kOwnFlagBase + W) arrival flags of the block-owner kernels
constexpr int kOwnFlagBase = 2 * kMaxPeers + 2;
constexpr int kFlagCount = 3 * kMaxPeers + 2;
struct PeerTable {
  int world = 0, rank = 0;
  int j0[kMaxPeers + 1] = {};               // column slice of rank g: [j0[g], j0[g+1])
  double* partial[kMaxPeers] = {};          // K_g' yhat_g of every rank
  double* xbar[kMaxPeers] = {};
  double* xhat_chk[kMaxPeers] = {};
  double* x_sol[kMaxPeers] = {};
  unsigned long long* flags[kMaxPeers] = {};
  // block-owner exchange (MIPCU_PARAM_SHARD_EXCHANGE = 2): every rank keeps the WHOLE yhat, each
  // rank's row kernel pushes its slice [i0[g], i0[g+1]) into all of them
  int i0[kMaxPeers + 1] = {};
  double* yhat_full[kMaxPeers] = {};
  // NVSwitch multicast mappings of xbar / yhat_full (mc.cu): one store reaches every rank's copy,
  // this rank's included; null = unicast pushes through the arrays above
  double* mc_xbar = nullptr;
  double* mc_yhat_full = nullptr;
  unsigned long long barrier_timeout_ns = 0;   // how long a block-owner kernel waits for its peers (0: forever)
};

// mc.cu: the window [xbar | yhat_full] of this rank, bound to the job's multicast object
struct McWindow {
  bool active = false, bound = false;
  size_t size = 0;
  unsigned long long mc_handle = 0, mem_handle = 0;   // CUmemGenericAllocationHandle
  void* uc_va = nullptr;                              // this rank's window (ordinary mapping)
  void* mc_va = nullptr;                              // multicast mapping of all ranks' windows
  double* saved_xbar = nullptr;                       // the cudaMalloc'd vectors the window replaces
  double* saved_yhat_full = nullptr;
};

}  // namespace mipcu

struct mipcu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::string err;
  double params[MIPCU_PARAM_COUNT_];

  // sharding (world == 1: plain single-GPU context)
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  double* row_sums = nullptr;     // 32 doubles: all-reduced row-side scalars of check iterations
  int64_t collectives = 0;        // NCCL calls issued by the current call
  mipcu::PeerTable peers;         // world == peers.world: the fused P2P exchange is usable
  std::vector<void*> peer_mappings;
  unsigned long long* p2p_flags = nullptr;
  mipcu::McWindow mc;
  // a re-load of a model of the same shape keeps the exported vectors, their peer mappings and the
  // multicast window (api.cu: mipcu_load_lp); set only while that load runs
  bool keep_exchange = false;
  double exchange_at_setup = -1.0, multicast_at_setup = -1.0;   // parameter values the mappings were built for
  bool own_ready = false;             // block-owner exchange: column block built, yhat_full mapped
  bool use_pdl = false;               // iteration kernels launched as programmatic dependents (MIPCU_PDL)
  int own_blocks_per_cta = 0;         // block-owner kernels: 256-row blocks one CTA works through (0 = automatic)
  int own_variant = 0;                // laboratory switches (p2p.cu), 0 = shipped protocol
  unsigned long long own_epoch = 0;   // barrier generation of the block-owner kernels (same on all ranks)

  // model as loaded (unscaled), device resident
  int m = 0, n = 0;
  int64_t nnz = 0;
  bool loaded = false, maximize = false;
  mipcu::Csr rows, cols;                 // scaled IN PLACE by prepare()
  // row-sharded, block-owner exchange: columns [j0, j1) of the WHOLE matrix (CSR of K'[j0:j1, :],
  // global row indices), assembled from all ranks' blocks at load; yhat_full: m_glob doubles
  mipcu::Csr colblk;
  double* yhat_full = nullptr;
  int m_glob = 0;
  int row_off[mipcu::kMaxPeers + 1] = {};
  // rows are kept sorted by decreasing length on the device (setup.cu:sort_rows_by_length): internal
  // row i is the caller's row row_perm[i].  Null = identity.  Every m-vector that crosses the ABI
  // (row sides, warm-start / solution duals, the test hooks) is permuted at the boundary.
  int* row_perm = nullptr;
  std::vector<int> h_row_perm;   // host copy, fetched on first use by the test hooks
  // columns binned at load (setup.cu:bin_columns, single-GPU contexts): internal column j is the caller's
  // column col_perm[j], col_inv is the inverse.  Null = identity.  Every n-vector that crosses the ABI
  // (objective, boxes, start point, solution, the test hooks) is permuted at the boundary.
  int *col_perm = nullptr, *col_inv = nullptr;
  std::vector<int> h_col_perm;
  double *c = nullptr, *lb = nullptr, *ub = nullptr, *lo = nullptr, *hi = nullptr;
  // scaled copies + diagonal scalings
  double *cs = nullptr, *ls = nullptr, *us = nullptr, *los = nullptr, *his = nullptr;
  double *d1 = nullptr, *d2 = nullptr;
  bool prepared = false;        // matrix scaled, eta known
  bool vectors_dirty = true;    // scaled vectors / norms must be refreshed (bounds or c changed)
  double eta = 0, w0 = 1, bnorm = 0, cnorm = 0;
  double setup_seconds = 0;

  // iterate buffers (scaled space)
  double *xbar = nullptr, *x0 = nullptr, *aty = nullptr, *aty0 = nullptr, *xhat_first = nullptr,
         *xhat_chk = nullptr, *aty_cand = nullptr;             // n each
  double *y = nullptr, *y0 = nullptr, *yhat = nullptr;         // m each
  double *tmp_n = nullptr, *tmp_m = nullptr;                   // scratch
  double *x_start = nullptr, *y_start = nullptr;               // warm start (unscaled) or null
  double *x_sol = nullptr, *y_sol = nullptr;                   // last solution (unscaled)
  bool has_solution = false;

  double* partials = nullptr;
  int max_blocks = 0;
  mipcu::Scalars* d_sc = nullptr;
  mipcu::Scalars* h_sc = nullptr;   // pinned, 2 slots

  cudaGraphExec_t graph_exec = nullptr;
  int graph_period = 0;
  cudaEvent_t ev[12] = {};
  void* flush_buf = nullptr;

  int64_t launches = 0;   // kernels launched by the current call
  cudaEvent_t split_event = nullptr;   // sharded col step: recorded between the partial SpMV and the exchange
  bool pool_alloc = false;   // model arrays come from the stream-ordered memory pool below
  void* pool = nullptr;      // cudaMemPool_t private to this context (api.cu: create_on)
  std::vector<void*> raw_allocs;   // arrays that had to come from cudaMalloc (exported over CUDA IPC)
};

// context whose call is executing on this thread: tells setup.cu's allocator which stream / pool
// to use (set by the entry points in api.cu)
extern thread_local const mipcu_ctx* g_alloc_ctx;

namespace mipcu {

// ---- setup.cu -------------------------------------------------------------------------------
void free_model(mipcu_ctx* ctx);
void load_lp(mipcu_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
             const int32_t* col_idx, const double* val, const double* c, const double* lb,
             const double* ub, const double* row_lo, const double* row_hi, int maximize);
void prepare(mipcu_ctx* ctx);           // scaling + step size (once per load)
void refresh_vectors(mipcu_ctx* ctx);   // scaled c / bounds / norms / w0 (after bound changes)
void spmv_plain(mipcu_ctx* ctx, const Csr& A, const double* v, double* out);
double device_norm2(mipcu_ctx* ctx, const double* v, int64_t len);
int choose_tpr(double avg_row_len, int override_tpr);
int kernel_format(const mipcu_ctx* ctx, const Csr& A);   // 0: SELL lane-per-row, else CSR lanes per row
void set_bounds(mipcu_ctx* ctx, int64_t n_changed, const int64_t* cols, const double* lb,
                const double* ub);
void spmv_host(mipcu_ctx* ctx, int transpose, const double* v, double* out);
void build_col_block(mipcu_ctx* ctx);   // collective: the column block of the block-owner exchange
const std::vector<int>& host_row_perm(mipcu_ctx* ctx);   // empty = identity
void cols_to_caller_order(mipcu_ctx* ctx, double* v);    // host n-vector, internal -> caller's column order
void cols_to_internal_order(mipcu_ctx* ctx, double* v);  // host n-vector, caller's -> internal column order
void rows_to_caller_order(mipcu_ctx* ctx, double* v);    // host m-vector, internal -> caller's row order
void rows_to_internal_order(mipcu_ctx* ctx, double* v);  // host m-vector, caller's -> internal row order

// ---- dist.cu --------------------------------------------------------------------------------
void dist_allreduce(mipcu_ctx* ctx, double* buf, size_t count, bool max_instead_of_sum);
double dist_sum_scalar(mipcu_ctx* ctx, double v);
void dist_allgather_bytes(mipcu_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes_per_rank);

// ---- p2p.cu ---------------------------------------------------------------------------------
struct ColStepArgs;
bool p2p_active(const mipcu_ctx* ctx);
bool own_active(const mipcu_ctx* ctx);   // block-owner exchange in use
struct RowStepArgs;
void own_setup(mipcu_ctx* ctx);          // collective, after p2p_setup: builds the column block
void own_rebuild(mipcu_ctx* ctx);        // collective: new column block for a re-loaded model, mappings kept
void own_row_step(mipcu_ctx* ctx, const RowStepArgs& a, bool check);
void own_col_step(mipcu_ctx* ctx, int iter_in_block, bool check, const double* xhat_in, double* xhat_out);
void p2p_setup(mipcu_ctx* ctx);
void p2p_teardown(mipcu_ctx* ctx);
void p2p_primal_step(mipcu_ctx* ctx, const ColStepArgs& a, bool check, bool store);
void p2p_allgather(mipcu_ctx* ctx, int which);   // 0 xbar, 1 xhat_chk, 2 x_sol

// ---- mc.cu ----------------------------------------------------------------------------------
bool mc_setup(mipcu_ctx* ctx);      // collective: xbar / yhat_full moved into an NVSwitch multicast window
void mc_teardown(mipcu_ctx* ctx);

// ---- solve.cu -------------------------------------------------------------------------------
void solve_lp(mipcu_ctx* ctx, mipcu_stats* out);
void debug_iterate(mipcu_ctx* ctx, int64_t iters, double* xbar_out, double* y_out);
void time_kernel(mipcu_ctx* ctx, int which, int reps, int flush_l2, double* mean_us,
                 int64_t* bytes);
int64_t row_kernel_bytes(const mipcu_ctx* ctx);
int64_t col_kernel_bytes(const mipcu_ctx* ctx);

}  // namespace mipcu
