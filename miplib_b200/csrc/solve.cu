// solve.cu - the iteration driver: blocks of `period` PDHG iterations enqueued (or replayed as a
// CUDA graph) with a device-side controller at the end of every block; the host only looks at a
// copy of the control block, one block behind the GPU, to decide termination.
// Restated on the CPU by oracle/pdhg_ref.py: solve().
#include <chrono>
#include <cmath>
#include <unordered_map>
#include <utility>

#include "internal.h"
#include "kernels.cuh"
#include "kernels_stream.cuh"
#include "kernels_tma.cuh"

namespace mipcu {

namespace {

inline int grid_for(int64_t n, int per_block = kThreads) {
  return (int)((n + per_block - 1) / per_block);
}

// deterministic sum of one partial array by a full CTA (fixed strided order + fixed tree)
__device__ double cta_sum_partials(const double* p, int count) {
  __shared__ double s[kThreads];
  double a = 0.0;
  for (int i = threadIdx.x; i < count; i += kThreads) a += p[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  double r = s[0];
  __syncthreads();
  return r;
}

__device__ double fixed_point_error(const Scalars* sc, double dx2, double dxdaty, double dy2) {
  double v = (sc->w / sc->eta) * dx2 - 2.0 * dxdaty + dy2 / (sc->eta * sc->w);
  return sqrt(fmax(v, 0.0));
}

// after the FIRST iteration of a block: records the fixed-point error right after a restart
// row-sharded: this rank's sums, to be all-reduced before the controllers read them.  Row-side
// quantities are always partial; column-side ones only when the primal step is sharded too (p2p).
__global__ void __launch_bounds__(kThreads) k_sum_partials(const Scalars* sc, const double* partials,
                                                           double* sums, int include_cols) {
  if (sc->done) return;
  const int mb = sc->max_blocks;
  for (int i = 0; i < Q_COUNT; ++i) {
    const bool is_row = (i == Q_DY2 || i == Q_DY0 || i == Q_DOBJ_ROW || i == Q_PRES2 ||
                         i == Q_RAYD_ROW || i == Q_RAYD_ROWV || i == Q_RAYP_ROWV);
    double v = 0.0;
    if (is_row || include_cols)
      v = cta_sum_partials(partials + (size_t)i * mb, is_row ? sc->nblk_row : sc->nblk_col);
    if (threadIdx.x == 0) sums[i] = v;
  }
}

__global__ void __launch_bounds__(kThreads) k_ctrl_first(Scalars* sc, const double* partials,
                                                         const double* row_sums, int cols_from_sums) {
  if (sc->done) return;
  const int mb = sc->max_blocks;
  double dy2 = row_sums ? row_sums[Q_DY2] : cta_sum_partials(partials + (size_t)Q_DY2 * mb, sc->nblk_row);
  double dx2 = cols_from_sums ? row_sums[Q_DX2] : cta_sum_partials(partials + (size_t)Q_DX2 * mb, sc->nblk_col);
  double dxd = cols_from_sums ? row_sums[Q_DXDATY]
                              : cta_sum_partials(partials + (size_t)Q_DXDATY * mb, sc->nblk_col);
  if (threadIdx.x == 0 && sc->k_base == 0) {
    double fp = fixed_point_error(sc, dx2, dxd, dy2);
    sc->fp0 = fp;
    sc->fp_prev = fp;
  }
}

// after the LAST iteration of a block: KKT test, restart test, primal-weight update
__global__ void __launch_bounds__(kThreads) k_ctrl_check(Scalars* sc, const double* partials,
                                                         const double* row_sums, int cols_from_sums) {
  if (sc->done) return;
  const int mb = sc->max_blocks;
  double q[Q_COUNT];
  for (int i = 0; i < Q_COUNT; ++i) {
    const bool is_row = (i == Q_DY2 || i == Q_DY0 || i == Q_DOBJ_ROW || i == Q_PRES2 ||
                         i == Q_RAYD_ROW || i == Q_RAYD_ROWV || i == Q_RAYP_ROWV);
    if (row_sums && (is_row || cols_from_sums)) q[i] = row_sums[i];
    else q[i] = cta_sum_partials(partials + (size_t)i * mb, is_row ? sc->nblk_row : sc->nblk_col);
  }
  if (threadIdx.x != 0) return;
  const double fp = fixed_point_error(sc, q[Q_DX2], q[Q_DXDATY], q[Q_DY2]);
  const double pobj = q[Q_POBJ], dobj = q[Q_DOBJ_ROW] + q[Q_DOBJ_COL];
  sc->fp = fp;
  sc->pobj = pobj;
  sc->dobj = dobj;
  sc->rel_p = sqrt(q[Q_PRES2]) / (1.0 + sc->bnorm);
  sc->rel_d = sqrt(q[Q_DRES2]) / (1.0 + sc->cnorm);
  sc->rel_g = fabs(pobj - dobj) / (1.0 + fabs(pobj) + fabs(dobj));
  const long long k = sc->k_base + sc->period;
  sc->total += sc->period;
  sc->restart_flag = 0;
  const double worst = fmax(sc->rel_p, fmax(sc->rel_d, sc->rel_g));
  if (!(worst == worst) || isinf(worst)) {   // NaN / inf
    sc->done = 1 + MIPCU_NUMERICAL_ERROR;
    return;
  }
  if (worst <= sc->eps) {
    sc->done = 1 + MIPCU_OPTIMAL;
    return;
  }
  if (sc->rel_d <= sc->eps && dobj >= sc->cutoff) {   // a dual-feasible y bounds the optimum
    sc->done = 1 + MIPCU_CUTOFF;
    return;
  }
  // Infeasibility certificates from the displacement T(z) - z, which converges to the infimal
  // displacement vector when the LP is inconsistent.  dy is a Farkas ray if its value is positive
  // and it prices nothing through a missing bound; dx is an improving recession direction if
  // c.dx < 0 and it moves no bounded side.  A test must pass on two consecutive checks.
  {
    const double nd = sqrt(q[Q_DY2]), nx = sqrt(q[Q_DX2]);
    const double dval = nd > 1e-12 ? (q[Q_RAYD_ROW] + q[Q_RAYD_COL]) / nd : 0.0;
    const double dviol = nd > 1e-12 ? sqrt(q[Q_RAYD_ROWV] + q[Q_RAYD_COLV]) / nd : 0.0;
    const double pval = nx > 1e-12 ? q[Q_RAYP_C] / nx : 0.0;
    const double pviol = nx > 1e-12 ? sqrt(q[Q_RAYP_COLV] + q[Q_RAYP_ROWV]) / nx : 0.0;
    sc->ray_dual_value = dval;
    sc->ray_primal_value = pval;
    sc->ray_hits_dual = (dval > 1e-9 && dviol <= RAY_TOL * dval) ? sc->ray_hits_dual + 1 : 0;
    sc->ray_hits_primal = (pval < -1e-9 && pviol <= RAY_TOL * -pval) ? sc->ray_hits_primal + 1 : 0;
    if (sc->ray_hits_dual >= 2) {
      sc->done = 1 + MIPCU_INFEASIBLE;
      return;
    }
    if (sc->ray_hits_primal >= 2) {
      sc->done = 1 + (sc->rel_p <= 1e-4 ? MIPCU_UNBOUNDED : MIPCU_INFEASIBLE_OR_UNBOUNDED);
      return;
    }
  }
  const bool restart = fp <= BETA_SUFFICIENT * sc->fp0 ||
                       (fp <= BETA_NECESSARY * sc->fp0 && fp > sc->fp_prev) ||
                       (double)k >= BETA_ARTIFICIAL * (double)sc->total;
  sc->fp_prev = fp;
  if (restart) {
    const double dxn = sqrt(q[Q_DX0]), dyn = sqrt(q[Q_DY0]);
    if (dxn > WEIGHT_EPS && dyn > WEIGHT_EPS)
      sc->w = exp(WEIGHT_SMOOTHING * log(dyn / dxn) + (1.0 - WEIGHT_SMOOTHING) * log(sc->w));
    sc->tau = sc->eta / sc->w;
    sc->sigma = sc->eta * sc->w;
    sc->k_base = 0;
    sc->restart_flag = 1;
    sc->restarts += 1;
  } else {
    sc->k_base = k;
  }
}

// new anchor z0 <- (xhat_chk, yhat); re-seed the iteration from it (also used at solve start)
__global__ void k_restart_apply(const Scalars* sc, int j_begin, int n, int m, const double* __restrict__ xcand,
                                const double* __restrict__ aty_cand, const double* __restrict__ yhat,
                                const double* __restrict__ c, const double* __restrict__ l,
                                const double* __restrict__ u, double* __restrict__ x0,
                                double* __restrict__ aty0, double* __restrict__ aty,
                                double* __restrict__ xhat_first, double* __restrict__ xbar,
                                double* __restrict__ y0, double* __restrict__ y) {
  if (sc->done || !sc->restart_flag) return;
  const double tau = sc->tau;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j_begin && i < n) {          // [j_begin, n): this rank's column slice (everything if not sharded)
    const double x = xcand[i], at = aty_cand[i];
    x0[i] = x;
    aty0[i] = at;
    aty[i] = at;
    const double xh = clampd(x - tau * (c[i] - at), l[i], u[i]);
    xhat_first[i] = xh;
    xbar[i] = 2.0 * xh - x;
  }
  if (i < m) {
    const double v = yhat[i];
    y0[i] = v;
    y[i] = v;
  }
}

// xs / ys: warm start in the caller's (unscaled) terms; row_perm / col_perm: internal row / column -> caller's
__global__ void k_start_point(int n, int m, const double* __restrict__ xs, const double* __restrict__ ys,
                              const int* __restrict__ row_perm, const int* __restrict__ col_perm,
                              const double* __restrict__ d1, const double* __restrict__ d2,
                              const double* __restrict__ l, const double* __restrict__ u, double sgn,
                              double* __restrict__ xcand, double* __restrict__ yhat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xcand[i] = clampd(xs ? xs[col_perm ? col_perm[i] : i] / d2[i] : 0.0, l[i], u[i]);
  if (i < m) yhat[i] = ys ? sgn * ys[row_perm ? row_perm[i] : i] / d1[i] : 0.0;
}

// solution in the caller's terms: unscaled, duals back in the caller's row order
__global__ void k_unscale(int n, int m, const double* __restrict__ xh, const double* __restrict__ yh,
                          const int* __restrict__ row_perm, const int* __restrict__ col_perm,
                          const double* __restrict__ d1, const double* __restrict__ d2, double sgn,
                          double* __restrict__ x, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[col_perm ? col_perm[i] : i] = xh[i] * d2[i];
  if (i < m) y[row_perm ? row_perm[i] : i] = sgn * yh[i] * d1[i];
}

__global__ void k_flush(double* buf, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = 0.0;
}

RowStepArgs row_args(mipcu_ctx* ctx, int i) {
  RowStepArgs a;
  const Csr& R = ctx->rows;
  if (kernel_format(ctx, R) == 0) { a.ptr = R.sptr; a.idx = R.sidx; a.val = R.sval; a.perm = R.perm; }
  else { a.ptr = R.ptr; a.idx = R.idx; a.val = R.val; a.perm = nullptr; }
  a.xbar = ctx->xbar; a.lo = ctx->los; a.hi = ctx->his; a.y0 = ctx->y0;
  a.y = ctx->y; a.yhat = ctx->yhat; a.sc = ctx->d_sc; a.partials = ctx->partials;
  a.m = ctx->m; a.iter_in_block = i;
  a.kx_save = nullptr;
  return a;
}

ColStepArgs col_args(mipcu_ctx* ctx, int i, const double* xhat_in, double* xhat_out) {
  ColStepArgs a;
  const Csr& C = ctx->cols;
  if (kernel_format(ctx, C) == 0) { a.ptr = C.sptr; a.idx = C.sidx; a.val = C.sval; a.perm = C.perm; }
  else { a.ptr = C.ptr; a.idx = C.idx; a.val = C.val; a.perm = nullptr; }
  a.yhat = ctx->yhat; a.x0 = ctx->x0; a.aty0 = ctx->aty0; a.c = ctx->cs; a.l = ctx->ls;
  a.u = ctx->us; a.d2 = ctx->d2; a.xbar = ctx->xbar; a.aty = ctx->aty; a.xhat_in = xhat_in;
  a.xhat_out = xhat_out; a.aty_cand = ctx->aty_cand; a.sc = ctx->d_sc;
  a.partials = ctx->partials; a.n = ctx->n; a.iter_in_block = i;
  return a;
}

// TMA-streamed variants are used when the matrix's 64-row tiles fit a shared-memory stage
bool tma_usable(const mipcu_ctx* ctx, const Csr& A, int nv) {
  // opt-in: on B200 the LSU-streamed kernels are faster on every config measured so far
  // (profiles/r01_kernel_variants.md), so "auto" (-1) means LSU
  const int mode = (int)ctx->params[MIPCU_PARAM_USE_TMA];
  if (mode != 1 || A.tile_cap <= 0 || ctx->world > 1) return false;
  return tma_smem_bytes(A.tile_cap, nv) <= 96 * 1024;   // keeps >= 2 CTAs per SM resident
}

// CSR-stream variants (kernels_stream.cuh): the tile's products live in shared memory
bool stream_usable(const mipcu_ctx* ctx, const Csr& A) {
  const int mode = (int)ctx->params[MIPCU_PARAM_USE_TMA];
  if (mode != 2 || A.stream_cap <= 0 || ctx->world > 1) return false;
  return stream_smem_bytes(A.stream_cap) <= 110 * 1024;   // >= 2 CTAs per SM
}

// cudaFuncSetAttribute / the occupancy query are driver calls: remember them per kernel FUNCTION
// (the template parameter is only the pointer type, shared by many instantiations)
std::unordered_map<const void*, std::pair<size_t, int>>& kernel_cache() {
  static std::unordered_map<const void*, std::pair<size_t, int>> cache;   // kernel -> (smem, CTAs/SM)
  return cache;
}

template <class Kernel>
void allow_smem(Kernel kernel, size_t smem) {
  auto& entry = kernel_cache()[reinterpret_cast<const void*>(kernel)];
  if (smem <= entry.first) return;
  MIPCU_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  entry.first = smem;
}

template <class Kernel>
int persistent_grid(mipcu_ctx* ctx, Kernel kernel, size_t smem, int ntiles) {
  auto& entry = kernel_cache()[reinterpret_cast<const void*>(kernel)];
  if (smem != entry.first || entry.second == 0) {
    MIPCU_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MIPCU_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&entry.second, kernel, kThreads, smem));
    entry.first = smem;
  }
  return std::max(1, std::min(ntiles, std::max(entry.second, 1) * ctx->sm_count));
}

template <int TPR>
void launch_row(mipcu_ctx* ctx, const RowStepArgs& a, bool check) {
  if (tma_usable(ctx, ctx->rows, 4)) {
    const int cap = ctx->rows.tile_cap, ntiles = grid_for(ctx->m, kTileRows);
    const size_t smem = tma_smem_bytes(cap, 4);
    if (check) {
      const int g = persistent_grid(ctx, k_row_step_tma<TPR, true>, smem, ntiles);
      k_row_step_tma<TPR, true><<<g, kThreads, smem, ctx->stream>>>(a, cap, ntiles);
    } else {
      const int g = persistent_grid(ctx, k_row_step_tma<TPR, false>, smem, ntiles);
      k_row_step_tma<TPR, false><<<g, kThreads, smem, ctx->stream>>>(a, cap, ntiles);
    }
    return;
  }
  const int g = grid_for(ctx->m, kRowsPerCta);
  if (stream_usable(ctx, ctx->rows)) {
    const size_t smem = stream_smem_bytes(ctx->rows.stream_cap);
    if (check) {
      allow_smem(k_row_step_stream<TPR, true>, smem);
      k_row_step_stream<TPR, true><<<g, kThreads, smem, ctx->stream>>>(a);
    } else {
      allow_smem(k_row_step_stream<TPR, false>, smem);
      k_row_step_stream<TPR, false><<<g, kThreads, smem, ctx->stream>>>(a);
    }
    return;
  }
  if (check) k_row_step<TPR, true><<<g, kThreads, 0, ctx->stream>>>(a);
  else k_row_step<TPR, false><<<g, kThreads, 0, ctx->stream>>>(a);
}

template <int TPR, bool CHECK, bool STORE>
void launch_col_stream(mipcu_ctx* ctx, const ColStepArgs& a) {
  const size_t smem = stream_smem_bytes(ctx->cols.stream_cap);
  allow_smem(k_col_step_stream<TPR, CHECK, STORE>, smem);
  k_col_step_stream<TPR, CHECK, STORE><<<grid_for(ctx->n, kRowsPerCta), kThreads, smem, ctx->stream>>>(a);
}

template <int TPR, bool CHECK, bool STORE>
void launch_col_tma(mipcu_ctx* ctx, const ColStepArgs& a) {
  const int cap = ctx->cols.tile_cap, ntiles = grid_for(ctx->n, kTileRows);
  const size_t smem = tma_smem_bytes(cap, 7);
  const int g = persistent_grid(ctx, k_col_step_tma<TPR, CHECK, STORE>, smem, ntiles);
  k_col_step_tma<TPR, CHECK, STORE><<<g, kThreads, smem, ctx->stream>>>(a, cap, ntiles);
}

template <int TPR>
void launch_col(mipcu_ctx* ctx, const ColStepArgs& a, bool check, bool store) {
  if (tma_usable(ctx, ctx->cols, 7)) {
    if (check && store) launch_col_tma<TPR, true, true>(ctx, a);
    else if (check) launch_col_tma<TPR, true, false>(ctx, a);
    else if (store) launch_col_tma<TPR, false, true>(ctx, a);
    else launch_col_tma<TPR, false, false>(ctx, a);
    return;
  }
  if (stream_usable(ctx, ctx->cols)) {
    if (check && store) launch_col_stream<TPR, true, true>(ctx, a);
    else if (check) launch_col_stream<TPR, true, false>(ctx, a);
    else if (store) launch_col_stream<TPR, false, true>(ctx, a);
    else launch_col_stream<TPR, false, false>(ctx, a);
    return;
  }
  const int g = grid_for(ctx->n, kRowsPerCta);
  cudaStream_t st = ctx->stream;
  if (check && store) k_col_step<TPR, true, true><<<g, kThreads, 0, st>>>(a);
  else if (check) k_col_step<TPR, true, false><<<g, kThreads, 0, st>>>(a);
  else if (store) k_col_step<TPR, false, true><<<g, kThreads, 0, st>>>(a);
  else k_col_step<TPR, false, false><<<g, kThreads, 0, st>>>(a);
}

template <int TPR>
void launch_kkt(mipcu_ctx* ctx) {
  const Csr& R = ctx->rows;
  const int g = grid_for(ctx->m, kRowsPerCta);
  if (TPR == 0)
    k_kkt_row<TPR><<<g, kThreads, 0, ctx->stream>>>(R.sptr, R.sidx, R.sval, R.perm, ctx->xhat_chk,
        ctx->los, ctx->his, ctx->d1, ctx->d_sc, ctx->partials, ctx->m, ctx->tmp_m);
  else
    k_kkt_row<TPR><<<g, kThreads, 0, ctx->stream>>>(R.ptr, R.idx, R.val, nullptr, ctx->xhat_chk,
        ctx->los, ctx->his, ctx->d1, ctx->d_sc, ctx->partials, ctx->m, ctx->tmp_m);
}

// SELL (lane per row) launches: no TMA / CSR-stream variants exist for this format
void launch_row_sell(mipcu_ctx* ctx, const RowStepArgs& a, bool check) {
  const int g = grid_for(ctx->m, kRowsPerCta);
  if (check) launch_kernel(k_row_step<0, true>, g, ctx->stream, ctx->use_pdl, a);
  else launch_kernel(k_row_step<0, false>, g, ctx->stream, ctx->use_pdl, a);
}

void launch_col_sell(mipcu_ctx* ctx, const ColStepArgs& a, bool check, bool store) {
  const int g = grid_for(ctx->n, kRowsPerCta);
  cudaStream_t st = ctx->stream;
  const bool pdl = ctx->use_pdl;
  if (check && store) launch_kernel(k_col_step<0, true, true>, g, st, pdl, a);
  else if (check) launch_kernel(k_col_step<0, true, false>, g, st, pdl, a);
  else if (store) launch_kernel(k_col_step<0, false, true>, g, st, pdl, a);
  else launch_kernel(k_col_step<0, false, false>, g, st, pdl, a);
}

#define MIPCU_DISPATCH_TPR(tpr, CALL)                                                          \
  switch (tpr) {                                                                               \
    case 2: { constexpr int T = 2; CALL; } break;                                              \
    case 4: { constexpr int T = 4; CALL; } break;                                              \
    case 8: { constexpr int T = 8; CALL; } break;                                              \
    case 16: { constexpr int T = 16; CALL; } break;                                            \
    default: { constexpr int T = 32; CALL; } break;                                            \
  }

void row_step(mipcu_ctx* ctx, int i, bool check, bool keep_kx = false) {
  // block-owner exchange: a rank that owns no rows still has to pass the kernel's barrier
  if (ctx->m == 0 && !own_active(ctx)) return;
  RowStepArgs a = row_args(ctx, i);
  if (keep_kx) a.kx_save = ctx->tmp_m;
  if (own_active(ctx)) { own_row_step(ctx, a, check); return; }
  const int fmt = kernel_format(ctx, ctx->rows);
  if (fmt == 0) launch_row_sell(ctx, a, check);
  else MIPCU_DISPATCH_TPR(fmt, launch_row<T>(ctx, a, check));
  ctx->launches++;
}

void col_step(mipcu_ctx* ctx, int i, bool check, const double* xhat_in, double* xhat_out) {
  if (ctx->n == 0) return;
  if (own_active(ctx)) {
    // row-sharded, block-owner exchange: this rank's columns of the WHOLE matrix against the whole
    // yhat the row kernels pushed together; xbar (and the KKT candidate) pushed to the peers
    own_col_step(ctx, i, check, xhat_in, xhat_out);
    return;
  }
  ColStepArgs a = col_args(ctx, i, xhat_in, xhat_out);
  if (p2p_active(ctx)) {
    // row-sharded, fused exchange: local partial K_g' yhat_g, then ONE kernel that pulls this rank's
    // slice of every partial over NVLink, runs the primal step on it and pushes xbar to the peers
    spmv_plain(ctx, ctx->cols, ctx->yhat, ctx->tmp_n);
    if (ctx->split_event) MIPCU_CK(cudaEventRecord(ctx->split_event, ctx->stream));
    p2p_primal_step(ctx, a, check, xhat_out != nullptr);
    if (xhat_out == ctx->xhat_chk) p2p_allgather(ctx, 1);   // gathered by the KKT pass: must be whole
    return;
  }
  if (ctx->world > 1) {
    // row-sharded, NCCL variant: local partial -> all-reduce over NVLink -> elementwise primal step
    // (replicated: every rank computes the same xbar from the same summed vector)
    spmv_plain(ctx, ctx->cols, ctx->yhat, ctx->tmp_n);
    if (ctx->split_event) MIPCU_CK(cudaEventRecord(ctx->split_event, ctx->stream));
    dist_allreduce(ctx, ctx->tmp_n, (size_t)ctx->n, false);
    const int g = grid_for(ctx->n, kRowsPerCta);
    const bool store = xhat_out != nullptr;
    cudaStream_t st = ctx->stream;
    if (check && store) k_primal_step<true, true><<<g, kThreads, 0, st>>>(a, ctx->tmp_n);
    else if (check) k_primal_step<true, false><<<g, kThreads, 0, st>>>(a, ctx->tmp_n);
    else if (store) k_primal_step<false, true><<<g, kThreads, 0, st>>>(a, ctx->tmp_n);
    else k_primal_step<false, false><<<g, kThreads, 0, st>>>(a, ctx->tmp_n);
    ctx->launches++;
    return;
  }
  const int fmt = kernel_format(ctx, ctx->cols);
  if (fmt == 0) launch_col_sell(ctx, a, check, xhat_out != nullptr);
  else MIPCU_DISPATCH_TPR(fmt, launch_col<T>(ctx, a, check, xhat_out != nullptr));
  ctx->launches++;
}

// row-sharded check iterations: make the row-side sums global before a controller runs
const double* global_row_sums(mipcu_ctx* ctx) {
  if (ctx->world <= 1) return nullptr;
  k_sum_partials<<<1, kThreads, 0, ctx->stream>>>(ctx->d_sc, ctx->partials, ctx->row_sums,
                                                  p2p_active(ctx) ? 1 : 0);
  ctx->launches++;
  dist_allreduce(ctx, ctx->row_sums, Q_COUNT, false);
  return ctx->row_sums;
}

void kkt_row(mipcu_ctx* ctx) {
  if (ctx->m == 0) return;
  NvtxRange nvtx("mipcu:kkt pass");
  const int fmt = kernel_format(ctx, ctx->rows);
  if (fmt == 0) launch_kkt<0>(ctx);
  else MIPCU_DISPATCH_TPR(fmt, launch_kkt<T>(ctx));
  ctx->launches++;
}

void restart_apply(mipcu_ctx* ctx) {
  const int nm = std::max(ctx->n, ctx->m);
  if (nm == 0) return;
  const bool p2p = p2p_active(ctx);
  const int j0 = p2p ? ctx->peers.j0[ctx->rank] : 0, j1 = p2p ? ctx->peers.j0[ctx->rank + 1] : ctx->n;
  k_restart_apply<<<grid_for(nm), kThreads, 0, ctx->stream>>>(
      ctx->d_sc, j0, j1, ctx->m, ctx->xhat_chk, ctx->aty_cand, ctx->yhat, ctx->cs, ctx->ls, ctx->us,
      ctx->x0, ctx->aty0, ctx->aty, ctx->xhat_first, ctx->xbar, ctx->y0, ctx->y);
  ctx->launches++;
  if (p2p) p2p_allgather(ctx, 0);   // every rank re-seeded its slice of xbar: make it whole again
}

// One block = `period` iterations.  Iteration 0 and period-1 are "check" iterations (partial sums);
// xhat of iteration period-1 is stored by the col step of iteration period-2 into xhat_chk and
// survives as the candidate solution / restart anchor; xhat of the next block's iteration 0 is
// stored into xhat_first by the last col step (or by k_restart_apply).
// time_events != null: record events around the row and col kernels of iteration 1.
void enqueue_block(mipcu_ctx* ctx, int period, cudaEvent_t* time_events, cudaEvent_t split = nullptr) {
  NvtxRange nvtx("mipcu:iteration block");
  cudaStream_t st = ctx->stream;
  for (int i = 0; i < period; ++i) {
    const bool first = (i == 0), last = (i == period - 1);
    const bool timed = time_events && i == (period > 2 ? 1 : 0);
    if (timed) MIPCU_CK(cudaEventRecord(time_events[0], st));
    row_step(ctx, i, first || last, last);
    if (timed) MIPCU_CK(cudaEventRecord(time_events[1], st));
    if (last) kkt_row(ctx);
    const double* xin = first ? ctx->xhat_first : ctx->xhat_chk;
    double* xout = last ? ctx->xhat_first : (i == period - 2 ? ctx->xhat_chk : nullptr);
    // period == 2: iteration 0 is both "first" (reads xhat_first) and period-2 (stores xhat_chk)
    ctx->split_event = (timed && ctx->world > 1) ? split : nullptr;
    col_step(ctx, i, first || last, xin, xout);
    ctx->split_event = nullptr;
    if (timed) MIPCU_CK(cudaEventRecord(time_events[2], st));
    if (first) {
      k_ctrl_first<<<1, kThreads, 0, st>>>(ctx->d_sc, ctx->partials, global_row_sums(ctx),
                                           p2p_active(ctx) ? 1 : 0);
      ctx->launches++;
    }
  }
  k_ctrl_check<<<1, kThreads, 0, st>>>(ctx->d_sc, ctx->partials, global_row_sums(ctx),
                                       p2p_active(ctx) ? 1 : 0);
  ctx->launches++;
  restart_apply(ctx);
}

void ensure_aux(mipcu_ctx* ctx) {
  ctx->use_pdl = ctx->params[MIPCU_PARAM_USE_PDL] != 0.0;
  if (!ctx->d_sc) MIPCU_CK(cudaMalloc(&ctx->d_sc, sizeof(Scalars)));
  if (!ctx->h_sc) MIPCU_CK(cudaMallocHost(&ctx->h_sc, sizeof(Scalars) * 2));
  for (auto& e : ctx->ev)
    if (!e) MIPCU_CK(cudaEventCreate(&e));
}

// seed (x0, y0) from the warm start (or zeros), aty0 = K'y0, first xhat / xbar
void init_iterate(mipcu_ctx* ctx, Scalars& h) {
  cudaStream_t st = ctx->stream;
  const int n = ctx->n, m = ctx->m, nm = std::max(n, m);
  const double sgn = ctx->maximize ? -1.0 : 1.0;
  h = Scalars();
  h.eta = ctx->eta;
  h.w = ctx->w0;
  h.tau = h.eta / h.w;
  h.sigma = h.eta * h.w;
  h.eps = ctx->params[MIPCU_PARAM_EPS_REL];
  {
    const double cut = ctx->params[MIPCU_PARAM_OBJ_CUTOFF];
    h.cutoff = (cut == cut) ? sgn * cut : INFINITY;
  }
  h.bnorm = ctx->bnorm;
  h.cnorm = ctx->cnorm;
  h.period = (int)ctx->params[MIPCU_PARAM_CHECK_PERIOD];
  h.restart_flag = 1;
  h.nblk_row = grid_for(m, tma_usable(ctx, ctx->rows, 4) ? kTileRows : kRowsPerCta);
  h.nblk_col = grid_for(n, tma_usable(ctx, ctx->cols, 7) ? kTileRows : kRowsPerCta);
  if (p2p_active(ctx))   // the sharded primal kernel reduces over this rank's column slice only
    h.nblk_col = std::max(1, grid_for(ctx->peers.j0[ctx->rank + 1] - ctx->peers.j0[ctx->rank], kRowsPerCta));
  h.max_blocks = ctx->max_blocks;
  // kernel variants write different ranges of the partial arrays: never sum a stale entry
  MIPCU_CK(cudaMemsetAsync(ctx->partials, 0, sizeof(double) * Q_COUNT * ctx->max_blocks, st));
  MIPCU_CK(cudaMemcpyAsync(ctx->d_sc, &h, sizeof(Scalars), cudaMemcpyHostToDevice, st));
  if (nm) {
    k_start_point<<<grid_for(nm), kThreads, 0, st>>>(n, m, ctx->x_start, ctx->y_start, ctx->row_perm, ctx->col_perm, ctx->d1,
                                                     ctx->d2, ctx->ls, ctx->us, sgn, ctx->xhat_chk,
                                                     ctx->yhat);
    ctx->launches++;
  }
  if (n) {
    if (m && ctx->nnz) spmv_plain(ctx, ctx->cols, ctx->yhat, ctx->aty_cand);
    else MIPCU_CK(cudaMemsetAsync(ctx->aty_cand, 0, sizeof(double) * n, st));
    dist_allreduce(ctx, ctx->aty_cand, (size_t)n, false);
  }
  restart_apply(ctx);
  MIPCU_CK(cudaStreamSynchronize(st));   // h is a stack object: the H2D copy must be done
}

}  // namespace

// Index structure besides the 12 bytes per nonzero: CSR reads a 4-byte pointer per row; SELL reads
// a 4-byte offset per 32-row slice and a 1-byte local permutation entry per row.  Padding entries
// of the SELL copy are waste, not algorithmic bytes, and are not counted.
int64_t index_bytes(const mipcu_ctx* ctx, const Csr& A) {
  if (kernel_format(ctx, A) == 0) {
    const int64_t nblk = grid_for(A.nrows, kRowsPerCta);
    return 4 * (nblk * (kRowsPerCta / 32) + 1) + nblk * kRowsPerCta;
  }
  return 4 * ((int64_t)A.nrows + 1);
}

int64_t row_kernel_bytes(const mipcu_ctx* ctx) {
  // matrix stream 12/nnz + index structure + gathered xbar once 8n + reads y,y0,lo,hi + writes y,yhat
  return 12 * ctx->nnz + index_bytes(ctx, ctx->rows) + 8 * (int64_t)ctx->n + 48 * (int64_t)ctx->m;
}

int64_t col_kernel_bytes(const mipcu_ctx* ctx) {
  if (own_active(ctx)) {
    // block-owner exchange: the rank's columns of the whole matrix, the whole yhat gathered once,
    // the primal vectors of the slice only
    const Csr& B = ctx->colblk;
    return 12 * B.nnz + index_bytes(ctx, B) + 8 * (int64_t)ctx->m_glob + 72 * (int64_t)B.nrows;
  }
  // CSC stream 12/nnz + col ptr 4(n+1) + gathered yhat once 8m + reads xbar,x0,aty,aty0,c,l,u
  // + writes xbar,aty
  return 12 * ctx->nnz + index_bytes(ctx, ctx->cols) + 8 * (int64_t)ctx->m + 72 * (int64_t)ctx->n;
}

void solve_lp(mipcu_ctx* ctx, mipcu_stats* out) {
  MIPCU_REQUIRE(ctx->loaded, "mipcu_solve_lp: no model loaded");
  MIPCU_REQUIRE(out, "mipcu_solve_lp: null stats");
  NvtxRange nvtx("mipcu:solve_lp");
  MIPCU_CK(cudaSetDevice(ctx->device));
  auto t0 = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  ctx->launches = 0;
  ensure_aux(ctx);
  const bool first_solve = !ctx->prepared;
  prepare(ctx);
  refresh_vectors(ctx);
  const int period = (int)ctx->params[MIPCU_PARAM_CHECK_PERIOD];
  MIPCU_REQUIRE(period >= 2, "check period must be >= 2");
  const int64_t iter_limit = (int64_t)ctx->params[MIPCU_PARAM_ITER_LIMIT];
  const double time_limit = ctx->params[MIPCU_PARAM_TIME_LIMIT];
  const bool verbose = ctx->params[MIPCU_PARAM_VERBOSE] != 0.0;
  const int graph_mode = (int)ctx->params[MIPCU_PARAM_USE_GRAPH];
  // small problems are launch-bound: replay the block as a CUDA graph
  // row-sharded solves enqueue directly (NCCL calls sit between the kernels)
  const bool use_graph = ctx->world <= 1 &&
                         (graph_mode == 1 || (graph_mode < 0 && ctx->nnz < (int64_t)4 << 20));
  // a wall-clock limit would let ranks disagree on when to stop issuing collectives
  MIPCU_REQUIRE(ctx->world <= 1 || time_limit <= 0.0,
                "a time limit is not supported on row-sharded contexts (use the iteration limit)");

  if (use_graph) ctx->use_pdl = false;
  Scalars h;
  init_iterate(ctx, h);

  if (use_graph && (!ctx->graph_exec || ctx->graph_period != period)) {
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    cudaGraph_t g = nullptr;
    int64_t saved = ctx->launches;
    MIPCU_CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_block(ctx, period, nullptr);
    } catch (...) {
      cudaStreamEndCapture(st, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    MIPCU_CK(cudaStreamEndCapture(st, &g));
    MIPCU_CK(cudaGraphInstantiate(&ctx->graph_exec, g, 0));
    cudaGraphDestroy(g);
    ctx->graph_period = period;
    ctx->launches = saved;
  }
  const int64_t launches_per_block = 2 * (int64_t)period + 4;

  cudaEvent_t ev_begin = ctx->ev[0], ev_end = ctx->ev[1];
  cudaEvent_t ev_slot[2] = {ctx->ev[2], ctx->ev[3]};
  cudaEvent_t* ev_time[2] = {&ctx->ev[4], &ctx->ev[7]};   // 3 events per slot
  MIPCU_CK(cudaEventRecord(ev_begin, st));

  int status = -1;
  int64_t blocks_enqueued = 0, blocks_seen = 0;
  double row_us = 0.0, col_us = 0.0, partial_us = 0.0;
  int64_t timed_samples = 0;
  Scalars last = h;
  auto can_enqueue = [&]() { return blocks_enqueued * (int64_t)period < iter_limit; };
  auto enqueue = [&]() {
    const int slot = (int)(blocks_enqueued & 1);
    if (use_graph) {
      MIPCU_CK(cudaGraphLaunch(ctx->graph_exec, st));
      ctx->launches += launches_per_block;
    } else {
      enqueue_block(ctx, period, ev_time[slot], ctx->ev[10 + slot]);
    }
    MIPCU_CK(cudaMemcpyAsync(ctx->h_sc + slot, ctx->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
    MIPCU_CK(cudaEventRecord(ev_slot[slot], st));
    ++blocks_enqueued;
  };
  enqueue();
  while (true) {
    // keep one block in flight behind the one about to be inspected: the GPU never idles while
    // the host reads the control block (a finished solve makes the extra block fall through)
    if (blocks_enqueued == blocks_seen + 1 && can_enqueue()) enqueue();
    const int slot = (int)(blocks_seen & 1);
    MIPCU_CK(cudaEventSynchronize(ev_slot[slot]));
    last = ctx->h_sc[slot];
    ++blocks_seen;
    if (last.fault) {
      // every later kernel of the queue falls through on the flag: drain, then report (p2p.cu:own_barrier)
      MIPCU_CK(cudaStreamSynchronize(st));
      throw Error("mipcu_solve_lp: a peer rank did not arrive at a cross-GPU barrier within " +
                  std::to_string(ctx->params[MIPCU_PARAM_PEER_TIMEOUT]) +
                  " s (MIPCU_PARAM_PEER_TIMEOUT): row-sharded load and solve are collective and fail-stop");
    }
    if (!use_graph && !last.done) {
      float a = 0.f, b = 0.f;
      MIPCU_CK(cudaEventElapsedTime(&a, ev_time[slot][0], ev_time[slot][1]));
      MIPCU_CK(cudaEventElapsedTime(&b, ev_time[slot][1], ev_time[slot][2]));
      row_us += 1e3 * a; col_us += 1e3 * b; ++timed_samples;
      if (ctx->world > 1 && !own_active(ctx)) {
        float c = 0.f;
        MIPCU_CK(cudaEventElapsedTime(&c, ev_time[slot][1], ctx->ev[10 + slot]));
        partial_us += 1e3 * c;
      }
    }
    if (verbose)
      fprintf(stderr,
              "[mipcu] it %8lld k %7lld  pobj %+.9e dobj %+.9e  rp %.2e rd %.2e gap %.2e  w %.3e fp %.2e%s\n",
              last.total, last.k_base, (ctx->maximize ? -1 : 1) * last.pobj,
              (ctx->maximize ? -1 : 1) * last.dobj, last.rel_p, last.rel_d, last.rel_g, last.w,
              last.fp, last.restart_flag ? "  restart" : "");
    if (last.done) { status = last.done - 1; break; }
    if (time_limit > 0.0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > time_limit) {
      status = MIPCU_TIME_LIMIT;
      break;
    }
    if (blocks_enqueued == blocks_seen) {
      if (!can_enqueue()) { status = MIPCU_ITERATION_LIMIT; break; }
      enqueue();
    }
  }
  MIPCU_CK(cudaStreamSynchronize(st));
  if (blocks_seen < blocks_enqueued) {
    // a block was still in flight when we stopped on a limit: report the state it left behind
    last = ctx->h_sc[(blocks_enqueued - 1) & 1];
    if (last.done) status = last.done - 1;
  }
  MIPCU_CK(cudaEventRecord(ev_end, st));
  const int nm = std::max(ctx->n, ctx->m);
  const double sgn = ctx->maximize ? -1.0 : 1.0;
  if (nm) {
    k_unscale<<<grid_for(nm), kThreads, 0, st>>>(ctx->n, ctx->m, ctx->xhat_chk, ctx->yhat, ctx->row_perm, ctx->col_perm, ctx->d1,
                                                 ctx->d2, sgn, ctx->x_sol, ctx->y_sol);
    ctx->launches++;
  }
  MIPCU_CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  MIPCU_CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
  ctx->has_solution = status != MIPCU_NUMERICAL_ERROR;

  out->status = status;
  out->restarts = last.restarts;
  out->iterations = last.total;
  out->kernel_launches = ctx->launches;
  out->primal_obj = sgn * last.pobj;
  out->dual_obj = sgn * last.dobj;
  out->rel_primal_res = last.rel_p;
  out->rel_dual_res = last.rel_d;
  out->rel_gap = last.rel_g;
  out->step_size = ctx->eta;
  out->primal_weight = last.w;
  out->iter_seconds = 1e-3 * ms;
  out->setup_seconds = first_solve ? ctx->setup_seconds : 0.0;
  out->row_kernel_us = timed_samples ? row_us / timed_samples : 0.0;
  out->col_kernel_us = timed_samples ? col_us / timed_samples : 0.0;
  out->row_kernel_bytes = row_kernel_bytes(ctx);
  out->col_kernel_bytes = col_kernel_bytes(ctx);
  // the block-owner exchange has no separate partial product: the split is reported as 0 / 0
  const bool split = timed_samples && ctx->world > 1 && !own_active(ctx);
  out->col_partial_us = split ? partial_us / timed_samples : 0.0;
  out->col_exchange_us = split ? out->col_kernel_us - out->col_partial_us : 0.0;
  out->solve_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void debug_iterate(mipcu_ctx* ctx, int64_t iters, double* xbar_out, double* y_out) {
  MIPCU_REQUIRE(ctx->loaded, "mipcu_debug_iterate: no model loaded");
  MIPCU_CK(cudaSetDevice(ctx->device));
  ensure_aux(ctx);
  prepare(ctx);
  refresh_vectors(ctx);
  Scalars h;
  init_iterate(ctx, h);
  for (int64_t i = 0; i < iters; ++i) {
    row_step(ctx, (int)i, false);
    col_step(ctx, (int)i, false, nullptr, nullptr);
  }
  if (xbar_out)
    MIPCU_CK(cudaMemcpyAsync(xbar_out, ctx->xbar, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
  if (y_out)
    MIPCU_CK(cudaMemcpyAsync(y_out, ctx->y, sizeof(double) * ctx->m, cudaMemcpyDeviceToHost, ctx->stream));
  MIPCU_CK(cudaStreamSynchronize(ctx->stream));
  if (y_out) rows_to_caller_order(ctx, y_out);
  if (xbar_out) cols_to_caller_order(ctx, xbar_out);
}

void time_kernel(mipcu_ctx* ctx, int which, int reps, int flush_l2, double* mean_us, int64_t* bytes) {
  MIPCU_REQUIRE(ctx->loaded, "mipcu_time_kernel: no model loaded");
  MIPCU_REQUIRE(which >= 0 && which <= 5 && reps > 0, "mipcu_time_kernel: bad arguments");
  MIPCU_CK(cudaSetDevice(ctx->device));
  ensure_aux(ctx);
  prepare(ctx);
  refresh_vectors(ctx);
  Scalars h;
  init_iterate(ctx, h);
  cudaStream_t st = ctx->stream;
  const int64_t flush_count = (int64_t)32 << 20;   // 256 MB of doubles
  if (flush_l2 && !ctx->flush_buf) MIPCU_CK(cudaMalloc(&ctx->flush_buf, sizeof(double) * flush_count));
  auto one = [&]() {
    switch (which) {
      case 0: spmv_plain(ctx, ctx->rows, ctx->xbar, ctx->tmp_m); break;
      case 1: spmv_plain(ctx, ctx->cols, ctx->yhat, ctx->tmp_n); break;
      case 2: row_step(ctx, 1, false); break;
      case 3: col_step(ctx, 1, false, nullptr, nullptr); break;
      case 4:   // one whole iteration as shipped: constant step, nothing between the two kernels
        row_step(ctx, 1, false);
        col_step(ctx, 1, false, nullptr, nullptr);
        break;
      default:  // what an adaptive step-size rule (PDLP) adds to every iteration: the three reductions
                // |dx|^2, |dy|^2, dx.K'dy in both epilogues and a single-CTA controller that must see them
                // before the next iteration may start (a rejected step would repeat all of it)
        row_step(ctx, 0, true);
        col_step(ctx, 0, true, ctx->xhat_first, nullptr);
        k_ctrl_first<<<1, kThreads, 0, st>>>(ctx->d_sc, ctx->partials, global_row_sums(ctx), p2p_active(ctx) ? 1 : 0);
        ctx->launches++;
        break;
    }
  };
  for (int i = 0; i < 3; ++i) one();
  MIPCU_CK(cudaStreamSynchronize(st));
  double total_ms = 0.0;
  if (flush_l2) {
    for (int r = 0; r < reps; ++r) {
      k_flush<<<1184, kThreads, 0, st>>>((double*)ctx->flush_buf, flush_count);
      MIPCU_CK(cudaEventRecord(ctx->ev[4], st));
      one();
      MIPCU_CK(cudaEventRecord(ctx->ev[5], st));
      MIPCU_CK(cudaEventSynchronize(ctx->ev[5]));
      float ms = 0.f;
      MIPCU_CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
      total_ms += ms;
    }
  } else {
    MIPCU_CK(cudaEventRecord(ctx->ev[4], st));
    for (int r = 0; r < reps; ++r) one();
    MIPCU_CK(cudaEventRecord(ctx->ev[5], st));
    MIPCU_CK(cudaEventSynchronize(ctx->ev[5]));
    float ms = 0.f;
    MIPCU_CK(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    total_ms = ms;
  }
  if (mean_us) *mean_us = 1e3 * total_ms / reps;
  if (bytes) {
    const int64_t nnz = ctx->nnz, m = ctx->m, n = ctx->n;
    switch (which) {
      case 0: *bytes = 12 * nnz + 4 * (m + 1) + 8 * n + 8 * m; break;   // B_Ax  
      case 1: *bytes = 12 * nnz + 4 * (n + 1) + 8 * m + 8 * n; break;   // B_ATy
      case 2: *bytes = row_kernel_bytes(ctx); break;
      case 3: *bytes = col_kernel_bytes(ctx); break;
      default: *bytes = row_kernel_bytes(ctx) + col_kernel_bytes(ctx); break;
    }
  }
}

}  // namespace mipcu
