// batch.cu - B node-LP relaxations that share K, c and the row sides and differ only in their
// column boxes (branch and bound), solved together (SURVEY.md section 2.2 "spmm_csr_batch +
// batched step kernels", section 8e).
//
// Layout: every iterate vector is stored [n][BP] / [m][BP] with the batch index fastest
// (BP = B rounded up to 32).  A warp owns one matrix row and its 32 lanes are 32 LPs: the
// (val, idx) pair of a nonzero is loaded once (warp-uniform) and used by 32 problems, and the
// gather `xbar[col][b .. b+32)` is one contiguous 256-byte segment - two full lines per nonzero
// for 32 problems instead of one line per problem as in the single-LP kernels, which is what
// makes node LPs cheap.  There is no sub-warp reduction at all: a lane accumulates its own row
// sum.  Every LP keeps its own control block (restart schedule, primal weight, done flag), the
// same arithmetic as kernels.cuh / solve.cu, one lane per problem.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "internal.h"
#include "kernels.cuh"

namespace mipcu {

namespace {

constexpr int kRowsPerCtaB = 64;          // rows per CTA: 8 warps x 8 rows
constexpr int kWarps = kThreads / 32;

inline int grid_for(int64_t n, int per_block = kThreads) {
  return (int)((n + per_block - 1) / per_block);
}

struct BatchArgs {
  // K by rows and by columns (shared by all LPs)
  const int *rptr, *ridx, *cptr, *cidx;
  const double *rval, *cval;
  const double *lo, *hi, *c, *d1, *d2;      // shared, scaled (d1, d2: the scalings)
  const double *l, *u;                      // [n][BP] scaled boxes
  double *xbar, *x0, *aty, *aty0, *xhat_first, *xhat_chk, *aty_cand;   // [n][BP]
  double *y, *y0, *yhat;                                              // [m][BP]
  double *kx_save;                          // [m][BP] K xbar of the last iteration of a block (ray test)
  Scalars* sc;                              // [BP]
  double* partials;                         // [Q_COUNT][nblk_max][BP]
  int m, n, BP, nblk_max;
};

// One row of the shared matrix times 32 problems' vectors (lane = problem).  val[k] / idx[k] are
// warp-uniform loads, the gather is one contiguous 256-byte segment per nonzero; two accumulator
// chains (even / odd entries), fixed summation order.
//
// Round 2 tried three ways of putting more loads in flight per warp, all measured on one B200 on
// configs 4 / 5 with 128 nodes (profiles/r02_batch_engine.md) and all SLOWER than this plain loop:
//   * row entries fetched one per lane and handed round by warp shuffle, four gathers in flight:
//     row / column kernel 96 / 112 us against 78 / 102 us on config 4 (removed);
//   * rows taken in pairs - step vectors of both rows requested before either product, the two
//     products interleaved (row_dot_batch2), 8 chains in the CTA-per-row kernel: 107 / 146 us against
//     80 / 115 us (MIPCU_BATCH_PAIRS=1);
//   * four accumulator chains in the one-row loop: 84 / 119 us against 80 / 115 us (MIPCU_BATCH_PAIRS=2).
// ncu on the shipped loops: DRAM 22 % (row) / 52 - 56 % (column) busy, L2 29 - 37 %, L1 -> XBAR port
// 19 - 25 %, 50 - 57 % active warps, no level saturated; every extra register the variants spend costs
// resident warps (48 registers = 5 CTAs per SM) and the trade has been negative three times.  The
// throughput that matters for a branch and bound comes from deciding nodes early (MIPCU_BRANCH), not
// from these loops.
__device__ int g_batch_pairs = 0;

__device__ __forceinline__ double row_dot_batch(const int* __restrict__ ptr, const int* __restrict__ idx,
                                                const double* __restrict__ val,
                                                const double* __restrict__ vec, int row, int BP, int b) {
  const int beg = ptr[row], end = ptr[row + 1];
  if (g_batch_pairs == 2) {      // four chains: twice the gathers in flight for four more registers
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = beg;
    for (; k + 3 < end; k += 4) {
      const int c0 = idx[k], c1 = idx[k + 1], c2 = idx[k + 2], c3 = idx[k + 3];
      const double x0 = vec[(size_t)c0 * BP + b], x1 = vec[(size_t)c1 * BP + b];
      const double x2 = vec[(size_t)c2 * BP + b], x3 = vec[(size_t)c3 * BP + b];
      a0 = fma(val[k], x0, a0);
      a1 = fma(val[k + 1], x1, a1);
      a2 = fma(val[k + 2], x2, a2);
      a3 = fma(val[k + 3], x3, a3);
    }
    for (; k < end; ++k) a0 = fma(val[k], vec[(size_t)idx[k] * BP + b], a0);
    return (a0 + a1) + (a2 + a3);
  }
  double a0 = 0.0, a1 = 0.0;
  int k = beg;
  for (; k + 1 < end; k += 2) {
    const double v0 = val[k], v1 = val[k + 1];
    const int c0 = idx[k], c1 = idx[k + 1];
    a0 = fma(v0, vec[(size_t)c0 * BP + b], a0);
    a1 = fma(v1, vec[(size_t)c1 * BP + b], a1);
  }
  if (k < end) a0 = fma(val[k], vec[(size_t)idx[k] * BP + b], a0);
  return a0 + a1;
}

// rows `row` and `row + 1` together: the common prefix of the two rows runs as four independent chains,
// the longer row finishes alone.  Each row's sum is formed exactly as in row_dot_batch.
__device__ __forceinline__ void row_dot_batch2(const int* __restrict__ ptr, const int* __restrict__ idx,
                                               const double* __restrict__ val, const double* __restrict__ vec,
                                               int row, int BP, int b, double& s0, double& s1) {
  const int b0 = ptr[row], b1 = ptr[row + 1], e1 = ptr[row + 2];
  const int e0 = b1;
  const int joint = min(e0 - b0, e1 - b1) & ~1;
  double a0 = 0.0, a1 = 0.0, c0 = 0.0, c1 = 0.0;
  int k = 0;
  for (; k < joint; k += 2) {
    const double v00 = val[b0 + k], v01 = val[b0 + k + 1], v10 = val[b1 + k], v11 = val[b1 + k + 1];
    const int i00 = idx[b0 + k], i01 = idx[b0 + k + 1], i10 = idx[b1 + k], i11 = idx[b1 + k + 1];
    const double x00 = vec[(size_t)i00 * BP + b], x01 = vec[(size_t)i01 * BP + b];
    const double x10 = vec[(size_t)i10 * BP + b], x11 = vec[(size_t)i11 * BP + b];
    a0 = fma(v00, x00, a0);
    a1 = fma(v01, x01, a1);
    c0 = fma(v10, x10, c0);
    c1 = fma(v11, x11, c1);
  }
  int p = b0 + k;
  for (; p + 1 < e0; p += 2) {
    a0 = fma(val[p], vec[(size_t)idx[p] * BP + b], a0);
    a1 = fma(val[p + 1], vec[(size_t)idx[p + 1] * BP + b], a1);
  }
  if (p < e0) a0 = fma(val[p], vec[(size_t)idx[p] * BP + b], a0);
  p = b1 + k;
  for (; p + 1 < e1; p += 2) {
    c0 = fma(val[p], vec[(size_t)idx[p] * BP + b], c0);
    c1 = fma(val[p + 1], vec[(size_t)idx[p + 1] * BP + b], c1);
  }
  if (p < e1) c0 = fma(val[p], vec[(size_t)idx[p] * BP + b], c0);
  s0 = a0 + a1;
  s1 = c0 + c1;
}

template <int NQ>
__device__ __forceinline__ void batch_reduce_store(double (&red)[NQ], const int (&q)[NQ], const BatchArgs& a,
                                                   int b) {
  __shared__ double s_red[NQ][kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NQ; ++i) s_red[i][warp][lane] = red[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += s_red[i][w][lane];
      a.partials[((size_t)q[i] * a.nblk_max + blockIdx.x) * a.BP + b] = t;
    }
  }
}

// the per-row part of the batched row step, given (K xbar)_row for this lane's problem
template <bool CHECK>
__device__ __forceinline__ void row_epilogue_batch(const BatchArgs& a, int row, size_t e, double kx, double yi,
                                                   double y0i, double sigma, double w_new, int keep_kx,
                                                   double (&red)[5]) {
  const double lo = a.lo[row], hi = a.hi[row];
  const double t = kx - yi / sigma;
  const double yh = sigma * (clampd(t, lo, hi) - t);
  a.yhat[e] = yh;
  a.y[e] = w_new * (2.0 * yh - yi) + (1.0 - w_new) * y0i;
  if (CHECK) {
    const double dy = yh - yi, d0 = yh - y0i;
    red[0] += dy * dy;
    red[1] += d0 * d0;
    red[2] += (isfinite(lo) ? lo * fmax(yh, 0.0) : 0.0) - (isfinite(hi) ? hi * fmax(-yh, 0.0) : 0.0);
    const double dp = fmax(dy, 0.0), dm = fmax(-dy, 0.0);
    red[3] += (isfinite(lo) ? lo * dp : 0.0) - (isfinite(hi) ? hi * dm : 0.0);
    const double rv = (isfinite(lo) ? 0.0 : dp) + (isfinite(hi) ? 0.0 : dm);
    red[4] += rv * rv;
    if (keep_kx) a.kx_save[e] = kx;
  }
}

template <bool CHECK>
__global__ void __launch_bounds__(kThreads) k_row_step_batch(const BatchArgs a, int iter_in_block, int keep_kx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const Scalars* sc = a.sc + b;
  const bool active = !sc->done;
  const double sigma = sc->sigma;
  const double k = (double)(sc->k_base + iter_in_block);
  const double w_new = (k + 1.0) / (k + 2.0);
  double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (__any_sync(0xffffffffu, active)) {
    constexpr int kPerWarp = kRowsPerCtaB / kWarps;
    const int first = blockIdx.x * kRowsPerCtaB + warp * kPerWarp;
    const int last = min(first + kPerWarp, a.m);
    int row = first;
    if (g_batch_pairs == 1) {
      for (; row + 1 < last; row += 2) {
        // both rows' step vectors are on their way before either product starts
        const size_t e0 = (size_t)row * a.BP + b, e1 = e0 + a.BP;
        double y_0 = 0.0, y0_0 = 0.0, y_1 = 0.0, y0_1 = 0.0;
        if (active) { y_0 = a.y[e0]; y0_0 = a.y0[e0]; y_1 = a.y[e1]; y0_1 = a.y0[e1]; }
        double kx0, kx1;
        row_dot_batch2(a.rptr, a.ridx, a.rval, a.xbar, row, a.BP, b, kx0, kx1);
        if (!active) continue;
        row_epilogue_batch<CHECK>(a, row, e0, kx0, y_0, y0_0, sigma, w_new, keep_kx, red);
        row_epilogue_batch<CHECK>(a, row + 1, e1, kx1, y_1, y0_1, sigma, w_new, keep_kx, red);
      }
    }
    for (; row < last; ++row) {
      const size_t e = (size_t)row * a.BP + b;
      double yi = 0.0, y0i = 0.0;
      if (active) { yi = a.y[e]; y0i = a.y0[e]; }
      const double kx = row_dot_batch(a.rptr, a.ridx, a.rval, a.xbar, row, a.BP, b);
      if (!active) continue;
      row_epilogue_batch<CHECK>(a, row, e, kx, yi, y0i, sigma, w_new, keep_kx, red);
    }
  }
  if (CHECK) {
    const int q[5] = {Q_DY2, Q_DY0, Q_DOBJ_ROW, Q_RAYD_ROW, Q_RAYD_ROWV};
    batch_reduce_store<5>(red, q, a, b);
  }
}

// Long rows (config 5: 500 rows of ~1600 nonzeros): one CTA per row, its 8 warps stride over the
// row's nonzeros and the partial sums meet in shared memory; warp 0 runs the epilogue.  Fixed
// summation order (warp 0..7), so still bit-reproducible.  MODE 0: row step, MODE 1: KKT residual.
template <bool CHECK, int MODE>
__global__ void __launch_bounds__(kThreads) k_row_batch_long(const BatchArgs a, int iter_in_block, int keep_kx) {
  __shared__ double s_part[kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const int row = blockIdx.x;
  const Scalars* sc = a.sc + b;
  const bool active = !sc->done;
  const double* vec = MODE == 0 ? a.xbar : a.xhat_chk;
  double acc0 = 0.0, acc1 = 0.0;
  if (__any_sync(0xffffffffu, active)) {
    const int beg = a.rptr[row], end = a.rptr[row + 1];
    if (g_batch_pairs) {
      // the 8 warps stride over the row's nonzeros; each warp keeps 8 gathers in flight (a row of 1600
      // entries is 200 per warp: two chains made that a dependent sequence of 100 round trips)
      double c[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      int k = beg + warp;
      for (; k + 7 * kWarps < end; k += 8 * kWarps) {
        double v[8], x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[u] = a.rval[k + u * kWarps];
          x[u] = vec[(size_t)a.ridx[k + u * kWarps] * a.BP + b];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = fma(v[u], x[u], c[u]);
      }
      for (; k < end; k += kWarps) c[0] = fma(a.rval[k], vec[(size_t)a.ridx[k] * a.BP + b], c[0]);
      acc0 = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
    } else {
      int k = beg + warp;
      for (; k + kWarps < end; k += 2 * kWarps) {
        const double v0 = a.rval[k], v1 = a.rval[k + kWarps];
        const int c0 = a.ridx[k], c1 = a.ridx[k + kWarps];
        acc0 = fma(v0, vec[(size_t)c0 * a.BP + b], acc0);
        acc1 = fma(v1, vec[(size_t)c1 * a.BP + b], acc1);
      }
      if (k < end) acc0 = fma(a.rval[k], vec[(size_t)a.ridx[k] * a.BP + b], acc0);
    }
  }
  s_part[warp][lane] = acc0 + acc1;
  __syncthreads();
  if (warp != 0) return;
  double kx = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) kx += s_part[w][lane];
  const size_t slot = (size_t)row * a.BP + b;       // partial index: one entry per row
  if (MODE == 1) {
    const double lo = a.lo[row], hi = a.hi[row];
    const double v = (kx - clampd(kx, lo, hi)) / a.d1[row];
    a.partials[((size_t)Q_PRES2 * a.nblk_max + row) * a.BP + b] = v * v;
    const double kdx = a.kx_save[slot] - kx;
    const double rv = (isfinite(lo) ? fmax(-kdx, 0.0) : 0.0) + (isfinite(hi) ? fmax(kdx, 0.0) : 0.0);
    a.partials[((size_t)Q_RAYP_ROWV * a.nblk_max + row) * a.BP + b] = rv * rv;
    return;
  }
  double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (active) {
    const double sigma = sc->sigma;
    const double kk = (double)(sc->k_base + iter_in_block);
    const double w_new = (kk + 1.0) / (kk + 2.0);
    const double yi = a.y[slot], y0i = a.y0[slot], lo = a.lo[row], hi = a.hi[row];
    const double t = kx - yi / sigma;
    const double yh = sigma * (clampd(t, lo, hi) - t);
    a.yhat[slot] = yh;
    a.y[slot] = w_new * (2.0 * yh - yi) + (1.0 - w_new) * y0i;
    if (CHECK) {
      const double dy = yh - yi, d0 = yh - y0i;
      red[0] = dy * dy;
      red[1] = d0 * d0;
      red[2] = (isfinite(lo) ? lo * fmax(yh, 0.0) : 0.0) - (isfinite(hi) ? hi * fmax(-yh, 0.0) : 0.0);
      const double dp = fmax(dy, 0.0), dm = fmax(-dy, 0.0);
      red[3] = (isfinite(lo) ? lo * dp : 0.0) - (isfinite(hi) ? hi * dm : 0.0);
      const double rv = (isfinite(lo) ? 0.0 : dp) + (isfinite(hi) ? 0.0 : dm);
      red[4] = rv * rv;
      if (keep_kx) a.kx_save[slot] = kx;
    }
  }
  if (CHECK) {
    a.partials[((size_t)Q_DY2 * a.nblk_max + row) * a.BP + b] = red[0];
    a.partials[((size_t)Q_DY0 * a.nblk_max + row) * a.BP + b] = red[1];
    a.partials[((size_t)Q_DOBJ_ROW * a.nblk_max + row) * a.BP + b] = red[2];
    a.partials[((size_t)Q_RAYD_ROW * a.nblk_max + row) * a.BP + b] = red[3];
    a.partials[((size_t)Q_RAYD_ROWV * a.nblk_max + row) * a.BP + b] = red[4];
  }
}

struct ColOperands { double xb, x0, at, at0, lj, uj, xh; };

template <bool CHECK>
__device__ __forceinline__ ColOperands col_load_batch(const BatchArgs& a, size_t e, const double* xhat_in, bool active) {
  ColOperands o = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (active) {
    o.xb = a.xbar[e]; o.x0 = a.x0[e]; o.at = a.aty[e]; o.at0 = a.aty0[e]; o.lj = a.l[e]; o.uj = a.u[e];
    if (CHECK) o.xh = xhat_in[e];
  }
  return o;
}

// the per-column part of the batched column step, given (K' yhat)_j for this lane's problem
template <bool CHECK, bool STORE>
__device__ __forceinline__ void col_epilogue_batch(const BatchArgs& a, int j, size_t e, double atyh,
                                                   const ColOperands& o, double tau, double w_new,
                                                   double* xhat_out, double (&red)[10]) {
  const double xb = o.xb, x0 = o.x0, at = o.at, at0 = o.at0, lj = o.lj, uj = o.uj;
  const double cj = a.c[j];
  if (CHECK) {
    const double xh = o.xh;
    const double dx = xb - xh, dat = atyh - at, d0 = xh - x0;
    red[0] += dx * dx;
    red[1] += dx * dat;
    red[2] += d0 * d0;
    const double r = cj - atyh;
    const double rp = fmax(r, 0.0), rm = fmax(-r, 0.0);
    const double viol = ((isfinite(lj) ? 0.0 : rp) + (isfinite(uj) ? 0.0 : rm)) / a.d2[j];
    red[3] += viol * viol;
    red[4] += cj * xh;
    red[5] += (isfinite(lj) ? lj * rp : 0.0) - (isfinite(uj) ? uj * rm : 0.0);
    a.aty_cand[e] = atyh;
    const double qp = fmax(-dat, 0.0), qm = fmax(dat, 0.0);
    red[6] += (isfinite(lj) ? lj * qp : 0.0) - (isfinite(uj) ? uj * qm : 0.0);
    const double qv = (isfinite(lj) ? 0.0 : qp) + (isfinite(uj) ? 0.0 : qm);
    red[7] += qv * qv;
    red[8] += cj * dx;
    const double pv = (isfinite(lj) ? fmax(-dx, 0.0) : 0.0) + (isfinite(uj) ? fmax(dx, 0.0) : 0.0);
    red[9] += pv * pv;
  }
  const double xn = w_new * xb + (1.0 - w_new) * x0;
  const double atn = w_new * (2.0 * atyh - at) + (1.0 - w_new) * at0;
  const double xh2 = clampd(xn - tau * (cj - atn), lj, uj);
  a.xbar[e] = 2.0 * xh2 - xn;
  a.aty[e] = atn;
  if (STORE) xhat_out[e] = xh2;
}

template <bool CHECK, bool STORE>
__global__ void __launch_bounds__(kThreads)
k_col_step_batch(const BatchArgs a, int iter_in_block, const double* xhat_in, double* xhat_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const Scalars* sc = a.sc + b;
  const bool active = !sc->done;
  const double tau = sc->tau;
  const double k = (double)(sc->k_base + iter_in_block);
  const double w_new = (k + 1.0) / (k + 2.0);
  double red[10] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (__any_sync(0xffffffffu, active)) {
    constexpr int kPerWarp = kRowsPerCtaB / kWarps;
    const int first = blockIdx.x * kRowsPerCtaB + warp * kPerWarp;
    const int last = min(first + kPerWarp, a.n);
    int j = first;
    if (g_batch_pairs == 1) {
      for (; j + 1 < last; j += 2) {
        // the seven streamed operands of both columns are requested before either product starts
        const size_t e0 = (size_t)j * a.BP + b, e1 = e0 + a.BP;
        const ColOperands o0 = col_load_batch<CHECK>(a, e0, xhat_in, active);
        const ColOperands o1 = col_load_batch<CHECK>(a, e1, xhat_in, active);
        double s0, s1;
        row_dot_batch2(a.cptr, a.cidx, a.cval, a.yhat, j, a.BP, b, s0, s1);
        if (!active) continue;
        col_epilogue_batch<CHECK, STORE>(a, j, e0, s0, o0, tau, w_new, xhat_out, red);
        col_epilogue_batch<CHECK, STORE>(a, j + 1, e1, s1, o1, tau, w_new, xhat_out, red);
      }
    }
    for (; j < last; ++j) {
      const size_t e = (size_t)j * a.BP + b;
      const ColOperands o = col_load_batch<CHECK>(a, e, xhat_in, active);
      const double atyh = row_dot_batch(a.cptr, a.cidx, a.cval, a.yhat, j, a.BP, b);
      if (!active) continue;
      col_epilogue_batch<CHECK, STORE>(a, j, e, atyh, o, tau, w_new, xhat_out, red);
    }
  }
  if (CHECK) {
    const int q[10] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL,
                       Q_RAYD_COL, Q_RAYD_COLV, Q_RAYP_C, Q_RAYP_COLV};
    batch_reduce_store<10>(red, q, a, b);
  }
}

__global__ void __launch_bounds__(kThreads) k_kkt_row_batch(const BatchArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  const bool active = !a.sc[b].done;
  double red[2] = {0.0, 0.0};
  if (__any_sync(0xffffffffu, active)) {
    for (int i = 0; i < kRowsPerCtaB / kWarps; ++i) {
      const int row = blockIdx.x * kRowsPerCtaB + warp * (kRowsPerCtaB / kWarps) + i;
      if (row >= a.m) break;
      const double kx = row_dot_batch(a.rptr, a.ridx, a.rval, a.xhat_chk, row, a.BP, b);
      const double lo = a.lo[row], hi = a.hi[row];
      const double v = (kx - clampd(kx, lo, hi)) / a.d1[row];
      red[0] += v * v;
      const double kdx = a.kx_save[(size_t)row * a.BP + b] - kx;      // K (xbar - xhat) = K dx
      const double rv = (isfinite(lo) ? fmax(-kdx, 0.0) : 0.0) + (isfinite(hi) ? fmax(kdx, 0.0) : 0.0);
      red[1] += rv * rv;
    }
  }
  const int q[2] = {Q_PRES2, Q_RAYP_ROWV};
  batch_reduce_store<2>(red, q, a, b);
}

// sum over CTAs of one quantity for the 32 problems of this chunk (lane = problem)
__device__ double batch_sum(const BatchArgs& a, int q, int nblk, int b) {
  __shared__ double s[kWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double t = 0.0;
  for (int blk = warp; blk < nblk; blk += kWarps) t += a.partials[((size_t)q * a.nblk_max + blk) * a.BP + b];
  s[warp][lane] = t;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < kWarps; ++w) r += s[w][lane];
  __syncthreads();
  return r;
}

__device__ double fp_error(const Scalars* sc, double dx2, double dxdaty, double dy2) {
  return sqrt(fmax((sc->w / sc->eta) * dx2 - 2.0 * dxdaty + dy2 / (sc->eta * sc->w), 0.0));
}

__global__ void __launch_bounds__(kThreads) k_ctrl_first_batch(const BatchArgs a, int nblk_row, int nblk_col) {
  const int b = blockIdx.x * 32 + (threadIdx.x & 31);
  const double dy2 = batch_sum(a, Q_DY2, nblk_row, b);
  const double dx2 = batch_sum(a, Q_DX2, nblk_col, b);
  const double dxd = batch_sum(a, Q_DXDATY, nblk_col, b);
  Scalars* sc = a.sc + b;
  if (threadIdx.x < 32 && !sc->done && sc->k_base == 0) {
    const double fp = fp_error(sc, dx2, dxd, dy2);
    sc->fp0 = fp;
    sc->fp_prev = fp;
  }
}

__global__ void __launch_bounds__(kThreads) k_ctrl_check_batch(const BatchArgs a, int nblk_row, int nblk_col) {
  const int b = blockIdx.x * 32 + (threadIdx.x & 31);
  double q[Q_COUNT];
  for (int i = 0; i < Q_COUNT; ++i) {
    const bool is_row = (i == Q_DY2 || i == Q_DY0 || i == Q_DOBJ_ROW || i == Q_PRES2 ||
                         i == Q_RAYD_ROW || i == Q_RAYD_ROWV || i == Q_RAYP_ROWV);
    q[i] = batch_sum(a, i, is_row ? nblk_row : nblk_col, b);
  }
  Scalars* sc = a.sc + b;
  if (threadIdx.x >= 32 || sc->done) return;
  // identical to k_ctrl_check in solve.cu, one lane per problem
  const double fp = fp_error(sc, q[Q_DX2], q[Q_DXDATY], q[Q_DY2]);
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
  if (!(worst == worst) || isinf(worst)) { sc->done = 1 + MIPCU_NUMERICAL_ERROR; return; }
  if (worst <= sc->eps) { sc->done = 1 + MIPCU_OPTIMAL; return; }
  if (sc->rel_d <= sc->eps && dobj >= sc->cutoff) { sc->done = 1 + MIPCU_CUTOFF; return; }
  // early branch: the iterate is branch_tol-accurate and its objective is below the cutoff by more than
  // the error it can still carry, so no accuracy will ever prune this node by bound (include/mipcu.h)
  if (sc->branch_tol > 0.0 && worst <= sc->branch_tol &&
      pobj + 2.0 * sc->branch_tol * (1.0 + fabs(pobj) + fabs(dobj)) < sc->cutoff) {
    sc->done = 1 + MIPCU_BRANCH;
    return;
  }
  {  // certificates of infeasibility from the displacement (same tests as k_ctrl_check)
    const double nd = sqrt(q[Q_DY2]), nx = sqrt(q[Q_DX2]);
    const double dval = nd > 1e-12 ? (q[Q_RAYD_ROW] + q[Q_RAYD_COL]) / nd : 0.0;
    const double dviol = nd > 1e-12 ? sqrt(q[Q_RAYD_ROWV] + q[Q_RAYD_COLV]) / nd : 0.0;
    const double pval = nx > 1e-12 ? q[Q_RAYP_C] / nx : 0.0;
    const double pviol = nx > 1e-12 ? sqrt(q[Q_RAYP_COLV] + q[Q_RAYP_ROWV]) / nx : 0.0;
    sc->ray_hits_dual = (dval > 1e-9 && dviol <= RAY_TOL * dval) ? sc->ray_hits_dual + 1 : 0;
    sc->ray_hits_primal = (pval < -1e-9 && pviol <= RAY_TOL * -pval) ? sc->ray_hits_primal + 1 : 0;
    if (sc->ray_hits_dual >= 2) { sc->done = 1 + MIPCU_INFEASIBLE; return; }
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

__global__ void k_restart_apply_batch(const BatchArgs a) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(e % a.BP);
  const Scalars* sc = a.sc + b;
  if (sc->done || !sc->restart_flag) return;
  const size_t row = e / a.BP;
  if (row < (size_t)a.n) {
    const double x = a.xhat_chk[e], at = a.aty_cand[e];
    a.x0[e] = x;
    a.aty0[e] = at;
    a.aty[e] = at;
    const double xh = clampd(x - sc->tau * (a.c[row] - at), a.l[e], a.u[e]);
    a.xhat_first[e] = xh;
    a.xbar[e] = 2.0 * xh - x;
  }
  if (row < (size_t)a.m) {
    const double v = a.yhat[e];
    a.y0[e] = v;
    a.y[e] = v;
  }
}

// boxes arrive [B][n] from the host; store them scaled and batch-fastest, pad lanes copy LP 0.
// x0 ([B][n], caller's terms, or null): the start point, clamped into the node's box.
__global__ void k_load_boxes(const double* __restrict__ lb, const double* __restrict__ ub,
                             const double* __restrict__ x0,
                             const double* __restrict__ d2, double* __restrict__ l, double* __restrict__ u,
                             double* __restrict__ xcand, int n, int B, int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n * BP) return;
  const int b = (int)(e % BP), j = (int)(e / BP);
  const int src = b < B ? b : 0;
  double lo = lb[(size_t)src * n + j], hi = ub[(size_t)src * n + j];
  if (lo <= -INF_THRESHOLD) lo = -INFINITY;
  if (hi >= INF_THRESHOLD) hi = INFINITY;
  lo /= d2[j];
  hi /= d2[j];
  l[e] = lo;
  u[e] = hi;
  xcand[e] = clampd(x0 ? x0[(size_t)src * n + j] / d2[j] : 0.0, lo, hi);
}

// y0 ([B][m], caller's terms and row order) -> yhat [m][BP] in the engine's (scaled, sorted-row) terms
__global__ void k_load_duals(const double* __restrict__ y0, const int* __restrict__ row_perm,
                             const double* __restrict__ d1, double sgn, double* __restrict__ yhat, int m, int B,
                             int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * BP) return;
  const int b = (int)(e % BP), i = (int)(e / BP);
  const int src = b < B ? b : 0;
  yhat[e] = sgn * y0[(size_t)src * m + (row_perm ? row_perm[i] : i)] / d1[i];
}

// aty_cand = K' yhat for every problem (seeds the warm start; a cold start has K'0 = 0)
__global__ void __launch_bounds__(kThreads) k_aty_batch(const BatchArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 32 + lane;
  for (int i = 0; i < kRowsPerCtaB / kWarps; ++i) {
    const int j = blockIdx.x * kRowsPerCtaB + warp * (kRowsPerCtaB / kWarps) + i;
    if (j >= a.n) break;
    a.aty_cand[(size_t)j * a.BP + b] = row_dot_batch(a.cptr, a.cidx, a.cval, a.yhat, j, a.BP, b);
  }
}

// node_of[slot]: the caller's node whose problem lives in lane `slot` (lanes are re-packed as nodes finish)
__global__ void k_unscale_batch(const double* __restrict__ xhat, const double* __restrict__ d2,
                                const int* __restrict__ node_of, double* __restrict__ x_out, int n, int B, int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n * BP) return;
  const int b = node_of[(int)(e % BP)], j = (int)(e / BP);
  if (b < B) x_out[(size_t)b * n + j] = xhat[e] * d2[j];
}

// Lane compaction: dst[r][s] = src[r][order[s]] for every row r of a [rows][BP] array.
__global__ void k_permute_lanes(const double* __restrict__ src, double* __restrict__ dst,
                                const int* __restrict__ order, size_t count, int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count) return;
  const size_t row = e / BP;
  dst[e] = src[row * BP + order[(int)(e % BP)]];
}
__global__ void k_permute_scalars(const Scalars* __restrict__ src, Scalars* __restrict__ dst,
                                  const int* __restrict__ order, int BP) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < BP) dst[s] = src[order[s]];
}

// duals back in the caller's terms and row order
__global__ void k_unscale_duals_batch(const double* __restrict__ yhat, const int* __restrict__ row_perm,
                                      const double* __restrict__ d1, double sgn, const int* __restrict__ node_of,
                                      double* __restrict__ y_out, int m, int B, int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * BP) return;
  const int b = node_of[(int)(e % BP)], i = (int)(e / BP);
  if (b < B) y_out[(size_t)b * m + (row_perm ? row_perm[i] : i)] = sgn * yhat[e] * d1[i];
}

struct DeviceBuffers {
  std::vector<void*> owned;
  template <class T>
  T* get(size_t count) {
    T* p = nullptr;
    MIPCU_CK(cudaMalloc(&p, sizeof(T) * std::max<size_t>(count, 1)));
    owned.push_back(p);
    return p;
  }
  ~DeviceBuffers() {
    for (void* p : owned) cudaFree(p);
  }
};

}  // namespace

// Algorithmic bytes of one launch of the batched kernels with B problems in flight:
// the matrix is streamed once per 32-problem chunk (12 bytes per nonzero + 4 per row pointer; it
// stays in L2 after the first chunk), the gathered vector is counted once per problem (8 n or 8 m),
// the step vectors as in the single-LP kernels (row: y, y0 read, y, yhat written, row sides shared;
// column: xbar, x0, aty, aty0, l, u read, xbar, aty written, c shared).
int64_t batch_row_bytes(const mipcu_ctx* ctx, int BP) {
  const int64_t chunks = BP / 32;
  return chunks * (12 * ctx->nnz + 4 * ((int64_t)ctx->m + 1) + 16 * (int64_t)ctx->m) +
         (int64_t)BP * (8 * (int64_t)ctx->n + 32 * (int64_t)ctx->m);
}
int64_t batch_col_bytes(const mipcu_ctx* ctx, int BP) {
  const int64_t chunks = BP / 32;
  return chunks * (12 * ctx->nnz + 4 * ((int64_t)ctx->n + 1) + 8 * (int64_t)ctx->n) +
         (int64_t)BP * (8 * (int64_t)ctx->m + 64 * (int64_t)ctx->n);
}

void solve_lp_batch(mipcu_ctx* ctx, int32_t B, const double* lb, const double* ub, const double* cutoff,
                    const double* x0, const double* y0, const double* weight0, const double* branch_tol,
                    mipcu_stats* out, double* x_out, double* y_out) {
  NvtxRange nvtx("mipcu:solve_lp_batch (round of node LPs)");
  MIPCU_REQUIRE(ctx->loaded, "mipcu_solve_lp_batch: no model loaded");
  MIPCU_REQUIRE(ctx->world <= 1, "mipcu_solve_lp_batch: node LPs are dealt to GPUs whole, not sharded");
  MIPCU_REQUIRE(B >= 1 && lb && ub && out, "mipcu_solve_lp_batch: bad arguments");
  MIPCU_CK(cudaSetDevice(ctx->device));
  auto t0 = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  ctx->launches = 0;
  {
    const char* v = std::getenv("MIPCU_BATCH_PAIRS");     // laboratory switch: 0 round-1 loops, 1 row pairs, 2 four chains
    const int pairs = v ? std::atoi(v) : 0;
    MIPCU_CK(cudaMemcpyToSymbolAsync(g_batch_pairs, &pairs, sizeof(int), 0, cudaMemcpyHostToDevice, st));
  }
  if (!ctx->d_sc) MIPCU_CK(cudaMalloc(&ctx->d_sc, sizeof(Scalars)));
  if (!ctx->h_sc) MIPCU_CK(cudaMallocHost(&ctx->h_sc, sizeof(Scalars) * 2));
  for (auto& e : ctx->ev)
    if (!e) MIPCU_CK(cudaEventCreate(&e));
  prepare(ctx);
  refresh_vectors(ctx);

  const int n = ctx->n, m = ctx->m;
  const int BP = (B + 31) / 32 * 32;
  const size_t nB = (size_t)n * BP, mB = (size_t)m * BP;
  // rows far longer than a warp can stream serially get a CTA each (config 5)
  const bool long_rows = m > 0 && ctx->nnz / std::max(m, 1) > 128;
  const int nblk_row = long_rows ? m : grid_for(m, kRowsPerCtaB), nblk_col = grid_for(n, kRowsPerCtaB);
  const int nblk_max = std::max(1, std::max(nblk_row, nblk_col));
  const int period = (int)ctx->params[MIPCU_PARAM_CHECK_PERIOD];
  const int64_t iter_limit = (int64_t)ctx->params[MIPCU_PARAM_ITER_LIMIT];
  const double time_limit = ctx->params[MIPCU_PARAM_TIME_LIMIT];
  const double sgn = ctx->maximize ? -1.0 : 1.0;

  DeviceBuffers buf;
  BatchArgs a{};
  a.rptr = ctx->rows.ptr; a.ridx = ctx->rows.idx; a.rval = ctx->rows.val;
  a.cptr = ctx->cols.ptr; a.cidx = ctx->cols.idx; a.cval = ctx->cols.val;
  a.lo = ctx->los; a.hi = ctx->his; a.c = ctx->cs; a.d1 = ctx->d1; a.d2 = ctx->d2;
  double* l = buf.get<double>(nB);
  double* u = buf.get<double>(nB);
  a.l = l; a.u = u;
  a.xbar = buf.get<double>(nB); a.x0 = buf.get<double>(nB); a.aty = buf.get<double>(nB);
  a.aty0 = buf.get<double>(nB); a.xhat_first = buf.get<double>(nB); a.xhat_chk = buf.get<double>(nB);
  a.aty_cand = buf.get<double>(nB);
  a.y = buf.get<double>(mB); a.y0 = buf.get<double>(mB); a.yhat = buf.get<double>(mB);
  a.kx_save = buf.get<double>(mB);
  a.sc = buf.get<Scalars>(BP);
  a.partials = buf.get<double>((size_t)Q_COUNT * nblk_max * BP);
  a.m = m; a.n = n; a.BP = BP; a.nblk_max = nblk_max;
  MIPCU_CK(cudaMemsetAsync(a.partials, 0, sizeof(double) * Q_COUNT * nblk_max * BP, st));

  // boxes -> device; start point: the caller's (x0, y0) - the parent node's solution in a branch and
  // bound - clamped into the node's box, or x0 = clamp(0), y0 = 0 (then K'y0 = 0: no product to seed)
  {
    double* h_l = buf.get<double>((size_t)B * n);
    double* h_u = buf.get<double>((size_t)B * n);
    MIPCU_CK(cudaMemcpyAsync(h_l, lb, sizeof(double) * (size_t)B * n, cudaMemcpyHostToDevice, st));
    MIPCU_CK(cudaMemcpyAsync(h_u, ub, sizeof(double) * (size_t)B * n, cudaMemcpyHostToDevice, st));
    double* h_x = nullptr;
    if (x0 && n) {
      h_x = buf.get<double>((size_t)B * n);
      MIPCU_CK(cudaMemcpyAsync(h_x, x0, sizeof(double) * (size_t)B * n, cudaMemcpyHostToDevice, st));
    }
    if (nB) {
      k_load_boxes<<<grid_for(nB), kThreads, 0, st>>>(h_l, h_u, h_x, ctx->d2, l, u, a.xhat_chk, n, B, BP);
      ctx->launches++;
    }
    if (y0 && mB) {
      double* h_y = buf.get<double>((size_t)B * m);
      MIPCU_CK(cudaMemcpyAsync(h_y, y0, sizeof(double) * (size_t)B * m, cudaMemcpyHostToDevice, st));
      k_load_duals<<<grid_for(mB), kThreads, 0, st>>>(h_y, ctx->row_perm, ctx->d1, sgn, a.yhat, m, B, BP);
      ctx->launches++;
      if (nB) {
        const dim3 g(std::max(grid_for(n, kRowsPerCtaB), 1), BP / 32);
        k_aty_batch<<<g, kThreads, 0, st>>>(a);
        ctx->launches++;
      }
    } else {
      MIPCU_CK(cudaMemsetAsync(a.aty_cand, 0, sizeof(double) * std::max<size_t>(nB, 1), st));
      MIPCU_CK(cudaMemsetAsync(a.yhat, 0, sizeof(double) * std::max<size_t>(mB, 1), st));
    }
  }
  std::vector<Scalars> h(BP);
  for (int b = 0; b < BP; ++b) {
    Scalars& s = h[b];
    s = Scalars();
    s.eta = ctx->eta;
    s.w = ctx->w0;
    if (weight0) {      // the primal weight the parent's solve ended with
      const double w = weight0[std::min(b, B - 1)];
      if (w > 0.0 && std::isfinite(w)) s.w = w;
    }
    s.tau = s.eta / s.w;
    s.sigma = s.eta * s.w;
    s.eps = ctx->params[MIPCU_PARAM_EPS_REL];
    double cut = cutoff ? cutoff[std::min(b, B - 1)] : ctx->params[MIPCU_PARAM_OBJ_CUTOFF];
    s.cutoff = (cut == cut) ? sgn * cut : INFINITY;
    s.branch_tol = branch_tol ? std::max(0.0, branch_tol[std::min(b, B - 1)]) : 0.0;
    s.bnorm = ctx->bnorm;
    s.cnorm = ctx->cnorm;
    s.period = period;
    s.restart_flag = 1;
    s.done = b < B ? 0 : 1 + MIPCU_OPTIMAL;     // pad lanes never run
  }
  MIPCU_CK(cudaMemcpyAsync(a.sc, h.data(), sizeof(Scalars) * BP, cudaMemcpyHostToDevice, st));
  // Lanes are re-packed as nodes finish (compact, below): `chunks` 32-lane groups hold every problem that
  // still runs, and only those are launched.
  int chunks = BP / 32;
  std::vector<int> slot_of(BP), node_of(BP);
  for (int i = 0; i < BP; ++i) slot_of[i] = node_of[i] = i;
  int* d_order = buf.get<int>(BP);
  int* d_node_of = buf.get<int>(BP);
  MIPCU_CK(cudaMemcpyAsync(d_node_of, node_of.data(), sizeof(int) * BP, cudaMemcpyHostToDevice, st));
  double* scratch_n = buf.get<double>(std::max<size_t>(nB, 1));   // the arrays rotate through these: one per shape
  double* scratch_m = buf.get<double>(std::max<size_t>(mB, 1));
  Scalars* sc_scratch = buf.get<Scalars>(BP);
  const bool allow_compaction = !(std::getenv("MIPCU_BATCH_COMPACT") && !std::atoi(std::getenv("MIPCU_BATCH_COMPACT")));
  const int grid_elem = grid_for(std::max(nB, mB));
  if (grid_elem) {
    k_restart_apply_batch<<<grid_elem, kThreads, 0, st>>>(a);
    ctx->launches++;
  }

  // in-situ kernel times: events around the row and the column kernel of iteration 1 of a block
  cudaEvent_t ev_t[3] = {ctx->ev[4], ctx->ev[5], ctx->ev[6]};
  double row_us = 0.0, col_us = 0.0;
  int64_t timed_blocks = 0;
  auto enqueue_block = [&]() {
    for (int i = 0; i < period; ++i) {
      const bool first = i == 0, last = i == period - 1;
      const bool timed = i == (period > 2 ? 1 : 0);
      if (timed) MIPCU_CK(cudaEventRecord(ev_t[0], st));
      if (m && long_rows) {
        const dim3 g(m, chunks);
        if (first || last) k_row_batch_long<true, 0><<<g, kThreads, 0, st>>>(a, i, last ? 1 : 0);
        else k_row_batch_long<false, 0><<<g, kThreads, 0, st>>>(a, i, 0);
        ctx->launches++;
        if (last) { k_row_batch_long<false, 1><<<g, kThreads, 0, st>>>(a, i, 0); ctx->launches++; }
      } else if (m) {
        const dim3 grid_rows(std::max(nblk_row, 1), chunks);
        if (first || last) k_row_step_batch<true><<<grid_rows, kThreads, 0, st>>>(a, i, last ? 1 : 0);
        else k_row_step_batch<false><<<grid_rows, kThreads, 0, st>>>(a, i, 0);
        ctx->launches++;
        if (last) { k_kkt_row_batch<<<grid_rows, kThreads, 0, st>>>(a); ctx->launches++; }
      }
      if (timed) MIPCU_CK(cudaEventRecord(ev_t[1], st));
      const double* xin = first ? a.xhat_first : a.xhat_chk;
      double* xo = last ? a.xhat_first : (i == period - 2 ? a.xhat_chk : nullptr);
      if (n) {
        const bool chk = first || last;
        const dim3 grid_cols(std::max(nblk_col, 1), chunks);
        if (chk && xo) k_col_step_batch<true, true><<<grid_cols, kThreads, 0, st>>>(a, i, xin, xo);
        else if (chk) k_col_step_batch<true, false><<<grid_cols, kThreads, 0, st>>>(a, i, xin, xo);
        else if (xo) k_col_step_batch<false, true><<<grid_cols, kThreads, 0, st>>>(a, i, xin, xo);
        else k_col_step_batch<false, false><<<grid_cols, kThreads, 0, st>>>(a, i, xin, xo);
        ctx->launches++;
      }
      if (timed) MIPCU_CK(cudaEventRecord(ev_t[2], st));
      if (first) { k_ctrl_first_batch<<<chunks, kThreads, 0, st>>>(a, nblk_row, nblk_col); ctx->launches++; }
    }
    k_ctrl_check_batch<<<chunks, kThreads, 0, st>>>(a, nblk_row, nblk_col);
    ctx->launches++;
    if (grid_elem) { k_restart_apply_batch<<<grid_elem, kThreads, 0, st>>>(a); ctx->launches++; }
  };

  // Re-pack the lanes once a whole 32-lane group can be retired: problems that still run move to the front
  // (stable), finished ones keep their final vectors further back.  A round is as long as its slowest
  // node (config 5: mean 6.8k iterations, max 21k), and a warp streams its rows for all 32 lanes as long
  // as one of them runs; after compaction the kernels are launched over the live groups only.
  auto compact = [&]() {
    if (!allow_compaction || chunks <= 1) return;
    std::vector<int> order;
    order.reserve(BP);
    for (int sl = 0; sl < BP; ++sl) if (!h[sl].done) order.push_back(sl);
    const int live = (int)order.size();
    if ((live + 31) / 32 >= chunks) return;
    for (int sl = 0; sl < BP; ++sl) if (h[sl].done) order.push_back(sl);
    MIPCU_CK(cudaMemcpyAsync(d_order, order.data(), sizeof(int) * BP, cudaMemcpyHostToDevice, st));
    double** nvec[] = {const_cast<double**>(&a.l), const_cast<double**>(&a.u), &a.xbar, &a.x0, &a.aty, &a.aty0,
                       &a.xhat_first, &a.xhat_chk, &a.aty_cand};
    double** mvec[] = {&a.y, &a.y0, &a.yhat, &a.kx_save};
    for (double** v : nvec) {
      if (!nB) break;
      k_permute_lanes<<<grid_for(nB), kThreads, 0, st>>>(*v, scratch_n, d_order, nB, BP);
      std::swap(*v, scratch_n);
      ctx->launches++;
    }
    for (double** v : mvec) {
      if (!mB) break;
      k_permute_lanes<<<grid_for(mB), kThreads, 0, st>>>(*v, scratch_m, d_order, mB, BP);
      std::swap(*v, scratch_m);
      ctx->launches++;
    }
    k_permute_scalars<<<grid_for(BP), kThreads, 0, st>>>(a.sc, sc_scratch, d_order, BP);
    std::swap(a.sc, sc_scratch);
    ctx->launches++;
    std::vector<int> new_node_of(BP);
    std::vector<Scalars> new_h(BP);
    for (int sl = 0; sl < BP; ++sl) {
      new_node_of[sl] = node_of[order[sl]];
      new_h[sl] = h[order[sl]];
      slot_of[new_node_of[sl]] = sl;
    }
    node_of.swap(new_node_of);
    h.swap(new_h);
    MIPCU_CK(cudaMemcpyAsync(d_node_of, node_of.data(), sizeof(int) * BP, cudaMemcpyHostToDevice, st));
    MIPCU_CK(cudaStreamSynchronize(st));      // `order` / `node_of` are host vectors about to change again
    chunks = std::max(1, (live + 31) / 32);
  };

  cudaEvent_t ev_begin = ctx->ev[0], ev_end = ctx->ev[1];
  MIPCU_CK(cudaEventRecord(ev_begin, st));
  int64_t total = 0;
  bool timed_out = false;
  while (true) {
    enqueue_block();
    total += period;
    MIPCU_CK(cudaMemcpyAsync(h.data(), a.sc, sizeof(Scalars) * BP, cudaMemcpyDeviceToHost, st));
    MIPCU_CK(cudaStreamSynchronize(st));
    if (timed_blocks == 0) {   // the first block: every problem of the batch is still in flight
      float ta = 0.f, tb = 0.f;
      MIPCU_CK(cudaEventElapsedTime(&ta, ev_t[0], ev_t[1]));
      MIPCU_CK(cudaEventElapsedTime(&tb, ev_t[1], ev_t[2]));
      row_us += 1e3 * ta; col_us += 1e3 * tb; ++timed_blocks;
    }
    bool all_done = true;
    for (int b = 0; b < B; ++b) all_done = all_done && h[slot_of[b]].done;
    if (all_done || total >= iter_limit) break;
    compact();
    if (time_limit > 0.0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > time_limit) {
      timed_out = true;
      break;
    }
  }
  MIPCU_CK(cudaEventRecord(ev_end, st));
  if (x_out && nB) {
    double* d_x = buf.get<double>((size_t)B * n);
    k_unscale_batch<<<grid_for(nB), kThreads, 0, st>>>(a.xhat_chk, ctx->d2, d_node_of, d_x, n, B, BP);
    ctx->launches++;
    MIPCU_CK(cudaMemcpyAsync(x_out, d_x, sizeof(double) * (size_t)B * n, cudaMemcpyDeviceToHost, st));
  }
  if (y_out && mB) {
    double* d_y = buf.get<double>((size_t)B * m);
    k_unscale_duals_batch<<<grid_for(mB), kThreads, 0, st>>>(a.yhat, ctx->row_perm, ctx->d1, sgn, d_node_of, d_y, m, B, BP);
    ctx->launches++;
    MIPCU_CK(cudaMemcpyAsync(y_out, d_y, sizeof(double) * (size_t)B * m, cudaMemcpyDeviceToHost, st));
  }
  MIPCU_CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  MIPCU_CK(cudaEventElapsedTime(&ms, ev_begin, ev_end));
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int b = 0; b < B; ++b) {
    const Scalars& s = h[slot_of[b]];
    mipcu_stats& o = out[b];
    o = mipcu_stats();
    o.status = s.done ? s.done - 1 : (timed_out ? MIPCU_TIME_LIMIT : MIPCU_ITERATION_LIMIT);
    o.restarts = s.restarts;
    o.iterations = s.total;
    o.kernel_launches = ctx->launches;      // of the whole batch call
    o.primal_obj = sgn * s.pobj;
    o.dual_obj = sgn * s.dobj;
    o.rel_primal_res = s.rel_p;
    o.rel_dual_res = s.rel_d;
    o.rel_gap = s.rel_g;
    o.step_size = ctx->eta;
    o.primal_weight = s.w;
    o.solve_seconds = secs;
    o.iter_seconds = 1e-3 * ms;
    // of the whole batch: one launch of the two batched kernels in the first block (all BP problems in flight)
    o.row_kernel_us = timed_blocks ? row_us / timed_blocks : 0.0;
    o.col_kernel_us = timed_blocks ? col_us / timed_blocks : 0.0;
    o.row_kernel_bytes = batch_row_bytes(ctx, BP);
    o.col_kernel_bytes = batch_col_bytes(ctx, BP);
  }
}

}  // namespace mipcu
