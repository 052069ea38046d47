// batch.cu - B node-LP relaxations that share K, c and the row sides and differ only in their
// column boxes (branch and bound), solved together (SURVEY.md section 2.2 "spmm_csr_batch +
// batched step kernels", section 8e).
//
// Layout: every iterate vector is stored [n][BP] / [m][BP] with the batch index fastest
// (BP = B rounded up to 32, 64 or 128).  A warp owns one matrix row and its 32 lanes hold P = 1, 2
// or 4 adjacent LPs each: the (val, idx) pair of a nonzero is loaded once (warp-uniform) and used by
// 32 P problems, and the gather `xbar[col][b .. b + 32 P)` is one contiguous segment of 256 P bytes -
// full lines shared by all problems instead of one line per problem as in the single-LP kernels,
// which is what makes node LPs cheap.  There is no sub-warp reduction at all: a lane accumulates
// its own row sums.  Every LP keeps its own control block (restart schedule, primal weight, done
// flag), the same arithmetic as kernels.cuh / solve.cu.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
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
  double* hot;                              // [4][BP]: sigma, tau, k_base, done - what the step kernels read
                                            // of the control blocks, as arrays (k_publish_batch)
  double* partials;                         // [Q_COUNT][nblk_max][BP]
  int m, n, BP, nblk_max;
};

// ---- P problems per lane -----------------------------------------------------------------------
// A warp owns one matrix row and covers 32 P problems: lane l holds problems b .. b + P - 1,
// b = (32 blockIdx.y + l) P, which are adjacent in memory (batch index fastest), so its part of the
// gather is one 8 P-byte load.  With P = 4 a warp moves 1 KB per nonzero with the same two warp-uniform
// (val, idx) loads and the same number of instructions that moved 256 bytes with P = 1: four times the
// bytes in flight per warp and a quarter of the L1 wavefronts spent on broadcast loads.  (Round 2 had
// tried more CHAINS per warp three times - shuffle-fed entries, rows in pairs, four accumulators - and
// lost each time to the registers they cost; wider lanes buy the same memory-level parallelism without
// more dependent-load chains.)  The layout does not depend on P: every launch picks its own width from
// the number of problems still running, and a problem's sums are formed in the same order for every P,
// so results are bit-identical across widths (tests/test_batch_gpu.py).
template <int P> __device__ __forceinline__ void ldP(double (&o)[P], const double* __restrict__ p);
template <> __device__ __forceinline__ void ldP<1>(double (&o)[1], const double* __restrict__ p) { o[0] = *p; }
template <> __device__ __forceinline__ void ldP<2>(double (&o)[2], const double* __restrict__ p) {
  const double2 t = *reinterpret_cast<const double2*>(p);
  o[0] = t.x; o[1] = t.y;
}
template <> __device__ __forceinline__ void ldP<4>(double (&o)[4], const double* __restrict__ p) {
  const double2 t0 = reinterpret_cast<const double2*>(p)[0], t1 = reinterpret_cast<const double2*>(p)[1];
  o[0] = t0.x; o[1] = t0.y; o[2] = t1.x; o[3] = t1.y;
}
template <int P> __device__ __forceinline__ void stP(double* __restrict__ p, const double (&v)[P]);
template <> __device__ __forceinline__ void stP<1>(double* __restrict__ p, const double (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void stP<2>(double* __restrict__ p, const double (&v)[2]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}
template <> __device__ __forceinline__ void stP<4>(double* __restrict__ p, const double (&v)[4]) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
// store the values of the problems that still run (one vector store unless the lane is mixed)
template <int P>
__device__ __forceinline__ void st_active(double* __restrict__ p, const double (&v)[P], const bool (&act)[P], bool all) {
  if (all) { stP<P>(p, v); return; }
#pragma unroll
  for (int q = 0; q < P; ++q)
    if (act[q]) p[q] = v[q];
}

// the control-block values a lane needs for its P problems
template <int P> struct LaneState {
  double step[P];      // sigma (row side) or tau (column side)
  double w[P];         // Halpern weight (k + 1) / (k + 2)
  bool act[P];
  bool any, all;
};
template <int P>
__device__ __forceinline__ LaneState<P> lane_state(const BatchArgs& a, int b, int iter_in_block, bool row_side) {
  LaneState<P> s;
  double kb[P], dn[P];
  ldP<P>(s.step, a.hot + (size_t)(row_side ? 0 : 1) * a.BP + b);
  ldP<P>(kb, a.hot + (size_t)2 * a.BP + b);
  ldP<P>(dn, a.hot + (size_t)3 * a.BP + b);
  s.any = false;
  s.all = true;
#pragma unroll
  for (int q = 0; q < P; ++q) {
    s.act[q] = dn[q] == 0.0;
    const double k = kb[q] + (double)iter_in_block;
    s.w[q] = (k + 1.0) / (k + 2.0);
    s.any = s.any || s.act[q];
    s.all = s.all && s.act[q];
  }
  return s;
}

// One row of the shared matrix times the vectors of the lane's P problems.  val[k] / idx[k] are
// warp-uniform loads; two accumulator chains per problem (even / odd entries), fixed summation order.
// Called under `if (lane has a running problem)`: a retired lane is masked off and requests nothing
// (a predicate on the loads instead made ptxas issue every gather right before its use).
template <int P>
__device__ __forceinline__ void row_dot_batch(const int* __restrict__ ptr, const int* __restrict__ idx,
                                              const double* __restrict__ val, const double* __restrict__ vec,
                                              int row, int BP, int b, double (&s)[P]) {
  const int beg = ptr[row], end = ptr[row + 1];
  double a0[P], a1[P];
#pragma unroll
  for (int q = 0; q < P; ++q) a0[q] = a1[q] = 0.0;
  int k = beg;
  // four entries per trip, all four gathers requested before the first product (the two chains still take
  // the even / the odd entries in order, so the sums are those of the two-entry loop below)
  for (; k + 3 < end; k += 4) {
    const double v0 = val[k], v1 = val[k + 1], v2 = val[k + 2], v3 = val[k + 3];
    const int c0 = idx[k], c1 = idx[k + 1], c2 = idx[k + 2], c3 = idx[k + 3];
    double x0[P], x1[P], x2[P], x3[P];
    ldP<P>(x0, vec + (size_t)c0 * BP + b);
    ldP<P>(x1, vec + (size_t)c1 * BP + b);
    ldP<P>(x2, vec + (size_t)c2 * BP + b);
    ldP<P>(x3, vec + (size_t)c3 * BP + b);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      a0[q] = fma(v0, x0[q], a0[q]);
      a1[q] = fma(v1, x1[q], a1[q]);
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
      a0[q] = fma(v2, x2[q], a0[q]);
      a1[q] = fma(v3, x3[q], a1[q]);
    }
  }
  for (; k + 1 < end; k += 2) {
    const double v0 = val[k], v1 = val[k + 1];
    const int c0 = idx[k], c1 = idx[k + 1];
    double x0[P], x1[P];
    ldP<P>(x0, vec + (size_t)c0 * BP + b);
    ldP<P>(x1, vec + (size_t)c1 * BP + b);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      a0[q] = fma(v0, x0[q], a0[q]);
      a1[q] = fma(v1, x1[q], a1[q]);
    }
  }
  if (k < end) {
    const double v0 = val[k];
    double x0[P];
    ldP<P>(x0, vec + (size_t)idx[k] * BP + b);
#pragma unroll
    for (int q = 0; q < P; ++q) a0[q] = fma(v0, x0[q], a0[q]);
  }
#pragma unroll
  for (int q = 0; q < P; ++q) s[q] = a0[q] + a1[q];
}

// per-CTA sums of NQ quantities for the lane's P problems -> partials[quantity][CTA][problem]; warps are
// added in a fixed order (0 .. 7) whatever P is
template <int NQ, int P>
__device__ __forceinline__ void batch_reduce_store(double (&red)[NQ][P], const int (&q)[NQ], const BatchArgs& a,
                                                   int b) {
  __shared__ double s_red[kWarps][32 * P];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
#pragma unroll
    for (int p = 0; p < P; ++p) s_red[warp][lane * P + p] = red[i][p];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += s_red[w][lane * P + p];
        a.partials[((size_t)q[i] * a.nblk_max + blockIdx.x) * a.BP + b + p] = t;
      }
    }
    __syncthreads();
  }
}

// The per-row part of the batched row step for one of the lane's problems, given (K xbar)_row.
// Every product and sum is an explicit IEEE operation (no contraction left to the compiler): a node's
// iterates must not depend on the lane width its launch happened to get, i.e. on how many other nodes
// of the round were still running.
template <bool CHECK>
__device__ __forceinline__ void row_update(double lo, double hi, double kx, double yi, double y0i, double sigma,
                                           double w_new, double& yh_out, double& y_out, double (&c)[5]) {
  const double t = __dsub_rn(kx, __ddiv_rn(yi, sigma));
  const double yh = __dmul_rn(sigma, __dsub_rn(clampd(t, lo, hi), t));
  yh_out = yh;
  y_out = __fma_rn(w_new, __dsub_rn(__dmul_rn(2.0, yh), yi), __dmul_rn(__dsub_rn(1.0, w_new), y0i));
  if (CHECK) {
    const double dy = __dsub_rn(yh, yi), d0 = __dsub_rn(yh, y0i);
    c[0] = __dmul_rn(dy, dy);
    c[1] = __dmul_rn(d0, d0);
    c[2] = __dsub_rn(isfinite(lo) ? __dmul_rn(lo, fmax(yh, 0.0)) : 0.0, isfinite(hi) ? __dmul_rn(hi, fmax(-yh, 0.0)) : 0.0);
    const double dp = fmax(dy, 0.0), dm = fmax(-dy, 0.0);
    c[3] = __dsub_rn(isfinite(lo) ? __dmul_rn(lo, dp) : 0.0, isfinite(hi) ? __dmul_rn(hi, dm) : 0.0);
    const double rv = __dadd_rn(isfinite(lo) ? 0.0 : dp, isfinite(hi) ? 0.0 : dm);
    c[4] = __dmul_rn(rv, rv);
  }
}

// rpw = rows per warp: a CTA owns 8 rpw consecutive rows.  Check iterations always run with rpw = 8
// (kRowsPerCtaB rows per CTA), which fixes the number of partial sums per quantity.
template <bool CHECK, int P>
__global__ void __launch_bounds__(kThreads) k_row_step_batch(const BatchArgs a, int iter_in_block, int keep_kx, int rpw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  const LaneState<P> ls = lane_state<P>(a, b, iter_in_block, true);
  double red[5][P];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int p = 0; p < P; ++p) red[i][p] = 0.0;
  if (ls.any) {     // per lane: a lane whose problems have all finished is masked off for the whole loop
    const int first = (blockIdx.x * kWarps + warp) * rpw;
    const int last = min(first + rpw, a.m);
    for (int row = first; row < last; ++row) {
      const size_t e = (size_t)row * a.BP + b;
      double yi[P], y0i[P], kx[P];
      ldP<P>(yi, a.y + e);
      ldP<P>(y0i, a.y0 + e);
      row_dot_batch<P>(a.rptr, a.ridx, a.rval, a.xbar, row, a.BP, b, kx);
      const double lo = a.lo[row], hi = a.hi[row];
      double yh[P], yn[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        double c[5];
        row_update<CHECK>(lo, hi, kx[p], yi[p], y0i[p], ls.step[p], ls.w[p], yh[p], yn[p], c);
        if (CHECK && ls.act[p]) {
#pragma unroll
          for (int i = 0; i < 5; ++i) red[i][p] += c[i];
        }
      }
      st_active<P>(a.yhat + e, yh, ls.act, ls.all);
      st_active<P>(a.y + e, yn, ls.act, ls.all);
      if (CHECK && keep_kx) st_active<P>(a.kx_save + e, kx, ls.act, ls.all);
    }
  }
  if (CHECK) {
    const int q[5] = {Q_DY2, Q_DY0, Q_DOBJ_ROW, Q_RAYD_ROW, Q_RAYD_ROWV};
    batch_reduce_store<5, P>(red, q, a, b);
  }
}

// Long rows (config 5: 500 rows of ~1600 nonzeros): one CTA per row, its 8 warps stride over the
// row's nonzeros and the partial sums meet in shared memory; warp 0 runs the epilogue.  Fixed
// summation order (warp 0..7), so still bit-reproducible.  MODE 0: row step, MODE 1: KKT residual.
template <bool CHECK, int MODE, int P>
__global__ void __launch_bounds__(kThreads) k_row_batch_long(const BatchArgs a, int iter_in_block, int keep_kx) {
  __shared__ double s_part[kWarps][32 * P];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  const int row = blockIdx.x;
  const LaneState<P> ls = lane_state<P>(a, b, iter_in_block, true);
  const double* vec = MODE == 0 ? a.xbar : a.xhat_chk;
  double acc0[P], acc1[P];
#pragma unroll
  for (int p = 0; p < P; ++p) acc0[p] = acc1[p] = 0.0;
  if (ls.any) {
    const int beg = a.rptr[row], end = a.rptr[row + 1];
    int k = beg + warp;
    for (; k + kWarps < end; k += 2 * kWarps) {
      const double v0 = a.rval[k], v1 = a.rval[k + kWarps];
      const int c0 = a.ridx[k], c1 = a.ridx[k + kWarps];
      double x0[P], x1[P];
      ldP<P>(x0, vec + (size_t)c0 * a.BP + b);
      ldP<P>(x1, vec + (size_t)c1 * a.BP + b);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        acc0[p] = fma(v0, x0[p], acc0[p]);
        acc1[p] = fma(v1, x1[p], acc1[p]);
      }
    }
    if (k < end) {
      const double v0 = a.rval[k];
      double x0[P];
      ldP<P>(x0, vec + (size_t)a.ridx[k] * a.BP + b);
#pragma unroll
      for (int p = 0; p < P; ++p) acc0[p] = fma(v0, x0[p], acc0[p]);
    }
  }
#pragma unroll
  for (int p = 0; p < P; ++p) s_part[warp][lane * P + p] = acc0[p] + acc1[p];
  __syncthreads();
  if (warp != 0) return;
  double kx[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += s_part[w][lane * P + p];
    kx[p] = t;
  }
  const size_t slot = (size_t)row * a.BP + b;       // partial index: one entry per row
  const double lo = a.lo[row], hi = a.hi[row];
  if constexpr (MODE == 1) {
    const double d1 = a.d1[row];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const double v = __ddiv_rn(__dsub_rn(kx[p], clampd(kx[p], lo, hi)), d1);
      a.partials[((size_t)Q_PRES2 * a.nblk_max + row) * a.BP + b + p] = __dmul_rn(v, v);
      const double kdx = __dsub_rn(a.kx_save[slot + p], kx[p]);
      const double rv = __dadd_rn(isfinite(lo) ? fmax(-kdx, 0.0) : 0.0, isfinite(hi) ? fmax(kdx, 0.0) : 0.0);
      a.partials[((size_t)Q_RAYP_ROWV * a.nblk_max + row) * a.BP + b + p] = __dmul_rn(rv, rv);
    }
  } else {
  double red[5][P];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int p = 0; p < P; ++p) red[i][p] = 0.0;
  if (ls.any) {
    double yi[P], y0i[P], yh[P], yn[P];
    ldP<P>(yi, a.y + slot);
    ldP<P>(y0i, a.y0 + slot);
#pragma unroll
    for (int p = 0; p < P; ++p) {
      double c[5];
      row_update<CHECK>(lo, hi, kx[p], yi[p], y0i[p], ls.step[p], ls.w[p], yh[p], yn[p], c);
      if (CHECK && ls.act[p]) {
#pragma unroll
        for (int i = 0; i < 5; ++i) red[i][p] += c[i];
      }
    }
    st_active<P>(a.yhat + slot, yh, ls.act, ls.all);
    st_active<P>(a.y + slot, yn, ls.act, ls.all);
    if (CHECK && keep_kx) st_active<P>(a.kx_save + slot, kx, ls.act, ls.all);
  }
  if (CHECK) {
    const int q[5] = {Q_DY2, Q_DY0, Q_DOBJ_ROW, Q_RAYD_ROW, Q_RAYD_ROWV};
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int p = 0; p < P; ++p) a.partials[((size_t)q[i] * a.nblk_max + row) * a.BP + b + p] = red[i][p];
  }
  }
}

// the per-column part of the batched column step for one problem, given (K' yhat)_j (explicit IEEE
// operations, see row_update)
template <bool CHECK>
__device__ __forceinline__ void col_update(double cj, double d2j, double atyh, double xb, double x0, double at,
                                           double at0, double lj, double uj, double xh, double tau, double w_new,
                                           double& xbar_out, double& aty_out, double& xhat_out, double (&c)[10]) {
  if (CHECK) {
    const double dx = __dsub_rn(xb, xh), dat = __dsub_rn(atyh, at), d0 = __dsub_rn(xh, x0);
    c[0] = __dmul_rn(dx, dx);
    c[1] = __dmul_rn(dx, dat);
    c[2] = __dmul_rn(d0, d0);
    const double r = __dsub_rn(cj, atyh);
    const double rp = fmax(r, 0.0), rm = fmax(-r, 0.0);
    const double viol = __ddiv_rn(__dadd_rn(isfinite(lj) ? 0.0 : rp, isfinite(uj) ? 0.0 : rm), d2j);
    c[3] = __dmul_rn(viol, viol);
    c[4] = __dmul_rn(cj, xh);
    c[5] = __dsub_rn(isfinite(lj) ? __dmul_rn(lj, rp) : 0.0, isfinite(uj) ? __dmul_rn(uj, rm) : 0.0);
    const double qp = fmax(-dat, 0.0), qm = fmax(dat, 0.0);
    c[6] = __dsub_rn(isfinite(lj) ? __dmul_rn(lj, qp) : 0.0, isfinite(uj) ? __dmul_rn(uj, qm) : 0.0);
    const double qv = __dadd_rn(isfinite(lj) ? 0.0 : qp, isfinite(uj) ? 0.0 : qm);
    c[7] = __dmul_rn(qv, qv);
    c[8] = __dmul_rn(cj, dx);
    const double pv = __dadd_rn(isfinite(lj) ? fmax(-dx, 0.0) : 0.0, isfinite(uj) ? fmax(dx, 0.0) : 0.0);
    c[9] = __dmul_rn(pv, pv);
  }
  const double one_w = __dsub_rn(1.0, w_new);
  const double xn = __fma_rn(w_new, xb, __dmul_rn(one_w, x0));
  const double atn = __fma_rn(w_new, __dsub_rn(__dmul_rn(2.0, atyh), at), __dmul_rn(one_w, at0));
  const double xh2 = clampd(__fma_rn(-tau, __dsub_rn(cj, atn), xn), lj, uj);
  xbar_out = __dsub_rn(__dmul_rn(2.0, xh2), xn);
  aty_out = atn;
  xhat_out = xh2;
}

template <bool CHECK, bool STORE, int P>
__global__ void __launch_bounds__(kThreads)
k_col_step_batch(const BatchArgs a, int iter_in_block, const double* xhat_in, double* xhat_out, int rpw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  const LaneState<P> ls = lane_state<P>(a, b, iter_in_block, false);
  double red[10][P];
#pragma unroll
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int p = 0; p < P; ++p) red[i][p] = 0.0;
  if (ls.any) {
    const int first = (blockIdx.x * kWarps + warp) * rpw;
    const int last = min(first + rpw, a.n);
    for (int j = first; j < last; ++j) {
      const size_t e = (size_t)j * a.BP + b;
      // the streamed operands are requested before the product starts
      double xb[P], x0[P], at[P], at0[P], lj[P], uj[P], xh[P], atyh[P];
      ldP<P>(xb, a.xbar + e); ldP<P>(x0, a.x0 + e); ldP<P>(at, a.aty + e); ldP<P>(at0, a.aty0 + e);
      ldP<P>(lj, a.l + e); ldP<P>(uj, a.u + e);
      if (CHECK) ldP<P>(xh, xhat_in + e);
      else {
#pragma unroll
        for (int p = 0; p < P; ++p) xh[p] = 0.0;
      }
      row_dot_batch<P>(a.cptr, a.cidx, a.cval, a.yhat, j, a.BP, b, atyh);
      const double cj = a.c[j], d2j = CHECK ? a.d2[j] : 1.0;
      double xbn[P], atn[P], xhn[P];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        double c[10];
        col_update<CHECK>(cj, d2j, atyh[p], xb[p], x0[p], at[p], at0[p], lj[p], uj[p], xh[p], ls.step[p], ls.w[p],
                          xbn[p], atn[p], xhn[p], c);
        if (CHECK && ls.act[p]) {
#pragma unroll
          for (int i = 0; i < 10; ++i) red[i][p] += c[i];
        }
      }
      if (CHECK) st_active<P>(a.aty_cand + e, atyh, ls.act, ls.all);
      st_active<P>(a.xbar + e, xbn, ls.act, ls.all);
      st_active<P>(a.aty + e, atn, ls.act, ls.all);
      if (STORE) st_active<P>(xhat_out + e, xhn, ls.act, ls.all);
    }
  }
  if (CHECK) {
    const int q[10] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL,
                       Q_RAYD_COL, Q_RAYD_COLV, Q_RAYP_C, Q_RAYP_COLV};
    batch_reduce_store<10, P>(red, q, a, b);
  }
}


// ---- entry-fed variants of the two step kernels (non-check iterations) ----------------------------------
// The loops above reach a nonzero through a dependent chain: warp-uniform idx[k] (an L1 / L2 round trip),
// then the gather (L2 or HBM round trip), trip after trip, row after row; ncu shows them waiting, not
// moving bytes.  Here the warp fetches up to 32 (idx, val) entries of a row with ONE coalesced load per
// array - lane l holds entry l - and hands them round by shuffle, so every gather of the row can be
// requested as soon as that single load has landed; and the entries of the NEXT row are requested before
// the current row's products start.  The chain per row is one gather round trip instead of
// (entries / 4) x (index + gather) round trips.  All 32 lanes stay converged (shuffles need them): a lane
// whose problems have all finished gathers at the address of a lane that still runs (same sectors, no extra
// traffic) and stores nothing.  The two accumulator chains take the even / odd entries in order exactly as
// row_dot_batch does, so both paths produce the same bits.
constexpr unsigned kFull = 0xffffffffu;

template <int P>
__device__ __forceinline__ void feed_chunk(int ci, double cv, int cnt, const double* __restrict__ vec, int BP,
                                           int b, double (&a0)[P], double (&a1)[P]) {
  int k = 0;
  for (; k + 3 < cnt; k += 4) {
    const int c0 = __shfl_sync(kFull, ci, k), c1 = __shfl_sync(kFull, ci, k + 1);
    const int c2 = __shfl_sync(kFull, ci, k + 2), c3 = __shfl_sync(kFull, ci, k + 3);
    const double v0 = __shfl_sync(kFull, cv, k), v1 = __shfl_sync(kFull, cv, k + 1);
    const double v2 = __shfl_sync(kFull, cv, k + 2), v3 = __shfl_sync(kFull, cv, k + 3);
    double x0[P], x1[P], x2[P], x3[P];
    ldP<P>(x0, vec + (size_t)c0 * BP + b);
    ldP<P>(x1, vec + (size_t)c1 * BP + b);
    ldP<P>(x2, vec + (size_t)c2 * BP + b);
    ldP<P>(x3, vec + (size_t)c3 * BP + b);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      a0[q] = fma(v0, x0[q], a0[q]);
      a1[q] = fma(v1, x1[q], a1[q]);
    }
#pragma unroll
    for (int q = 0; q < P; ++q) {
      a0[q] = fma(v2, x2[q], a0[q]);
      a1[q] = fma(v3, x3[q], a1[q]);
    }
  }
  for (; k + 1 < cnt; k += 2) {
    const int c0 = __shfl_sync(kFull, ci, k), c1 = __shfl_sync(kFull, ci, k + 1);
    const double v0 = __shfl_sync(kFull, cv, k), v1 = __shfl_sync(kFull, cv, k + 1);
    double x0[P], x1[P];
    ldP<P>(x0, vec + (size_t)c0 * BP + b);
    ldP<P>(x1, vec + (size_t)c1 * BP + b);
#pragma unroll
    for (int q = 0; q < P; ++q) {
      a0[q] = fma(v0, x0[q], a0[q]);
      a1[q] = fma(v1, x1[q], a1[q]);
    }
  }
  if (k < cnt) {
    const int c0 = __shfl_sync(kFull, ci, k);
    const double v0 = __shfl_sync(kFull, cv, k);
    double x0[P];
    ldP<P>(x0, vec + (size_t)c0 * BP + b);
#pragma unroll
    for (int q = 0; q < P; ++q) a0[q] = fma(v0, x0[q], a0[q]);
  }
}

// the warp's rows [first, first + nr): pointers once (lane r holds ptr[first + r]), then per row the
// product with entries fed by shuffle; `next` prefetches the following row's first 32 entries
template <int P>
struct RowFeed {
  const int* __restrict__ ptr; const int* __restrict__ idx; const double* __restrict__ val;
  int lane, myptr, beg, end, ci, nbeg, nend, nci;
  double cv, ncv;
  __device__ __forceinline__ void start(int first, int nr) {
    myptr = (nr > 0 && lane <= nr) ? ptr[first + lane] : 0;
    beg = __shfl_sync(kFull, myptr, 0);
    end = __shfl_sync(kFull, myptr, 1);
    fetch(beg, end, ci, cv);
  }
  __device__ __forceinline__ void fetch(int from, int to, int& i_out, double& v_out) const {
    i_out = 0; v_out = 0.0;
    if (from + lane < to) { i_out = idx[from + lane]; v_out = val[from + lane]; }
  }
  __device__ __forceinline__ void prefetch_next(int r, int nr) {
    nbeg = nend = 0; nci = 0; ncv = 0.0;
    if (r + 1 < nr) {
      nbeg = __shfl_sync(kFull, myptr, r + 1);
      nend = __shfl_sync(kFull, myptr, r + 2);
      fetch(nbeg, nend, nci, ncv);
    }
  }
  __device__ __forceinline__ void dot(const double* __restrict__ vec, int BP, int b, double (&s)[P]) {
    double a0[P], a1[P];
#pragma unroll
    for (int q = 0; q < P; ++q) a0[q] = a1[q] = 0.0;
    int base = beg;
    while (true) {
      feed_chunk<P>(ci, cv, min(32, end - base), vec, BP, b, a0, a1);
      base += 32;
      if (base >= end) break;
      fetch(base, end, ci, cv);
    }
#pragma unroll
    for (int q = 0; q < P; ++q) s[q] = a0[q] + a1[q];
  }
  __device__ __forceinline__ void advance() { beg = nbeg; end = nend; ci = nci; cv = ncv; }
};

template <int P>
__global__ void __launch_bounds__(kThreads) k_row_step_batch_fed(const BatchArgs a, int iter_in_block, int rpw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  const LaneState<P> ls = lane_state<P>(a, b, iter_in_block, true);
  const unsigned live = __ballot_sync(kFull, ls.any);
  if (live == 0) return;
  const int b_ld = __shfl_sync(kFull, b, ls.any ? lane : __ffs(live) - 1);
  const int first = (blockIdx.x * kWarps + warp) * rpw;
  const int nr = min(first + rpw, a.m) - first;
  RowFeed<P> rf{a.rptr, a.ridx, a.rval, lane};
  rf.start(first, nr);
  for (int r = 0; r < nr; ++r) {
    const int row = first + r;
    rf.prefetch_next(r, nr);
    double yi[P], y0i[P], kx[P];
    ldP<P>(yi, a.y + (size_t)row * a.BP + b_ld);
    ldP<P>(y0i, a.y0 + (size_t)row * a.BP + b_ld);
    rf.dot(a.xbar, a.BP, b_ld, kx);
    const double lo = a.lo[row], hi = a.hi[row];
    double yh[P], yn[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      double c[5];
      row_update<false>(lo, hi, kx[p], yi[p], y0i[p], ls.step[p], ls.w[p], yh[p], yn[p], c);
    }
    if (ls.any) {
      const size_t e = (size_t)row * a.BP + b;
      st_active<P>(a.yhat + e, yh, ls.act, ls.all);
      st_active<P>(a.y + e, yn, ls.act, ls.all);
    }
    rf.advance();
  }
}

template <bool STORE, int P>
__global__ void __launch_bounds__(kThreads)
k_col_step_batch_fed(const BatchArgs a, int iter_in_block, double* xhat_out, int rpw) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  const LaneState<P> ls = lane_state<P>(a, b, iter_in_block, false);
  const unsigned live = __ballot_sync(kFull, ls.any);
  if (live == 0) return;
  const int b_ld = __shfl_sync(kFull, b, ls.any ? lane : __ffs(live) - 1);
  const int first = (blockIdx.x * kWarps + warp) * rpw;
  const int nr = min(first + rpw, a.n) - first;
  RowFeed<P> rf{a.cptr, a.cidx, a.cval, lane};
  rf.start(first, nr);
  for (int r = 0; r < nr; ++r) {
    const int j = first + r;
    rf.prefetch_next(r, nr);
    const size_t el = (size_t)j * a.BP + b_ld;
    double xb[P], x0[P], at[P], at0[P], lj[P], uj[P], atyh[P];
    ldP<P>(xb, a.xbar + el); ldP<P>(x0, a.x0 + el); ldP<P>(at, a.aty + el); ldP<P>(at0, a.aty0 + el);
    ldP<P>(lj, a.l + el); ldP<P>(uj, a.u + el);
    rf.dot(a.yhat, a.BP, b_ld, atyh);
    const double cj = a.c[j];
    double xbn[P], atn[P], xhn[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      double c[10];
      col_update<false>(cj, 1.0, atyh[p], xb[p], x0[p], at[p], at0[p], lj[p], uj[p], 0.0, ls.step[p], ls.w[p],
                        xbn[p], atn[p], xhn[p], c);
    }
    if (ls.any) {
      const size_t e = (size_t)j * a.BP + b;
      st_active<P>(a.xbar + e, xbn, ls.act, ls.all);
      st_active<P>(a.aty + e, atn, ls.act, ls.all);
      if (STORE) st_active<P>(xhat_out + e, xhn, ls.act, ls.all);
    }
    rf.advance();
  }
}

template <int P>
__global__ void __launch_bounds__(kThreads) k_kkt_row_batch(const BatchArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  double dn[P];
  ldP<P>(dn, a.hot + (size_t)3 * a.BP + b);
  bool any = false;
#pragma unroll
  for (int p = 0; p < P; ++p) any = any || dn[p] == 0.0;
  double red[2][P];
#pragma unroll
  for (int p = 0; p < P; ++p) red[0][p] = red[1][p] = 0.0;
  if (any) {
    for (int i = 0; i < kRowsPerCtaB / kWarps; ++i) {
      const int row = blockIdx.x * kRowsPerCtaB + warp * (kRowsPerCtaB / kWarps) + i;
      if (row >= a.m) break;
      double kx[P], ks[P];
      row_dot_batch<P>(a.rptr, a.ridx, a.rval, a.xhat_chk, row, a.BP, b, kx);
      ldP<P>(ks, a.kx_save + (size_t)row * a.BP + b);
      const double lo = a.lo[row], hi = a.hi[row], d1 = a.d1[row];
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const double v = __ddiv_rn(__dsub_rn(kx[p], clampd(kx[p], lo, hi)), d1);
        red[0][p] += __dmul_rn(v, v);
        const double kdx = __dsub_rn(ks[p], kx[p]);      // K (xbar - xhat) = K dx
        const double rv = __dadd_rn(isfinite(lo) ? fmax(-kdx, 0.0) : 0.0, isfinite(hi) ? fmax(kdx, 0.0) : 0.0);
        red[1][p] += __dmul_rn(rv, rv);
      }
    }
  }
  const int q[2] = {Q_PRES2, Q_RAYP_ROWV};
  batch_reduce_store<2, P>(red, q, a, b);
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

// The four control-block fields the step kernels read, as arrays [4][BP] (sigma, tau, k_base, done): a
// lane picks up the state of its P problems with three vector loads instead of 3 P loads strided by
// sizeof(Scalars).  Runs after everything that changes them (start, controller, lane compaction).
__global__ void k_publish_batch(const BatchArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.BP) return;
  const Scalars* sc = a.sc + b;
  a.hot[b] = sc->sigma;
  a.hot[(size_t)a.BP + b] = sc->tau;
  a.hot[(size_t)2 * a.BP + b] = (double)sc->k_base;
  a.hot[(size_t)3 * a.BP + b] = sc->done ? 1.0 : 0.0;
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
                             const double* __restrict__ x0, const int* __restrict__ col_perm,
                             const double* __restrict__ d2, double* __restrict__ l, double* __restrict__ u,
                             double* __restrict__ xcand, int n, int B, int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n * BP) return;
  const int b = (int)(e % BP), j = (int)(e / BP);
  const int src = b < B ? b : 0;
  const int jc = col_perm ? col_perm[j] : j;      // the caller's column
  double lo = lb[(size_t)src * n + jc], hi = ub[(size_t)src * n + jc];
  if (lo <= -INF_THRESHOLD) lo = -INFINITY;
  if (hi >= INF_THRESHOLD) hi = INFINITY;
  lo /= d2[j];
  hi /= d2[j];
  l[e] = lo;
  u[e] = hi;
  xcand[e] = clampd(x0 ? x0[(size_t)src * n + jc] / d2[j] : 0.0, lo, hi);
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
template <int P>
__global__ void __launch_bounds__(kThreads) k_aty_batch(const BatchArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = (blockIdx.y * 32 + lane) * P;
  for (int i = 0; i < kRowsPerCtaB / kWarps; ++i) {
    const int j = blockIdx.x * kRowsPerCtaB + warp * (kRowsPerCtaB / kWarps) + i;
    if (j >= a.n) break;
    double s[P];
    row_dot_batch<P>(a.cptr, a.cidx, a.cval, a.yhat, j, a.BP, b, s);
    stP<P>(a.aty_cand + (size_t)j * a.BP + b, s);
  }
}

// node_of[slot]: the caller's node whose problem lives in lane `slot` (lanes are re-packed as nodes finish)
__global__ void k_unscale_batch(const double* __restrict__ xhat, const double* __restrict__ d2,
                                const int* __restrict__ node_of, const int* __restrict__ col_perm,
                                double* __restrict__ x_out, int n, int B, int BP) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)n * BP) return;
  const int b = node_of[(int)(e % BP)], j = (int)(e / BP);
  if (b < B) x_out[(size_t)b * n + (col_perm ? col_perm[j] : j)] = xhat[e] * d2[j];
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

// f(std::integral_constant<int, P>) for P = 1, 2, 4: picks the kernel instance of a lane width
template <class F>
void with_width(int P, F&& f) {
  if (P == 4) f(std::integral_constant<int, 4>{});
  else if (P == 2) f(std::integral_constant<int, 2>{});
  else f(std::integral_constant<int, 1>{});
}

// Scratch of one batch call.  From the context's stream-ordered pool when it has one (a branch and bound
// issues a round every few hundred milliseconds: ~25 arrays of up to 50 MB each through cudaMalloc /
// cudaFree were 15 - 25 ms of every round, and cudaFree synchronises the device); the pool keeps the
// memory between rounds.  Freed in stream order when the call returns.
struct DeviceBuffers {
  mipcu_ctx* ctx;
  std::vector<void*> owned;
  explicit DeviceBuffers(mipcu_ctx* c) : ctx(c) {}
  template <class T>
  T* get(size_t count) {
    T* p = nullptr;
    const size_t bytes = sizeof(T) * std::max<size_t>(count, 1);
    if (ctx->pool_alloc) MIPCU_CK(cudaMallocFromPoolAsync(&p, bytes, (cudaMemPool_t)ctx->pool, ctx->stream));
    else MIPCU_CK(cudaMalloc(&p, bytes));
    owned.push_back(p);
    return p;
  }
  ~DeviceBuffers() {
    for (void* p : owned) {
      if (ctx->pool_alloc) cudaFreeAsync(p, ctx->stream);
      else cudaFree(p);
    }
  }
};

}  // namespace

// Algorithmic bytes of one launch of the batched kernels with B problems in flight:
// the shared matrix once (12 bytes per nonzero + 4 per row pointer; round 2 counted it once per
// 32-problem chunk, which is what the kernels did then), the shared row sides / objective once, the
// gathered vector once per problem (8 n or 8 m), the step vectors as in the single-LP kernels
// (row: y, y0 read, y, yhat written; column: xbar, x0, aty, aty0, l, u read, xbar, aty written).
int64_t batch_row_bytes(const mipcu_ctx* ctx, int BP) {
  return 12 * ctx->nnz + 4 * ((int64_t)ctx->m + 1) + 16 * (int64_t)ctx->m +
         (int64_t)BP * (8 * (int64_t)ctx->n + 32 * (int64_t)ctx->m);
}
int64_t batch_col_bytes(const mipcu_ctx* ctx, int BP) {
  return 12 * ctx->nnz + 4 * ((int64_t)ctx->n + 1) + 8 * (int64_t)ctx->n +
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
  const bool timing = std::getenv("MIPCU_TIMING") != nullptr;      // wall-clock laps of the call on stderr
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(st);
    std::fprintf(stderr, "[mipcu_solve_lp_batch] %s %.1f ms\n", what,
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  };
  if (!ctx->d_sc) MIPCU_CK(cudaMalloc(&ctx->d_sc, sizeof(Scalars)));
  if (!ctx->h_sc) MIPCU_CK(cudaMallocHost(&ctx->h_sc, sizeof(Scalars) * 2));
  for (auto& e : ctx->ev)
    if (!e) MIPCU_CK(cudaEventCreate(&e));
  prepare(ctx);
  refresh_vectors(ctx);

  const int n = ctx->n, m = ctx->m;
  // lanes hold up to max_width problems (laboratory switch MIPCU_BATCH_WIDTH = 1 | 2 | 4; default 4)
  int max_width = 4;
  if (const char* v = std::getenv("MIPCU_BATCH_WIDTH")) max_width = std::atoi(v) >= 4 ? 4 : std::atoi(v) >= 2 ? 2 : 1;
  // laboratory switches: MIPCU_BATCH_FEED (row step; default 1) / MIPCU_BATCH_FEED_COL (column step; default 0):
  // 1 = entries fetched one per lane and fed by shuffle, 0 = warp-uniform entry loads.  Measured on configs 4 / 5
  // (profiles/r02_batch_engine.md): the entry-fed row kernel with 4 problems per lane and the uniform-load
  // column kernel with 2 are the fastest pair; the CTA-per-row kernel of long-row models is best with 2.
  const bool fed = !(std::getenv("MIPCU_BATCH_FEED") && !std::atoi(std::getenv("MIPCU_BATCH_FEED")));
  const bool fed_col = std::getenv("MIPCU_BATCH_FEED_COL") && std::atoi(std::getenv("MIPCU_BATCH_FEED_COL"));
  const int fixed_rpw = std::getenv("MIPCU_BATCH_RPW") ? std::max(1, std::min(8, std::atoi(std::getenv("MIPCU_BATCH_RPW")))) : 0;
  const int pad_to = B > 64 ? 32 * max_width : B > 32 ? 32 * std::min(max_width, 2) : 32;
  const int BP = (B + pad_to - 1) / pad_to * pad_to;
  const size_t nB = (size_t)n * BP, mB = (size_t)m * BP;
  // rows far longer than a warp can stream serially get a CTA each (config 5)
  const bool long_rows = m > 0 && ctx->nnz / std::max(m, 1) > 128;
  const int nblk_row = long_rows ? m : grid_for(m, kRowsPerCtaB), nblk_col = grid_for(n, kRowsPerCtaB);
  const int nblk_max = std::max(1, std::max(nblk_row, nblk_col));
  const int period = (int)ctx->params[MIPCU_PARAM_CHECK_PERIOD];
  const int64_t iter_limit = (int64_t)ctx->params[MIPCU_PARAM_ITER_LIMIT];
  const double time_limit = ctx->params[MIPCU_PARAM_TIME_LIMIT];
  const double sgn = ctx->maximize ? -1.0 : 1.0;

  DeviceBuffers buf(ctx);
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
  a.hot = buf.get<double>((size_t)4 * BP);
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
      k_load_boxes<<<grid_for(nB), kThreads, 0, st>>>(h_l, h_u, h_x, ctx->col_perm, ctx->d2, l, u, a.xhat_chk, n, B, BP);
      ctx->launches++;
    }
    if (y0 && mB) {
      double* h_y = buf.get<double>((size_t)B * m);
      MIPCU_CK(cudaMemcpyAsync(h_y, y0, sizeof(double) * (size_t)B * m, cudaMemcpyHostToDevice, st));
      k_load_duals<<<grid_for(mB), kThreads, 0, st>>>(h_y, ctx->row_perm, ctx->d1, sgn, a.yhat, m, B, BP);
      ctx->launches++;
      if (nB) {
        const int P0 = (BP % 128 == 0 && max_width >= 4) ? 4 : (BP % 64 == 0 && max_width >= 2) ? 2 : 1;
        const dim3 g(std::max(grid_for(n, kRowsPerCtaB), 1), BP / (32 * P0));
        with_width(P0, [&](auto w) { k_aty_batch<decltype(w)::value><<<g, kThreads, 0, st>>>(a); });
        ctx->launches++;
      }
    } else {
      MIPCU_CK(cudaMemsetAsync(a.aty_cand, 0, sizeof(double) * std::max<size_t>(nB, 1), st));
      MIPCU_CK(cudaMemsetAsync(a.yhat, 0, sizeof(double) * std::max<size_t>(mB, 1), st));
    }
  }
  lap("scratch allocated, boxes and start points on the device");
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
  k_publish_batch<<<grid_for(BP), kThreads, 0, st>>>(a);
  ctx->launches++;
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
  // Lane width and rows per warp of a launch.  Width: the widest lanes the running problems fill more than
  // half of (check iterations: at most 2 - their 5 / 10 running sums per problem are registers).  Rows per
  // warp (non-check launches only; check launches keep 8 so that the partial sums have a fixed shape): as
  // many as leave about two waves of warps - with wide lanes a launch has few, long warps otherwise.
  auto width_for = [&](int cap) {
    const int live_lanes = chunks * 32;
    for (int P : {4, 2})
      if (P <= cap && P <= max_width && live_lanes > 16 * P && (live_lanes + 32 * P - 1) / (32 * P) * (32 * P) <= BP)
        return P;
    return 1;
  };
  auto rpw_for = [&](int rows, int gy) {
    if (fixed_rpw) return fixed_rpw;
    int r = kRowsPerCtaB / kWarps;
    while (r > 1 && (long long)rows * gy / r < 8192) r >>= 1;
    return r;
  };
  auto enqueue_block = [&]() {
    const int live_lanes = chunks * 32;
    const int Pr = width_for(long_rows ? 2 : fed ? 4 : 2), Pj = width_for(fed_col ? 4 : 2), Pc = width_for(2);
    auto groups = [&](int P) { return (live_lanes + 32 * P - 1) / (32 * P); };
    const int gyr = groups(Pr), gyj = groups(Pj), gyc = groups(Pc);
    const int rpw_row = rpw_for(m, gyr), rpw_col = rpw_for(n, gyj);
    for (int i = 0; i < period; ++i) {
      const bool first = i == 0, last = i == period - 1;
      const bool timed = i == (period > 2 ? 1 : 0);
      const bool chk = first || last;
      if (timed) MIPCU_CK(cudaEventRecord(ev_t[0], st));
      if (m && long_rows) {
        if (chk) {
          const dim3 g(m, gyc);
          const int keep = last ? 1 : 0;
          with_width(Pc, [&](auto w) { k_row_batch_long<true, 0, decltype(w)::value><<<g, kThreads, 0, st>>>(a, i, keep); });
        } else {
          const dim3 g(m, gyr);
          with_width(Pr, [&](auto w) { k_row_batch_long<false, 0, decltype(w)::value><<<g, kThreads, 0, st>>>(a, i, 0); });
        }
        ctx->launches++;
        if (last) {
          const dim3 g(m, gyc);
          with_width(Pc, [&](auto w) { k_row_batch_long<false, 1, decltype(w)::value><<<g, kThreads, 0, st>>>(a, i, 0); });
          ctx->launches++;
        }
      } else if (m) {
        if (chk) {
          const dim3 g(std::max(nblk_row, 1), gyc);
          const int keep = last ? 1 : 0;
          with_width(Pc, [&](auto w) {
            k_row_step_batch<true, decltype(w)::value><<<g, kThreads, 0, st>>>(a, i, keep, kRowsPerCtaB / kWarps);
          });
        } else {
          const dim3 g(std::max(grid_for(m, kWarps * rpw_row), 1), gyr);
          with_width(Pr, [&](auto w) {
            if (fed) k_row_step_batch_fed<decltype(w)::value><<<g, kThreads, 0, st>>>(a, i, rpw_row);
            else k_row_step_batch<false, decltype(w)::value><<<g, kThreads, 0, st>>>(a, i, 0, rpw_row);
          });
        }
        ctx->launches++;
        if (last) {
          const dim3 g(std::max(nblk_row, 1), gyc);
          with_width(Pc, [&](auto w) { k_kkt_row_batch<decltype(w)::value><<<g, kThreads, 0, st>>>(a); });
          ctx->launches++;
        }
      }
      if (timed) MIPCU_CK(cudaEventRecord(ev_t[1], st));
      const double* xin = first ? a.xhat_first : a.xhat_chk;
      double* xo = last ? a.xhat_first : (i == period - 2 ? a.xhat_chk : nullptr);
      if (n) {
        if (chk) {
          const dim3 g(std::max(nblk_col, 1), gyc);
          with_width(Pc, [&](auto w) {
            constexpr int P = decltype(w)::value;
            if (xo) k_col_step_batch<true, true, P><<<g, kThreads, 0, st>>>(a, i, xin, xo, kRowsPerCtaB / kWarps);
            else k_col_step_batch<true, false, P><<<g, kThreads, 0, st>>>(a, i, xin, xo, kRowsPerCtaB / kWarps);
          });
        } else {
          const dim3 g(std::max(grid_for(n, kWarps * rpw_col), 1), gyj);
          with_width(Pj, [&](auto w) {
            constexpr int P = decltype(w)::value;
            if (fed_col && xo) k_col_step_batch_fed<true, P><<<g, kThreads, 0, st>>>(a, i, xo, rpw_col);
            else if (fed_col) k_col_step_batch_fed<false, P><<<g, kThreads, 0, st>>>(a, i, xo, rpw_col);
            else if (xo) k_col_step_batch<false, true, P><<<g, kThreads, 0, st>>>(a, i, xin, xo, rpw_col);
            else k_col_step_batch<false, false, P><<<g, kThreads, 0, st>>>(a, i, xin, xo, rpw_col);
          });
        }
        ctx->launches++;
      }
      if (timed) MIPCU_CK(cudaEventRecord(ev_t[2], st));
      if (first) { k_ctrl_first_batch<<<chunks, kThreads, 0, st>>>(a, nblk_row, nblk_col); ctx->launches++; }
    }
    k_ctrl_check_batch<<<chunks, kThreads, 0, st>>>(a, nblk_row, nblk_col);
    k_publish_batch<<<grid_for(BP), kThreads, 0, st>>>(a);
    ctx->launches += 2;
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
    k_publish_batch<<<grid_for(BP), kThreads, 0, st>>>(a);
    ctx->launches += 2;
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
  lap("iterations");
  if (x_out && nB) {
    double* d_x = buf.get<double>((size_t)B * n);
    k_unscale_batch<<<grid_for(nB), kThreads, 0, st>>>(a.xhat_chk, ctx->d2, d_node_of, ctx->col_perm, d_x, n, B, BP);
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
  lap("solutions read back");
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
