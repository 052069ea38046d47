// kernels.cuh - the per-iteration kernels of the restarted reflected-Halpern PDHG method.
//
// One PDHG iteration is TWO kernel launches and nothing else:
//   k_row_step : (K xbar)_i by a sub-warp per row, then - in the same CTA, coalesced over the
//                CTA's 256 rows - the dual proximal step, the Halpern/reflection update of y and,
//                on check iterations, the per-CTA partial sums the restart test needs;
//   k_col_step : (K' yhat)_j by a sub-warp per column of the CSC copy, then the Halpern update of
//                x and K'y, the projected primal step and the reflected point xbar that the next
//                k_row_step gathers.
// All scalars (tau, sigma, the Halpern weight) are read from the device-resident control block,
// so a block of `period` iterations runs with no host involvement and can be graph-captured.
//
// The arithmetic is restated line for line in oracle/pdhg_ref.py (row_step / col_step /
// kkt_scaled); tests/test_parity_gpu.py compares the two.
#pragma once

#include "internal.h"

namespace mipcu {

__device__ __forceinline__ double ld_stream(const double* p) {
  // matrix values are touched once per launch: keep them out of L1
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// Programmatic dependent launch (both no-ops in a kernel launched without the attribute).
// pdl_trigger: the next kernel of the stream may start filling SM slots this grid no longer needs;
// pdl_wait: everything the previous kernels wrote is visible from here on.  The iteration kernels
// read only the (constant) matrix before pdl_wait, so the head of their nonzero stream is in
// flight while the previous kernel drains.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Host side: launch `kernel` on kThreads-wide CTAs, optionally as a programmatic dependent of the
// previous kernel in the stream.
template <class... Params, class... Args>
inline void launch_kernel(void (*kernel)(Params...), int grid, cudaStream_t stream, bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  MIPCU_CK(cudaLaunchKernelEx(&cfg, kernel, Params(args)...));
}

// `pre` hook of cta_row_sums for kernels with nothing to wait for
struct NoWait {
  __device__ __forceinline__ bool operator()() const { return true; }
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) {
  return fmin(fmax(v, lo), hi);
}

// Sum of A[row, :] * vec for the kRowsPerCta rows owned by this CTA, left in s_sum[0..kRowsPerCta).
//
// Two storage formats share this entry point (the template parameter selects one):
//
//  TPR == 0: sliced ELL ("SELL-32-256", built by setup.cu:build_sell once the matrix is scaled).
//    The 256 rows of a CTA are sorted by length; 32 consecutive rows of that order form a slice
//    whose entries are stored lane-major (entry j of lane l at ptr[slice] + 32 j + l, padded to
//    the slice's longest row with idx = -1).  One lane owns one row: the val / idx streams are
//    perfectly coalesced, there is no shuffle reduction, no row-pointer dependency and no
//    divergence inside a warp; kSellUnroll independent (val, idx, gather) chains per lane cover
//    the L2 gather latency.  perm[] maps the sorted position back to the CTA-local row, so the
//    vectors of the epilogue keep their natural order.  Measured on config 3 (scripts/spmv_lab.cu,
//    profiles/r01_spmv_lab.md): row step 174 -> 140 us, col step 165 -> 137 us.
//
//  TPR in {2,4,8,16,32}: CSR with TPR lanes cooperating on one row (long or few rows, where one
//    lane per row would leave the machine empty); a warp streams 32/TPR consecutive rows, two
//    independent chains per lane.
constexpr int kSellUnroll = 8;

//
// `pre()` is called by every thread exactly once, after the first loads of the matrix stream have
// been issued and before anything of `vec` is read: the place to wait for the producer of `vec`
// (programmatic dependent launch, a cross-GPU barrier).  It returns false when the kernel has
// nothing to do (solve finished); the function then returns false with s_sum undefined.
template <int TPR, class Pre>
__device__ __forceinline__ bool cta_row_sums(const int* __restrict__ ptr,
                                             const int* __restrict__ idx,
                                             const double* __restrict__ val,
                                             const unsigned char* __restrict__ perm,
                                             const double* __restrict__ vec, int nrows,
                                             double* s_sum, Pre pre, const int blk) {
  if constexpr (TPR == 0) {
    constexpr int U = kSellUnroll;
    const int lane = threadIdx.x & 31;
    const int slice = blk * (kRowsPerCta / 32) + (threadIdx.x >> 5);
    const int beg = ptr[slice], end = ptr[slice + 1];
    const int slot = perm[blk * kRowsPerCta + threadIdx.x];
    double acc0 = 0.0, acc1 = 0.0;
    double v[U];
    int c[U];
    int p = beg + lane;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = p + 32 * u < end;       // warp-uniform: a slice is padded to whole lanes
      v[u] = ok ? ld_stream(val + p + 32 * u) : 0.0;
      c[u] = ok ? ld_stream(idx + p + 32 * u) : -1;
    }
    if (!pre()) return false;
    while (true) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double x = c[u] >= 0 ? __ldg(vec + c[u]) : 0.0;
        if (u & 1) acc1 = fma(v[u], x, acc1); else acc0 = fma(v[u], x, acc0);
      }
      p += 32 * U;
      if (p - lane >= end) break;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = p + 32 * u < end;
        v[u] = ok ? ld_stream(val + p + 32 * u) : 0.0;
        c[u] = ok ? ld_stream(idx + p + 32 * u) : -1;
      }
    }
    s_sum[slot] = acc0 + acc1;
    __syncthreads();
  } else {
    if (!pre()) return false;
    constexpr int SUBS = kThreads / TPR;        // rows in flight per pass
    const int lane = threadIdx.x % TPR;
    const int sub = threadIdx.x / TPR;
    const int row0 = blk * kRowsPerCta;
#pragma unroll 2
    for (int r = sub; r < kRowsPerCta; r += SUBS) {
      const int row = row0 + r;
      double acc0 = 0.0, acc1 = 0.0;
      if (row < nrows) {
        const int b = ptr[row], e = ptr[row + 1];
        int k = b + lane;
        for (; k + TPR < e; k += 2 * TPR) {
          const double v0 = ld_stream(val + k);
          const int c0 = ld_stream(idx + k);
          const double v1 = ld_stream(val + k + TPR);
          const int c1 = ld_stream(idx + k + TPR);
          acc0 = fma(v0, __ldg(vec + c0), acc0);
          acc1 = fma(v1, __ldg(vec + c1), acc1);
        }
        if (k < e) acc0 = fma(ld_stream(val + k), __ldg(vec + ld_stream(idx + k)), acc0);
      }
      double acc = acc0 + acc1;
#pragma unroll
      for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, TPR);
      if (lane == 0) s_sum[r] = acc;
    }
    __syncthreads();
  }
  return true;
}

template <int TPR, class Pre>
__device__ __forceinline__ bool cta_row_sums(const int* __restrict__ ptr, const int* __restrict__ idx,
                                             const double* __restrict__ val,
                                             const unsigned char* __restrict__ perm,
                                             const double* __restrict__ vec, int nrows, double* s_sum, Pre pre) {
  return cta_row_sums<TPR>(ptr, idx, val, perm, vec, nrows, s_sum, pre, (int)blockIdx.x);
}

template <int TPR>
__device__ __forceinline__ void cta_row_sums(const int* __restrict__ ptr, const int* __restrict__ idx,
                                             const double* __restrict__ val,
                                             const unsigned char* __restrict__ perm,
                                             const double* __restrict__ vec, int nrows, double* s_sum) {
  cta_row_sums<TPR>(ptr, idx, val, perm, vec, nrows, s_sum, NoWait());
}

// Deterministic CTA reduction of NQ running sums; thread 0 stores them to
// partials[(q0 + q) * max_blocks + blockIdx.x].
template <int NQ>
__device__ __forceinline__ void cta_reduce_store(double (&v)[NQ], const int (&q)[NQ],
                                                 double* partials, int max_blocks, const int blk) {
  __shared__ double s_red[NQ][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    double a = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) s_red[i][warp] = a;
  }
  __syncthreads();
  if (threadIdx.x < NQ) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) a += s_red[threadIdx.x][w];
    partials[(size_t)q[threadIdx.x] * max_blocks + blk] = a;
  }
}

template <int NQ>
__device__ __forceinline__ void cta_reduce_store(double (&v)[NQ], const int (&q)[NQ],
                                                 double* partials, int max_blocks) {
  cta_reduce_store<NQ>(v, q, partials, max_blocks, (int)blockIdx.x);
}

struct RowStepArgs {
  const int* ptr; const int* idx; const double* val;   // K by rows (CSR, or SELL when TPR == 0)
  const unsigned char* perm;                            // SELL: sorted position -> CTA-local row
  const double* xbar;                                   // gathered
  const double* lo; const double* hi; const double* y0;
  double* y; double* yhat;
  const Scalars* sc; double* partials;
  int m; int iter_in_block;
  double* kx_save;   // CHECK, last iteration of a block: K xbar kept for the primal-ray test (or null)
};

// The per-row part of the row step, given (K xbar)_i: dual proximal step, Halpern / reflection
// update of y; on CHECK iterations the five partial sums of the restart / certificate tests.
// Returns yhat_i.  Shared by the single-GPU kernel and the block-owner kernel of p2p.cu.
template <bool CHECK>
__device__ __forceinline__ double row_epilogue(const RowStepArgs& a, const int i, const double kx,
                                               const double sigma, const double w_new, double (&red)[5]) {
  const double yi = a.y[i], y0i = a.y0[i], lo = a.lo[i], hi = a.hi[i];
  const double t = kx - yi / sigma;
  const double yh = sigma * (clampd(t, lo, hi) - t);
  a.yhat[i] = yh;
  a.y[i] = w_new * (2.0 * yh - yi) + (1.0 - w_new) * y0i;
  if (CHECK) {
    const double dy = yh - yi, d0 = yh - y0i;
    red[0] = dy * dy;
    red[1] = d0 * d0;
    const double yp = fmax(yh, 0.0), ym = fmax(-yh, 0.0);
    red[2] = (isfinite(lo) ? lo * yp : 0.0) - (isfinite(hi) ? hi * ym : 0.0);
    // dy as a candidate Farkas ray
    const double dp = fmax(dy, 0.0), dm = fmax(-dy, 0.0);
    red[3] = (isfinite(lo) ? lo * dp : 0.0) - (isfinite(hi) ? hi * dm : 0.0);
    const double rv = (isfinite(lo) ? 0.0 : dp) + (isfinite(hi) ? 0.0 : dm);
    red[4] = rv * rv;
    if (a.kx_save) a.kx_save[i] = kx;
  }
  return yh;
}

template <int TPR, bool CHECK>
__global__ void __launch_bounds__(kThreads)
k_row_step(const RowStepArgs a) {
  __shared__ double s_sum[kRowsPerCta];
  const Scalars* sc = a.sc;
  pdl_trigger();
  if (!cta_row_sums<TPR>(a.ptr, a.idx, a.val, a.perm, a.xbar, a.m, s_sum,
                         [&] { pdl_wait(); return !sc->done; }))
    return;
  const double sigma = sc->sigma;
  const double k = (double)(sc->k_base + a.iter_in_block);
  const double w_new = (k + 1.0) / (k + 2.0);
  const int i = blockIdx.x * kRowsPerCta + threadIdx.x;
  double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < a.m) row_epilogue<CHECK>(a, i, s_sum[threadIdx.x], sigma, w_new, red);
  if (CHECK) {
    const int q[5] = {Q_DY2, Q_DY0, Q_DOBJ_ROW, Q_RAYD_ROW, Q_RAYD_ROWV};
    cta_reduce_store<5>(red, q, a.partials, sc->max_blocks);
  }
}

struct ColStepArgs {
  const int* ptr; const int* idx; const double* val;   // K by columns (CSR of K', or SELL)
  const unsigned char* perm;
  const double* yhat;                                   // gathered
  const double* x0; const double* aty0; const double* c; const double* l; const double* u;
  const double* d2;
  double* xbar; double* aty;
  const double* xhat_in;     // CHECK: xhat of this iteration (stored by the previous col step)
  double* xhat_out;          // STORE: xhat of the NEXT iteration
  double* aty_cand;          // CHECK: K' yhat (restart candidate)
  const Scalars* sc; double* partials;
  int n; int iter_in_block;
};

// The per-column part of the col step, given (K' yhat)_j: Halpern update of x and K'y, projected
// primal step, reflected point; on CHECK iterations the six partial sums of the restart / KKT test.
// Shared by the fused single-GPU kernel and by k_primal_step of the row-sharded path.
template <bool CHECK, bool STORE>
__device__ __forceinline__ void col_epilogue(const ColStepArgs& a, const int j, const double atyh,
                                             const double tau, const double w_new, double (&red)[10]) {
  const double xb = a.xbar[j], x0 = a.x0[j], at = a.aty[j], at0 = a.aty0[j];
  const double cj = a.c[j], lj = a.l[j], uj = a.u[j];
  if (CHECK) {
    const double xh = a.xhat_in[j];
    const double dx = xb - xh;            // = xhat_k - x_k  because xbar = 2 xhat - x
    const double dat = atyh - at;
    const double d0 = xh - x0;
    red[0] = dx * dx;
    red[1] = dx * dat;
    red[2] = d0 * d0;
    const double r = cj - atyh;
    const double rp = fmax(r, 0.0), rm = fmax(-r, 0.0);
    const double viol = ((isfinite(lj) ? 0.0 : rp) + (isfinite(uj) ? 0.0 : rm)) / a.d2[j];
    red[3] = viol * viol;
    red[4] = cj * xh;
    red[5] = (isfinite(lj) ? lj * rp : 0.0) - (isfinite(uj) ? uj * rm : 0.0);
    a.aty_cand[j] = atyh;
    // dual ray dy: its reduced costs are -K'dy = -(K'yhat - K'y)
    const double qp = fmax(-dat, 0.0), qm = fmax(dat, 0.0);
    red[6] = (isfinite(lj) ? lj * qp : 0.0) - (isfinite(uj) ? uj * qm : 0.0);
    const double qv = (isfinite(lj) ? 0.0 : qp) + (isfinite(uj) ? 0.0 : qm);
    red[7] = qv * qv;
    // primal ray dx: objective along it, and its components that leave a bounded side
    red[8] = cj * dx;
    const double pv = (isfinite(lj) ? fmax(-dx, 0.0) : 0.0) + (isfinite(uj) ? fmax(dx, 0.0) : 0.0);
    red[9] = pv * pv;
  }
  const double xn = w_new * xb + (1.0 - w_new) * x0;
  const double atn = w_new * (2.0 * atyh - at) + (1.0 - w_new) * at0;
  const double xh2 = clampd(xn - tau * (cj - atn), lj, uj);
  a.xbar[j] = 2.0 * xh2 - xn;
  a.aty[j] = atn;
  if (STORE) a.xhat_out[j] = xh2;
}

template <int TPR, bool CHECK, bool STORE>
__global__ void __launch_bounds__(kThreads)
k_col_step(const ColStepArgs a) {
  __shared__ double s_sum[kRowsPerCta];
  const Scalars* sc = a.sc;
  pdl_trigger();
  if (!cta_row_sums<TPR>(a.ptr, a.idx, a.val, a.perm, a.yhat, a.n, s_sum,
                         [&] { pdl_wait(); return !sc->done; }))
    return;
  const double tau = sc->tau;
  const double k = (double)(sc->k_base + a.iter_in_block);
  const double w_new = (k + 1.0) / (k + 2.0);
  const int j = blockIdx.x * kRowsPerCta + threadIdx.x;
  double red[10] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (j < a.n) col_epilogue<CHECK, STORE>(a, j, s_sum[threadIdx.x], tau, w_new, red);
  if (CHECK) {
    const int q[10] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL,
                       Q_RAYD_COL, Q_RAYD_COLV, Q_RAYP_C, Q_RAYP_COLV};
    cta_reduce_store<10>(red, q, a.partials, sc->max_blocks);
  }
}

// Row-sharded path: K' yhat arrives already summed over the ranks (NCCL all-reduce of the local
// partial products), so the primal side is a plain elementwise kernel with the same arithmetic
// and the same CTA-to-column mapping (hence the same partial sums) as the fused kernel.
template <bool CHECK, bool STORE>
__global__ void __launch_bounds__(kThreads)
k_primal_step(const ColStepArgs a, const double* __restrict__ atyh) {
  const Scalars* sc = a.sc;
  if (sc->done) return;
  const double tau = sc->tau;
  const double k = (double)(sc->k_base + a.iter_in_block);
  const double w_new = (k + 1.0) / (k + 2.0);
  const int j = blockIdx.x * kRowsPerCta + threadIdx.x;
  double red[10] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (j < a.n) col_epilogue<CHECK, STORE>(a, j, atyh[j], tau, w_new, red);
  if (CHECK) {
    const int q[10] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL,
                       Q_RAYD_COL, Q_RAYD_COLV, Q_RAYP_C, Q_RAYP_COLV};
    cta_reduce_store<10>(red, q, a.partials, sc->max_blocks);
  }
}

// Primal residual of xhat on check iterations: sum ((K xhat - clamp(K xhat, lo, hi)) / d1)^2.
template <int TPR>
__global__ void __launch_bounds__(kThreads)
k_kkt_row(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
          const unsigned char* __restrict__ perm, const double* __restrict__ xhat, const double* __restrict__ lo,
          const double* __restrict__ hi, const double* __restrict__ d1, const Scalars* sc,
          double* partials, int m, const double* __restrict__ kxbar) {
  __shared__ double s_sum[kRowsPerCta];
  if (sc->done) return;
  cta_row_sums<TPR>(ptr, idx, val, perm, xhat, m, s_sum);
  const int i = blockIdx.x * kRowsPerCta + threadIdx.x;
  double red[2] = {0.0, 0.0};
  if (i < m) {
    const double kx = s_sum[threadIdx.x];
    const double v = (kx - clampd(kx, lo[i], hi[i])) / d1[i];
    red[0] = v * v;
    // K dx = K xbar - K xhat (dx = xhat - x = xbar - xhat): must stay inside the rows' recession cone
    const double kdx = kxbar[i] - kx;
    const double rv = (isfinite(lo[i]) ? fmax(-kdx, 0.0) : 0.0) + (isfinite(hi[i]) ? fmax(kdx, 0.0) : 0.0);
    red[1] = rv * rv;
  }
  const int q[2] = {Q_PRES2, Q_RAYP_ROWV};
  cta_reduce_store<2>(red, q, partials, sc->max_blocks);
}

// Plain out = A v through the same row-sum routine (setup, power iteration, test hook).
template <int TPR>
__global__ void __launch_bounds__(kThreads)
k_spmv(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
       const unsigned char* __restrict__ perm, const double* __restrict__ v, double* __restrict__ out, int nrows) {
  __shared__ double s_sum[kRowsPerCta];
  cta_row_sums<TPR>(ptr, idx, val, perm, v, nrows, s_sum);
  const int i = blockIdx.x * kRowsPerCta + threadIdx.x;
  if (i < nrows) out[i] = s_sum[threadIdx.x];
}

}  // namespace mipcu
