// kernels_stream.cuh - "CSR-stream" form of the two per-iteration kernels.
//
// profiles/r01_k_col_step_v1_ncu_full.md: the sub-warp-per-row kernels are bound by the per-SM L1
// miss-request port (one 128-byte-line request per gathered double) and keep it 76 % busy; the
// rest is lost because a lane must wait for its (val, idx) pair before it can issue the gather, and
// because rows of different length make the four sub-warps of a warp diverge.
//
// Here the CTA's tile of 256 rows is treated as ONE contiguous run of nonzeros:
//   phase 1: thread t multiplies entries t, t+256, ... of the run by the gathered vector entry and
//            parks the products in shared memory - perfectly coalesced streams, no dependence on
//            the row structure, 8 independent gathers in flight per thread;
//   phase 2: sub-warps sum each row's segment of products out of shared memory;
//   phase 3: the same coalesced epilogue as kernels.cuh.
// Results differ from kernels.cuh only by rounding (product then add instead of fma).
#pragma once

#include "internal.h"
#include "kernels.cuh"

namespace mipcu {

constexpr int kStreamUnroll = 8;

// phases 1 and 2; s_prod has room for the tile's nonzeros, s_sum for kRowsPerCta row sums
template <int TPR>
__device__ __forceinline__ void stream_row_sums(const int* __restrict__ ptr, const int* __restrict__ idx,
                                                const double* __restrict__ val,
                                                const double* __restrict__ vec, int nrows,
                                                double* s_prod, double* s_sum) {
  const int row0 = blockIdx.x * kRowsPerCta;
  const int row1 = min(row0 + kRowsPerCta, nrows);
  const int k0 = ptr[row0], k1 = ptr[row1];
  // ---- phase 1: products of the whole run -------------------------------------------------------
  int k = k0 + threadIdx.x;
  for (; k + (kStreamUnroll - 1) * kThreads < k1; k += kStreamUnroll * kThreads) {
    double v[kStreamUnroll];
    int c[kStreamUnroll];
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) {
      v[u] = ld_stream(val + k + u * kThreads);
      c[u] = ld_stream(idx + k + u * kThreads);
    }
    double g[kStreamUnroll];
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) g[u] = __ldg(vec + c[u]);
#pragma unroll
    for (int u = 0; u < kStreamUnroll; ++u) s_prod[k - k0 + u * kThreads] = v[u] * g[u];
  }
  for (; k < k1; k += kThreads) s_prod[k - k0] = ld_stream(val + k) * __ldg(vec + ld_stream(idx + k));
  __syncthreads();
  // ---- phase 2: per-row sums of the products ----------------------------------------------------
  constexpr int SUBS = kThreads / TPR;
  const int lane = threadIdx.x % TPR;
  const int sub = threadIdx.x / TPR;
#pragma unroll 2
  for (int r = sub; r < kRowsPerCta; r += SUBS) {
    const int row = row0 + r;
    double acc = 0.0;
    if (row < nrows) {
      const int b = ptr[row] - k0, e = ptr[row + 1] - k0;
      for (int j = b + lane; j < e; j += TPR) acc += s_prod[j];
    }
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, TPR);
    if (lane == 0) s_sum[r] = acc;
  }
  __syncthreads();
}

template <int TPR, bool CHECK>
__global__ void __launch_bounds__(kThreads)
k_row_step_stream(const RowStepArgs a) {
  extern __shared__ __align__(16) double s_dyn[];
  double* s_sum = s_dyn;
  double* s_prod = s_dyn + kRowsPerCta;
  const Scalars* sc = a.sc;
  if (sc->done) return;
  stream_row_sums<TPR>(a.ptr, a.idx, a.val, a.xbar, a.m, s_prod, s_sum);
  const double sigma = sc->sigma;
  const double k = (double)(sc->k_base + a.iter_in_block);
  const double w_new = (k + 1.0) / (k + 2.0);
  const int i = blockIdx.x * kRowsPerCta + threadIdx.x;
  double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (i < a.m) {
    const double kx = s_sum[threadIdx.x];
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
      const double dp = fmax(dy, 0.0), dm = fmax(-dy, 0.0);
      red[3] = (isfinite(lo) ? lo * dp : 0.0) - (isfinite(hi) ? hi * dm : 0.0);
      const double rv = (isfinite(lo) ? 0.0 : dp) + (isfinite(hi) ? 0.0 : dm);
      red[4] = rv * rv;
      if (a.kx_save) a.kx_save[i] = kx;
    }
  }
  if (CHECK) {
    const int q[5] = {Q_DY2, Q_DY0, Q_DOBJ_ROW, Q_RAYD_ROW, Q_RAYD_ROWV};
    cta_reduce_store<5>(red, q, a.partials, sc->max_blocks);
  }
}

template <int TPR, bool CHECK, bool STORE>
__global__ void __launch_bounds__(kThreads)
k_col_step_stream(const ColStepArgs a) {
  extern __shared__ __align__(16) double s_dyn[];
  double* s_sum = s_dyn;
  double* s_prod = s_dyn + kRowsPerCta;
  const Scalars* sc = a.sc;
  if (sc->done) return;
  stream_row_sums<TPR>(a.ptr, a.idx, a.val, a.yhat, a.n, s_prod, s_sum);
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

inline size_t stream_smem_bytes(int cap) { return sizeof(double) * ((size_t)cap + kRowsPerCta); }

}  // namespace mipcu
