// kernels_tma.cuh - the two per-iteration kernels again, with the STREAMING half of their traffic
// (val / idx of the CSR tile and the epilogue vectors) moved by the TMA engine into a
// shared-memory ring (cp.async.bulk + mbarrier), so that the LSU path carries only the x / yhat
// gathers and the stores.
//
// Why (profiles/r01_k_col_step_v1_ncu_full.md): on uniformly random columns the LSU kernels are
// bound by the per-SM L1 miss-request port (~1 line request per cycle), not by HBM.  In v1 the same
// warps alternate between waiting ~1 us for the HBM stream and issuing gathers, so the port idles
// 24 % of the time.  Here one elected thread keeps the next tile's stream in flight through the
// async proxy while every warp issues gathers back to back out of shared memory.
//
// Persistent CTAs (grid = SMs x resident CTAs), static round-robin over tiles of kTileRows rows.
// Arithmetic and summation order are identical to kernels.cuh (same TPR, same lane order), so the
// two variants produce bit-identical iterates (tested).
#pragma once

#include <cstdint>

#include "internal.h"
#include "kernels.cuh"

namespace mipcu {

constexpr int kTileRows = 64;     // rows per tile
constexpr int kStages = 2;        // ring depth

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// One stage of the ring: val[cap] | idx[cap] | NV vectors of kTileRows doubles
template <int NV>
struct StageView {
  double* val;
  int* idx;
  double* vec[NV > 0 ? NV : 1];
  __device__ StageView(unsigned char* base, int cap) {
    val = reinterpret_cast<double*>(base);
    idx = reinterpret_cast<int*>(base + (size_t)cap * 8);
    unsigned char* v = base + (size_t)cap * 12;
#pragma unroll
    for (int i = 0; i < NV; ++i) vec[i] = reinterpret_cast<double*>(v + (size_t)i * kTileRows * 8);
  }
};

inline size_t tma_stage_bytes(int cap, int nv) { return (size_t)cap * 12 + (size_t)nv * kTileRows * 8; }
inline size_t tma_smem_bytes(int cap, int nv) { return kStages * tma_stage_bytes(cap, nv) + 64; }

// issued by ONE thread: the matrix slice of a tile plus NV per-row vectors
template <int NV>
__device__ __forceinline__ void issue_tile(int tile, int nrows, const int* __restrict__ ptr,
                                           const int* __restrict__ idx, const double* __restrict__ val,
                                           const double* const (&vecs)[NV > 0 ? NV : 1],
                                           StageView<NV>& st, uint64_t* bar, int* k0_out) {
  const int r0 = tile * kTileRows;
  const int r1 = min(r0 + kTileRows, nrows);
  const int kb = ptr[r0] & ~3;                       // 16-byte aligned start of both streams
  const int cnt = ((ptr[r1] - kb) + 3) & ~3;
  const int nr = ((r1 - r0) + 1) & ~1;               // vectors carry 2 doubles of padding
  mbar_expect_tx(bar, (uint32_t)cnt * 12u + (uint32_t)NV * (uint32_t)nr * 8u);
  if (cnt) {
    bulk_g2s(st.val, val + kb, (uint32_t)cnt * 8u, bar);
    bulk_g2s(st.idx, idx + kb, (uint32_t)cnt * 4u, bar);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) bulk_g2s(st.vec[i], vecs[i] + r0, (uint32_t)nr * 8u, bar);
  *k0_out = kb;
}

// row sums of one tile out of shared memory; gathers are the only global loads
template <int TPR>
__device__ __forceinline__ void tile_row_sums(const int* __restrict__ ptr, const double* s_val,
                                              const int* s_idx, int kbase, const double* __restrict__ vec,
                                              int r0, int nrows, double* s_sum) {
  constexpr int SUBS = kThreads / TPR;
  const int lane = threadIdx.x % TPR;
  const int sub = threadIdx.x / TPR;
#pragma unroll
  for (int r = sub; r < kTileRows; r += SUBS) {
    const int row = r0 + r;
    double acc0 = 0.0, acc1 = 0.0;
    if (row < nrows) {
      const int b = ptr[row] - kbase, e = ptr[row + 1] - kbase;
      int k = b + lane;
      for (; k + TPR < e; k += 2 * TPR) {
        const int c0 = s_idx[k], c1 = s_idx[k + TPR];
        const double g0 = __ldg(vec + c0), g1 = __ldg(vec + c1);
        acc0 = fma(s_val[k], g0, acc0);
        acc1 = fma(s_val[k + TPR], g1, acc1);
      }
      if (k < e) acc0 = fma(s_val[k], __ldg(vec + s_idx[k]), acc0);
    }
    double acc = acc0 + acc1;
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, TPR);
    if (lane == 0) s_sum[r] = acc;
  }
}

template <int NQ>
__device__ __forceinline__ void tile_reduce_store(double (&v)[NQ], const int (&q)[NQ], double* partials,
                                                  int max_blocks, int tile, double (*s_red)[kThreads / 32]) {
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
    partials[(size_t)q[threadIdx.x] * max_blocks + tile] = a;
  }
}

// ---- row step -----------------------------------------------------------------------------------
template <int TPR, bool CHECK>
__global__ void __launch_bounds__(kThreads)
k_row_step_tma(const RowStepArgs a, const int cap, const int ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kStages];
  __shared__ int s_k0[kStages];
  __shared__ double s_sum[kTileRows];
  __shared__ double s_red[3][kThreads / 32];
  const Scalars* sc = a.sc;
  if (sc->done) return;
  constexpr int NV = 4;   // y, y0, lo, hi
  const double* const vecs[NV] = {a.y, a.y0, a.lo, a.hi};
  const size_t stage_bytes = (size_t)cap * 12 + (size_t)NV * kTileRows * 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double sigma = sc->sigma;
  const double kk = (double)(sc->k_base + a.iter_in_block);
  const double w_new = (kk + 1.0) / (kk + 2.0);
  int it = 0;
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) {
    StageView<NV> st(smem, cap);
    issue_tile<NV>(blockIdx.x, a.m, a.ptr, a.idx, a.val, vecs, st, &full[0], &s_k0[0]);
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const int next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) {
      const int sn = (it + 1) % kStages;
      StageView<NV> stn(smem + sn * stage_bytes, cap);
      issue_tile<NV>(next, a.m, a.ptr, a.idx, a.val, vecs, stn, &full[sn], &s_k0[sn]);
    }
    mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
    __syncthreads();   // s_k0[s] written by thread 0 before the issue of this stage
    StageView<NV> st(smem + s * stage_bytes, cap);
    const int r0 = tile * kTileRows;
    tile_row_sums<TPR>(a.ptr, st.val, st.idx, s_k0[s], a.xbar, r0, a.m, s_sum);
    __syncthreads();
    double red[3] = {0.0, 0.0, 0.0};
    const int t = threadIdx.x;
    if (t < kTileRows && r0 + t < a.m) {
      const int i = r0 + t;
      const double kx = s_sum[t];
      const double yi = st.vec[0][t], y0i = st.vec[1][t], lo = st.vec[2][t], hi = st.vec[3][t];
      const double tt = kx - yi / sigma;
      const double yh = sigma * (clampd(tt, lo, hi) - tt);
      a.yhat[i] = yh;
      a.y[i] = w_new * (2.0 * yh - yi) + (1.0 - w_new) * y0i;
      if (CHECK) {
        const double dy = yh - yi, d0 = yh - y0i;
        red[0] = dy * dy;
        red[1] = d0 * d0;
        red[2] = (isfinite(lo) ? lo * fmax(yh, 0.0) : 0.0) - (isfinite(hi) ? hi * fmax(-yh, 0.0) : 0.0);
      }
    }
    if (CHECK) {
      const int q[3] = {Q_DY2, Q_DY0, Q_DOBJ_ROW};
      tile_reduce_store<3>(red, q, a.partials, sc->max_blocks, tile, s_red);
    }
    __syncthreads();   // everyone is done with stage s before it is refilled (two iterations later)
  }
}

// ---- col step -----------------------------------------------------------------------------------
template <int TPR, bool CHECK, bool STORE>
__global__ void __launch_bounds__(kThreads)
k_col_step_tma(const ColStepArgs a, const int cap, const int ntiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full[kStages];
  __shared__ int s_k0[kStages];
  __shared__ double s_sum[kTileRows];
  __shared__ double s_red[6][kThreads / 32];
  const Scalars* sc = a.sc;
  if (sc->done) return;
  constexpr int NV = 7;   // xbar, x0, aty, aty0, c, l, u
  const double* const vecs[NV] = {a.xbar, a.x0, a.aty, a.aty0, a.c, a.l, a.u};
  const size_t stage_bytes = (size_t)cap * 12 + (size_t)NV * kTileRows * 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double tau = sc->tau;
  const double kk = (double)(sc->k_base + a.iter_in_block);
  const double w_new = (kk + 1.0) / (kk + 2.0);
  int it = 0;
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) {
    StageView<NV> st(smem, cap);
    issue_tile<NV>(blockIdx.x, a.n, a.ptr, a.idx, a.val, vecs, st, &full[0], &s_k0[0]);
  }
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % kStages;
    const int next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) {
      const int sn = (it + 1) % kStages;
      StageView<NV> stn(smem + sn * stage_bytes, cap);
      issue_tile<NV>(next, a.n, a.ptr, a.idx, a.val, vecs, stn, &full[sn], &s_k0[sn]);
    }
    mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
    __syncthreads();
    StageView<NV> st(smem + s * stage_bytes, cap);
    const int r0 = tile * kTileRows;
    tile_row_sums<TPR>(a.ptr, st.val, st.idx, s_k0[s], a.yhat, r0, a.n, s_sum);
    __syncthreads();
    double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const int t = threadIdx.x;
    if (t < kTileRows && r0 + t < a.n) {
      const int j = r0 + t;
      const double atyh = s_sum[t];
      const double xb = st.vec[0][t], x0 = st.vec[1][t], at = st.vec[2][t], at0 = st.vec[3][t];
      const double cj = st.vec[4][t], lj = st.vec[5][t], uj = st.vec[6][t];
      if (CHECK) {
        const double xh = a.xhat_in[j];
        const double dx = xb - xh, dat = atyh - at, d0 = xh - x0;
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
      }
      const double xn = w_new * xb + (1.0 - w_new) * x0;
      const double atn = w_new * (2.0 * atyh - at) + (1.0 - w_new) * at0;
      const double xh2 = clampd(xn - tau * (cj - atn), lj, uj);
      a.xbar[j] = 2.0 * xh2 - xn;
      a.aty[j] = atn;
      if (STORE) a.xhat_out[j] = xh2;
    }
    if (CHECK) {
      const int q[6] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL};
      tile_reduce_store<6>(red, q, a.partials, sc->max_blocks, tile, s_red);
    }
    __syncthreads();
  }
}

}  // namespace mipcu
