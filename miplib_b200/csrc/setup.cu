// setup.cu - everything that happens once per model: upload, CSC (= CSR of K') construction,
// Ruiz + Pock-Chambolle equilibration of both copies in place, step-size estimate.
// Restated on the CPU by oracle/pdhg_ref.py: ruiz_pc_scale / spectral_norm / prepare.
#include <cub/cub.cuh>

#include <chrono>
#include <cmath>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "internal.h"
#include "kernels.cuh"

namespace mipcu {

namespace {

inline int grid_for(int64_t n, int per_block = kThreads) {
  return (int)((n + per_block - 1) / per_block);
}

// Device allocations of this file.  A context allocates stream-ordered from its own memory
// pool (created with an unlimited release threshold, api.cu): cudaMalloc / cudaFree synchronise the
// device and map / unmap memory on every call, which made RE-loading a config-3 model cost 0.7 s;
// from the pool the same buffers are reused.  The few vectors a row-sharded context exports over
// CUDA IPC (pool memory cannot be exported) come from cudaMalloc and are remembered as such.
thread_local bool g_alloc_exported = false;

template <class T>
T* dalloc(int64_t count) {
  T* p = nullptr;
  const size_t bytes = sizeof(T) * (size_t)(count > 0 ? count : 1);
  mipcu_ctx* ctx = const_cast<mipcu_ctx*>(g_alloc_ctx);
  if (ctx && ctx->pool_alloc && !g_alloc_exported) {
    MIPCU_CK(cudaMallocFromPoolAsync(&p, bytes, (cudaMemPool_t)ctx->pool, ctx->stream));
  } else {
    MIPCU_CK(cudaMalloc(&p, bytes));
    if (ctx && ctx->pool_alloc) ctx->raw_allocs.push_back(p);
  }
  return p;
}

template <class T>
void dfree(T*& p) {
  if (p) {
    mipcu_ctx* ctx = const_cast<mipcu_ctx*>(g_alloc_ctx);
    bool raw = !(ctx && ctx->pool_alloc);
    if (!raw) {
      auto it = std::find(ctx->raw_allocs.begin(), ctx->raw_allocs.end(), (void*)p);
      if (it != ctx->raw_allocs.end()) { raw = true; ctx->raw_allocs.erase(it); }
    }
    if (raw) cudaFree(p);
    else cudaFreeAsync(p, ctx->stream);
  }
  p = nullptr;
}

__global__ void k_narrow_ptr(const long long* __restrict__ in, int* __restrict__ out, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (int)in[i];
}

// |v| >= INF_THRESHOLD (the MIPCU_INF sentinel, SCIP's 1e20, lp_solve's 1e30) -> IEEE infinity
__global__ void k_sanitize_inf(double* v, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    double x = v[i];
    if (x >= INF_THRESHOLD) v[i] = INFINITY;
    else if (x <= -INF_THRESHOLD) v[i] = -INFINITY;
  }
}

__global__ void k_expand_rows(const int* __restrict__ ptr, int* __restrict__ row_of, int nrows) {
  // one warp per row (rows are short; long rows just loop)
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp < nrows)
    for (int k = ptr[warp] + lane; k < ptr[warp + 1]; k += 32) row_of[k] = warp;
}

__global__ void k_iota(int* v, int64_t count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) v[i] = (int)i;
}

__global__ void k_count(const int* __restrict__ keys, int* __restrict__ counts, int64_t count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) atomicAdd(counts + keys[i], 1);
}

__global__ void k_permute(const int* __restrict__ perm, const int* __restrict__ row_of,
                          const double* __restrict__ val, int* __restrict__ out_idx,
                          double* __restrict__ out_val, int64_t count) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    int p = perm[i];
    out_idx[i] = row_of[p];
    out_val[i] = val[p];
  }
}

__global__ void k_check_sorted(const int* __restrict__ ptr, const int* __restrict__ idx, int nrows,
                               int ncols, int* bad) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp < nrows) {
    int b = ptr[warp], e = ptr[warp + 1];
    if (e < b) { *bad = 1; return; }
    for (int k = b + lane; k < e; k += 32) {
      int c = idx[k];
      if (c < 0 || c >= ncols) *bad = 2;
      if (k > b && idx[k - 1] >= c) *bad = 3;
    }
  }
}

// per-row max |a| (MODE 0) or sum |a| (MODE 1), one sub-warp of 8 lanes per row
template <int MODE>
__global__ void k_row_norm(const int* __restrict__ ptr, const double* __restrict__ val,
                           double* __restrict__ out, int nrows) {
  constexpr int TPR = 8;
  int row = (blockIdx.x * blockDim.x + threadIdx.x) / TPR, lane = threadIdx.x % TPR;
  double a = 0.0;
  if (row < nrows)
    for (int k = ptr[row] + lane; k < ptr[row + 1]; k += TPR) {
      double v = fabs(val[k]);
      a = MODE == 0 ? fmax(a, v) : a + v;
    }
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) {
    double b = __shfl_down_sync(0xffffffffu, a, o, TPR);
    a = MODE == 0 ? fmax(a, b) : a + b;
  }
  if (row < nrows && lane == 0) out[row] = a;
}

// norm -> factor 1/sqrt(norm) (1 for empty rows); d *= factor
__global__ void k_factor(double* __restrict__ norm_io, double* __restrict__ d, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    double nv = norm_io[i];
    double f = nv > 0.0 ? 1.0 / sqrt(nv) : 1.0;
    norm_io[i] = f;
    d[i] *= f;
  }
}

// val <- (val * rfac[row_of_entry]) * cfac[col_of_entry], the same expression on both copies.
// OUTER_IS_ROW: this copy's outer index is the matrix row (CSR of K); else it is the column.
template <bool OUTER_IS_ROW>
__global__ void k_scale_matrix(const int* __restrict__ ptr, const int* __restrict__ idx,
                               double* __restrict__ val, const double* __restrict__ rfac,
                               const double* __restrict__ cfac, int nouter) {
  constexpr int TPR = 8;
  int o = (blockIdx.x * blockDim.x + threadIdx.x) / TPR, lane = threadIdx.x % TPR;
  if (o < nouter) {
    const double fo = OUTER_IS_ROW ? rfac[o] : cfac[o];
    for (int k = ptr[o] + lane; k < ptr[o + 1]; k += TPR) {
      const double fi = OUTER_IS_ROW ? cfac[idx[k]] : rfac[idx[k]];
      val[k] = OUTER_IS_ROW ? (val[k] * fo) * fi : (val[k] * fi) * fo;
    }
  }
}

__global__ void k_scatter_bounds(const long long* __restrict__ cols, const double* __restrict__ lb,
                                 const double* __restrict__ ub, double* __restrict__ dlb,
                                 double* __restrict__ dub, const int* __restrict__ col_inv, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    double l = lb[i], u = ub[i];
    if (l <= -INF_THRESHOLD) l = -INFINITY; else if (l >= INF_THRESHOLD) l = INFINITY;
    if (u >= INF_THRESHOLD) u = INFINITY; else if (u <= -INF_THRESHOLD) u = -INFINITY;
    const long long j = col_inv ? col_inv[cols[i]] : cols[i];
    dlb[j] = l;
    dub[j] = u;
  }
}

__global__ void k_fill(double* v, double a, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) v[i] = a;
}

__global__ void k_mul(const double* __restrict__ a, const double* __restrict__ b,
                      double* __restrict__ out, int count, double s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = s * a[i] * b[i];
}

__global__ void k_div(const double* __restrict__ a, const double* __restrict__ b,
                      double* __restrict__ out, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = a[i] / b[i];
}

__global__ void k_div_inplace(double* v, const double* __restrict__ b, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) v[i] /= b[i];
}

__global__ void k_scale_inplace(double* v, double s, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) v[i] *= s;
}

__global__ void k_hash_vector(double* v, int count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    unsigned long long h = ((unsigned long long)i * 2654435761ull) & 0xFFFFFFFFull;
    v[i] = (double)h / 4294967296.0 + 0.5;
  }
}

// MODE 0: sum v^2 ; MODE 1: sum max(|lo| finite, |hi| finite)^2 with b = second array
template <int MODE>
__global__ void k_sumsq_partial(const double* __restrict__ a, const double* __restrict__ b,
                                double* __restrict__ partials, int64_t count) {
  __shared__ double s[kThreads / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v;
    if (MODE == 0) v = a[i];
    else {
      double lo = a[i], hi = b[i];
      v = fmax(isfinite(lo) ? fabs(lo) : 0.0, isfinite(hi) ? fabs(hi) : 0.0);
    }
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += s[w];
    partials[blockIdx.x] = t;
  }
}

template <int MODE>
double sumsq(mipcu_ctx* ctx, const double* a, const double* b, int64_t count) {
  if (count <= 0) return 0.0;
  int blocks = (int)std::min<int64_t>(grid_for(count), 1024);
  k_sumsq_partial<MODE><<<blocks, kThreads, 0, ctx->stream>>>(a, b, ctx->partials, count);
  ctx->launches++;
  std::vector<double> h(blocks);
  MIPCU_CK(cudaMemcpyAsync(h.data(), ctx->partials, sizeof(double) * blocks, cudaMemcpyDeviceToHost,
                           ctx->stream));
  MIPCU_CK(cudaStreamSynchronize(ctx->stream));
  double t = 0.0;
  for (double v : h) t += v;
  return t;
}

template <int TPR>
void launch_spmv(mipcu_ctx* ctx, const Csr& A, const double* v, double* out) {
  if (A.nrows == 0) return;
  const int g = grid_for(A.nrows, kRowsPerCta);
  if (TPR == 0)
    k_spmv<TPR><<<g, kThreads, 0, ctx->stream>>>(A.sptr, A.sidx, A.sval, A.perm, v, out, A.nrows);
  else
    k_spmv<TPR><<<g, kThreads, 0, ctx->stream>>>(A.ptr, A.idx, A.val, nullptr, v, out, A.nrows);
  ctx->launches++;
}

// ---- sliced ELL (SELL-32-256) ----------------------------------------------------------------
// One CTA per block of 256 rows: stable sort of the rows by decreasing length (rank by counting,
// 256 x 256 comparisons in shared memory), perm[sorted position] = CTA-local row, and the padded
// size of each of the block's 8 slices (32 x the length of the slice's first = longest row).
__global__ void __launch_bounds__(kThreads)
k_sell_plan(const int* __restrict__ ptr, int nrows, unsigned char* __restrict__ perm,
            long long* __restrict__ slice_size) {
  __shared__ int s_len[kRowsPerCta];
  __shared__ int s_sorted[kRowsPerCta];
  const int t = threadIdx.x;
  const int r = blockIdx.x * kRowsPerCta + t;
  const int len = r < nrows ? ptr[r + 1] - ptr[r] : -1;    // rows past the end sort last
  s_len[t] = len;
  __syncthreads();
  int rank = 0;
  for (int o = 0; o < kRowsPerCta; ++o) {
    const int lo = s_len[o];
    rank += (lo > len) || (lo == len && o < t);
  }
  perm[blockIdx.x * kRowsPerCta + rank] = (unsigned char)t;
  s_sorted[rank] = max(len, 0);
  __syncthreads();
  if (t < kRowsPerCta / 32) slice_size[blockIdx.x * (kRowsPerCta / 32) + t] = 32ll * s_sorted[32 * t];
}

__global__ void __launch_bounds__(kThreads)
k_sell_fill(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
            int nrows, const unsigned char* __restrict__ perm, const int* __restrict__ sptr,
            int* __restrict__ sidx, double* __restrict__ sval) {
  const int t = threadIdx.x, lane = t & 31;
  const int r = blockIdx.x * kRowsPerCta + perm[blockIdx.x * kRowsPerCta + t];
  const int slice = blockIdx.x * (kRowsPerCta / 32) + (t >> 5);
  const int beg = sptr[slice], width = (sptr[slice + 1] - beg) / 32;
  const int k0 = r < nrows ? ptr[r] : 0;
  const int len = r < nrows ? ptr[r + 1] - k0 : 0;
  for (int j = 0; j < width; ++j) {
    sidx[beg + 32 * j + lane] = j < len ? idx[k0 + j] : -1;
    sval[beg + 32 * j + lane] = j < len ? val[k0 + j] : 0.0;
  }
}

__global__ void k_narrow_ll(const long long* in, int* out, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = (int)in[i];
}

void free_sell(Csr& A) {
  dfree(A.sptr); dfree(A.sidx); dfree(A.sval);
  dfree(A.perm);
  A.sell_entries = 0;
  A.sell = false;
}

// Builds the SELL copy of the (already scaled) matrix.  Eligible: short rows on average (one lane
// per row must still fill the machine) and a padded size within 35 % of the nonzero count.
void build_sell(mipcu_ctx* ctx, Csr& A) {
  free_sell(A);
  if (A.nrows == 0 || A.nnz == 0) return;
  if ((double)A.nnz / (double)A.nrows > kSellMaxAvgRow) return;
  cudaStream_t st = ctx->stream;
  const int nblk = grid_for(A.nrows, kRowsPerCta), nslices = nblk * (kRowsPerCta / 32);
  A.perm = dalloc<unsigned char>((int64_t)nblk * kRowsPerCta);
  long long* sizes = dalloc<long long>(nslices + 1);
  long long* offs = dalloc<long long>(nslices + 1);
  MIPCU_CK(cudaMemsetAsync(sizes, 0, sizeof(long long) * (nslices + 1), st));
  k_sell_plan<<<nblk, kThreads, 0, st>>>(A.ptr, A.nrows, A.perm, sizes);
  ctx->launches++;
  size_t tmp_bytes = 0;
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sizes, offs, nslices + 1, st));
  char* tmp = dalloc<char>((int64_t)tmp_bytes + 16);
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, sizes, offs, nslices + 1, st));
  long long total = 0;
  MIPCU_CK(cudaMemcpyAsync(&total, offs + nslices, sizeof(long long), cudaMemcpyDeviceToHost, st));
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(tmp);
  const bool ok = total < (1ll << 31) - 1024 && (double)total <= 1.35 * (double)A.nnz + 8192.0;
  if (ok) {
    A.sptr = dalloc<int>(nslices + 1);
    A.sidx = dalloc<int>(total + 4);
    A.sval = dalloc<double>(total + 4);
    k_narrow_ll<<<grid_for(nslices + 1), kThreads, 0, st>>>(offs, A.sptr, nslices + 1);
    k_sell_fill<<<nblk, kThreads, 0, st>>>(A.ptr, A.idx, A.val, A.nrows, A.perm, A.sptr, A.sidx, A.sval);
    ctx->launches += 2;
    MIPCU_CK(cudaStreamSynchronize(st));
    A.sell_entries = total;
    A.sell = true;
  } else {
    dfree(A.perm);
  }
  dfree(sizes); dfree(offs);
}

// CSR of the transpose on the device: stable radix sort of (col, position) pairs, so the entries
// of every column come out in increasing row order (deterministic summation order).
void build_transpose(mipcu_ctx* ctx, const Csr& A, Csr& T) {
  cudaStream_t st = ctx->stream;
  T.nrows = A.ncols;
  T.ncols = A.nrows;
  T.nnz = A.nnz;
  T.ptr = dalloc<int>(T.nrows + 1);
  T.idx = dalloc<int>(T.nnz + 4);
  T.val = dalloc<double>(T.nnz + 4);
  MIPCU_CK(cudaMemsetAsync(T.ptr, 0, sizeof(int) * (T.nrows + 1), st));
  MIPCU_CK(cudaMemsetAsync(T.idx + T.nnz, 0, sizeof(int) * 4, st));
  MIPCU_CK(cudaMemsetAsync(T.val + T.nnz, 0, sizeof(double) * 4, st));
  if (A.nnz == 0) return;
  int* row_of = dalloc<int>(A.nnz);
  int* pos = dalloc<int>(A.nnz);
  int* keys_out = dalloc<int>(A.nnz);
  int* perm = dalloc<int>(A.nnz);
  int* counts = dalloc<int>(T.nrows + 1);
  k_expand_rows<<<grid_for((int64_t)A.nrows * 32), kThreads, 0, st>>>(A.ptr, row_of, A.nrows);
  k_iota<<<grid_for(A.nnz), kThreads, 0, st>>>(pos, A.nnz);
  MIPCU_CK(cudaMemsetAsync(counts, 0, sizeof(int) * (T.nrows + 1), st));
  k_count<<<grid_for(A.nnz), kThreads, 0, st>>>(A.idx, counts, A.nnz);
  ctx->launches += 3;
  int end_bit = 1;
  while ((1ll << end_bit) < (long long)A.ncols && end_bit < 31) ++end_bit;
  size_t tmp_bytes = 0, tmp2 = 0;
  MIPCU_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, A.idx, keys_out, pos, perm,
                                           (int)A.nnz, 0, end_bit, st));
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, counts, T.ptr, T.nrows + 1, st));
  char* tmp = dalloc<char>((int64_t)std::max(tmp_bytes, tmp2) + 16);
  MIPCU_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, A.idx, keys_out, pos, perm, (int)A.nnz,
                                           0, end_bit, st));
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp2, counts, T.ptr, T.nrows + 1, st));
  k_permute<<<grid_for(A.nnz), kThreads, 0, st>>>(perm, row_of, A.val, T.idx, T.val, A.nnz);
  ctx->launches++;
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(tmp);
  dfree(row_of); dfree(pos); dfree(keys_out); dfree(perm); dfree(counts);
}

// entries a 64-row tile of A occupies in the TMA kernels' shared-memory stage (its nonzeros plus the
// slack of aligning the slice to 16 bytes), maximised over the tiles
__global__ void k_tile_max(const int* __restrict__ ptr, int nrows, int tile_rows, int* out) {
  const int tile = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = tile * tile_rows;
  if (r0 >= nrows) return;
  const int r1 = min(r0 + tile_rows, nrows);
  const int kb = ptr[r0] & ~3;
  atomicMax(out, ((ptr[r1] - kb) + 3) & ~3);
}

// ---- rows sorted by decreasing length ----------------------------------------------------------
__global__ void k_len_keys(const int* __restrict__ ptr, unsigned* __restrict__ keys, int* __restrict__ ids, int nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) {
    keys[i] = 0x7fffffffu - (unsigned)(ptr[i + 1] - ptr[i]);   // ascending key = descending length
    ids[i] = i;
  }
}

// (length, first index): rows of equal length ordered by their first entry, so that the 32 rows of a slice
// start in neighbouring lines of the gathered vector (one coalesced request instead of 32 for that step)
__global__ void k_len_first_keys(const int* __restrict__ ptr, const int* __restrict__ idx,
                                 unsigned long long* __restrict__ keys, int* __restrict__ ids, int nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) {
    const int beg = ptr[i], len = ptr[i + 1] - beg;
    keys[i] = ((unsigned long long)(0x7fffffffu - (unsigned)len) << 32) | (unsigned)(len ? idx[beg] : 0);
    ids[i] = i;
  }
}

// out[i] = in[perm[i]] (composition of two row permutations)
__global__ void k_gather_ints(const int* __restrict__ in, const int* __restrict__ perm, int* __restrict__ out, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = in[perm[i]];
}

__global__ void k_invert_perm(const int* __restrict__ perm, int* __restrict__ inv, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) inv[perm[i]] = i;
}

__global__ void k_perm_len(const int* __restrict__ ptr, const int* __restrict__ perm, int* __restrict__ len, int nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) len[i] = ptr[perm[i] + 1] - ptr[perm[i]];
  if (i == nrows) len[i] = 0;
}

// new row i = old row perm[i]; 8 lanes per row
__global__ void k_permute_rows(const int* __restrict__ optr, const int* __restrict__ oidx, const double* __restrict__ oval,
                               const int* __restrict__ perm, const int* __restrict__ nptr, int* __restrict__ nidx,
                               double* __restrict__ nval, int nrows) {
  constexpr int TPR = 8;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) / TPR, lane = threadIdx.x % TPR;
  if (i >= nrows) return;
  const int src = optr[perm[i]], dst = nptr[i], len = nptr[i + 1] - dst;
  for (int k = lane; k < len; k += TPR) {
    nidx[dst + k] = oidx[src + k];
    nval[dst + k] = oval[src + k];
  }
}

__global__ void k_gather_rows(const double* __restrict__ in, const int* __restrict__ perm, double* __restrict__ out, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = in[perm[i]];
}

int tile_capacity(mipcu_ctx* ctx, const Csr& A, int tile_rows = 64) {
  if (A.nrows == 0) return 0;
  int* d = dalloc<int>(1);
  MIPCU_CK(cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
  k_tile_max<<<grid_for(grid_for(A.nrows, tile_rows)), kThreads, 0, ctx->stream>>>(A.ptr, A.nrows, tile_rows, d);
  ctx->launches++;
  int h = 0;
  MIPCU_CK(cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MIPCU_CK(cudaStreamSynchronize(ctx->stream));
  dfree(d);
  return std::max(h, 4);
}

void free_csr(Csr& A) {
  free_sell(A);
  dfree(A.ptr); dfree(A.idx); dfree(A.val);
  A = Csr();
}

// ---- column block of the block-owner exchange ---------------------------------------------------
// ptr_all / idx_all / val_all: every rank's local transpose (CSR of K_g', n rows, LOCAL row ids),
// all-gathered.  Column j of the whole matrix is the concatenation over g of row j of those, with
// row ids shifted by the rank's row offset: rows come out increasing (deterministic sums).
__global__ void k_colblk_count(const int* __restrict__ ptr_all, int W, int n1, int j0, int nloc,
                               int* __restrict__ cnt) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nloc) return;
  int c = 0;
  for (int g = 0; g < W; ++g) {
    const int* P = ptr_all + (size_t)g * n1;
    c += P[j0 + j + 1] - P[j0 + j];
  }
  cnt[j] = c;
}

__global__ void k_colblk_fill(const int* __restrict__ ptr_all, const int* __restrict__ idx_all,
                              const double* __restrict__ val_all, size_t stride, int W, int n1, int j0,
                              int nloc, const int* __restrict__ row_off, const int* __restrict__ bptr,
                              int* __restrict__ bidx, double* __restrict__ bval) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nloc) return;
  int pos = bptr[j];
  for (int g = 0; g < W; ++g) {
    const int* P = ptr_all + (size_t)g * n1;
    for (int k = P[j0 + j]; k < P[j0 + j + 1]; ++k) {
      bidx[pos] = idx_all[(size_t)g * stride + k] + row_off[g];
      bval[pos] = val_all[(size_t)g * stride + k];
      ++pos;
    }
  }
}

__global__ void k_place(const double* __restrict__ src, double* __restrict__ dst, int offset, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[offset + i] = src[i];
}

// Stable sort of the rows of A by decreasing length (ties keep the caller's order; with `by_first`,
// ties are ordered by the row's first index), in place; returns the permutation (new row -> old row),
// null for models with fewer than two rows.  With rows of equal length next to each other the 32-row
// slices of the SELL copy carry no padding (config 3: 6.4 % of the row-side stream) and the rows of a
// warp finish together.
int* sort_rows_by_length(mipcu_ctx* ctx, Csr& A, bool by_first = false) {
  const int m = A.nrows;
  if (m < 2 || A.nnz == 0) return nullptr;
  cudaStream_t st = ctx->stream;
  int* ids = dalloc<int>(m);
  int* perm = dalloc<int>(m);
  int* nlen = dalloc<int>(m + 1);
  int* nptr = dalloc<int>(m + 1);
  size_t tmp_bytes = 0, tmp2 = 0;
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, nlen, nptr, m + 1, st));
  char* tmp = nullptr;
  if (by_first) {
    unsigned long long* keys = dalloc<unsigned long long>(m);
    unsigned long long* keys_out = dalloc<unsigned long long>(m);
    k_len_first_keys<<<grid_for(m), kThreads, 0, st>>>(A.ptr, A.idx, keys, ids, m);
    MIPCU_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, ids, perm, m, 0, 64, st));
    tmp = dalloc<char>((int64_t)std::max(tmp_bytes, tmp2) + 16);
    MIPCU_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, ids, perm, m, 0, 64, st));
    MIPCU_CK(cudaStreamSynchronize(st));
    dfree(keys); dfree(keys_out);
  } else {
    unsigned* keys = dalloc<unsigned>(m);
    unsigned* keys_out = dalloc<unsigned>(m);
    k_len_keys<<<grid_for(m), kThreads, 0, st>>>(A.ptr, keys, ids, m);
    MIPCU_CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys_out, ids, perm, m, 0, 32, st));
    tmp = dalloc<char>((int64_t)std::max(tmp_bytes, tmp2) + 16);
    MIPCU_CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, ids, perm, m, 0, 32, st));
    MIPCU_CK(cudaStreamSynchronize(st));
    dfree(keys); dfree(keys_out);
  }
  k_perm_len<<<grid_for(m + 1), kThreads, 0, st>>>(A.ptr, perm, nlen, m);
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp2, nlen, nptr, m + 1, st));
  int* nidx = dalloc<int>(A.nnz + 4);
  double* nval = dalloc<double>(A.nnz + 4);
  MIPCU_CK(cudaMemsetAsync(nidx + A.nnz, 0, sizeof(int) * 4, st));
  MIPCU_CK(cudaMemsetAsync(nval + A.nnz, 0, sizeof(double) * 4, st));
  k_permute_rows<<<grid_for((int64_t)m * 8), kThreads, 0, st>>>(A.ptr, A.idx, A.val, perm, nptr, nidx, nval, m);
  ctx->launches += 3;
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(A.ptr); dfree(A.idx); dfree(A.val);
  A.ptr = nptr; A.idx = nidx; A.val = nval;
  dfree(ids); dfree(nlen); dfree(tmp);
  return perm;
}

// Columns binned at load (MIPCU_PARAM_ROW_SORT = 2, single-GPU contexts).  T = K' by rows (one row per
// column of the model).  Its rows are sorted by (decreasing length, first row index) - columns of a
// 32-column slice then start in neighbouring lines of yhat - and A is rebuilt from the permuted T, so the
// column indices inside a row of A are the NEW ones, increasing.  Leaves ctx->col_perm / col_inv.
// Measured on config 3 (profiles/r02_locality.md): requests per gathered double 0.999 -> 0.935 on the
// column step, 0.978 on the row step after the second row sort below; column kernel -5 %, row kernel -4 %.
void bin_columns(mipcu_ctx* ctx, Csr& A, Csr& T) {
  const int n = T.nrows, m = A.nrows;
  // perm (new -> previous numbering) applied on top of *total (previous -> caller's); takes ownership of perm
  auto compose = [&](int*& total, int* perm, int count) {
    if (!total) { total = perm; return; }
    int* composed = dalloc<int>(count);
    k_gather_ints<<<grid_for(count), kThreads, 0, ctx->stream>>>(total, perm, composed, count);
    ctx->launches++;
    MIPCU_CK(cudaStreamSynchronize(ctx->stream));
    dfree(total);
    dfree(perm);
    total = composed;
  };
  // MIPCU_PARAM_ROW_SORT: 2 = columns once; 3 = columns, then rows; 4 = two (columns, rows) rounds.  Every
  // round re-sorts one side by (length, first index in the OTHER side's current numbering) and rebuilds the
  // other copy; requests per gathered double on config 3 (scripts/locality_lab.py), row / column side:
  // 0.990 / 0.939 (2), 0.978 / 0.935 (3), 0.971 / 0.929 (4), 0.970 / 0.927 (a third round: not worth its 6 ms).
  const int mode = (int)ctx->params[MIPCU_PARAM_ROW_SORT];
  const int rounds = mode >= 4 ? 2 : 1;
  for (int r = 0; r < rounds; ++r) {
    int* cperm = sort_rows_by_length(ctx, T, true);
    if (!cperm) break;
    compose(ctx->col_perm, cperm, n);
    const int tpr = A.tpr;
    free_csr(A);
    build_transpose(ctx, T, A);       // entries of a row come out sorted by the new column index
    A.tpr = tpr;
    if (mode < 3) break;
    // the rows, by (length, first NEW column): the first entries of a slice's rows then sit in neighbouring
    // lines of xbar as well.  The rows get new numbers, so K' is rebuilt from the re-sorted rows (same column
    // order), and the row permutation is the composition of the sorts.
    int* again = sort_rows_by_length(ctx, A, true);
    if (!again) break;
    compose(ctx->row_perm, again, m);
    const int ttpr = T.tpr;
    free_csr(T);
    build_transpose(ctx, A, T);
    T.tpr = ttpr;
  }
  if (ctx->col_perm) {
    ctx->col_inv = dalloc<int>(n);
    k_invert_perm<<<grid_for(n), kThreads, 0, ctx->stream>>>(ctx->col_perm, ctx->col_inv, n);
    ctx->launches++;
  }
}

}  // namespace

const std::vector<int>& host_row_perm(mipcu_ctx* ctx) {
  if (ctx->row_perm && ctx->h_row_perm.empty() && ctx->m) {
    ctx->h_row_perm.resize((size_t)ctx->m);
    MIPCU_CK(cudaMemcpy(ctx->h_row_perm.data(), ctx->row_perm, sizeof(int) * (size_t)ctx->m, cudaMemcpyDeviceToHost));
  }
  return ctx->h_row_perm;
}

void rows_to_caller_order(mipcu_ctx* ctx, double* v) {
  const std::vector<int>& p = host_row_perm(ctx);
  if (p.empty()) return;
  std::vector<double> t(v, v + p.size());
  for (size_t i = 0; i < p.size(); ++i) v[p[i]] = t[i];
}

static const std::vector<int>& host_col_perm(mipcu_ctx* ctx) {
  if (ctx->col_perm && ctx->h_col_perm.empty() && ctx->n) {
    ctx->h_col_perm.resize((size_t)ctx->n);
    MIPCU_CK(cudaMemcpy(ctx->h_col_perm.data(), ctx->col_perm, sizeof(int) * (size_t)ctx->n, cudaMemcpyDeviceToHost));
  }
  return ctx->h_col_perm;
}

void cols_to_caller_order(mipcu_ctx* ctx, double* v) {
  const std::vector<int>& p = host_col_perm(ctx);
  if (p.empty()) return;
  std::vector<double> t(v, v + p.size());
  for (size_t j = 0; j < p.size(); ++j) v[p[j]] = t[j];
}

void cols_to_internal_order(mipcu_ctx* ctx, double* v) {
  const std::vector<int>& p = host_col_perm(ctx);
  if (p.empty()) return;
  std::vector<double> t(v, v + p.size());
  for (size_t j = 0; j < p.size(); ++j) v[j] = t[p[j]];
}

void rows_to_internal_order(mipcu_ctx* ctx, double* v) {
  const std::vector<int>& p = host_row_perm(ctx);
  if (p.empty()) return;
  std::vector<double> t(v, v + p.size());
  for (size_t i = 0; i < p.size(); ++i) v[i] = t[p[i]];
}

// Collective.  Rank g already holds its row block and that block's transpose; the block-owner
// exchange also needs columns [j0, j1) of the WHOLE matrix.  Every rank all-gathers the local
// transposes (setup only: the matrix crosses NVLink once) and cuts its column range out of them.
void build_col_block(mipcu_ctx* ctx) {
  free_csr(ctx->colblk);
  if (ctx->world <= 1) return;
  cudaStream_t st = ctx->stream;
  const int W = ctx->world, n = ctx->n, n1 = n + 1;
  const int j0 = (int)(((int64_t)n * ctx->rank) / W), j1 = (int)(((int64_t)n * (ctx->rank + 1)) / W);
  const int nloc = j1 - j0;
  const Csr& T = ctx->cols;
  // sizes of everybody's block
  long long* d_meta = dalloc<long long>(W + 1);
  const long long mine = T.nnz;
  MIPCU_CK(cudaMemcpyAsync(d_meta + W, &mine, sizeof(long long), cudaMemcpyHostToDevice, st));
  dist_allgather_bytes(ctx, d_meta + W, d_meta, sizeof(long long));
  std::vector<long long> nnz_of(W);
  MIPCU_CK(cudaMemcpyAsync(nnz_of.data(), d_meta, sizeof(long long) * W, cudaMemcpyDeviceToHost, st));
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(d_meta);
  size_t stride = 4;
  for (long long v : nnz_of) stride = std::max(stride, (size_t)((v + 3) & ~3ll));
  int* send_idx = dalloc<int>(stride);
  double* send_val = dalloc<double>(stride);
  MIPCU_CK(cudaMemsetAsync(send_idx, 0, sizeof(int) * stride, st));
  MIPCU_CK(cudaMemsetAsync(send_val, 0, sizeof(double) * stride, st));
  if (T.nnz) {
    MIPCU_CK(cudaMemcpyAsync(send_idx, T.idx, sizeof(int) * T.nnz, cudaMemcpyDeviceToDevice, st));
    MIPCU_CK(cudaMemcpyAsync(send_val, T.val, sizeof(double) * T.nnz, cudaMemcpyDeviceToDevice, st));
  }
  int* ptr_all = dalloc<int>((int64_t)W * n1);
  int* idx_all = dalloc<int>((int64_t)W * stride);
  double* val_all = dalloc<double>((int64_t)W * stride);
  dist_allgather_bytes(ctx, T.ptr, ptr_all, sizeof(int) * (size_t)n1);
  dist_allgather_bytes(ctx, send_idx, idx_all, sizeof(int) * stride);
  dist_allgather_bytes(ctx, send_val, val_all, sizeof(double) * stride);
  int* d_off = dalloc<int>(W + 1);
  MIPCU_CK(cudaMemcpyAsync(d_off, ctx->row_off, sizeof(int) * (W + 1), cudaMemcpyHostToDevice, st));

  Csr& B = ctx->colblk;
  B.nrows = nloc;
  B.ncols = ctx->m_glob;
  B.ptr = dalloc<int>(nloc + 1);
  int* cnt = dalloc<int>(nloc + 1);
  MIPCU_CK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nloc + 1), st));
  if (nloc) {
    k_colblk_count<<<grid_for(nloc), kThreads, 0, st>>>(ptr_all, W, n1, j0, nloc, cnt);
    ctx->launches++;
  }
  size_t tmp_bytes = 0;
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, B.ptr, nloc + 1, st));
  char* tmp = dalloc<char>((int64_t)tmp_bytes + 16);
  MIPCU_CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, B.ptr, nloc + 1, st));
  int total = 0;
  MIPCU_CK(cudaMemcpyAsync(&total, B.ptr + nloc, sizeof(int), cudaMemcpyDeviceToHost, st));
  MIPCU_CK(cudaStreamSynchronize(st));
  B.nnz = total;
  B.idx = dalloc<int>((int64_t)total + 4);
  B.val = dalloc<double>((int64_t)total + 4);
  MIPCU_CK(cudaMemsetAsync(B.idx + total, 0, sizeof(int) * 4, st));
  MIPCU_CK(cudaMemsetAsync(B.val + total, 0, sizeof(double) * 4, st));
  if (nloc) {
    k_colblk_fill<<<grid_for(nloc), kThreads, 0, st>>>(ptr_all, idx_all, val_all, stride, W, n1, j0, nloc,
                                                      d_off, B.ptr, B.idx, B.val);
    ctx->launches++;
  }
  B.tpr = choose_tpr(nloc ? (double)total / (double)nloc : 0.0, (int)ctx->params[MIPCU_PARAM_THREADS_PER_ROW]);
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(tmp); dfree(cnt); dfree(d_off); dfree(send_idx); dfree(send_val);
  dfree(ptr_all); dfree(idx_all); dfree(val_all);
}

int choose_tpr(double avg, int override_tpr) {
  if (override_tpr == 2 || override_tpr == 4 || override_tpr == 8 || override_tpr == 16 ||
      override_tpr == 32)
    return override_tpr;
  if (avg <= 3.0) return 2;
  if (avg <= 6.0) return 4;
  if (avg <= 48.0) return 8;     // measured on config 3 rows (avg 32): 8 lanes 175 us, 16 lanes 185 us
  if (avg <= 96.0) return 16;
  return 32;
}

// Which kernel family a matrix runs through: the SELL copy when it exists and nothing asks for a
// CSR variant (an explicit lanes-per-row override, or one of the opt-in TMA / CSR-stream kernels).
int kernel_format(const mipcu_ctx* ctx, const Csr& A) {
  const int tma = (int)ctx->params[MIPCU_PARAM_USE_TMA];
  const int lanes = (int)ctx->params[MIPCU_PARAM_THREADS_PER_ROW];
  if (A.sell && lanes == 0 && tma != 1 && tma != 2) return 0;
  return A.tpr;
}

void spmv_plain(mipcu_ctx* ctx, const Csr& A, const double* v, double* out) {
  switch (kernel_format(ctx, A)) {
    case 0: launch_spmv<0>(ctx, A, v, out); break;
    case 2: launch_spmv<2>(ctx, A, v, out); break;
    case 4: launch_spmv<4>(ctx, A, v, out); break;
    case 8: launch_spmv<8>(ctx, A, v, out); break;
    case 16: launch_spmv<16>(ctx, A, v, out); break;
    default: launch_spmv<32>(ctx, A, v, out); break;
  }
}

double device_norm2(mipcu_ctx* ctx, const double* v, int64_t len) {
  return std::sqrt(sumsq<0>(ctx, v, nullptr, len));
}

void free_model(mipcu_ctx* ctx) {
  free_csr(ctx->rows);
  free_csr(ctx->cols);
  free_csr(ctx->colblk);
  // keep_exchange (re-load of the same shape on a row-sharded context): the vectors the peers have
  // mapped - and the multicast window xbar / yhat_full may live in - stay where they are
  const bool keep = ctx->keep_exchange;
  if (!keep) dfree(ctx->yhat_full);
  dfree(ctx->row_perm);
  ctx->h_row_perm.clear();
  dfree(ctx->col_perm);
  dfree(ctx->col_inv);
  ctx->h_col_perm.clear();
  double** vecs[] = {&ctx->c, &ctx->lb, &ctx->ub, &ctx->lo, &ctx->hi, &ctx->cs, &ctx->ls, &ctx->us,
                     &ctx->los, &ctx->his, &ctx->d1, &ctx->d2, &ctx->xbar, &ctx->x0, &ctx->aty,
                     &ctx->aty0, &ctx->xhat_first, &ctx->xhat_chk, &ctx->aty_cand, &ctx->y,
                     &ctx->y0, &ctx->yhat, &ctx->tmp_n, &ctx->tmp_m, &ctx->x_start, &ctx->y_start,
                     &ctx->x_sol, &ctx->y_sol, &ctx->partials};
  for (double** p : vecs) {
    if (keep && (p == &ctx->xbar || p == &ctx->xhat_chk || p == &ctx->tmp_n || p == &ctx->x_sol)) continue;
    dfree(*p);
  }
  if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
  ctx->graph_period = 0;
  ctx->loaded = ctx->prepared = ctx->has_solution = false;
  ctx->vectors_dirty = true;
}

void load_lp(mipcu_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
             const int32_t* col_idx, const double* val, const double* c, const double* lb,
             const double* ub, const double* row_lo, const double* row_hi, int maximize) {
  NvtxRange nvtx("mipcu:load_lp (H2D, transpose, row sort)");
  // MIPCU_TIMING=1: wall-clock laps of the load on stderr (each lap ends with a stream synchronisation)
  const bool timing = std::getenv("MIPCU_TIMING") != nullptr;
  const auto lap_t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    cudaStreamSynchronize(ctx->stream);
    std::fprintf(stderr, "[mipcu_load_lp] %s %.1f ms\n", what,
                 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - lap_t0).count());
  };
  MIPCU_REQUIRE(m >= 0 && n >= 0 && nnz >= 0, "mipcu_load_lp: negative dimension");
  MIPCU_REQUIRE(m < (1ll << 31) - 1024 && n < (1ll << 31) - 1024 && nnz < (1ll << 31) - 1024,
                "mipcu_load_lp: dimensions beyond int32 indexing are not supported");
  MIPCU_REQUIRE(row_ptr && (nnz == 0 || (col_idx && val)), "mipcu_load_lp: null matrix array");
  MIPCU_REQUIRE(n == 0 || (c && lb && ub), "mipcu_load_lp: null column array");
  MIPCU_REQUIRE(m == 0 || (row_lo && row_hi), "mipcu_load_lp: null row array");
  MIPCU_REQUIRE(row_ptr[0] == 0 && row_ptr[m] == nnz, "mipcu_load_lp: row_ptr[0] != 0 or row_ptr[m] != nnz");
  MIPCU_CK(cudaSetDevice(ctx->device));
  free_model(ctx);
  cudaStream_t st = ctx->stream;
  ctx->m = (int)m; ctx->n = (int)n; ctx->nnz = nnz; ctx->maximize = maximize != 0;

  Csr& A = ctx->rows;
  A.nrows = (int)m; A.ncols = (int)n; A.nnz = nnz;
  A.ptr = dalloc<int>(m + 1);
  A.idx = dalloc<int>(nnz + 4);
  A.val = dalloc<double>(nnz + 4);
  {
    long long* tmp64 = dalloc<long long>(m + 1);
    MIPCU_CK(cudaMemcpyAsync(tmp64, row_ptr, sizeof(int64_t) * (m + 1), cudaMemcpyHostToDevice, st));
    k_narrow_ptr<<<grid_for(m + 1), kThreads, 0, st>>>(tmp64, A.ptr, (int)(m + 1));
    ctx->launches++;
    MIPCU_CK(cudaStreamSynchronize(st));
    dfree(tmp64);
  }
  MIPCU_CK(cudaMemsetAsync(A.idx + nnz, 0, sizeof(int) * 4, st));
  MIPCU_CK(cudaMemsetAsync(A.val + nnz, 0, sizeof(double) * 4, st));
  if (nnz) {
    MIPCU_CK(cudaMemcpyAsync(A.idx, col_idx, sizeof(int) * nnz, cudaMemcpyHostToDevice, st));
    MIPCU_CK(cudaMemcpyAsync(A.val, val, sizeof(double) * nnz, cudaMemcpyHostToDevice, st));
  }
  // structural validation on the device (sorted, in range)
  {
    int* bad = dalloc<int>(1);
    MIPCU_CK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (m) {
      k_check_sorted<<<grid_for((int64_t)m * 32), kThreads, 0, st>>>(A.ptr, A.idx, (int)m, (int)n, bad);
      ctx->launches++;
    }
    int hbad = 0;
    MIPCU_CK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    MIPCU_CK(cudaStreamSynchronize(st));
    dfree(bad);
    if (hbad) {
      free_model(ctx);
      throw Error(hbad == 1 ? "mipcu_load_lp: row_ptr is not non-decreasing"
                  : hbad == 2 ? "mipcu_load_lp: column index out of range"
                              : "mipcu_load_lp: column indices must be strictly increasing inside a row");
    }
  }
  lap("matrix on the device, validated");
  // (rows of equal length by their first column from ROW_SORT = 3 on: on row-sharded contexts, which keep the
  // caller's column numbering, this is the part of the load-time ordering that is purely local to a rank)
  if (ctx->params[MIPCU_PARAM_ROW_SORT] != 0.0)
    ctx->row_perm = sort_rows_by_length(ctx, A, ctx->params[MIPCU_PARAM_ROW_SORT] >= 3.0);
  lap("rows sorted by length");
  build_transpose(ctx, A, ctx->cols);
  lap("transpose");
  // columns binned by their first row (row-sharded contexts keep the caller's column numbering: the ranks'
  // column slices are cut in it)
  if (ctx->params[MIPCU_PARAM_ROW_SORT] >= 2.0 && ctx->world <= 1 && n >= 2) bin_columns(ctx, A, ctx->cols);
  lap("columns binned, row copy rebuilt");
  int tpr_override = (int)ctx->params[MIPCU_PARAM_THREADS_PER_ROW];
  A.tpr = choose_tpr(m ? (double)nnz / (double)m : 0.0, tpr_override);
  ctx->cols.tpr = choose_tpr(n ? (double)nnz / (double)n : 0.0, tpr_override);

  auto upload = [&](double*& dst, const double* src, int64_t count, bool sanitize) {
    dst = dalloc<double>(count + 2);
    if (count) {
      MIPCU_CK(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyHostToDevice, st));
      if (sanitize) {
        k_sanitize_inf<<<grid_for(count), kThreads, 0, st>>>(dst, (int)count);
        ctx->launches++;
      }
    }
  };
  upload(ctx->c, c, n, false);
  upload(ctx->lb, lb, n, true);
  upload(ctx->ub, ub, n, true);
  upload(ctx->lo, row_lo, m, true);
  upload(ctx->hi, row_hi, m, true);
  if (ctx->row_perm) {            // row sides follow their rows
    for (double** side : {&ctx->lo, &ctx->hi}) {
      double* sorted = dalloc<double>(m + 2);
      k_gather_rows<<<grid_for(m), kThreads, 0, st>>>(*side, ctx->row_perm, sorted, (int)m);
      ctx->launches++;
      dfree(*side);
      *side = sorted;
    }
  }
  if (ctx->col_perm) {            // objective and box follow their columns
    for (double** vec : {&ctx->c, &ctx->lb, &ctx->ub}) {
      double* sorted = dalloc<double>(n + 2);
      k_gather_rows<<<grid_for(n), kThreads, 0, st>>>(*vec, ctx->col_perm, sorted, (int)n);
      ctx->launches++;
      dfree(*vec);
      *vec = sorted;
    }
  }

  double** nvecs[] = {&ctx->cs, &ctx->ls, &ctx->us, &ctx->d2, &ctx->xbar, &ctx->x0, &ctx->aty,
                      &ctx->aty0, &ctx->xhat_first, &ctx->xhat_chk, &ctx->aty_cand, &ctx->tmp_n,
                      &ctx->x_sol};
  for (double** p : nvecs) {
    // row-sharded: the exchange buffers are mapped by the peers (p2p.cu)
    const bool exported = ctx->world > 1 && (p == &ctx->xbar || p == &ctx->xhat_chk || p == &ctx->tmp_n || p == &ctx->x_sol);
    if (exported && ctx->keep_exchange) continue;     // same shape as the model before: still mapped
    g_alloc_exported = exported;
    *p = dalloc<double>(n + 2);   // +2: TMA reads an even number of doubles
    g_alloc_exported = false;
  }
  double** mvecs[] = {&ctx->los, &ctx->his, &ctx->d1, &ctx->y, &ctx->y0, &ctx->yhat, &ctx->tmp_m,
                      &ctx->y_sol};
  for (double** p : mvecs) *p = dalloc<double>(m + 2);
  const int m_glob_before = ctx->m_glob;
  ctx->m_glob = (int)m;
  for (int& v : ctx->row_off) v = 0;
  ctx->row_off[1] = (int)m;
  if (ctx->world > 1) {
    // every rank's row count -> row offsets of the blocks (collective, like the rest of a sharded load)
    const int W = ctx->world;
    long long* d_rows = dalloc<long long>(W + 1);
    const long long mine = m;
    MIPCU_CK(cudaMemcpyAsync(d_rows + W, &mine, sizeof(long long), cudaMemcpyHostToDevice, st));
    dist_allgather_bytes(ctx, d_rows + W, d_rows, sizeof(long long));
    std::vector<long long> rows_of(W);
    MIPCU_CK(cudaMemcpyAsync(rows_of.data(), d_rows, sizeof(long long) * W, cudaMemcpyDeviceToHost, st));
    MIPCU_CK(cudaStreamSynchronize(st));
    dfree(d_rows);
    long long acc = 0;
    for (int g = 0; g < W; ++g) { ctx->row_off[g] = (int)acc; acc += rows_of[g]; }
    MIPCU_REQUIRE(acc < (1ll << 31) - 1024, "mipcu_load_lp: total row count beyond int32 indexing");
    ctx->row_off[W] = (int)acc;
    MIPCU_REQUIRE(!ctx->keep_exchange || m_glob_before == (int)acc, "mipcu_load_lp: row blocks changed under kept mappings");
    ctx->m_glob = (int)acc;
    if (!ctx->keep_exchange) {
      g_alloc_exported = true;
      ctx->yhat_full = dalloc<double>(acc + 2);
      g_alloc_exported = false;
    }
    MIPCU_CK(cudaMemsetAsync(ctx->yhat_full, 0, sizeof(double) * (acc + 2), st));
  }
  ctx->max_blocks = std::max(1024, std::max(grid_for(m, 64), grid_for(n, 64)));   // 64 = TMA tile rows
  A.tile_cap = tile_capacity(ctx, A);
  ctx->cols.tile_cap = tile_capacity(ctx, ctx->cols);
  A.stream_cap = tile_capacity(ctx, A, kRowsPerCta);
  ctx->cols.stream_cap = tile_capacity(ctx, ctx->cols, kRowsPerCta);
  ctx->partials = dalloc<double>((int64_t)Q_COUNT * ctx->max_blocks);
  MIPCU_CK(cudaMemsetAsync(ctx->partials, 0, sizeof(double) * Q_COUNT * ctx->max_blocks, st));
  if (n) { k_fill<<<grid_for(n), kThreads, 0, st>>>(ctx->d2, 1.0, (int)n); ctx->launches++; }
  if (m) { k_fill<<<grid_for(m), kThreads, 0, st>>>(ctx->d1, 1.0, (int)m); ctx->launches++; }
  MIPCU_CK(cudaStreamSynchronize(st));
  lap("vectors uploaded and allocated");
  ctx->loaded = true;
  ctx->prepared = false;
  ctx->vectors_dirty = true;
}

// ---- MIPCU_PARAM_VALIDATE: every index structure the iteration kernels dereference ------------------
// The kernels' memory accesses are driven by these arrays (loop bounds from ptr / sptr, gather
// addresses from idx / sidx, scatter slots from perm); the epilogues are guarded by i < nrows.  So
// "arrays consistent and in range" is "no out-of-bounds access", checked here by brute force.
// bad[0] counts violations, bad[1] keeps the first offending code * 2^32 + position.
__device__ __forceinline__ void report_bad(unsigned long long* bad, int code, long long where) {
  atomicAdd(bad, 1ull);
  atomicCAS(bad + 1, 0ull, ((unsigned long long)code << 40) | (unsigned long long)(where & 0xffffffffffll));
}

__global__ void k_validate_ptr(const int* __restrict__ ptr, int count, long long total, int code,
                               unsigned long long* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > count) return;
  const int v = ptr[i];
  if (i == 0 && v != 0) report_bad(bad, code, 0);
  if (i == count && (long long)v != total) report_bad(bad, code, i);
  if (i < count && ptr[i + 1] < v) report_bad(bad, code, i);
}

__global__ void k_validate_idx(const int* __restrict__ idx, long long count, int lo, int hi, int code,
                               unsigned long long* bad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int v = idx[i];
  if (v < lo || v >= hi) report_bad(bad, code, i);
}

// slices of a SELL copy hold whole multiples of 32 entries; perm is a permutation inside each 256-row block
__global__ void k_validate_sell(const int* __restrict__ sptr, const unsigned char* __restrict__ perm, int nblk,
                                int code, unsigned long long* bad) {
  __shared__ int seen[kRowsPerCta];
  const int blk = blockIdx.x;
  seen[threadIdx.x] = 0;
  __syncthreads();
  atomicAdd(&seen[perm[(size_t)blk * kRowsPerCta + threadIdx.x]], 1);
  __syncthreads();
  if (seen[threadIdx.x] != 1) report_bad(bad, code, (long long)blk * kRowsPerCta + threadIdx.x);
  if (threadIdx.x < kRowsPerCta / 32) {
    const int s = blk * (kRowsPerCta / 32) + threadIdx.x;
    if ((sptr[s + 1] - sptr[s]) % 32 != 0) report_bad(bad, code + 1, s);
  }
}

void validate_csr(mipcu_ctx* ctx, const Csr& A, int code, unsigned long long* bad) {
  cudaStream_t st = ctx->stream;
  if (A.nrows <= 0 || !A.ptr) return;
  k_validate_ptr<<<grid_for(A.nrows + 1), kThreads, 0, st>>>(A.ptr, A.nrows, A.nnz, code, bad);
  if (A.nnz) k_validate_idx<<<grid_for(A.nnz), kThreads, 0, st>>>(A.idx, A.nnz, 0, A.ncols, code + 1, bad);
  ctx->launches += 2;
  if (A.sptr) {
    const int nblk = grid_for(A.nrows, kRowsPerCta), nslices = nblk * (kRowsPerCta / 32);
    k_validate_ptr<<<grid_for(nslices + 1), kThreads, 0, st>>>(A.sptr, nslices, A.sell_entries, code + 2, bad);
    if (A.sell_entries)
      k_validate_idx<<<grid_for(A.sell_entries), kThreads, 0, st>>>(A.sidx, A.sell_entries, -1, A.ncols, code + 3, bad);
    k_validate_sell<<<nblk, kRowsPerCta, 0, st>>>(A.sptr, A.perm, nblk, code + 4, bad);
    ctx->launches += 3;
  }
}

void validate_structures(mipcu_ctx* ctx) {
  NvtxRange nvtx("mipcu:validate");
  cudaStream_t st = ctx->stream;
  unsigned long long* bad = dalloc<unsigned long long>(2);
  MIPCU_CK(cudaMemsetAsync(bad, 0, 2 * sizeof(unsigned long long), st));
  if (ctx->params[MIPCU_PARAM_VALIDATE] == 2.0 && ctx->rows.nnz > 0) {
    // self-test of the validator (tests/test_validate_gpu.py): plant one out-of-range column index
    const int wrong = ctx->rows.ncols;
    int* where = ctx->rows.sidx ? ctx->rows.sidx : ctx->rows.idx;
    MIPCU_CK(cudaMemcpyAsync(where, &wrong, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  validate_csr(ctx, ctx->rows, 10, bad);
  validate_csr(ctx, ctx->cols, 20, bad);
  validate_csr(ctx, ctx->colblk, 30, bad);
  unsigned long long h[2] = {0, 0};
  MIPCU_CK(cudaMemcpyAsync(h, bad, sizeof(h), cudaMemcpyDeviceToHost, st));
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(bad);
  // shape relations the kernels rely on
  bool shape_ok = ctx->rows.nrows == ctx->m && ctx->rows.ncols == ctx->n && ctx->cols.nrows == ctx->n &&
                  ctx->cols.ncols == ctx->m && ctx->rows.nnz == ctx->cols.nnz;
  if (ctx->world > 1 && ctx->peers.world == ctx->world) {
    const PeerTable& t = ctx->peers;
    shape_ok = shape_ok && t.i0[ctx->rank + 1] - t.i0[ctx->rank] == ctx->m && t.i0[t.world] == ctx->m_glob &&
               t.j0[0] == 0 && t.j0[t.world] == ctx->n;
    for (int g = 0; g < t.world; ++g) shape_ok = shape_ok && t.i0[g] <= t.i0[g + 1] && t.j0[g] <= t.j0[g + 1];
    if (ctx->colblk.nrows > 0)
      shape_ok = shape_ok && ctx->colblk.nrows == t.j0[ctx->rank + 1] - t.j0[ctx->rank] && ctx->colblk.ncols == ctx->m_glob;
  }
  MIPCU_REQUIRE(shape_ok, "MIPCU_PARAM_VALIDATE: matrix copies / slice extents are inconsistent");
  if (h[0])
    throw Error("MIPCU_PARAM_VALIDATE: " + std::to_string(h[0]) + " index structure violation(s); first: structure code " +
                std::to_string((unsigned)(h[1] >> 40)) + " at position " + std::to_string((long long)(h[1] & 0xffffffffffull)) +
                " (10 rows, 20 columns, 30 column block; +0 ptr, +1 idx, +2 slice ptr, +3 slice idx, +4 perm, +5 slice size)");
}

// Ruiz x RUIZ_PASSES (row/col max-abs) then Pock-Chambolle alpha = 1 (row/col 1-norms), applied
// in place to both copies; then eta = STEP_SAFETY / sigma_max(K) by power iteration.
void prepare(mipcu_ctx* ctx) {
  MIPCU_REQUIRE(ctx->loaded, "no model loaded");
  if (ctx->prepared) return;
  NvtxRange nvtx("mipcu:setup (Ruiz / Pock-Chambolle scaling, SELL copies, step size)");
  auto t0 = std::chrono::steady_clock::now();
  cudaStream_t st = ctx->stream;
  const int m = ctx->m, n = ctx->n;
  Csr &R = ctx->rows, &C = ctx->cols;
  const bool scaling = ctx->params[MIPCU_PARAM_SCALING] != 0.0;
  if (scaling && (ctx->nnz > 0 || ctx->world > 1)) {
    double* rf = ctx->tmp_m;
    double* cf = ctx->tmp_n;
    for (int sweep = 0; sweep < RUIZ_PASSES + 1; ++sweep) {
      const int gr = grid_for((int64_t)m * 8), gc = grid_for((int64_t)n * 8);
      if (sweep < RUIZ_PASSES) {
        k_row_norm<0><<<gr, kThreads, 0, st>>>(R.ptr, R.val, rf, m);
        k_row_norm<0><<<gc, kThreads, 0, st>>>(C.ptr, C.val, cf, n);
      } else {
        k_row_norm<1><<<gr, kThreads, 0, st>>>(R.ptr, R.val, rf, m);
        k_row_norm<1><<<gc, kThreads, 0, st>>>(C.ptr, C.val, cf, n);
      }
      // row-sharded: a column's norm spans all ranks (max for Ruiz, sum for Pock-Chambolle)
      dist_allreduce(ctx, cf, (size_t)n, sweep < RUIZ_PASSES);
      k_factor<<<grid_for(m), kThreads, 0, st>>>(rf, ctx->d1, m);
      k_factor<<<grid_for(n), kThreads, 0, st>>>(cf, ctx->d2, n);
      k_scale_matrix<true><<<gr, kThreads, 0, st>>>(R.ptr, R.idx, R.val, rf, cf, m);
      k_scale_matrix<false><<<gc, kThreads, 0, st>>>(C.ptr, C.idx, C.val, rf, cf, n);
      ctx->launches += 6;
      if (ctx->colblk.nrows > 0 && ctx->yhat_full) {
        // block-owner exchange: the column block holds entries of every rank's rows, so it needs
        // all row factors; each row is owned by exactly one rank: place + sum assembles them
        // (yhat_full is free until the first iteration).  Same (val * r_i) * c_j as the row copy.
        Csr& B = ctx->colblk;
        double* rf_all = ctx->yhat_full;
        MIPCU_CK(cudaMemsetAsync(rf_all, 0, sizeof(double) * (size_t)ctx->m_glob, st));
        if (m) k_place<<<grid_for(m), kThreads, 0, st>>>(rf, rf_all, ctx->row_off[ctx->rank], m);
        dist_allreduce(ctx, rf_all, (size_t)ctx->m_glob, false);
        const int j0 = (int)(((int64_t)n * ctx->rank) / ctx->world);
        k_scale_matrix<false><<<grid_for((int64_t)B.nrows * 8), kThreads, 0, st>>>(B.ptr, B.idx, B.val, rf_all,
                                                                                cf + j0, B.nrows);
        ctx->launches += 2;
      }
    }
  }
  const bool verbose = ctx->params[MIPCU_PARAM_VERBOSE] != 0.0;
  auto lap = [&](const char* what) {
    if (!verbose) return;
    MIPCU_CK(cudaStreamSynchronize(st));
    fprintf(stderr, "[mipcu] prepare: %-16s at %.1f ms\n", what,
            1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  };
  lap("scaling");
  // lane-per-row copies of the scaled matrix for the iteration kernels
  build_sell(ctx, R);
  build_sell(ctx, C);
  if (ctx->colblk.nrows > 0) build_sell(ctx, ctx->colblk);
  lap("sell copies");
  // step size
  double step = ctx->params[MIPCU_PARAM_STEP_SIZE];
  if (step > 0.0) {
    ctx->eta = step;
  } else if (ctx->world <= 1 && (ctx->nnz == 0 || n == 0 || m == 0)) {
    ctx->eta = 1.0;
  } else {
    double* v = ctx->tmp_n;
    double* u = ctx->tmp_m;
    k_hash_vector<<<grid_for(n), kThreads, 0, st>>>(v, n);
    ctx->launches++;
    double nv = device_norm2(ctx, v, n);
    k_scale_inplace<<<grid_for(n), kThreads, 0, st>>>(v, 1.0 / nv, n);
    ctx->launches++;
    double prev = 0.0, s = 0.0;
    int it = 0;
    bool zero = false;
    while (it < POWER_MAX && !zero) {
      double s2 = 0.0;
      for (int b = 0; b < POWER_BATCH; ++b) {
        spmv_plain(ctx, R, v, u);
        spmv_plain(ctx, C, u, v);
        dist_allreduce(ctx, v, (size_t)n, false);
        s2 = device_norm2(ctx, v, n);
        if (s2 == 0.0) { zero = true; break; }
        k_scale_inplace<<<grid_for(n), kThreads, 0, st>>>(v, 1.0 / s2, n);
        ctx->launches++;
        ++it;
      }
      if (zero) break;
      s = std::sqrt(s2);
      if (std::fabs(s - prev) <= POWER_TOL * s) break;
      prev = s;
    }
    ctx->eta = (!zero && s > 0.0) ? STEP_SAFETY / s : 1.0;
    if (verbose) fprintf(stderr, "[mipcu] prepare: %d power iterations, sigma_max %.9g\n", it, s);
  }
  lap("step size");
  MIPCU_CK(cudaStreamSynchronize(st));
  if (ctx->params[MIPCU_PARAM_VALIDATE] != 0.0) validate_structures(ctx);
  ctx->prepared = true;
  ctx->vectors_dirty = true;
  ctx->setup_seconds =
      std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// scaled c / bounds / row sides, the norms of the termination test and the initial primal weight
void refresh_vectors(mipcu_ctx* ctx) {
  if (!ctx->vectors_dirty) return;
  cudaStream_t st = ctx->stream;
  const int m = ctx->m, n = ctx->n;
  const double sgn = ctx->maximize ? -1.0 : 1.0;
  if (n) {
    k_mul<<<grid_for(n), kThreads, 0, st>>>(ctx->c, ctx->d2, ctx->cs, n, sgn);
    k_div<<<grid_for(n), kThreads, 0, st>>>(ctx->lb, ctx->d2, ctx->ls, n);
    k_div<<<grid_for(n), kThreads, 0, st>>>(ctx->ub, ctx->d2, ctx->us, n);
    ctx->launches += 3;
  }
  if (m) {
    k_mul<<<grid_for(m), kThreads, 0, st>>>(ctx->lo, ctx->d1, ctx->los, m, 1.0);
    k_mul<<<grid_for(m), kThreads, 0, st>>>(ctx->hi, ctx->d1, ctx->his, m, 1.0);
    ctx->launches += 2;
  }
  ctx->cnorm = std::sqrt(sumsq<0>(ctx, ctx->c, nullptr, n));
  ctx->bnorm = std::sqrt(dist_sum_scalar(ctx, sumsq<1>(ctx, ctx->lo, ctx->hi, m)));
  double cs = std::sqrt(sumsq<0>(ctx, ctx->cs, nullptr, n));
  double bs = std::sqrt(dist_sum_scalar(ctx, sumsq<1>(ctx, ctx->los, ctx->his, m)));
  ctx->w0 = (cs > 0.0 && bs > 0.0) ? cs / bs : 1.0;
  ctx->vectors_dirty = false;
}

// IVar::set_lb / set_ub and branch-and-bound: overwrite the bounds of a few columns
void set_bounds(mipcu_ctx* ctx, int64_t n_changed, const int64_t* cols, const double* lb,
                const double* ub) {
  MIPCU_REQUIRE(ctx->loaded, "mipcu_set_bounds: no model loaded");
  MIPCU_REQUIRE(n_changed >= 0 && (n_changed == 0 || (cols && lb && ub)), "mipcu_set_bounds: null array");
  if (n_changed == 0) return;
  for (int64_t k = 0; k < n_changed; ++k)
    MIPCU_REQUIRE(cols[k] >= 0 && cols[k] < ctx->n, "mipcu_set_bounds: column out of range");
  MIPCU_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  long long* dcols = dalloc<long long>(n_changed);
  double* dl = dalloc<double>(2 * n_changed);
  double* du = dl + n_changed;
  MIPCU_CK(cudaMemcpyAsync(dcols, cols, sizeof(int64_t) * n_changed, cudaMemcpyHostToDevice, st));
  MIPCU_CK(cudaMemcpyAsync(dl, lb, sizeof(double) * n_changed, cudaMemcpyHostToDevice, st));
  MIPCU_CK(cudaMemcpyAsync(du, ub, sizeof(double) * n_changed, cudaMemcpyHostToDevice, st));
  k_scatter_bounds<<<grid_for(n_changed), kThreads, 0, st>>>(dcols, dl, du, ctx->lb, ctx->ub, ctx->col_inv, (int)n_changed);
  ctx->launches++;
  MIPCU_CK(cudaStreamSynchronize(st));
  dfree(dcols);
  dfree(dl);
  ctx->vectors_dirty = true;
}

// test hook: out = A v (or A' v) on the matrix AS LOADED.  Once prepare() has equilibrated the
// copies in place, A = D1^-1 K D2^-1, so the input is divided by one diagonal and the output by
// the other around the same SpMV kernel.
void spmv_host(mipcu_ctx* ctx, int transpose, const double* v, double* out) {
  MIPCU_REQUIRE(ctx->loaded, "mipcu_spmv: no model loaded");
  MIPCU_REQUIRE(v && out, "mipcu_spmv: null array");
  MIPCU_CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int n_in = transpose ? ctx->m : ctx->n, n_out = transpose ? ctx->n : ctx->m;
  double* d_in = transpose ? ctx->tmp_m : ctx->tmp_n;
  double* d_out = transpose ? ctx->tmp_n : ctx->tmp_m;
  const double* din_scale = transpose ? ctx->d1 : ctx->d2;
  const double* dout_scale = transpose ? ctx->d2 : ctx->d1;
  std::vector<double> staged;
  if (n_in && transpose && ctx->row_perm) {      // the caller's m-vector, in internal row order
    staged.assign(v, v + n_in);
    rows_to_internal_order(ctx, staged.data());
    v = staged.data();
  }
  if (n_in && !transpose && ctx->col_perm) {     // the caller's n-vector, in internal column order
    staged.assign(v, v + n_in);
    cols_to_internal_order(ctx, staged.data());
    v = staged.data();
  }
  if (n_in) MIPCU_CK(cudaMemcpyAsync(d_in, v, sizeof(double) * n_in, cudaMemcpyHostToDevice, st));
  if (ctx->prepared && n_in) {
    k_div_inplace<<<grid_for(n_in), kThreads, 0, st>>>(d_in, din_scale, n_in);
    ctx->launches++;
  }
  if (n_out) {
    if (n_in && ctx->nnz) spmv_plain(ctx, transpose ? ctx->cols : ctx->rows, d_in, d_out);
    else MIPCU_CK(cudaMemsetAsync(d_out, 0, sizeof(double) * n_out, st));
    if (ctx->prepared) {
      k_div_inplace<<<grid_for(n_out), kThreads, 0, st>>>(d_out, dout_scale, n_out);
      ctx->launches++;
    }
    MIPCU_CK(cudaMemcpyAsync(out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
  }
  MIPCU_CK(cudaStreamSynchronize(st));
  if (n_out && !transpose) rows_to_caller_order(ctx, out);
  if (n_out && transpose) cols_to_caller_order(ctx, out);
}

}  // namespace mipcu
