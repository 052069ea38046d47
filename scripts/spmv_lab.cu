// spmv_lab.cu - standalone kernel laboratory for the two fused SpMV + step kernels of the PDHG
// iteration (config 3 shape: 1M x 2M, 16 nnz per column, uniformly random rows).  It times storage
// formats / load hints / unroll depths side by side in ONE GPU call so that the winner can be
// moved into miplib_b200/csrc/.  Not part of the product; nothing links against it.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scripts/_lab/spmv_lab scripts/spmv_lab.cu
//   ./scripts/_lab/spmv_lab [m] [n] [nnz_per_col] [reps]
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int kThreads = 256;
constexpr int kRowsPerCta = 256;

// ------------------------------------------------------------------------------------------
// load flavours
// ------------------------------------------------------------------------------------------
enum { G_LDG = 0, G_NOALLOC = 1, G_EVLAST = 2, G_PLAIN = 3, G_TEX = 4 };
enum { S_NOALLOC = 0, S_EVFIRST = 1, S_LDG = 2 };

template <int SH> __device__ __forceinline__ double ld_sd(const double* p, uint64_t pol) {
  double v;
  if (SH == S_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (SH == S_EVFIRST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  else v = __ldg(p);
  return v;
}
template <int SH> __device__ __forceinline__ int ld_si(const int* p, uint64_t pol) {
  int v;
  if (SH == S_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  else if (SH == S_EVFIRST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  else v = __ldg(p);
  return v;
}
template <int GH> __device__ __forceinline__ double ld_g(const double* p, uint64_t pol) {
  double v;
  if (GH == G_LDG) v = __ldg(p);
  else if (GH == G_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (GH == G_EVLAST) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  else v = *p;
  return v;
}
template <int GH> __device__ __forceinline__ double ld_gt(const double* vec, cudaTextureObject_t tex, int c, uint64_t pol) {
  if (GH == G_TEX) { const int2 t = tex1Dfetch<int2>(tex, c); return __hiloint2double(t.y, t.x); }
  return ld_g<GH>(vec + c, pol);
}
__device__ __forceinline__ uint64_t pol_evict_first() {
  uint64_t p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t pol_evict_last() {
  uint64_t p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// ------------------------------------------------------------------------------------------
// epilogues with the memory traffic of the real kernels
// ------------------------------------------------------------------------------------------
struct Vecs {      // row mode uses v0..v3 (read) w0,w1 (write); col mode v0..v6, w0, w1
  const double* v[7];
  double* w[2];
  double s0, s1;
};
template <int MODE> struct EpiRegs { double r[MODE == 0 ? 4 : 7]; };

template <int MODE> __device__ __forceinline__ void epi_load(const Vecs& a, int i, EpiRegs<MODE>& e) {
  constexpr int N = MODE == 0 ? 4 : 7;
#pragma unroll
  for (int q = 0; q < N; ++q) e.r[q] = a.v[q][i];
}
template <int MODE> __device__ __forceinline__ void epi_apply(const Vecs& a, int i, double s, const EpiRegs<MODE>& e) {
  if (MODE == 0) {   // row step: y, y0, lo, hi -> yhat (w0), y (w1)
    const double yi = e.r[0], y0 = e.r[1], lo = e.r[2], hi = e.r[3];
    const double t = s - yi / a.s0;
    const double yh = a.s0 * (clampd(t, lo, hi) - t);
    a.w[0][i] = yh;
    a.w[1][i] = a.s1 * (2.0 * yh - yi) + (1.0 - a.s1) * y0;
  } else {           // col step: xbar, x0, aty, aty0, c, l, u -> xbar (w0), aty (w1)
    const double xb = e.r[0], x0 = e.r[1], at = e.r[2], at0 = e.r[3], c = e.r[4], l = e.r[5], u = e.r[6];
    const double xn = a.s1 * xb + (1.0 - a.s1) * x0;
    const double atn = a.s1 * (2.0 * s - at) + (1.0 - a.s1) * at0;
    const double xh = clampd(xn - a.s0 * (c - atn), l, u);
    a.w[0][i] = 2.0 * xh - xn;
    a.w[1][i] = atn;
  }
}

// ------------------------------------------------------------------------------------------
// A. the product kernel as it is today: CSR, TPR lanes per row
// ------------------------------------------------------------------------------------------
template <int TPR, int MODE, int GH, int SH>
__global__ void __launch_bounds__(kThreads)
k_csr(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
      const double* __restrict__ vec, int nrows, Vecs a) {
  __shared__ double s_sum[kRowsPerCta];
  constexpr int SUBS = kThreads / TPR;
  const uint64_t pf = SH == S_EVFIRST ? pol_evict_first() : 0, pl = GH == G_EVLAST ? pol_evict_last() : 0;
  const int lane = threadIdx.x % TPR, sub = threadIdx.x / TPR;
  const int row0 = blockIdx.x * kRowsPerCta;
#pragma unroll 2
  for (int r = sub; r < kRowsPerCta; r += SUBS) {
    const int row = row0 + r;
    double acc0 = 0.0, acc1 = 0.0;
    if (row < nrows) {
      const int b = ptr[row], e = ptr[row + 1];
      int k = b + lane;
      for (; k + TPR < e; k += 2 * TPR) {
        const double v0 = ld_sd<SH>(val + k, pf);
        const int c0 = ld_si<SH>(idx + k, pf);
        const double v1 = ld_sd<SH>(val + k + TPR, pf);
        const int c1 = ld_si<SH>(idx + k + TPR, pf);
        acc0 = fma(v0, ld_g<GH>(vec + c0, pl), acc0);
        acc1 = fma(v1, ld_g<GH>(vec + c1, pl), acc1);
      }
      if (k < e) acc0 = fma(ld_sd<SH>(val + k, pf), ld_g<GH>(vec + ld_si<SH>(idx + k, pf), pl), acc0);
    }
    double acc = acc0 + acc1;
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, TPR);
    if (lane == 0) s_sum[r] = acc;
  }
  __syncthreads();
  const int i = row0 + threadIdx.x;
  if (i < nrows) {
    EpiRegs<MODE> e;
    epi_load<MODE>(a, i, e);
    epi_apply<MODE>(a, i, s_sum[threadIdx.x], e);
  }
}

// ------------------------------------------------------------------------------------------
// B. sliced ELL: slices of 32 rows stored lane-major (entry j of lane l at off + 32 j + l), rows
//    sorted by length inside a window of SIGMA rows (256 = one CTA; a local byte permutation maps
//    the sorted position back to the CTA's row), one lane per row, U entries in flight per lane.
// ------------------------------------------------------------------------------------------
template <int U, int MODE, int GH, int SH, bool PRE, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_sell(const int* __restrict__ sptr, const int* __restrict__ sidx, const double* __restrict__ sval,
       const unsigned char* __restrict__ perm, const double* __restrict__ vec, int nrows, Vecs a, cudaTextureObject_t tex) {
  __shared__ double s_sum[kRowsPerCta];
  const uint64_t pf = SH == S_EVFIRST ? pol_evict_first() : 0, pl = GH == G_EVLAST ? pol_evict_last() : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * kRowsPerCta;
  const int i = row0 + threadIdx.x;
  EpiRegs<MODE> e;
  if (PRE && i < nrows) epi_load<MODE>(a, i, e);
  const int slice = blockIdx.x * (kRowsPerCta / 32) + warp;
  const int beg = sptr[slice], end = sptr[slice + 1];
  double acc[2] = {0.0, 0.0};
  for (int p = beg + lane; p < end; p += 32 * U) {
    double v[U]; int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = p + 32 * u < end;
      v[u] = ok ? ld_sd<SH>(sval + p + 32 * u, pf) : 0.0;
      c[u] = ok ? ld_si<SH>(sidx + p + 32 * u, pf) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double x = c[u] >= 0 ? ld_gt<GH>(vec, tex, c[u], pl) : 0.0;
      acc[u & 1] = fma(v[u], x, acc[u & 1]);
    }
  }
  const int r = threadIdx.x;
  const int dst = perm ? (int)perm[row0 + r] : r;
  s_sum[dst] = acc[0] + acc[1];
  __syncthreads();
  if (i < nrows) {
    if (!PRE) epi_load<MODE>(a, i, e);
    epi_apply<MODE>(a, i, s_sum[threadIdx.x], e);
  }
}

// ------------------------------------------------------------------------------------------
// C. ceilings: gather only (indices streamed, no values) and stream only (no gather)
// ------------------------------------------------------------------------------------------
template <int U, int GH>
__global__ void __launch_bounds__(kThreads)
k_gather_only(const int* __restrict__ sidx, const double* __restrict__ vec, long long total, double* out, cudaTextureObject_t tex) {
  const uint64_t pl = GH == G_EVLAST ? pol_evict_last() : 0;
  long long p = ((long long)blockIdx.x * kThreads + threadIdx.x);
  const long long stride = (long long)gridDim.x * kThreads;
  double acc = 0.0;
  for (; p + (U - 1) * stride < total; p += U * stride) {
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ld_si<S_NOALLOC>(sidx + p + u * stride, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += c[u] >= 0 ? ld_gt<GH>(vec, tex, c[u], pl) : 0.0;
  }
  if (acc == 123.456) out[0] = acc;
}
template <int U>
__global__ void __launch_bounds__(kThreads)
k_stream_only(const int* __restrict__ sidx, const double* __restrict__ sval, long long total, double* out) {
  long long p = ((long long)blockIdx.x * kThreads + threadIdx.x);
  const long long stride = (long long)gridDim.x * kThreads;
  double acc = 0.0;
  for (; p + (U - 1) * stride < total; p += U * stride) {
#pragma unroll
    for (int u = 0; u < U; ++u) acc += ld_sd<S_NOALLOC>(sval + p + u * stride, 0) * (double)ld_si<S_NOALLOC>(sidx + p + u * stride, 0);
  }
  if (acc == 123.456) out[0] = acc;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct HostCsr { int nrows, ncols; std::vector<int> ptr, idx; std::vector<double> val; };
struct HostSell { int nrows, nslices; std::vector<int> sptr, sidx; std::vector<double> sval; std::vector<unsigned char> perm; std::vector<int> gperm; double pad_frac; };

static HostSell to_sell(const HostCsr& A, int sigma /* 256 or 0 = global */) {
  HostSell S; S.nrows = A.nrows;
  const int nblk = (A.nrows + kRowsPerCta - 1) / kRowsPerCta;
  S.nslices = nblk * (kRowsPerCta / 32);
  S.sptr.assign(S.nslices + 1, 0);
  std::vector<int> order(nblk * kRowsPerCta);   // sorted position -> row (or -1)
  for (int i = 0; i < nblk * kRowsPerCta; ++i) order[i] = i < A.nrows ? i : -1;
  auto len = [&](int r) { return r < 0 ? -1 : A.ptr[r + 1] - A.ptr[r]; };
  if (sigma == 256) {
    for (int b = 0; b < nblk; ++b)
      std::stable_sort(order.begin() + b * 256, order.begin() + (b + 1) * 256, [&](int x, int y) { return len(x) > len(y); });
    S.perm.resize(nblk * 256);
    for (int i = 0; i < nblk * 256; ++i) S.perm[i] = order[i] < 0 ? (unsigned char)(i & 255) : (unsigned char)(order[i] & 255);
    // padded rows keep distinct targets: fix up so that perm is a permutation inside a block
    for (int b = 0; b < nblk; ++b) {
      bool used[256] = {};
      for (int r = 0; r < 256; ++r) if (order[b * 256 + r] >= 0) used[S.perm[b * 256 + r]] = true;
      int nxt = 0;
      for (int r = 0; r < 256; ++r) if (order[b * 256 + r] < 0) { while (used[nxt]) ++nxt; S.perm[b * 256 + r] = nxt; used[nxt] = true; }
    }
  } else {
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return len(x) > len(y); });
    S.gperm = order;
  }
  long long tot = 0;
  for (int s = 0; s < S.nslices; ++s) {
    int w = 0;
    for (int l = 0; l < 32; ++l) w = std::max(w, len(order[s * 32 + l]));
    S.sptr[s] = (int)tot; tot += 32LL * w;
  }
  S.sptr[S.nslices] = (int)tot;
  S.sidx.assign(tot, -1); S.sval.assign(tot, 0.0);
  for (int s = 0; s < S.nslices; ++s)
    for (int l = 0; l < 32; ++l) {
      const int r = order[s * 32 + l];
      if (r < 0) continue;
      for (int j = 0; j < len(r); ++j) { S.sidx[S.sptr[s] + 32 * j + l] = A.idx[A.ptr[r] + j]; S.sval[S.sptr[s] + 32 * j + l] = A.val[A.ptr[r] + j]; }
    }
  S.pad_frac = (double)(tot - A.ptr[A.nrows]) / (double)A.ptr[A.nrows];
  return S;
}

template <class T> static T* up(const std::vector<T>& v) { T* d; CK(cudaMalloc(&d, std::max<size_t>(1, v.size()) * sizeof(T) + 4096)); CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

struct Dev { int* ptr; int* idx; double* val; int* sptr; int* sidx; double* sval; unsigned char* perm; int* gsptr; int* gsidx; double* gsval; int nrows, ncols; long long nnz, sell_tot, gsell_tot; std::vector<int> gperm; };

int main(int argc, char** argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 1000000;
  const int n = argc > 2 ? atoi(argv[2]) : 2000000;
  const int k = argc > 3 ? atoi(argv[3]) : 16;
  const int reps = argc > 4 ? atoi(argv[4]) : 20;
  const long long nnz = (long long)n * k;
  printf("lab: m=%d n=%d nnz=%lld reps=%d\n", m, n, nnz, reps);

  // K by columns (CSR of K'): exactly k sorted random rows per column
  HostCsr C; C.nrows = n; C.ncols = m; C.ptr.resize(n + 1); C.idx.resize(nnz); C.val.resize(nnz);
  std::mt19937_64 rng(20263);
  std::uniform_real_distribution<double> uv(0.1, 1.0);
  for (int j = 0; j < n; ++j) {
    C.ptr[j] = j * k;
    int* p = &C.idx[(size_t)j * k];
    for (int q = 0; q < k; ++q) p[q] = (int)(rng() % (uint64_t)m);
    std::sort(p, p + k);
    for (int q = 0; q < k; ++q) C.val[(size_t)j * k + q] = uv(rng);
  }
  C.ptr[n] = (int)nnz;
  // K by rows
  HostCsr R; R.nrows = m; R.ncols = n; R.ptr.assign(m + 1, 0); R.idx.resize(nnz); R.val.resize(nnz);
  for (long long q = 0; q < nnz; ++q) R.ptr[C.idx[q] + 1]++;
  for (int i = 0; i < m; ++i) R.ptr[i + 1] += R.ptr[i];
  { std::vector<int> cur(R.ptr.begin(), R.ptr.end() - 1);
    for (int j = 0; j < n; ++j) for (int q = C.ptr[j]; q < C.ptr[j + 1]; ++q) { const int d = cur[C.idx[q]]++; R.idx[d] = j; R.val[d] = C.val[q]; } }
  printf("generated\n"); fflush(stdout);

  std::vector<double> hx(n), hy(m);
  for (auto& v : hx) v = uv(rng);
  for (auto& v : hy) v = uv(rng) - 0.5;
  double* dx = up(hx); double* dy = up(hy);
  const int big = std::max(m, n);
  std::vector<double> hv(big);
  Vecs proto{};
  double* scratch[9];
  for (int q = 0; q < 9; ++q) { for (auto& v : hv) v = uv(rng) - 0.3; scratch[q] = up(hv); }

  auto mk = [&](const HostCsr& A) {
    Dev D{}; D.nrows = A.nrows; D.ncols = A.ncols; D.nnz = A.ptr[A.nrows];
    D.ptr = up(A.ptr); D.idx = up(A.idx); D.val = up(A.val);
    HostSell S = to_sell(A, 256);
    D.sptr = up(S.sptr); D.sidx = up(S.sidx); D.sval = up(S.sval); D.perm = up(S.perm); D.sell_tot = S.sptr[S.nslices];
    printf("  SELL-32-256: padding %.2f %%\n", 100 * S.pad_frac);
    HostSell G = to_sell(A, 0);
    D.gsptr = up(G.sptr); D.gsidx = up(G.sidx); D.gsval = up(G.sval); D.gsell_tot = G.sptr[G.nslices]; D.gperm = G.gperm;
    printf("  SELL-32-global: padding %.2f %%\n", 100 * G.pad_frac);
    return D;
  };
  printf("rows:\n"); Dev DR = mk(R);
  printf("cols:\n"); Dev DC = mk(C);
  fflush(stdout);

  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double* dout[2]; CK(cudaMalloc(&dout[0], big * 8)); CK(cudaMalloc(&dout[1], big * 8));
  double* dref; CK(cudaMalloc(&dref, big * 8));
  std::vector<double> href(big), hres(big);

  for (int side = 0; side < 2; ++side) {
    const Dev& D = side == 0 ? DR : DC;
    const double* vec = side == 0 ? dx : dy;
    const int nr = D.nrows;
    cudaTextureObject_t tex = 0;
    { cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = (void*)vec;
      rd.res.linear.desc = cudaCreateChannelDesc<int2>(); rd.res.linear.sizeInBytes = (size_t)D.ncols * 8;
      cudaTextureDesc td{}; td.readMode = cudaReadModeElementType; td.addressMode[0] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
      CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr)); }
    const int nblk = (nr + kRowsPerCta - 1) / kRowsPerCta;
    const long long abytes = side == 0 ? 12 * D.nnz + 4LL * (nr + 1) + 8LL * D.ncols + 48LL * nr
                                       : 12 * D.nnz + 4LL * (nr + 1) + 8LL * D.ncols + 72LL * nr;
    Vecs a{};
    for (int q = 0; q < 7; ++q) a.v[q] = scratch[q];
    a.w[0] = dout[0]; a.w[1] = dout[1]; a.s0 = 0.37; a.s1 = 0.9;
    printf("=== side %s: rows=%d nnz=%lld algorithmic MB=%.1f\n", side == 0 ? "ROW (K xbar)" : "COL (K' yhat)", nr, D.nnz, abytes / 1e6);
    bool have_ref = false;
    auto run = [&](const char* name, std::function<void()> launch, bool check, const std::vector<int>* gperm) {
      for (int w = 0; w < 3; ++w) launch();
      CK(cudaStreamSynchronize(st));
      CK(cudaEventRecord(e0, st));
      for (int r = 0; r < reps; ++r) launch();
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1e3 / reps;
      double err = -1;
      if (check) {
        CK(cudaMemcpy(hres.data(), dout[0], nr * 8, cudaMemcpyDeviceToHost));
        if (!have_ref) { href = hres; have_ref = true; err = 0; }
        else {
          err = 0;
          if (!gperm) for (int i = 0; i < nr; ++i) err = std::max(err, fabs(hres[i] - href[i]));
          else err = -2;   // permuted space: epilogue vectors differ, not comparable element-wise
        }
      }
      printf("%-44s %8.1f us  %7.0f GB/s  %5.1f %% of 6539   maxdiff %.2e\n", name, us, abytes / us / 1e3, 100 * abytes / us / 1e3 / 6538.9, err);
      fflush(stdout);
    };
#define CSR(TPR, GH, SH) run("csr tpr" #TPR " g" #GH " s" #SH, [&] { if (side == 0) k_csr<TPR, 0, GH, SH><<<nblk, kThreads, 0, st>>>(D.ptr, D.idx, D.val, vec, nr, a); else k_csr<TPR, 1, GH, SH><<<nblk, kThreads, 0, st>>>(D.ptr, D.idx, D.val, vec, nr, a); }, true, nullptr)
#define SELL(U, GH, SH, PRE, MINB) run("sell256 U" #U " g" #GH " s" #SH " pre" #PRE " minb" #MINB, [&] { if (side == 0) k_sell<U, 0, GH, SH, PRE, MINB><<<nblk, kThreads, 0, st>>>(D.sptr, D.sidx, D.sval, D.perm, vec, nr, a, tex); else k_sell<U, 1, GH, SH, PRE, MINB><<<nblk, kThreads, 0, st>>>(D.sptr, D.sidx, D.sval, D.perm, vec, nr, a, tex); }, true, nullptr)
#define SELLG(U, GH, SH, PRE, MINB) run("sellG   U" #U " g" #GH " s" #SH " pre" #PRE " minb" #MINB, [&] { if (side == 0) k_sell<U, 0, GH, SH, PRE, MINB><<<nblk, kThreads, 0, st>>>(D.gsptr, D.gsidx, D.gsval, nullptr, vec, nr, a, tex); else k_sell<U, 1, GH, SH, PRE, MINB><<<nblk, kThreads, 0, st>>>(D.gsptr, D.gsidx, D.gsval, nullptr, vec, nr, a, tex); }, true, &D.gperm)
    CSR(8, 0, 0);
    CSR(4, 0, 0);
    SELL(2, 0, 0, false, 8);
    SELL(4, 0, 0, false, 8);
    SELL(8, 0, 0, false, 6);
    SELL(8, 0, 0, false, 4);
    SELL(16, 0, 0, false, 4);
    SELL(8, 0, 0, true, 6);
    if (getenv("LAB_SHORT")) continue;
    // ceilings
    const int g = 148 * 8;
    run("ceiling: gather only U4", [&] { k_gather_only<4, 0><<<g, kThreads, 0, st>>>(D.sidx, vec, D.sell_tot, dref, tex); }, false, nullptr);
    run("ceiling: gather only U8", [&] { k_gather_only<8, 0><<<g, kThreads, 0, st>>>(D.sidx, vec, D.sell_tot, dref, tex); }, false, nullptr);
    run("ceiling: gather only U8 noalloc", [&] { k_gather_only<8, 1><<<g, kThreads, 0, st>>>(D.sidx, vec, D.sell_tot, dref, tex); }, false, nullptr);
    run("ceiling: gather only TEX U8", [&] { k_gather_only<8, 4><<<g, kThreads, 0, st>>>(D.sidx, vec, D.sell_tot, dref, tex); }, false, nullptr);
    run("ceiling: gather only TEX U16", [&] { k_gather_only<16, 4><<<g, kThreads, 0, st>>>(D.sidx, vec, D.sell_tot, dref, tex); }, false, nullptr);
    run("ceiling: gather only U16", [&] { k_gather_only<16, 0><<<g, kThreads, 0, st>>>(D.sidx, vec, D.sell_tot, dref, tex); }, false, nullptr);
    run("ceiling: stream only U8", [&] { k_stream_only<8><<<g, kThreads, 0, st>>>(D.sidx, D.sval, D.sell_tot, dref); }, false, nullptr);
  }
  printf("done\n");
  return 0;
}
