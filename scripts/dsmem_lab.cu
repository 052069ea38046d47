// dsmem_lab.cu - microbenchmark behind the "vector panel in distributed shared memory" question
// (VERDICT round 1, item 2): how many random 8-byte gathers per SM-cycle can a thread-block cluster
// serve out of its CTAs' shared memory (ld.shared::cluster), alone and side by side with the usual
// gathers through L1 -> XBAR -> L2?  The fused SpMV kernels of miplib_b200/csrc/kernels.cuh are
// bound by one L1-miss request per gathered double (about 0.93 requests per SM-cycle); a panel of
// the gathered vector held in DSMEM only helps if the SM-to-SM path adds request slots.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scripts/_lab/dsmem_lab scripts/dsmem_lab.cu
//   ./scripts/_lab/dsmem_lab [total_gathers]
//
// One kernel: every thread streams 4-byte indices (coalesced, L1::no_allocate) and gathers one
// double per index: index c < P comes out of the cluster's panel (element c lives in CTA
// (c >> 4) & (CS - 1) of the cluster, 128-byte chunks dealt round-robin), c >= P out of global
// memory.  Indices are uniform in [0, N): the DSMEM share of the gathers is P / N.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ int ld_si(const int* p) {
  int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem(uint32_t local_addr, uint32_t rank) {
  uint32_t ra; double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra));
  return v;
}

struct Clk { long long cyc; unsigned long long ns; };

template <int U, int LOGCS>
__global__ void k_mix(const int* __restrict__ idx, long long total, const double* __restrict__ vec,
                      int P, int W, double* out, Clk* clk) {
  extern __shared__ double s_panel[];
  constexpr int CS = 1 << LOGCS;
  const uint32_t rank = LOGCS ? cluster_rank() : 0;
  long long c0 = 0; unsigned long long t0 = 0;
  if (threadIdx.x == 0) { c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
  // this CTA's share of the panel: chunks of 16 doubles, chunk q of the panel lives in CTA q % CS
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const int chunk = i >> 4, in = i & 15;
    const int src = ((chunk * CS + (int)rank) << 4) | in;
    s_panel[i] = src < P ? vec[src] : 0.0;
  }
  if (LOGCS) cluster_sync_all(); else __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_panel);
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  for (; p + (U - 1) * stride < total; p += U * stride) {
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ld_si(idx + p + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double x;
      if (c[u] < P) {
        const uint32_t r = ((uint32_t)c[u] >> 4) & (CS - 1);
        const uint32_t off = (((uint32_t)c[u] >> (4 + LOGCS)) << 4) | ((uint32_t)c[u] & 15u);
        if (LOGCS) x = ld_dsmem(sbase + off * 8u, r); else x = s_panel[off];
      } else {
        x = __ldg(vec + c[u]);
      }
      if (u & 1) acc1 += x; else acc0 += x;
    }
  }
  if (LOGCS) cluster_sync_all();
  if (threadIdx.x == 0 && clk) {
    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    clk[blockIdx.x].cyc = clock64() - c0; clk[blockIdx.x].ns = t1 - t0;
  }
  if (acc0 + acc1 == 123.456) out[0] = acc0;
  // checksum for the correctness test
  if (out && total <= (1 << 20)) atomicAdd(out + 1, acc0 + acc1);
}


// ------------------------------------------------------------------------------------------
// Other ways to issue the same random gathers out of L2 (no shared-memory panel): does any path
// avoid the one-line-per-cycle tag stage of the L1?
//   F_LDG  ld.global.nc (what the product uses)      F_CG   ld.global.cg (L2 only, no L1 line)
//   F_CV   ld.volatile.global                        F_LDGSTS cp.async 8 B into shared memory
//   F_BULK cp.async.bulk 16 B (TMA engine) into shared memory, completion on an mbarrier
// ------------------------------------------------------------------------------------------
enum { F_LDG = 0, F_CG = 1, F_CV = 2, F_LDGSTS = 3, F_BULK = 4, F_MIX2 = 5, F_MIX4 = 6 };   // F_MIXk: k of the U gathers by cp.async.bulk, the rest by ld.global.nc

template <int U, int F>
__global__ void k_flavour(const int* __restrict__ idx, long long total, const double* __restrict__ vec, double* out, Clk* clk) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  long long c0 = 0; unsigned long long t0 = 0;
  if (threadIdx.x == 0) { c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  // F_LDGSTS / F_BULK: a private landing zone of U 16-byte slots per thread; F_BULK: one mbarrier per warp
  double2* slots = reinterpret_cast<double2*>(s_raw) + (size_t)threadIdx.x * (F == F_MIX2 ? 2 : F == F_MIX4 ? 4 : U);
  const uint32_t slot_addr = (uint32_t)__cvta_generic_to_shared(slots);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_raw + (size_t)blockDim.x * (F == F_MIX2 ? 2 : F == F_MIX4 ? 4 : U) * 16);
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bars + (threadIdx.x >> 5));
  uint32_t phase = 0;
  constexpr int UT = F == F_BULK ? U : F == F_MIX2 ? 2 : F == F_MIX4 ? 4 : 0;   // gathers by the TMA engine
  if (UT > 0) {
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  for (; p + (U - 1) * stride < total; p += U * stride) {
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ld_si(idx + p + u * stride);
    if (F == F_LDG || F == F_CG || F == F_CV) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        double x;
        if (F == F_LDG) x = __ldg(vec + c[u]);
        else if (F == F_CG) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(x) : "l"(vec + c[u]));
        else asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(x) : "l"(vec + c[u]));
        if (u & 1) acc1 += x; else acc0 += x;
      }
    } else if (F == F_LDGSTS) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(slot_addr + 16u * u), "l"(vec + c[u]) : "memory");
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int u = 0; u < U; ++u) { const double x = slots[u].x; if (u & 1) acc1 += x; else acc0 += x; }
    } else {
      if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_addr), "r"(32u * UT * 16u) : "memory");
      __syncwarp();
#pragma unroll
      for (int u = 0; u < UT; ++u)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                     :: "r"(slot_addr + 16u * u), "l"(vec + (c[u] & ~1)), "r"(bar_addr) : "memory");
#pragma unroll
      for (int u = UT; u < U; ++u) { const double x = __ldg(vec + c[u]); if (u & 1) acc1 += x; else acc0 += x; }
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar_addr), "r"(phase) : "memory");
      phase ^= 1;
#pragma unroll
      for (int u = 0; u < UT; ++u) { const double x = (c[u] & 1) ? slots[u].y : slots[u].x; if (u & 1) acc1 += x; else acc0 += x; }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0 && clk) {
    unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    clk[blockIdx.x].cyc = clock64() - c0; clk[blockIdx.x].ns = t1 - t0;
  }
  if (acc0 + acc1 == 123.456) out[0] = acc0;
  if (out && total <= (1 << 20)) atomicAdd(out + 1, acc0 + acc1);
}

template <class T> static T* up(const std::vector<T>& v) { T* d; CK(cudaMalloc(&d, v.size() * sizeof(T) + 4096)); CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

static int g_sms = 148;

template <int U, int LOGCS>
static void run(const char* tag, int threads, int W, int N, long long total, const double* dvec, const std::vector<double>& hvec, int reps, int p_override = -1) {
  constexpr int CS = 1 << LOGCS;
  const int P = p_override >= 0 ? p_override : (int)std::min<long long>((long long)CS * W, N);
  // indices
  static std::vector<int> hidx; static int last_n = -1; static int* didx = nullptr; static long long last_total = 0;
  if (last_n != N || last_total != total) {
    hidx.resize(total);
    std::mt19937_64 rng(7 + N);
    for (auto& v : hidx) v = (int)(rng() % (uint64_t)N);
    if (didx) CK(cudaFree(didx));
    didx = up(hidx); last_n = N; last_total = total;
  }
  auto kern = k_mix<U, LOGCS>;
  const size_t smem = (size_t)W * 8;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {(unsigned)CS, 1, 1};
  cfg.attrs = at; cfg.numAttrs = LOGCS ? 1 : 0;
  int nclusters = 0, ctas_per_sm = 0;
  if (LOGCS) {
    cfg.gridDim = dim3(CS * 64);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e != cudaSuccess || nclusters == 0) { printf("%-34s cluster %2d x %4d thr, %3zu KB smem: not launchable (%s)\n", tag, CS, threads, smem >> 10, cudaGetErrorString(e)); cudaGetLastError(); return; }
    cfg.gridDim = dim3(nclusters * CS);
  } else {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, threads, smem));
    cfg.gridDim = dim3(g_sms * ctas_per_sm);
  }
  double* dout; CK(cudaMalloc(&dout, 64)); CK(cudaMemset(dout, 0, 64));
  Clk* dclk; CK(cudaMalloc(&dclk, sizeof(Clk) * cfg.gridDim.x));
  // correctness on a small prefix: checksum against the host
  {
    const long long small = std::min<long long>(total, 1 << 20);
    const long long stride = (long long)cfg.gridDim.x * threads;
    CK(cudaLaunchKernelEx(&cfg, kern, (const int*)didx, small, dvec, P, W, dout, (Clk*)nullptr));
    CK(cudaDeviceSynchronize());
    double got[2]; CK(cudaMemcpy(got, dout, 16, cudaMemcpyDeviceToHost));
    double want = 0;   // the kernel drops the ragged tail of every thread
    for (long long t = 0; t < stride; ++t)
      for (long long q = t; q + (U - 1) * stride < small; q += U * stride)
        for (int u = 0; u < U; ++u) want += hvec[hidx[q + u * stride]];
    if (fabs(got[1] - want) > 1e-6 * (1 + fabs(want))) printf("   !! checksum mismatch: got %.9g want %.9g\n", got[1], want);
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) CK(cudaLaunchKernelEx(&cfg, kern, (const int*)didx, total, dvec, P, W, dout, dclk));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) CK(cudaLaunchKernelEx(&cfg, kern, (const int*)didx, total, dvec, P, W, dout, dclk));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<Clk> hc(cfg.gridDim.x); CK(cudaMemcpy(hc.data(), dclk, sizeof(Clk) * hc.size(), cudaMemcpyDeviceToHost));
  double mhz = 0; for (auto& c : hc) mhz += (double)c.cyc / (double)c.ns * 1e3; mhz /= hc.size();
  const double us = ms * 1e3 / reps;
  const int sms_used = LOGCS ? std::min(g_sms, nclusters * CS) : g_sms;
  const double per_sm_cycle = (double)total / (us * mhz) / sms_used;
  printf("%-34s CS %2d thr %4d smem %3zu KB grid %4u (%3d SMs) N %8d f_dsmem %.2f U%d: %8.1f us  %5.2f gathers/SM-cycle (DSMEM %.2f + L2 %.2f)  %.0f MHz\n",
         tag, CS, threads, smem >> 10, cfg.gridDim.x, sms_used, N, (double)P / N, U, us, per_sm_cycle,
         per_sm_cycle * P / N, per_sm_cycle * (1 - (double)P / N), mhz);
  fflush(stdout);
  CK(cudaFree(dout)); CK(cudaFree(dclk)); CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}


template <int U, int F>
static void run_flavour(const char* tag, int threads, int ctas_per_sm_cap, int N, long long total, const double* dvec, const std::vector<double>& hvec, int reps) {
  static std::vector<int> hidx; static int last_n = -1; static int* didx = nullptr;
  if (last_n != N) {
    hidx.resize(total);
    std::mt19937_64 rng(7 + N);
    for (auto& v : hidx) v = (int)(rng() % (uint64_t)N);
    if (didx) CK(cudaFree(didx));
    didx = up(hidx); last_n = N;
  }
  auto kern = k_flavour<U, F>;
  const size_t smem = F >= F_LDGSTS ? (size_t)threads * (F == F_MIX2 ? 2 : F == F_MIX4 ? 4 : U) * 16 + 8 * (threads / 32) : 0;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (ctas_per_sm_cap > 0) occ = std::min(occ, ctas_per_sm_cap);
  const int grid = g_sms * occ;
  double* dout; CK(cudaMalloc(&dout, 64)); CK(cudaMemset(dout, 0, 64));
  Clk* dclk; CK(cudaMalloc(&dclk, sizeof(Clk) * grid));
  {
    const long long small = std::min<long long>(total, 1 << 20);
    const long long stride = (long long)grid * threads;
    kern<<<grid, threads, smem>>>(didx, small, dvec, dout, nullptr);
    CK(cudaDeviceSynchronize());
    double got[2]; CK(cudaMemcpy(got, dout, 16, cudaMemcpyDeviceToHost));
    double want = 0;
    for (long long t = 0; t < stride; ++t)
      for (long long q = t; q + (U - 1) * stride < small; q += U * stride)
        for (int u = 0; u < U; ++u) want += hvec[hidx[q + u * stride]];
    if (fabs(got[1] - want) > 1e-6 * (1 + fabs(want))) printf("   !! checksum mismatch: got %.9g want %.9g\n", got[1], want);
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) kern<<<grid, threads, smem>>>(didx, total, dvec, dout, dclk);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) kern<<<grid, threads, smem>>>(didx, total, dvec, dout, dclk);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<Clk> hc(grid); CK(cudaMemcpy(hc.data(), dclk, sizeof(Clk) * hc.size(), cudaMemcpyDeviceToHost));
  double mhz = 0; for (auto& c : hc) mhz += (double)c.cyc / (double)c.ns * 1e3; mhz /= hc.size();
  const double us = ms * 1e3 / reps;
  printf("%-34s thr %4d x %d CTAs/SM smem %3zu KB/CTA N %8d U%-2d: %8.1f us  %5.2f gathers/SM-cycle  %.0f MHz\n",
         tag, threads, occ, smem >> 10, N, U, us, (double)total / (us * mhz) / g_sms, mhz);
  fflush(stdout);
  CK(cudaFree(dout)); CK(cudaFree(dclk)); CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
  const long long total = argc > 1 ? atoll(argv[1]) : 32000000LL;
  const int reps = 10;
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); g_sms = pr.multiProcessorCount;
  printf("dsmem_lab: %s, %d SMs, %lld gathers per launch\n", pr.name, g_sms, total);
  const int NMAX = 2000000;
  std::vector<double> hvec(NMAX);
  std::mt19937_64 rng(1); std::uniform_real_distribution<double> uv(0.1, 1.0);
  for (auto& v : hvec) v = uv(rng);
  double* dvec = up(hvec);
  const int W128 = 16384, W200 = 25600, W64 = 8192;
  const char* mode = argc > 2 ? argv[2] : "all";
  if (!strcmp(mode, "flavours") || !strcmp(mode, "all")) {
    run_flavour<8, F_LDG>("ld.global.nc", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<16, F_LDG>("ld.global.nc", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<8, F_CG>("ld.global.cg", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<16, F_CG>("ld.global.cg", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<8, F_CV>("ld.volatile.global", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<8, F_LDGSTS>("cp.async 8 B -> smem", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<8, F_LDGSTS>("cp.async 8 B -> smem, 2 CTAs/SM", 256, 2, 2000000, total, dvec, hvec, reps);
    run_flavour<4, F_BULK>("cp.async.bulk 16 B (TMA)", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<8, F_BULK>("cp.async.bulk 16 B (TMA)", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<8, F_BULK>("cp.async.bulk 16 B (TMA), 2 CTAs/SM", 256, 2, 2000000, total, dvec, hvec, reps);
  }
  if (!strcmp(mode, "mix")) {   // LSU gathers and TMA gathers side by side: are the two paths additive?
    run_flavour<8, F_LDG>("ld.global.nc", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<10, F_MIX2>("8 ld.global.nc + 2 cp.async.bulk", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<12, F_MIX4>("8 ld.global.nc + 4 cp.async.bulk", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<16, F_MIX2>("14 ld.global.nc + 2 cp.async.bulk", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<16, F_MIX4>("12 ld.global.nc + 4 cp.async.bulk", 256, 0, 2000000, total, dvec, hvec, reps);
    run_flavour<10, F_MIX2>("8 + 2, 4 CTAs/SM", 256, 4, 2000000, total, dvec, hvec, reps);
    run_flavour<10, F_MIX2>("8 + 2, 512 thr", 512, 0, 2000000, total, dvec, hvec, reps);
    printf("done\n"); return 0;
  }
  if (!strcmp(mode, "ncu")) {   // the three launches worth a counter capture
    run<8, 0>("L2 only, 1024 thr", 1024, 0, 2000000, total, dvec, hvec, 1);
    run<8, 4>("DSMEM only", 1024, W128, 16 * W128, total, dvec, hvec, 1);
    run<8, 4>("mixed, yhat-sized (1M)", 1024, W200, 1000000, total, dvec, hvec, 1);
    run<8, 4>("L2 only under 200 KB smem", 1024, W200, 2000000, total, dvec, hvec, 1, 0);
    return 0;
  }
  if (strcmp(mode, "all") && strcmp(mode, "dsmem")) { printf("done\n"); return 0; }
  // 1. baseline: everything through L2 (no smem; then with 128 / 200 KB of smem taken away from L1)
  run<8, 0>("L2 only, 256 thr", 256, 0, 2000000, total, dvec, hvec, reps);
  run<8, 0>("L2 only, 1024 thr", 1024, 0, 2000000, total, dvec, hvec, reps);
  run<8, 0>("L2 only N=1M", 1024, 0, 1000000, total, dvec, hvec, reps);
  // 2. local shared memory only (no cluster): the smem port itself
  run<8, 0>("local smem only (CS 1)", 1024, W128, W128, total, dvec, hvec, reps);
  // 3. DSMEM only: the whole index range inside the cluster's panel
  run<8, 1>("DSMEM only", 1024, W128, 2 * W128, total, dvec, hvec, reps);
  run<8, 2>("DSMEM only", 1024, W128, 4 * W128, total, dvec, hvec, reps);
  run<8, 3>("DSMEM only", 1024, W128, 8 * W128, total, dvec, hvec, reps);
  run<8, 4>("DSMEM only", 1024, W128, 16 * W128, total, dvec, hvec, reps);
  run<4, 4>("DSMEM only", 1024, W128, 16 * W128, total, dvec, hvec, reps);
  run<8, 4>("DSMEM only 512 thr", 512, W128, 16 * W128, total, dvec, hvec, reps);
  run<8, 4>("DSMEM only 64 KB", 1024, W64, 16 * W64, total, dvec, hvec, reps);
  run<8, 4>("DSMEM only 200 KB", 1024, W200, 16 * W200, total, dvec, hvec, reps);
  run<8, 3>("DSMEM only 200 KB", 1024, W200, 8 * W200, total, dvec, hvec, reps);
  // 4. mixed: panel of the 2M (row step) and 1M (col step) vectors
  run<8, 4>("mixed, xbar-sized (2M)", 1024, W200, 2000000, total, dvec, hvec, reps);
  run<8, 4>("mixed, yhat-sized (1M)", 1024, W200, 1000000, total, dvec, hvec, reps);
  run<8, 3>("mixed, xbar-sized (2M)", 1024, W200, 2000000, total, dvec, hvec, reps);
  run<8, 3>("mixed, yhat-sized (1M)", 1024, W200, 1000000, total, dvec, hvec, reps);
  run<8, 4>("mixed, 0.5M", 1024, W200, 500000, total, dvec, hvec, reps);
  run<8, 4>("mixed 128 KB, yhat-sized", 1024, W128, 1000000, total, dvec, hvec, reps);
  run<8, 2>("mixed CS4 200 KB, yhat-sized", 1024, W200, 1000000, total, dvec, hvec, reps);
  // 5. what the smem carve-out alone costs the L2 path (panel allocated, never hit)
  run<8, 4>("L2 only under 200 KB smem", 1024, W200, 2000000, total, dvec, hvec, reps, 0);
  run<8, 0>("L2 only under 200 KB smem (CS 1)", 1024, W200, 2000000, total, dvec, hvec, reps, 0);
  printf("done\n");
  return 0;
}
