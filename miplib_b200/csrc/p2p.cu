// p2p.cu - the row-sharded exchange as ONE kernel over NVLink peer memory, replacing
// "partial SpMV -> ncclAllReduce(n doubles) -> replicated primal step".
//
// Rank g owns the column slice [j0[g], j0[g+1]).  After its local partial product K_g' yhat_g sits
// in its own `partial` buffer, k_primal_step_p2p on every rank
//   1. waits until every peer has entered the kernel (their partials are complete, and they are
//      done reading the previous xbar),
//   2. PULLS its slice of all W partial vectors straight out of the peers' HBM (P2P loads, summed
//      in rank order => deterministic), i.e. a reduce-scatter with no intermediate buffer,
//   3. runs the primal step (the same col_epilogue as the single-GPU kernel) on its slice only -
//      the 80 n-byte replicated step of the all-reduce variant becomes 80 n / W,
//   4. PUSHES the new xbar slice into every peer's xbar (P2P stores), i.e. the all-gather,
//   5. leaves through a cross-GPU barrier so the next row step sees a complete xbar.
// Wire traffic per rank is the same 2 (W-1)/W n doubles as an all-reduce, but it is issued tile by
// tile from inside the compute kernel, there is one launch instead of three plus two NCCL kernels,
// and nothing n-proportional is replicated any more.
//
// Peer buffers are mapped with CUDA IPC (one process per GPU); handles travel through one
// ncclAllGather at load time.  All ranks execute the same sequence of barrier kernels, so a
// per-rank epoch counter stays in lock step.
#include <cstring>
#include <vector>

#include "internal.h"
#include "kernels.cuh"

namespace mipcu {

void dist_allgather_bytes(mipcu_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes_per_rank);

namespace {

inline int grid_for(int64_t n, int per_block = kThreads) {
  return (int)((n + per_block - 1) / per_block);
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// flags layout per rank: [0, W) ready, [W, 2W) done, [2W] CTA counter, [2W+1] epoch
__device__ __forceinline__ unsigned long long barrier_epoch(const PeerTable& t) {
  return t.flags[t.rank][2 * t.world + 1] + 1;
}

__device__ __forceinline__ void entry_barrier(const PeerTable& t, unsigned long long epoch) {
  if (blockIdx.x == 0 && threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(t.flags[threadIdx.x] + t.rank, epoch);          // "rank has arrived" on every peer
  }
  if (threadIdx.x < t.world)
    while (ld_acquire_sys(t.flags[t.rank] + threadIdx.x) < epoch) {}
  __syncthreads();
}

__device__ __forceinline__ void exit_barrier(const PeerTable& t, unsigned long long epoch) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                                        // this CTA's pushes are out
    unsigned long long* counter = t.flags[t.rank] + 2 * t.world;
    s_last = (atomicAdd(counter, 1ull) == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // the last CTA of this rank: tell every peer we are done, wait until they all are
  if (threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(t.flags[threadIdx.x] + t.world + t.rank, epoch);
    while (ld_acquire_sys(t.flags[t.rank] + t.world + threadIdx.x) < epoch) {}
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    t.flags[t.rank][2 * t.world] = 0;
    t.flags[t.rank][2 * t.world + 1] = epoch;
  }
}

template <bool CHECK, bool STORE>
__global__ void __launch_bounds__(kThreads)
k_primal_step_p2p(const ColStepArgs a, const PeerTable t) {
  const Scalars* sc = a.sc;
  // NOTE: no early exit on sc->done: the barriers must be executed by every rank, and done is
  // identical on all ranks, so finished solves simply skip the arithmetic below.
  const bool live = !sc->done;
  const unsigned long long epoch = barrier_epoch(t);
  entry_barrier(t, epoch);
  const int j = t.j0[t.rank] + blockIdx.x * kRowsPerCta + threadIdx.x;
  double red[10] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (live && j < t.j0[t.rank + 1]) {
    double atyh = 0.0;
    for (int p = 0; p < t.world; ++p) atyh += __ldcv(t.partial[p] + j);     // reduce-scatter by pull
    const double tau = sc->tau;
    const double k = (double)(sc->k_base + a.iter_in_block);
    const double w_new = (k + 1.0) / (k + 2.0);
    col_epilogue<CHECK, STORE>(a, j, atyh, tau, w_new, red);
    const double xb = a.xbar[j];
    for (int p = 0; p < t.world; ++p)
      if (p != t.rank) t.xbar[p][j] = xb;                                     // all-gather by push
  }
  if (CHECK && live) {
    const int q[10] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL,
                       Q_RAYD_COL, Q_RAYD_COLV, Q_RAYP_C, Q_RAYP_COLV};
    cta_reduce_store<10>(red, q, a.partials, sc->max_blocks);
  }
  exit_barrier(t, epoch);
}

// all-gather of one replicated vector: every rank pushes its slice into all peers
__global__ void __launch_bounds__(kThreads) k_push_slice(const PeerTable t, int which) {
  const unsigned long long epoch = barrier_epoch(t);
  entry_barrier(t, epoch);
  const int j = t.j0[t.rank] + blockIdx.x * kThreads + threadIdx.x;
  if (j < t.j0[t.rank + 1]) {
    double* const* dst = which == 0 ? t.xbar : which == 1 ? t.xhat_chk : t.x_sol;
    const double v = dst[t.rank][j];
    for (int p = 0; p < t.world; ++p)
      if (p != t.rank) dst[p][j] = v;
  }
  exit_barrier(t, epoch);
}

}  // namespace

bool p2p_active(const mipcu_ctx* ctx) { return ctx->world > 1 && ctx->peers.world == ctx->world; }

void p2p_teardown(mipcu_ctx* ctx) {
  for (void* p : ctx->peer_mappings) cudaIpcCloseMemHandle(p);
  ctx->peer_mappings.clear();
  if (ctx->p2p_flags) cudaFree(ctx->p2p_flags);
  ctx->p2p_flags = nullptr;
  ctx->peers = PeerTable();
}

// Collective (every rank calls it from mipcu_load_lp): exchange IPC handles of the four shared
// vectors and the flag block, map the peers' copies.  On any failure the context silently stays on
// the NCCL all-reduce path.
void p2p_setup(mipcu_ctx* ctx) {
  p2p_teardown(ctx);
  if (ctx->world <= 1 || ctx->world > kMaxPeers || (int)ctx->params[MIPCU_PARAM_SHARD_EXCHANGE] == 0) return;
  const int W = ctx->world, n = ctx->n;
  cudaStream_t st = ctx->stream;
  MIPCU_CK(cudaMalloc(&ctx->p2p_flags, sizeof(unsigned long long) * (2 * kMaxPeers + 2)));
  MIPCU_CK(cudaMemsetAsync(ctx->p2p_flags, 0, sizeof(unsigned long long) * (2 * kMaxPeers + 2), st));
  constexpr int kBufs = 5;
  void* mine[kBufs] = {ctx->tmp_n, ctx->xbar, ctx->xhat_chk, ctx->x_sol, ctx->p2p_flags};
  std::vector<cudaIpcMemHandle_t> send(kBufs), recv((size_t)kBufs * W);
  int ok = 1;
  for (int i = 0; i < kBufs; ++i)
    if (cudaIpcGetMemHandle(&send[i], mine[i]) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  // handles (and whether this rank could make them) travel in one all-gather
  const size_t bytes = sizeof(cudaIpcMemHandle_t) * kBufs + 8;
  std::vector<unsigned char> hsend(bytes), hrecv(bytes * W);
  std::memcpy(hsend.data(), send.data(), sizeof(cudaIpcMemHandle_t) * kBufs);
  std::memcpy(hsend.data() + sizeof(cudaIpcMemHandle_t) * kBufs, &ok, sizeof(int));
  unsigned char *dsend = nullptr, *drecv = nullptr;
  MIPCU_CK(cudaMalloc(&dsend, bytes));
  MIPCU_CK(cudaMalloc(&drecv, bytes * W));
  MIPCU_CK(cudaMemcpyAsync(dsend, hsend.data(), bytes, cudaMemcpyHostToDevice, st));
  dist_allgather_bytes(ctx, dsend, drecv, bytes);
  MIPCU_CK(cudaMemcpyAsync(hrecv.data(), drecv, bytes * W, cudaMemcpyDeviceToHost, st));
  MIPCU_CK(cudaStreamSynchronize(st));
  cudaFree(dsend);
  cudaFree(drecv);
  PeerTable t;
  t.world = W;
  t.rank = ctx->rank;
  for (int g = 0; g <= W; ++g) t.j0[g] = (int)(((int64_t)n * g) / W);
  bool all_ok = true;
  for (int p = 0; p < W && all_ok; ++p) {
    int peer_ok = 0;
    std::memcpy(&peer_ok, hrecv.data() + bytes * p + sizeof(cudaIpcMemHandle_t) * kBufs, sizeof(int));
    all_ok = all_ok && peer_ok;
  }
  for (int p = 0; p < W && all_ok; ++p) {
    void* ptr[kBufs];
    if (p == ctx->rank) {
      for (int i = 0; i < kBufs; ++i) ptr[i] = mine[i];
    } else {
      for (int i = 0; i < kBufs && all_ok; ++i) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hrecv.data() + bytes * p + sizeof(cudaIpcMemHandle_t) * i, sizeof(h));
        if (cudaIpcOpenMemHandle(&ptr[i], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          all_ok = false;
        } else {
          ctx->peer_mappings.push_back(ptr[i]);
        }
      }
    }
    if (!all_ok) break;
    t.partial[p] = (double*)ptr[0];
    t.xbar[p] = (double*)ptr[1];
    t.xhat_chk[p] = (double*)ptr[2];
    t.x_sol[p] = (double*)ptr[3];
    t.flags[p] = (unsigned long long*)ptr[4];
  }
  // every rank must take the same path: agree on success
  const double agreed = dist_sum_scalar(ctx, all_ok ? 1.0 : 0.0);
  if (agreed < W - 0.5) {
    for (void* p : ctx->peer_mappings) cudaIpcCloseMemHandle(p);
    ctx->peer_mappings.clear();
    ctx->params[MIPCU_PARAM_SHARD_EXCHANGE] = 0;   // visible through mipcu_get_param: NCCL path in use
    return;
  }
  ctx->peers = t;
}

void p2p_primal_step(mipcu_ctx* ctx, const ColStepArgs& a, bool check, bool store) {
  const PeerTable& t = ctx->peers;
  const int g = std::max(1, grid_for(t.j0[t.rank + 1] - t.j0[t.rank], kRowsPerCta));
  cudaStream_t st = ctx->stream;
  if (check && store) k_primal_step_p2p<true, true><<<g, kThreads, 0, st>>>(a, t);
  else if (check) k_primal_step_p2p<true, false><<<g, kThreads, 0, st>>>(a, t);
  else if (store) k_primal_step_p2p<false, true><<<g, kThreads, 0, st>>>(a, t);
  else k_primal_step_p2p<false, false><<<g, kThreads, 0, st>>>(a, t);
  ctx->launches++;
}

void p2p_allgather(mipcu_ctx* ctx, int which) {
  const PeerTable& t = ctx->peers;
  const int g = std::max(1, grid_for(t.j0[t.rank + 1] - t.j0[t.rank]));
  k_push_slice<<<g, kThreads, 0, ctx->stream>>>(t, which);
  ctx->launches++;
}

}  // namespace mipcu
