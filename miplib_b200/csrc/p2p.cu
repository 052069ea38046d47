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
#include <cstdlib>
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

// ---- block-owner exchange (MIPCU_PARAM_SHARD_EXCHANGE = 2) --------------------------------------
// Rank g owns row block g of K (whole rows) AND column block g of K' (whole columns, cut out of all
// ranks' blocks at load: setup.cu:build_col_block).  Both kernels of an iteration then stream 1/W of
// the nonzeros, nothing n-proportional is written besides the iterate itself, and the exchange is two
// all-gathers issued from the kernels' epilogues straight into peer memory:
//   k_row_step_own : K_g xbar (gathers the whole xbar), dual step on the rank's rows, yhat_i pushed
//                    into every rank's yhat_full (m_glob doubles per rank);
//   k_col_step_own : (K' yhat)_j for the rank's columns (gathers the whole yhat_full), primal step,
//                    xbar_j pushed into every rank's xbar.
// One cross-GPU barrier per kernel, at its entry: "every rank has finished the previous kernel"
// covers both hazards (the vector about to be gathered is complete; the vector about to be pushed
// is no longer being read).  Arrival flags carry a generation number the host counts (all ranks
// launch the same sequence).  The pushes of kernel t-1 are complete when the grid has retired (stream
// order), the arrival flag of kernel t is a system-scope release store issued after that, and the
// gathers of kernel t are issued after the acquire that saw the flag.  An extra system fence per
// CTA after its pushes is NOT needed and costs 8 us (row) / 21 us (column) per launch on config 3
// at two GPUs (profiles/r01_block_owner_exchange.md); it stays available as a laboratory switch.
// Laboratory switches (environment variable MIPCU_OWN_VARIANT, read at load; scripts/own_lab.py):
// what each piece of the protocol costs.  OWN_NO_PUBLISH alone (1) is the shipped protocol.
enum { OWN_NO_PUBLISH = 1, OWN_RELAXED_POLL = 2, OWN_NO_PUSH = 4, OWN_NO_BARRIER = 8 };

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Entry barrier of the block-owner kernels.  Returns false when the kernel must not run: a peer did
// not arrive within PeerTable::barrier_timeout_ns (MIPCU_PARAM_PEER_TIMEOUT), now or in an earlier
// kernel of this solve.  The expiry is recorded in the control block (Scalars::fault) - no trap, the
// context stays usable - and every later kernel of the queue falls through at once; the host turns it
// into rc = 1 + mipcu_last_error (solve.cu).
__device__ __forceinline__ bool own_barrier(const PeerTable& t, const unsigned long long epoch, const int variant,
                                            const Scalars* sc) {
  if (variant & OWN_NO_BARRIER) return true;
  int* fault = const_cast<int*>(&sc->fault);
  if (threadIdx.x < t.world) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(t.flags[threadIdx.x] + kOwnFlagBase + t.rank, epoch);
    }
    const unsigned long long* f = t.flags[t.rank] + kOwnFlagBase + threadIdx.x;
    const bool relaxed = variant & OWN_RELAXED_POLL;
    unsigned long long deadline = 0;
    int spins = 0;
    while ((relaxed ? ld_relaxed_sys(f) : ld_acquire_sys(f)) < epoch) {
      if ((++spins & 1023) != 0) continue;
      if (*(volatile int*)fault) break;                       // an earlier kernel already gave up
      if (t.barrier_timeout_ns == 0) continue;
      const unsigned long long now = global_timer_ns();
      if (deadline == 0) deadline = now + t.barrier_timeout_ns;
      else if (now > deadline) { atomicExch(fault, 1); break; }
    }
    if (relaxed) __threadfence_system();
  }
  __syncthreads();
  return *(volatile int*)fault == 0;
}

__device__ __forceinline__ void own_publish(const int variant) {
  if (variant & OWN_NO_PUBLISH) return;
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
}

// A CTA owns the blocks blockIdx.x, blockIdx.x + gridDim.x, ... (nblk in all): with more than one
// block per CTA the pushes of block k travel while block k + 1 is multiplied.
template <int TPR, bool CHECK, bool MULTI>
__global__ void __launch_bounds__(kThreads)
k_row_step_own(const RowStepArgs a, const PeerTable t, const unsigned long long epoch, const int variant,
               const int nblk) {
  __shared__ double s_sum[kRowsPerCta];
  const Scalars* sc = a.sc;
  pdl_trigger();
  double sigma = 0.0, w_new = 0.0;
  for (int blk = blockIdx.x; MULTI ? blk < nblk : blk == (int)blockIdx.x; blk += MULTI ? gridDim.x : 1) {
    // first block: the head of the nonzero stream is requested before the wait for the peers;
    // `done` is identical on every rank and is looked at only after the barrier has been passed
    const bool first = !MULTI || blk == (int)blockIdx.x;
    if (!cta_row_sums<TPR>(a.ptr, a.idx, a.val, a.perm, a.xbar, a.m, s_sum, [&] {
          if (!first) return true;
          pdl_wait();
          return own_barrier(t, epoch, variant, sc) && !sc->done;
        }, blk))
      return;
    if (first) {
      sigma = sc->sigma;
      const double k = (double)(sc->k_base + a.iter_in_block);
      w_new = (k + 1.0) / (k + 2.0);
    }
    const int i = blk * kRowsPerCta + threadIdx.x;
    double red[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < a.m) {
      const double yh = row_epilogue<CHECK>(a, i, s_sum[threadIdx.x], sigma, w_new, red);
      const int gi = t.i0[t.rank] + i;
      if (!(variant & OWN_NO_PUSH)) {
        if (t.mc_yhat_full) t.mc_yhat_full[gi] = yh;                      // one store, replicated by the switch
        else for (int p = 0; p < t.world; ++p) t.yhat_full[p][gi] = yh;   // all-gather by push (self too)
      }
    }
    if (CHECK) {
      const int q[5] = {Q_DY2, Q_DY0, Q_DOBJ_ROW, Q_RAYD_ROW, Q_RAYD_ROWV};
      cta_reduce_store<5>(red, q, a.partials, sc->max_blocks, blk);
    }
    if (MULTI) __syncthreads();      // s_sum (and the reduction scratch) are reused by the next block
  }
  own_publish(variant);
}

// `a` addresses the rank's column slice: every per-column pointer is already offset by j0, a.n is
// the slice length, a.yhat is this rank's whole yhat_full.  push_xhat: the stored xhat is the KKT
// candidate, which the row-side KKT pass of every rank gathers.
template <int TPR, bool CHECK, bool STORE, bool MULTI>
__global__ void __launch_bounds__(kThreads)
k_col_step_own(const ColStepArgs a, const PeerTable t, const unsigned long long epoch, const int push_xhat,
               const int variant, const int nblk) {
  __shared__ double s_sum[kRowsPerCta];
  const Scalars* sc = a.sc;
  pdl_trigger();
  double tau = 0.0, w_new = 0.0;
  for (int blk = blockIdx.x; MULTI ? blk < nblk : blk == (int)blockIdx.x; blk += MULTI ? gridDim.x : 1) {
    const bool first = !MULTI || blk == (int)blockIdx.x;
    if (!cta_row_sums<TPR>(a.ptr, a.idx, a.val, a.perm, a.yhat, a.n, s_sum, [&] {
          if (!first) return true;
          pdl_wait();
          return own_barrier(t, epoch, variant, sc) && !sc->done;
        }, blk))
      return;
    if (first) {
      tau = sc->tau;
      const double k = (double)(sc->k_base + a.iter_in_block);
      w_new = (k + 1.0) / (k + 2.0);
    }
    const int j = blk * kRowsPerCta + threadIdx.x;
    double red[10] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (j < a.n) {
      col_epilogue<CHECK, STORE>(a, j, s_sum[threadIdx.x], tau, w_new, red);
      const int gj = t.j0[t.rank] + j;
      const double xb = a.xbar[j];
      if (!(variant & OWN_NO_PUSH)) {
        if (t.mc_xbar) t.mc_xbar[gj] = xb;
        else
          for (int p = 0; p < t.world; ++p)
            if (p != t.rank) t.xbar[p][gj] = xb;
      }
      if (STORE && push_xhat) {
        const double xh = a.xhat_out[j];
        for (int p = 0; p < t.world; ++p)
          if (p != t.rank) t.xhat_chk[p][gj] = xh;
      }
    }
    if (CHECK) {
      const int q[10] = {Q_DX2, Q_DXDATY, Q_DX0, Q_DRES2, Q_POBJ, Q_DOBJ_COL,
                         Q_RAYD_COL, Q_RAYD_COLV, Q_RAYP_C, Q_RAYP_COLV};
      cta_reduce_store<10>(red, q, a.partials, sc->max_blocks, blk);
    }
    if (MULTI) __syncthreads();
  }
  own_publish(variant);
}

// Blocks per CTA.  One while a grid of one block per CTA is longer than about a wave (the waves
// already stagger the pushes over the kernel).  Around one wave or less - the 4 and 8 GPU shards of
// config 3 - every CTA would finish, and push, at the same moment at the end of the kernel; a CTA
// that works through about 14k nonzeros in 2 - 4 blocks pushes block k while it multiplies block
// k + 1.  Measured on config 3 (profiles/r01_block_owner_exchange.md): 8 GPUs, row step (488
// blocks of 8k nonzeros) 36.1 / 35.0 / 44.5 us and column step (977 blocks of 4k) 45.0 / 40.4 /
// 38.1 us for 1 / 2 / 3 blocks per CTA; 4 GPUs, row step (977 blocks) 57.3 / 51.6 / 56.3 us.
// MIPCU_OWN_BLOCKS overrides (laboratory).
int own_grid(const mipcu_ctx* ctx, int nblk, int64_t nnz) {
  int per_cta = ctx->own_blocks_per_cta;
  if (per_cta <= 0) {
    per_cta = 1;
    const int slots = 6 * ctx->sm_count;                 // resident CTAs of these kernels (40 registers)
    if (nblk > 1 && nblk <= slots + slots / 4) {
      const double per_block = (double)nnz / nblk;
      per_cta = (int)(14000.0 / std::max(per_block, 1.0) + 0.5);
      per_cta = std::max(1, std::min(per_cta, 4));
      while (per_cta > 1 && (nblk + per_cta - 1) / per_cta < ctx->sm_count + ctx->sm_count / 2) --per_cta;
    }
  }
  return std::max(1, (nblk + per_cta - 1) / per_cta);
}

template <int TPR>
void launch_row_own(mipcu_ctx* ctx, const RowStepArgs& a, bool check) {
  const int nblk = std::max(1, grid_for(ctx->m, kRowsPerCta)), g = own_grid(ctx, nblk, ctx->rows.nnz);
  const unsigned long long e = ++ctx->own_epoch;
  const bool pdl = ctx->use_pdl;
  const int v = ctx->own_variant;
  cudaStream_t st = ctx->stream;
  const PeerTable& t = ctx->peers;
  if (g < nblk) {
    if (check) launch_kernel(k_row_step_own<TPR, true, true>, g, st, pdl, a, t, e, v, nblk);
    else launch_kernel(k_row_step_own<TPR, false, true>, g, st, pdl, a, t, e, v, nblk);
  } else {
    if (check) launch_kernel(k_row_step_own<TPR, true, false>, g, st, pdl, a, t, e, v, nblk);
    else launch_kernel(k_row_step_own<TPR, false, false>, g, st, pdl, a, t, e, v, nblk);
  }
}

template <int TPR>
void launch_col_own(mipcu_ctx* ctx, const ColStepArgs& a, bool check, bool store, int push_xhat) {
  const int nblk = std::max(1, grid_for(a.n, kRowsPerCta)), g = own_grid(ctx, nblk, ctx->colblk.nnz);
  const unsigned long long e = ++ctx->own_epoch;
  cudaStream_t st = ctx->stream;
  const PeerTable& t = ctx->peers;
  const bool pdl = ctx->use_pdl;
  const int v = ctx->own_variant;
  if (g < nblk) {
    if (check && store) launch_kernel(k_col_step_own<TPR, true, true, true>, g, st, pdl, a, t, e, push_xhat, v, nblk);
    else if (check) launch_kernel(k_col_step_own<TPR, true, false, true>, g, st, pdl, a, t, e, push_xhat, v, nblk);
    else if (store) launch_kernel(k_col_step_own<TPR, false, true, true>, g, st, pdl, a, t, e, push_xhat, v, nblk);
    else launch_kernel(k_col_step_own<TPR, false, false, true>, g, st, pdl, a, t, e, push_xhat, v, nblk);
  } else {
    if (check && store) launch_kernel(k_col_step_own<TPR, true, true, false>, g, st, pdl, a, t, e, push_xhat, v, nblk);
    else if (check) launch_kernel(k_col_step_own<TPR, true, false, false>, g, st, pdl, a, t, e, push_xhat, v, nblk);
    else if (store) launch_kernel(k_col_step_own<TPR, false, true, false>, g, st, pdl, a, t, e, push_xhat, v, nblk);
    else launch_kernel(k_col_step_own<TPR, false, false, false>, g, st, pdl, a, t, e, push_xhat, v, nblk);
  }
}

// all-gather of one replicated vector: every rank pushes its slice into all peers
__global__ void __launch_bounds__(kThreads) k_push_slice(const PeerTable t, int which) {
  const unsigned long long epoch = barrier_epoch(t);
  entry_barrier(t, epoch);
  const int j = t.j0[t.rank] + blockIdx.x * kThreads + threadIdx.x;
  if (j < t.j0[t.rank + 1]) {
    double* const* dst = which == 0 ? t.xbar : which == 1 ? t.xhat_chk : t.x_sol;
    const double v = dst[t.rank][j];
    if (which == 0 && t.mc_xbar) t.mc_xbar[j] = v;
    else
      for (int p = 0; p < t.world; ++p)
        if (p != t.rank) dst[p][j] = v;
  }
  exit_barrier(t, epoch);
}

}  // namespace

bool p2p_active(const mipcu_ctx* ctx) { return ctx->world > 1 && ctx->peers.world == ctx->world; }
bool own_active(const mipcu_ctx* ctx) { return p2p_active(ctx) && ctx->own_ready; }

// Collective (from mipcu_load_lp, after p2p_setup).  The block-owner exchange needs the peer table
// of the fused exchange plus this rank's column block of the whole matrix.
void own_setup(mipcu_ctx* ctx) {
  ctx->own_ready = false;
  ctx->own_epoch = 0;
  if (!p2p_active(ctx) || (int)ctx->params[MIPCU_PARAM_SHARD_EXCHANGE] != 2) return;
  build_col_block(ctx);
  // optional: pushes through the NVSwitch multicast window when the box has one; the parameter
  // reads back whether it is in use
  ctx->params[MIPCU_PARAM_MULTICAST] = mc_setup(ctx) ? 1.0 : 0.0;
  const char* v = std::getenv("MIPCU_OWN_VARIANT");
  ctx->own_variant = v ? std::atoi(v) : OWN_NO_PUBLISH;
  const char* b = std::getenv("MIPCU_OWN_BLOCKS");
  ctx->own_blocks_per_cta = b ? std::atoi(b) : 0;     // 0: chosen per kernel (own_grid)
  ctx->peers.barrier_timeout_ns = (unsigned long long)(std::max(0.0, ctx->params[MIPCU_PARAM_PEER_TIMEOUT]) * 1e9);
  ctx->own_ready = true;
}

// Re-load of a model of the same shape (mipcu_load_lp kept the exported vectors and every mapping):
// only the column block depends on the matrix.
void own_rebuild(mipcu_ctx* ctx) {
  if (!ctx->own_ready) return;
  build_col_block(ctx);
}

#define MIPCU_OWN_DISPATCH(fmt, CALL)                                                          \
  switch (fmt) {                                                                               \
    case 0: { constexpr int T = 0; CALL; } break;                                              \
    case 2: { constexpr int T = 2; CALL; } break;                                              \
    case 4: { constexpr int T = 4; CALL; } break;                                              \
    case 8: { constexpr int T = 8; CALL; } break;                                              \
    case 16: { constexpr int T = 16; CALL; } break;                                            \
    default: { constexpr int T = 32; CALL; } break;                                            \
  }

void own_row_step(mipcu_ctx* ctx, const RowStepArgs& a, bool check) {
  MIPCU_OWN_DISPATCH(kernel_format(ctx, ctx->rows), launch_row_own<T>(ctx, a, check));
  ctx->launches++;
}

void own_col_step(mipcu_ctx* ctx, int iter_in_block, bool check, const double* xhat_in, double* xhat_out) {
  const PeerTable& t = ctx->peers;
  const int j0 = t.j0[t.rank];
  const Csr& B = ctx->colblk;
  ColStepArgs a;
  if (kernel_format(ctx, B) == 0) { a.ptr = B.sptr; a.idx = B.sidx; a.val = B.sval; a.perm = B.perm; }
  else { a.ptr = B.ptr; a.idx = B.idx; a.val = B.val; a.perm = nullptr; }
  a.yhat = ctx->yhat_full;
  a.x0 = ctx->x0 + j0; a.aty0 = ctx->aty0 + j0; a.c = ctx->cs + j0; a.l = ctx->ls + j0; a.u = ctx->us + j0;
  a.d2 = ctx->d2 + j0; a.xbar = ctx->xbar + j0; a.aty = ctx->aty + j0;
  a.xhat_in = xhat_in ? xhat_in + j0 : nullptr;
  a.xhat_out = xhat_out ? xhat_out + j0 : nullptr;
  a.aty_cand = ctx->aty_cand + j0;
  a.sc = ctx->d_sc; a.partials = ctx->partials;
  a.n = t.j0[t.rank + 1] - j0; a.iter_in_block = iter_in_block;
  const int push_xhat = xhat_out == ctx->xhat_chk ? 1 : 0;
  MIPCU_OWN_DISPATCH(kernel_format(ctx, B), launch_col_own<T>(ctx, a, check, xhat_out != nullptr, push_xhat));
  ctx->launches++;
}

void p2p_teardown(mipcu_ctx* ctx) {
  mc_teardown(ctx);
  for (void* p : ctx->peer_mappings) cudaIpcCloseMemHandle(p);
  ctx->peer_mappings.clear();
  if (ctx->p2p_flags) cudaFree(ctx->p2p_flags);
  ctx->p2p_flags = nullptr;
  ctx->peers = PeerTable();
  ctx->own_ready = false;
}

// Collective (every rank calls it from mipcu_load_lp): exchange IPC handles of the four shared
// vectors and the flag block, map the peers' copies.  On any failure the context silently stays on
// the NCCL all-reduce path.
void p2p_setup(mipcu_ctx* ctx) {
  p2p_teardown(ctx);
  if (ctx->world <= 1 || ctx->world > kMaxPeers || (int)ctx->params[MIPCU_PARAM_SHARD_EXCHANGE] == 0) return;
  const int W = ctx->world, n = ctx->n;
  cudaStream_t st = ctx->stream;
  MIPCU_CK(cudaMalloc(&ctx->p2p_flags, sizeof(unsigned long long) * kFlagCount));
  MIPCU_CK(cudaMemsetAsync(ctx->p2p_flags, 0, sizeof(unsigned long long) * kFlagCount, st));
  constexpr int kBufs = 6;
  void* mine[kBufs] = {ctx->tmp_n, ctx->xbar, ctx->xhat_chk, ctx->x_sol, ctx->p2p_flags, ctx->yhat_full};
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
  for (int g = 0; g <= W; ++g) t.i0[g] = ctx->row_off[g];
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
    t.yhat_full[p] = (double*)ptr[5];
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
