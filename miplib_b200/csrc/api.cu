// api.cu - the extern "C" surface declared in include/mipcu.h.  Every entry point converts
// C++ exceptions into rc = 1 + a message retrievable with mipcu_last_error(), the convention of
// the reference's own C API (reference miplib/capi.cpp:7-28).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "internal.h"

using namespace mipcu;

namespace mipcu {
// dist.cu
void dist_unique_id(void* out);
void dist_init(mipcu_ctx* ctx, int rank, int world, const void* unique_id);
void dist_destroy(mipcu_ctx* ctx);
// batch.cu
void solve_lp_batch(mipcu_ctx* ctx, int32_t B, const double* lb, const double* ub, const double* cutoff,
                    const double* x0, const double* y0, const double* weight0, const double* branch_tol,
                    mipcu_stats* out, double* x_out, double* y_out);
}  // namespace mipcu

thread_local const mipcu_ctx* g_alloc_ctx = nullptr;

namespace {

std::string g_create_error;   // errors raised before a context exists
std::mutex g_mutex;

struct AllocScope {
  const mipcu_ctx* saved;
  explicit AllocScope(const mipcu_ctx* ctx) : saved(g_alloc_ctx) { g_alloc_ctx = ctx; }
  ~AllocScope() { g_alloc_ctx = saved; }
};

template <class F>
int guarded(mipcu_ctx* ctx, F&& f) {
  AllocScope scope(ctx);
  try {
    f();
    return 0;
  } catch (std::exception const& e) {
    if (ctx) ctx->err = e.what();
    else {
      std::lock_guard<std::mutex> lk(g_mutex);
      g_create_error = e.what();
    }
    cudaGetLastError();   // clear a sticky-less error so the next call starts clean
    return 1;
  }
}

void set_default_params(mipcu_ctx* ctx) {
  ctx->params[MIPCU_PARAM_EPS_REL] = 1e-6;
  ctx->params[MIPCU_PARAM_ITER_LIMIT] = 2e6;
  ctx->params[MIPCU_PARAM_TIME_LIMIT] = 0.0;
  ctx->params[MIPCU_PARAM_CHECK_PERIOD] = 64;
  ctx->params[MIPCU_PARAM_VERBOSE] = 0;
  ctx->params[MIPCU_PARAM_STEP_SIZE] = 0.0;
  ctx->params[MIPCU_PARAM_USE_GRAPH] = -1;
  ctx->params[MIPCU_PARAM_SCALING] = 1;
  ctx->params[MIPCU_PARAM_THREADS_PER_ROW] = 0;
  ctx->params[MIPCU_PARAM_OBJ_CUTOFF] = NAN;
  ctx->params[MIPCU_PARAM_USE_TMA] = -1;
  ctx->params[MIPCU_PARAM_SHARD_EXCHANGE] = 2;
  ctx->params[MIPCU_PARAM_USE_PDL] = 0;
  ctx->params[MIPCU_PARAM_MULTICAST] = 1;
  ctx->params[MIPCU_PARAM_ROW_SORT] = 4;
  ctx->params[MIPCU_PARAM_PEER_TIMEOUT] = 120.0;
  ctx->params[MIPCU_PARAM_VALIDATE] = 0;
}

mipcu_ctx* create_on(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    throw Error("no CUDA device available: the Cuda backend has no CPU fallback");
  }
  MIPCU_REQUIRE(device >= 0 && device < count, "mipcu_create: device index out of range");
  cudaDeviceProp prop;
  MIPCU_CK(cudaGetDeviceProperties(&prop, device));
  MIPCU_REQUIRE(prop.major >= 10, std::string("mipcu_create: device '") + prop.name +
                                      "' is not sm_100-class; this library ships sm_100a code only");
  MIPCU_CK(cudaSetDevice(device));
  mipcu_ctx* ctx = new mipcu_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  set_default_params(ctx);
  try {
    MIPCU_CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    // model arrays come stream-ordered from a memory pool PRIVATE to this context that keeps what is
    // freed (see setup.cu: dalloc); the device's default pool - which a host application such as
    // PyTorch's cudaMallocAsync allocator may share - is left alone
    int pools = 0;
    if (cudaDeviceGetAttribute(&pools, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && pools) {
      cudaMemPoolProps props;
      std::memset(&props, 0, sizeof(props));
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      cudaMemPool_t pool = nullptr;
      if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess) {
          ctx->pool = pool;
          ctx->pool_alloc = true;
        } else {
          cudaMemPoolDestroy(pool);
        }
      }
    }
    cudaGetLastError();
  } catch (...) {
    delete ctx;
    throw;
  }
  return ctx;
}

}  // namespace

extern "C" {

int mipcu_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int usable = 0;
  for (int d = 0; d < count; ++d) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major >= 10) ++usable;
  }
  return usable;
}

int mipcu_create(mipcu_ctx** out, int device) {
  if (!out) return 1;
  *out = nullptr;
  return guarded(nullptr, [&] { *out = create_on(device); });
}

int mipcu_dist_unique_id(void* id_out) {
  return guarded(nullptr, [&] {
    MIPCU_REQUIRE(id_out, "mipcu_dist_unique_id: null output");
    dist_unique_id(id_out);
  });
}

int mipcu_create_sharded(mipcu_ctx** out, int device, int rank, int world, const void* unique_id) {
  if (!out) return 1;
  *out = nullptr;
  return guarded(nullptr, [&] {
    MIPCU_REQUIRE(world >= 1 && rank >= 0 && rank < world, "mipcu_create_sharded: bad rank/world");
    mipcu_ctx* ctx = create_on(device);
    try {
      if (world > 1) dist_init(ctx, rank, world, unique_id);
    } catch (...) {
      mipcu_destroy(ctx);
      throw;
    }
    *out = ctx;
  });
}

void mipcu_destroy(mipcu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  AllocScope scope(ctx);
  try { p2p_teardown(ctx); } catch (...) {}
  try { free_model(ctx); } catch (...) {}
  if (ctx->nccl_comm) { try { dist_destroy(ctx); } catch (...) {} }
  if (ctx->d_sc) cudaFree(ctx->d_sc);
  if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
  if (ctx->flush_buf) cudaFree(ctx->flush_buf);
  for (auto& e : ctx->ev)
    if (e) cudaEventDestroy(e);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
    if (ctx->pool) cudaMemPoolDestroy((cudaMemPool_t)ctx->pool);   // hands the cached memory back to the device
  }
  cudaGetLastError();
  delete ctx;
}

int mipcu_reset(mipcu_ctx* ctx) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(ctx->world <= 1, "mipcu_reset: not supported on row-sharded contexts");
    MIPCU_CK(cudaSetDevice(ctx->device));
    MIPCU_CK(cudaStreamSynchronize(ctx->stream));
    free_model(ctx);
    set_default_params(ctx);
    ctx->err.clear();
  });
}

const char* mipcu_last_error(const mipcu_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  // a copy private to the calling thread: the pointer stays valid after the lock is released
  thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_mutex);
  copy = g_create_error;
  return copy.c_str();
}

int mipcu_load_lp(mipcu_ctx* ctx, int64_t m, int64_t n, int64_t nnz, const int64_t* row_ptr,
                  const int32_t* col_idx, const double* val, const double* c, const double* lb,
                  const double* ub, const double* row_lo, const double* row_hi, int maximize) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    ctx->launches = 0;
    // Row-sharded contexts: a model of the same shape as the one before (every rank: same n, same
    // row count, same exchange parameters) keeps the exported vectors, the CUDA-IPC mappings and the
    // multicast window - re-creating them is most of the cost of a re-load.  All ranks must agree.
    bool keep = false;
    if (ctx->world > 1) {
      MIPCU_CK(cudaSetDevice(ctx->device));
      const bool same = ctx->loaded && p2p_active(ctx) && n == ctx->n && m == ctx->m &&
                        ctx->params[MIPCU_PARAM_SHARD_EXCHANGE] == ctx->exchange_at_setup &&
                        ctx->params[MIPCU_PARAM_MULTICAST] == ctx->multicast_at_setup;
      keep = dist_sum_scalar(ctx, same ? 1.0 : 0.0) > ctx->world - 0.5;
    }
    const bool timing = std::getenv("MIPCU_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (timing)
        std::fprintf(stderr, "[mipcu_load_lp rank %d] %s %.1f ms\n", ctx->rank, what,
                     1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    };
    if (!keep) p2p_teardown(ctx);
    ctx->keep_exchange = keep;
    lap(keep ? "exchange mappings kept" : "exchange torn down");
    try {
      load_lp(ctx, m, n, nnz, row_ptr, col_idx, val, c, lb, ub, row_lo, row_hi, maximize);
    } catch (...) {
      ctx->keep_exchange = false;
      throw;
    }
    ctx->keep_exchange = false;
    lap("load_lp");
    if (keep) {
      own_rebuild(ctx);   // only the column block of the block-owner exchange depends on the matrix
      lap("column block rebuilt");
    } else {
      p2p_setup(ctx);   // collective on sharded contexts: maps the peers' exchange buffers
      own_setup(ctx);   // block-owner exchange (MIPCU_PARAM_SHARD_EXCHANGE = 2): column block of the whole matrix
      ctx->exchange_at_setup = ctx->params[MIPCU_PARAM_SHARD_EXCHANGE];
      ctx->multicast_at_setup = ctx->params[MIPCU_PARAM_MULTICAST];
    }
  });
}

int mipcu_set_bounds(mipcu_ctx* ctx, int64_t n_changed, const int64_t* cols, const double* lb,
                     const double* ub) {
  if (!ctx) return 1;
  return guarded(ctx, [&] { set_bounds(ctx, n_changed, cols, lb, ub); });
}

int mipcu_set_objective(mipcu_ctx* ctx, const double* c, int maximize) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(ctx->loaded, "mipcu_set_objective: no model loaded");
    MIPCU_REQUIRE(ctx->n == 0 || c, "mipcu_set_objective: null array");
    MIPCU_CK(cudaSetDevice(ctx->device));
    if (ctx->n && ctx->col_perm) {       // the objective follows its columns (binned at load)
      std::vector<double> staged(c, c + ctx->n);
      cols_to_internal_order(ctx, staged.data());
      MIPCU_CK(cudaMemcpy(ctx->c, staged.data(), sizeof(double) * ctx->n, cudaMemcpyHostToDevice));
    } else if (ctx->n) {
      MIPCU_CK(cudaMemcpyAsync(ctx->c, c, sizeof(double) * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    }
    MIPCU_CK(cudaStreamSynchronize(ctx->stream));
    ctx->maximize = maximize != 0;
    ctx->vectors_dirty = true;
  });
}

int mipcu_set_param(mipcu_ctx* ctx, int key, double value) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(key >= 0 && key < MIPCU_PARAM_COUNT_, "mipcu_set_param: unknown key");
    if (key == MIPCU_PARAM_CHECK_PERIOD) MIPCU_REQUIRE(value >= 2 && value <= 65536, "check period must be in [2, 65536]");
    if (key == MIPCU_PARAM_EPS_REL) MIPCU_REQUIRE(value > 0, "eps_rel must be positive");
    if (key == MIPCU_PARAM_ITER_LIMIT) MIPCU_REQUIRE(value >= 1, "iteration limit must be >= 1");
    if ((key == MIPCU_PARAM_SCALING || key == MIPCU_PARAM_STEP_SIZE) && ctx->prepared)
      throw Error("scaling / step size must be set before the first solve of a loaded model");
    ctx->params[key] = value;
    if (key == MIPCU_PARAM_USE_TMA && ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    if (key == MIPCU_PARAM_THREADS_PER_ROW && ctx->loaded) {
      int o = (int)value;
      ctx->rows.tpr = choose_tpr(ctx->m ? (double)ctx->nnz / ctx->m : 0.0, o);
      ctx->cols.tpr = choose_tpr(ctx->n ? (double)ctx->nnz / ctx->n : 0.0, o);
      if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; }
    }
  });
}

int mipcu_get_param(const mipcu_ctx* ctx, int key, double* value) {
  if (!ctx || !value || key < 0 || key >= MIPCU_PARAM_COUNT_) return 1;
  *value = ctx->params[key];
  return 0;
}

int mipcu_warm_start(mipcu_ctx* ctx, const double* x, const double* y) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(ctx->loaded, "mipcu_warm_start: no model loaded");
    MIPCU_CK(cudaSetDevice(ctx->device));
    auto set = [&](double*& dst, const double* src, int count) {
      if (!src) { if (dst) cudaFree(dst); dst = nullptr; return; }
      if (!dst) MIPCU_CK(cudaMalloc(&dst, sizeof(double) * (size_t)std::max(count, 1)));
      if (count) MIPCU_CK(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    };
    set(ctx->x_start, x, ctx->n);
    set(ctx->y_start, y, ctx->m);
    MIPCU_CK(cudaStreamSynchronize(ctx->stream));
  });
}

int mipcu_solve_lp(mipcu_ctx* ctx, mipcu_stats* out) {
  if (!ctx) return 1;
  return guarded(ctx, [&] { solve_lp(ctx, out); });
}

int mipcu_solve_lp_batch(mipcu_ctx* ctx, int32_t B, const double* lb, const double* ub,
                         const double* cutoff, mipcu_stats* out, double* x_out) {
  if (!ctx) return 1;
  return guarded(ctx, [&] { solve_lp_batch(ctx, B, lb, ub, cutoff, nullptr, nullptr, nullptr, nullptr, out, x_out, nullptr); });
}

int mipcu_solve_lp_batch_warm(mipcu_ctx* ctx, int32_t B, const double* lb, const double* ub,
                              const double* cutoff, const double* x0, const double* y0,
                              const double* primal_weight0, const double* branch_tol, mipcu_stats* out,
                              double* x_out, double* y_out) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    solve_lp_batch(ctx, B, lb, ub, cutoff, x0, y0, primal_weight0, branch_tol, out, x_out, y_out);
  });
}

int mipcu_get_x(mipcu_ctx* ctx, double* x) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(ctx->has_solution, "mipcu_get_x: no solution available");
    MIPCU_REQUIRE(x || ctx->n == 0, "mipcu_get_x: null output");
    MIPCU_CK(cudaSetDevice(ctx->device));
    if (ctx->n) MIPCU_CK(cudaMemcpy(x, ctx->x_sol, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost));
  });
}

int mipcu_get_y(mipcu_ctx* ctx, double* y) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(ctx->has_solution, "mipcu_get_y: no solution available");
    MIPCU_REQUIRE(y || ctx->m == 0, "mipcu_get_y: null output");
    MIPCU_CK(cudaSetDevice(ctx->device));
    if (ctx->m) MIPCU_CK(cudaMemcpy(y, ctx->y_sol, sizeof(double) * ctx->m, cudaMemcpyDeviceToHost));
  });
}

int mipcu_spmv(mipcu_ctx* ctx, int transpose, const double* v, double* out) {
  if (!ctx) return 1;
  return guarded(ctx, [&] { spmv_host(ctx, transpose, v, out); });
}

int mipcu_prepare(mipcu_ctx* ctx) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_CK(cudaSetDevice(ctx->device));
    prepare(ctx);
    refresh_vectors(ctx);
  });
}

int mipcu_get_scaling(mipcu_ctx* ctx, double* d1, double* d2) {
  if (!ctx) return 1;
  return guarded(ctx, [&] {
    MIPCU_REQUIRE(ctx->loaded, "mipcu_get_scaling: no model loaded");
    MIPCU_CK(cudaSetDevice(ctx->device));
    if (d1 && ctx->m) {
      MIPCU_CK(cudaMemcpy(d1, ctx->d1, sizeof(double) * ctx->m, cudaMemcpyDeviceToHost));
      rows_to_caller_order(ctx, d1);
    }
    if (d2 && ctx->n) {
      MIPCU_CK(cudaMemcpy(d2, ctx->d2, sizeof(double) * ctx->n, cudaMemcpyDeviceToHost));
      cols_to_caller_order(ctx, d2);
    }
  });
}

int mipcu_debug_iterate(mipcu_ctx* ctx, int64_t iters, double* xbar_out, double* y_out) {
  if (!ctx) return 1;
  return guarded(ctx, [&] { debug_iterate(ctx, iters, xbar_out, y_out); });
}

int mipcu_time_kernel(mipcu_ctx* ctx, int which, int reps, int flush_l2, double* mean_us,
                      int64_t* algorithmic_bytes) {
  if (!ctx) return 1;
  return guarded(ctx, [&] { time_kernel(ctx, which, reps, flush_l2, mean_us, algorithmic_bytes); });
}

}  // extern "C"
