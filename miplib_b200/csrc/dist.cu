// dist.cu - the collective side of the row-sharded solve: one process per
// GPU, one NCCL communicator over NVLink 5 / NVSwitch.  The only data-path exchange of the method
// is the sum over ranks of the partial products K_g' yhat_g (n doubles per iteration); everything
// else that crosses ranks is a handful of scalars on check iterations and the column norms during
// the one-off equilibration.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a PyTorch process this resolves to the
// copy torch already loaded, so there is exactly one NCCL in the address space; stand-alone C++
// hosts get the system library.  Only the long-stable part of the API is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "internal.h"

namespace mipcu {

namespace {

struct Nccl {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Nccl& nccl() {
  static Nccl api;
  static std::once_flag once;
  static std::string error;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      error = std::string("cannot load NCCL: ") + dlerror();
      return;
    }
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (!p) error = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  if (!error.empty()) throw Error(error);
  return api;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw Error(std::string(what) + ": " + nccl().GetErrorString(r));
}

static_assert(sizeof(ncclUniqueId) == MIPCU_UNIQUE_ID_BYTES, "unique id size");

}  // namespace

void dist_unique_id(void* out) {
  ncclUniqueId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(out, &id, sizeof(id));
}

void dist_init(mipcu_ctx* ctx, int rank, int world, const void* unique_id) {
  MIPCU_REQUIRE(unique_id, "mipcu_create_sharded: null unique id");
  ncclUniqueId id;
  std::memcpy(&id, unique_id, sizeof(id));
  MIPCU_CK(cudaSetDevice(ctx->device));
  ncclComm_t comm = nullptr;
  nccl_check(nccl().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->world = world;
  MIPCU_CK(cudaMalloc(&ctx->row_sums, sizeof(double) * 32));
}

void dist_destroy(mipcu_ctx* ctx) {
  if (ctx->nccl_comm) nccl().CommDestroy(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  if (ctx->row_sums) cudaFree(ctx->row_sums);
  ctx->row_sums = nullptr;
}

// in-place all-reduce of a device vector on the solver's stream; no-op for a single rank
void dist_allreduce(mipcu_ctx* ctx, double* buf, size_t count, bool max_instead_of_sum) {
  if (ctx->world <= 1 || count == 0) return;
  nccl_check(nccl().AllReduce(buf, buf, count, ncclFloat64, max_instead_of_sum ? ncclMax : ncclSum,
                              static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream),
             "ncclAllReduce");
  ctx->collectives++;
}

// raw bytes from every rank, in rank order (setup only: IPC handle exchange)
void dist_allgather_bytes(mipcu_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes_per_rank) {
  nccl_check(nccl().AllGather(send_dev, recv_dev, bytes_per_rank, ncclUint8,
                              static_cast<ncclComm_t>(ctx->nccl_comm), ctx->stream),
             "ncclAllGather");
  ctx->collectives++;
}

// sum of one host scalar over the ranks (setup only)
double dist_sum_scalar(mipcu_ctx* ctx, double v) {
  if (ctx->world <= 1) return v;
  MIPCU_CK(cudaMemcpyAsync(ctx->row_sums + 31, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  dist_allreduce(ctx, ctx->row_sums + 31, 1, false);
  double out = 0;
  MIPCU_CK(cudaMemcpyAsync(&out, ctx->row_sums + 31, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MIPCU_CK(cudaStreamSynchronize(ctx->stream));
  return out;
}

}  // namespace mipcu
