// mc.cu - NVSwitch multicast window for the block-owner exchange (p2p.cu).
//
// The two vectors every rank keeps whole - xbar (n doubles) and yhat_full (m_glob doubles) - are
// moved into one physical allocation per GPU (cuMemCreate) that is bound to ONE multicast object
// shared by all ranks.  A store to the multicast mapping of that window is replicated by the
// NVSwitch into the window of every GPU, so the all-gather push of a slice costs one store per
// element instead of W - 1: at 8 GPUs the egress of an iteration drops from 21 MB to 3 MB per rank,
// which was the largest single cost of the exchange (profiles/r01_block_owner_exchange.md).
// Reads go through the ordinary (unicast) mapping of the rank's own window.
//
// Setup is collective: rank 0 creates the multicast object and hands its POSIX file descriptor to
// the other ranks over abstract unix datagram sockets (SCM_RIGHTS); everything else a rank needs is
// its own.  Driver entry points are resolved through the runtime (no link-time libcuda).  If any
// step fails on any rank, every rank stays on unicast pushes into the CUDA-IPC mappings.
#include <cuda.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <random>

#include "internal.h"

namespace mipcu {

namespace {

struct Driver {
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  bool ok = false;
};

const Driver& driver() {
  static Driver d = [] {
    Driver r;
    bool all = true;
    auto get = [&](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess ||
          q != cudaDriverEntryPointSuccess || !*fn) {
        cudaGetLastError();
        all = false;
      }
    };
    get("cuDeviceGet", (void**)&r.DeviceGet);
    get("cuDeviceGetAttribute", (void**)&r.DeviceGetAttribute);
    get("cuMulticastCreate", (void**)&r.MulticastCreate);
    get("cuMulticastAddDevice", (void**)&r.MulticastAddDevice);
    get("cuMulticastBindMem", (void**)&r.MulticastBindMem);
    get("cuMulticastUnbind", (void**)&r.MulticastUnbind);
    get("cuMulticastGetGranularity", (void**)&r.MulticastGetGranularity);
    get("cuMemCreate", (void**)&r.MemCreate);
    get("cuMemRelease", (void**)&r.MemRelease);
    get("cuMemGetAllocationGranularity", (void**)&r.MemGetAllocationGranularity);
    get("cuMemAddressReserve", (void**)&r.MemAddressReserve);
    get("cuMemAddressFree", (void**)&r.MemAddressFree);
    get("cuMemMap", (void**)&r.MemMap);
    get("cuMemUnmap", (void**)&r.MemUnmap);
    get("cuMemSetAccess", (void**)&r.MemSetAccess);
    get("cuMemExportToShareableHandle", (void**)&r.MemExportToShareableHandle);
    get("cuMemImportFromShareableHandle", (void**)&r.MemImportFromShareableHandle);
    r.ok = all;
    return r;
  }();
  return d;
}

// ---- one file descriptor from rank 0 to every other rank ------------------------------------------
sockaddr_un socket_name(unsigned long long token, int rank, socklen_t* len) {
  sockaddr_un a;
  std::memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  char name[64];
  const int k = std::snprintf(name, sizeof(name), "mipcu-mc-%016llx-%d", token, rank);
  std::memcpy(a.sun_path + 1, name, (size_t)k);          // abstract namespace: sun_path[0] == 0
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + k);
  return a;
}

int open_receiver(unsigned long long token, int rank) {
  const int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (s < 0) return -1;
  socklen_t len;
  sockaddr_un a = socket_name(token, rank, &len);
  timeval tv = {20, 0};
  setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
  if (bind(s, (sockaddr*)&a, len) != 0) { close(s); return -1; }
  return s;
}

bool send_fd(unsigned long long token, int to_rank, int fd) {
  const int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  if (s < 0) return false;
  socklen_t len;
  sockaddr_un a = socket_name(token, to_rank, &len);
  char payload = 'm';
  iovec iov = {&payload, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  std::memset(ctrl, 0, sizeof(ctrl));
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_name = &a;
  msg.msg_namelen = len;
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  cmsghdr* c = CMSG_FIRSTHDR(&msg);
  c->cmsg_level = SOL_SOCKET;
  c->cmsg_type = SCM_RIGHTS;
  c->cmsg_len = CMSG_LEN(sizeof(int));
  std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  const bool ok = sendmsg(s, &msg, 0) == 1;
  close(s);
  return ok;
}

int recv_fd(int s) {
  char payload = 0;
  iovec iov = {&payload, 1};
  alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
  msghdr msg;
  std::memset(&msg, 0, sizeof(msg));
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  if (recvmsg(s, &msg, 0) != 1) return -1;
  for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c))
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      int fd = -1;
      std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
      return fd;
    }
  return -1;
}

size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

// every rank must agree before the next collective step: true only if all ranks pass true
bool all_ranks(mipcu_ctx* ctx, bool mine) {
  return dist_sum_scalar(ctx, mine ? 1.0 : 0.0) > ctx->world - 0.5;
}

}  // namespace

void mc_teardown(mipcu_ctx* ctx) {
  McWindow& w = ctx->mc;
  if (w.saved_xbar) {          // hand the iteration its CUDA-IPC-exported vectors back
    ctx->xbar = w.saved_xbar;
    ctx->yhat_full = w.saved_yhat_full;
  }
  if (w.active || w.mc_va || w.uc_va || w.mc_handle || w.mem_handle) {
    const Driver& d = driver();
    if (d.ok) {
      if (ctx->stream) cudaStreamSynchronize(ctx->stream);
      if (w.mc_va) { d.MemUnmap((CUdeviceptr)w.mc_va, w.size); d.MemAddressFree((CUdeviceptr)w.mc_va, w.size); }
      if (w.uc_va) { d.MemUnmap((CUdeviceptr)w.uc_va, w.size); d.MemAddressFree((CUdeviceptr)w.uc_va, w.size); }
      if (w.mc_handle && w.bound) {
        CUdevice dev;
        if (d.DeviceGet(&dev, ctx->device) == CUDA_SUCCESS) d.MulticastUnbind(w.mc_handle, dev, 0, w.size);
      }
      if (w.mem_handle) d.MemRelease(w.mem_handle);
      if (w.mc_handle) d.MemRelease(w.mc_handle);
    }
  }
  w = McWindow();
  ctx->peers.mc_xbar = nullptr;
  ctx->peers.mc_yhat_full = nullptr;
}

// Collective, from own_setup.  Returns true when every rank has the window mapped.
bool mc_setup(mipcu_ctx* ctx) {
  mc_teardown(ctx);
  if (ctx->params[MIPCU_PARAM_MULTICAST] == 0.0) return false;   // the same on every rank (set before load)
  const int W = ctx->world;
  McWindow& w = ctx->mc;
  const Driver& d = driver();
  CUdevice dev = 0;
  bool ok = d.ok && d.DeviceGet(&dev, ctx->device) == CUDA_SUCCESS;
  if (ok) {
    int supported = 0;
    ok = d.DeviceGetAttribute(&supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) == CUDA_SUCCESS && supported;
  }
  if (!all_ranks(ctx, ok)) return false;

  // window layout: [ xbar : n + 2 | yhat_full : m_glob + 2 ], each part starting on a 4 KiB boundary
  const size_t xbar_bytes = round_up(sizeof(double) * ((size_t)ctx->n + 2), 4096);
  const size_t yhat_bytes = round_up(sizeof(double) * ((size_t)ctx->m_glob + 2), 4096);
  CUmulticastObjectProp mprop;
  std::memset(&mprop, 0, sizeof(mprop));
  mprop.numDevices = (unsigned)W;
  mprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CUmemAllocationProp aprop;
  std::memset(&aprop, 0, sizeof(aprop));
  aprop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  aprop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  aprop.location.id = ctx->device;
  aprop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran_mc = 0, gran_mem = 0;
  mprop.size = xbar_bytes + yhat_bytes;
  ok = d.MulticastGetGranularity(&gran_mc, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
       d.MemGetAllocationGranularity(&gran_mem, &aprop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS &&
       gran_mc > 0 && gran_mem > 0;
  if (!all_ranks(ctx, ok)) return false;
  const size_t gran = std::max(gran_mc, gran_mem);
  w.size = round_up(round_up(xbar_bytes + yhat_bytes, gran_mc), gran_mem);
  mprop.size = w.size;

  // a token only this job knows names the sockets; rank 0's choice is used by everybody
  unsigned long long token = 0, *d_tok = nullptr;
  {
    std::random_device rd;
    token = ((unsigned long long)rd() << 32) ^ rd() ^ ((unsigned long long)getpid() << 16);
    std::vector<unsigned long long> all((size_t)W);
    MIPCU_CK(cudaMalloc(&d_tok, sizeof(unsigned long long) * (W + 1)));
    MIPCU_CK(cudaMemcpyAsync(d_tok + W, &token, sizeof(token), cudaMemcpyHostToDevice, ctx->stream));
    dist_allgather_bytes(ctx, d_tok + W, d_tok, sizeof(token));
    MIPCU_CK(cudaMemcpyAsync(all.data(), d_tok, sizeof(token) * W, cudaMemcpyDeviceToHost, ctx->stream));
    MIPCU_CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_tok);
    token = all[0];
  }
  int sock = -1;
  if (ctx->rank != 0) {
    sock = open_receiver(token, ctx->rank);
    ok = sock >= 0;
  } else {
    ok = d.MulticastCreate(&w.mc_handle, &mprop) == CUDA_SUCCESS;
    if (!ok) w.mc_handle = 0;
  }
  if (!all_ranks(ctx, ok)) {            // receivers are bound, the object exists: nothing sent yet
    if (sock >= 0) close(sock);
    mc_teardown(ctx);
    return false;
  }
  if (ctx->rank == 0) {
    int fd = -1;
    ok = d.MemExportToShareableHandle(&fd, w.mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS;
    for (int p = 1; p < W && ok; ++p) ok = send_fd(token, p, fd);
    if (fd >= 0) close(fd);
  } else {
    const int fd = recv_fd(sock);
    close(sock);
    ok = fd >= 0 && d.MemImportFromShareableHandle(&w.mc_handle, (void*)(uintptr_t)fd,
                                                   CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) == CUDA_SUCCESS;
    if (!ok) w.mc_handle = 0;
    if (fd >= 0) close(fd);
  }
  if (!all_ranks(ctx, ok)) { mc_teardown(ctx); return false; }

  ok = d.MulticastAddDevice(w.mc_handle, dev) == CUDA_SUCCESS;
  if (!all_ranks(ctx, ok)) { mc_teardown(ctx); return false; }      // also: every device is in the team

  ok = d.MemCreate(&w.mem_handle, w.size, &aprop, 0) == CUDA_SUCCESS;
  if (!ok) w.mem_handle = 0;
  if (ok) {
    ok = d.MulticastBindMem(w.mc_handle, 0, w.mem_handle, 0, w.size, 0) == CUDA_SUCCESS;
    w.bound = ok;
  }
  CUmemAccessDesc access;
  std::memset(&access, 0, sizeof(access));
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = ctx->device;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  auto map = [&](CUmemGenericAllocationHandle h, void** va) {
    CUdeviceptr p = 0;
    if (d.MemAddressReserve(&p, w.size, gran, 0, 0) != CUDA_SUCCESS) return false;
    if (d.MemMap(p, w.size, 0, h, 0) != CUDA_SUCCESS) { d.MemAddressFree(p, w.size); return false; }
    *va = (void*)p;
    return d.MemSetAccess(p, w.size, &access, 1) == CUDA_SUCCESS;
  };
  if (ok) ok = map(w.mem_handle, &w.uc_va);
  if (!all_ranks(ctx, ok)) { mc_teardown(ctx); return false; }      // every rank has bound its memory
  ok = map(w.mc_handle, &w.mc_va);
  if (ok) ok = cudaMemsetAsync(w.uc_va, 0, w.size, ctx->stream) == cudaSuccess &&
               cudaStreamSynchronize(ctx->stream) == cudaSuccess;
  if (!all_ranks(ctx, ok)) { cudaGetLastError(); mc_teardown(ctx); return false; }

  // from here on the iteration reads its own window and pushes through the multicast mapping
  w.saved_xbar = ctx->xbar;
  w.saved_yhat_full = ctx->yhat_full;
  ctx->xbar = (double*)w.uc_va;
  ctx->yhat_full = (double*)((char*)w.uc_va + xbar_bytes);
  ctx->peers.mc_xbar = (double*)w.mc_va;
  ctx->peers.mc_yhat_full = (double*)((char*)w.mc_va + xbar_bytes);
  ctx->peers.xbar[ctx->rank] = ctx->xbar;
  ctx->peers.yhat_full[ctx->rank] = ctx->yhat_full;
  w.active = true;
  return true;
}

}  // namespace mipcu
