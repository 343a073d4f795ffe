// tc_comm.cu — NCCL bootstrap for the sharded ICP reduction (one process per GPU).
//
// The reference has no distributed backend at all; this is the only
// collective on the path: a per-iteration sum-all-reduce of the 29 f64 normal-equation scalars
// over NVLink.  libnccl is resolved at run time (dlopen) so the library loads — and every
// single-GPU entry point works — on hosts without NCCL; the unique id is exchanged out of band
// by the host (torch.distributed / MPI / a Rust channel).
#include <dlfcn.h>
#include <nccl.h>

#include "tc_internal.cuh"

struct tc_comm {
  tc_context* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
  // NVLink peer exchange for the fused ICP all-reduce: one small cudaMalloc'ed buffer per rank,
  // IPC-mapped into every other rank (2 parities x n_ranks slots x 32 doubles)
  double* xbuf = nullptr;
  double* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool peers_open = false;
  unsigned long long epoch = 1;  // advances identically on every rank (same call sequence)
};

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
  api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce &&
           api.GetErrorString;
  return api;
}

}  // namespace

static_assert(TC_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

extern "C" int tc_comm_get_unique_id(tc_context* ctx, void* id_out) {
  if (!ctx || !id_out) return TC_INVALID_DATA;
  NcclApi& api = nccl();
  if (!api.ok) return tc_fail(ctx, TC_GPU, "libnccl.so.2 not found (multi-GPU path unavailable)");
  ncclUniqueId id;
  const ncclResult_t r = api.GetUniqueId(&id);
  if (r != ncclSuccess)
    return tc_fail(ctx, TC_GPU, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
  memcpy(id_out, id.internal, NCCL_UNIQUE_ID_BYTES);
  return TC_OK;
}

extern "C" int tc_comm_init_rank(tc_context* ctx, const void* id, int n_ranks, int rank,
                                 tc_comm** out) {
  if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return TC_INVALID_DATA;
  *out = nullptr;
  NcclApi& api = nccl();
  if (!api.ok) return tc_fail(ctx, TC_GPU, "libnccl.so.2 not found (multi-GPU path unavailable)");
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(uid.internal, id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t c = nullptr;
  const ncclResult_t r = api.CommInitRank(&c, n_ranks, uid, rank);
  if (r != ncclSuccess)
    return tc_fail(ctx, TC_GPU, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
  tc_comm* cm = new tc_comm();
  cm->ctx = ctx;
  cm->comm = c;
  cm->n_ranks = n_ranks;
  cm->rank = rank;
  *out = cm;
  return TC_OK;
}

extern "C" int tc_comm_peer_handle(tc_comm* comm, void* handle_out) {
  if (!comm || !handle_out) return TC_INVALID_DATA;
  tc_context* ctx = comm->ctx;
  if (comm->n_ranks > 8) return tc_fail(ctx, TC_INVALID_DATA, "peer exchange supports <= 8 ranks");
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!comm->xbuf) {
    const size_t bytes = (size_t)2 * comm->n_ranks * 32 * sizeof(double);
    TC_CUDA(ctx, cudaMalloc((void**)&comm->xbuf, bytes));
    TC_CUDA(ctx, cudaMemset(comm->xbuf, 0, bytes));
  }
  cudaIpcMemHandle_t h;
  TC_CUDA(ctx, cudaIpcGetMemHandle(&h, comm->xbuf));
  static_assert(sizeof(cudaIpcMemHandle_t) == TC_IPC_HANDLE_BYTES, "ipc handle size");
  memcpy(handle_out, &h, sizeof(h));
  return TC_OK;
}

extern "C" int tc_comm_peer_open(tc_comm* comm, const void* all_handles) {
  if (!comm || !all_handles) return TC_INVALID_DATA;
  tc_context* ctx = comm->ctx;
  if (!comm->xbuf) return tc_fail(ctx, TC_INVALID_DATA, "call tc_comm_peer_handle first");
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < comm->n_ranks; ++r) {
    if (r == comm->rank) {
      comm->peer[r] = comm->xbuf;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + (size_t)r * TC_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    TC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    comm->peer[r] = (double*)p;
  }
  comm->peers_open = true;
  return TC_OK;
}

bool tci_comm_peers(tc_comm* comm, double** peers8, int* world, int* rank,
                    unsigned long long* epoch_base, unsigned long long epochs_needed) {
  if (!comm || !comm->peers_open) return false;
  for (int r = 0; r < 8; ++r) peers8[r] = comm->peer[r];
  *world = comm->n_ranks;
  *rank = comm->rank;
  *epoch_base = comm->epoch;
  comm->epoch += epochs_needed;
  return true;
}
// Epochs are consumed one per executed exchange, so consecutive calls use consecutive epochs and
// the two parity slots alternate across call boundaries as well: a rank that has started the next
// call can only be one epoch ahead of a rank still reading the previous call's last slot.
void tci_comm_commit_epochs(tc_comm* comm, unsigned long long used) {
  if (comm) comm->epoch += used;
}

extern "C" void tc_comm_destroy(tc_comm* comm) {
  if (!comm) return;
  if (comm->peers_open)
    for (int r = 0; r < comm->n_ranks; ++r)
      if (r != comm->rank && comm->peer[r]) cudaIpcCloseMemHandle(comm->peer[r]);
  if (comm->xbuf) cudaFree(comm->xbuf);
  if (comm->comm && nccl().ok) nccl().CommDestroy(comm->comm);
  delete comm;
}

int tci_comm_allreduce(tc_comm* comm, double* d_buf, uint64_t count) {
  NcclApi& api = nccl();
  const ncclResult_t r =
      api.AllReduce(d_buf, d_buf, count, ncclDouble, ncclSum, comm->comm, comm->ctx->stream);
  if (r != ncclSuccess)
    return tc_fail(comm->ctx, TC_GPU, std::string("ncclAllReduce: ") + api.GetErrorString(r));
  return TC_OK;
}

extern "C" int tc_comm_allreduce_f64(tc_comm* comm, double* d_buf, uint64_t count) {
  if (!comm || !d_buf) return TC_INVALID_DATA;
  return tci_comm_allreduce(comm, d_buf, count);
}
