// tc_comm.cu — multi-GPU plumbing (one process per GPU): the NCCL bootstrap of the sharded ICP
// reduction, and the NVLink peer-memory window of the distributed normals.
//
// The reference has no distributed backend at all.  The exchange steps on
// the path are (1) the per-iteration sum-all-reduce of the 29 f64 normal-equation scalars of the
// sharded ICP, and (2) for tc_estimate_normals_distributed, the all-gather of the cloud's chunks
// and the scatter of every rank's normal rows to the ranks that own them - both done by this
// library's own kernels over peer memory, not by collective calls.  libnccl is resolved at run time (dlopen) so the library loads — and every
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
  // distributed normals: one cudaMalloc'ed window per rank, IPC-mapped into every other rank:
  // [0, 256) barrier flags (one u64 per rank) | the whole cloud, n x 12 B | this rank's chunk of
  // the result, chunk x 24 B
  char* win = nullptr;
  char* win_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t win_n = 0, win_chunk = 0, win_out_off = 0, win_bytes = 0;
  bool win_open = false;
  char* win_retired = nullptr;       // the previous window: freed once every peer has let go of it
  unsigned long long bar_epoch = 0;  // barriers passed so far (identical on every rank)
  uint32_t* d_err = nullptr;         // a barrier that timed out sets this
};
constexpr uint64_t kWinFlagsBytes = 256;

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

// ------------------------------------------------------------------ distributed normals
extern "C" void tc_dist_chunk(uint64_t n, int n_ranks, int rank, uint64_t* lo, uint64_t* hi) {
  // contiguous row ranges of equal length (a multiple of 4 rows, so chunk boundaries stay
  // 16-byte aligned in the 12-byte point array); the last ranks may hold fewer rows, or none
  const uint64_t w = (uint64_t)(n_ranks < 1 ? 1 : n_ranks);
  const uint64_t chunk = std::max<uint64_t>(4, (((n + w - 1) / w) + 3) & ~(uint64_t)3);
  const uint64_t a = std::min<uint64_t>(n, (uint64_t)rank * chunk);
  const uint64_t b = std::min<uint64_t>(n, a + chunk);
  if (lo) *lo = a;
  if (hi) *hi = b;
}
static uint64_t dist_chunk_len(uint64_t n, int n_ranks) {
  const uint64_t w = (uint64_t)n_ranks;
  return std::max<uint64_t>(4, (((n + w - 1) / w) + 3) & ~(uint64_t)3);
}

// Lets go of the peers' windows and retires the own one.  An exported allocation must not be
// freed while a peer still has it mapped: the own window is only freed by the NEXT
// tc_comm_window_open (the host has gathered every rank's new handle by then, so every rank has
// been through here and closed its mapping) or at tc_comm_destroy.
static void window_close(tc_comm* comm, bool free_now) {
  if (comm->win_open)
    for (int r = 0; r < comm->n_ranks; ++r)
      if (r != comm->rank && comm->win_peer[r]) cudaIpcCloseMemHandle(comm->win_peer[r]);
  for (int r = 0; r < 8; ++r) comm->win_peer[r] = nullptr;
  comm->win_open = false;
  if (comm->win_retired) cudaFree(comm->win_retired);  // (two generations old: nobody maps it)
  comm->win_retired = nullptr;
  if (free_now) {
    if (comm->win) cudaFree(comm->win);
  } else {
    comm->win_retired = comm->win;
  }
  comm->win = nullptr;
  comm->win_n = 0;
}

extern "C" int tc_comm_window_handle(tc_comm* comm, uint64_t n_points, void* handle_out) {
  if (!comm || !handle_out) return TC_INVALID_DATA;
  tc_context* ctx = comm->ctx;
  if (comm->n_ranks > 8) return tc_fail(ctx, TC_INVALID_DATA, "peer exchange supports <= 8 ranks");
  if (n_points == 0 || n_points >= 0xFFFFFFFFull)
    return tc_fail(ctx, TC_INVALID_DATA, "window: 0 < n_points < 2^32-1");
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  window_close(comm, false);
  comm->win_n = n_points;
  comm->win_chunk = dist_chunk_len(n_points, comm->n_ranks);
  // (the cloud region is sized for whole chunks so that every rank's chunk copy stays inside)
  const uint64_t pts_bytes = (comm->win_chunk * comm->n_ranks * 12 + 255) & ~(uint64_t)255;
  comm->win_out_off = kWinFlagsBytes + pts_bytes;
  comm->win_bytes = comm->win_out_off + comm->win_chunk * 24;
  TC_CUDA(ctx, cudaMalloc((void**)&comm->win, comm->win_bytes));
  TC_CUDA(ctx, cudaMemset(comm->win, 0, kWinFlagsBytes));
  if (!comm->d_err) {
    TC_CUDA(ctx, cudaMalloc((void**)&comm->d_err, sizeof(uint32_t)));
    TC_CUDA(ctx, cudaMemset(comm->d_err, 0, sizeof(uint32_t)));
  }
  comm->bar_epoch = 0;
  cudaIpcMemHandle_t h;
  TC_CUDA(ctx, cudaIpcGetMemHandle(&h, comm->win));
  memcpy(handle_out, &h, sizeof(h));
  return TC_OK;
}

extern "C" int tc_comm_window_open(tc_comm* comm, const void* all_handles) {
  if (!comm || !all_handles) return TC_INVALID_DATA;
  tc_context* ctx = comm->ctx;
  if (!comm->win) return tc_fail(ctx, TC_INVALID_DATA, "call tc_comm_window_handle first");
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  if (comm->win_retired) {  // every rank has created its new window, i.e. closed the old mappings
    cudaFree(comm->win_retired);
    comm->win_retired = nullptr;
  }
  for (int r = 0; r < comm->n_ranks; ++r) {
    if (r == comm->rank) {
      comm->win_peer[r] = comm->win;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)all_handles + (size_t)r * TC_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    TC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    comm->win_peer[r] = (char*)p;
  }
  comm->win_open = true;
  return TC_OK;
}

namespace {

struct PeerWin {
  char* peer[8];
  int world, rank;
};

// Stream-ordered barrier over the ranks' windows: thread t tells rank t "this rank has passed
// barrier number `epoch`" (everything this rank enqueued before - its upload, its kernel's rows
// in the peers' windows - is complete: stream order, then a system-scope fence) and waits until
// rank t has said the same here.  A peer that never arrives fails the call instead of hanging.
__global__ void k_peer_barrier(PeerWin w, unsigned long long epoch, unsigned long long spin_limit,
                               uint32_t* __restrict__ err) {
  const int t = threadIdx.x;
  if (t >= w.world) return;
  __threadfence_system();
  reinterpret_cast<volatile unsigned long long*>(w.peer[t])[w.rank] = epoch;
  const volatile unsigned long long* mine =
      reinterpret_cast<const volatile unsigned long long*>(w.peer[w.rank]);
  unsigned long long spins = 0;
  while (mine[t] < epoch) {
    __nanosleep(64);
    if (++spins > spin_limit) {
      *err = 1u;
      break;
    }
  }
  __threadfence_system();
}

// All-gather of the cloud over NVLink: every rank PULLS the chunks of the other ranks out of
// their windows into its own (128-bit loads from all peers at once, so the ingress of this GPU's
// NVLink ports is what bounds it, not one peer link at a time).
__global__ void __launch_bounds__(256) k_gather_chunks(PeerWin w, uint64_t pts_off,
                                                       uint64_t chunk_vec /* float4 per chunk */,
                                                       uint64_t total_vec) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4* mine = reinterpret_cast<float4*>(w.peer[w.rank] + pts_off);
  for (int j = 1; j < w.world; ++j) {
    const int r = (w.rank + j) % w.world;  // (every rank starts at a different peer)
    const uint64_t a = (uint64_t)r * chunk_vec;
    if (a >= total_vec) continue;
    const uint64_t b = min(a + chunk_vec, total_vec);
    const float4* src = reinterpret_cast<const float4*>(w.peer[r] + pts_off);
    uint64_t i = a + tid;
    for (; i + stride < b; i += 2 * stride) {  // two 128-bit peer loads in flight per thread
      const float4 u = src[i], v = src[i + stride];
      mine[i] = u;
      mine[i + stride] = v;
    }
    if (i < b) mine[i] = src[i];
  }
}

int peer_barrier(tc_comm* comm) {
  tc_context* ctx = comm->ctx;
  PeerWin w{};
  for (int r = 0; r < 8; ++r) w.peer[r] = comm->win_peer[r];
  w.world = comm->n_ranks;
  w.rank = comm->rank;
  double secs = 30.0;
  if (const char* e = std::getenv("TC_PEER_TIMEOUT_S")) secs = std::max(0.1, atof(e));
  k_peer_barrier<<<1, 32, 0, ctx->stream>>>(w, ++comm->bar_epoch,
                                            (unsigned long long)(secs / 100e-9), comm->d_err);
  TC_LAUNCHED(ctx);
  return TC_OK;
}

}  // namespace

// estimate_normals (normals.rs:238-268) over the ranks of `comm`: rank r passes rows
// tc_dist_chunk(n_total, n_ranks, r) of the cloud and receives the NormalPoint3f rows of the same
// range.  Per call and rank: one upload of the chunk, a pull of the other chunks over NVLink, the
// slab-sharded index build, the normals kernel - which writes each row straight into the window
// of the rank that owns it - and one download of the chunk's rows.
extern "C" int tc_estimate_normals_distributed(tc_context* ctx, tc_comm* comm,
                                               const float* chunk_xyz, uint64_t n_total, uint32_t k,
                                               int consistent_orientation, const float* viewpoint3,
                                               float* chunk_out) {
  if (!ctx || !comm) return TC_INVALID_DATA;
  TcRange nvtx_range("tc_estimate_normals_distributed");
  if (n_total == 0) return TC_OK;  // empty -> Ok(empty), before the k check (normals.rs:261-263)
  if (k < 3) return tc_fail(ctx, TC_INVALID_DATA, "k_neighbors must be at least 3");
  if (!comm->win_open || comm->win_n != n_total)
    return tc_fail(ctx, TC_INVALID_DATA, "open a window for this cloud size first (tc_comm_window_*)");
  uint64_t lo, hi;
  tc_dist_chunk(n_total, comm->n_ranks, comm->rank, &lo, &hi);
  if (hi > lo && (!chunk_xyz || !chunk_out)) return TC_INVALID_DATA;
  TC_CUDA(ctx, cudaSetDevice(ctx->device));
  float* d_pts = reinterpret_cast<float*>(comm->win + kWinFlagsBytes);
  float* d_rows = reinterpret_cast<float*>(comm->win + comm->win_out_off);
  if (hi > lo)
    TC_CUDA(ctx, cudaMemcpyAsync(d_pts + 3 * lo, chunk_xyz, (hi - lo) * 12, cudaMemcpyHostToDevice,
                                 ctx->stream));
  TC_TRY(peer_barrier(comm));  // every rank's chunk is in its window
  if (comm->n_ranks > 1) {
    PeerWin w{};
    for (int r = 0; r < 8; ++r) w.peer[r] = comm->win_peer[r];
    w.world = comm->n_ranks;
    w.rank = comm->rank;
    const uint64_t chunk_vec = comm->win_chunk * 12 / 16;
    const uint64_t total_vec = (n_total * 12 + 15) / 16;  // (inside the whole-chunk region)
    const int blocks = (int)std::min<uint64_t>((total_vec + 255) / 256, (uint64_t)ctx->sm_count * 8);
    k_gather_chunks<<<blocks, 256, 0, ctx->stream>>>(w, kWinFlagsBytes, chunk_vec, total_vec);
    TC_LAUNCHED(ctx);
  }
  tc_cloud view;
  view.ctx = ctx;
  view.n = n_total;
  view.d_xyz = d_pts;
  tc_index* ix = nullptr;
  int st = tci_index_build(ctx, &view, k, 0.0f, comm->rank, comm->n_ranks, &ix);
  if (st == TC_OK) {
    // rows of original index i go to rank i / chunk, at row i - rank * chunk of its window
    for (int r = 0; r < comm->n_ranks; ++r)
      ctx->route.base[r] = reinterpret_cast<float*>(comm->win_peer[r] + comm->win_out_off) -
                           6 * (int64_t)((uint64_t)r * comm->win_chunk);
    ctx->route.chunk = (uint32_t)comm->win_chunk;
    ctx->route.magic = (uint32_t)((1ull << 32) / comm->win_chunk);
    st = tc_estimate_normals_device(ctx, ix, k, 0.0f, consistent_orientation, viewpoint3, 0,
                                    UINT64_MAX, d_rows);
    ctx->route = tc_context::OutRoute{};
  }
  if (st == TC_OK) st = peer_barrier(comm);  // every rank's rows have landed
  if (st == TC_OK && hi > lo &&
      cudaMemcpyAsync(chunk_out, d_rows, (hi - lo) * 24, cudaMemcpyDeviceToHost, ctx->stream) !=
          cudaSuccess)
    st = tc_fail(ctx, TC_GPU, "distributed normals: download failed");
  uint32_t err = 0;
  if (st == TC_OK &&
      (cudaMemcpyAsync(&err, comm->d_err, sizeof(err), cudaMemcpyDeviceToHost, ctx->stream) !=
           cudaSuccess ||
       cudaStreamSynchronize(ctx->stream) != cudaSuccess))
    st = tc_fail(ctx, TC_GPU, "distributed normals failed");
  if (ix) tc_index_free(ix);
  if (st == TC_OK && err) {
    cudaMemsetAsync(comm->d_err, 0, sizeof(uint32_t), ctx->stream);
    st = tc_fail(ctx, TC_GPU, "distributed normals: a peer rank did not reach the barrier");
  }
  return st;
}

extern "C" void tc_comm_destroy(tc_comm* comm) {
  if (!comm) return;
  window_close(comm, true);
  if (comm->d_err) cudaFree(comm->d_err);
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
