// tc_api.cu — C ABI glue: context, device-resident cloud, and the host-buffer entry points that
// mirror the reference's public functions (see include/threecrate_cuda.h for the citations).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "tc_internal.cuh"

#define TC_ENTER(ctx)                                   \
  do {                                                  \
    if (!(ctx)) return TC_INVALID_DATA;                 \
    TC_CUDA((ctx), cudaSetDevice((ctx)->device));       \
  } while (0)

extern "C" const char* tc_version(void) { return "threecrate_cuda 0.1.0 (sm_100a)"; }

// ------------------------------------------------------------------------------------ context
extern "C" void tc_debug_set_search_flags(int flags);
extern "C" void tc_debug_set_fine_cap(int cells_per_point);
extern "C" void tc_debug_set_icp_keep(int on);

extern "C" int tc_context_create(int device, tc_context** out) {
  if (!out) return TC_INVALID_DATA;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count) {
    cudaGetLastError();
    return TC_GPU;  // no CUDA device: there is no CPU fallback
  }
  tc_context* ctx = new tc_context();
  ctx->device = device;
  auto fail = [&](const char* what) {
    fprintf(stderr, "tc_context_create: %s failed: %s\n", what,
            cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return (int)TC_GPU;
  };
  if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice");
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
    return fail("cudaStreamCreate");
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess)
    return fail("cudaEventCreate");

  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (cudaMalloc((void**)&ctx->d_scratch, tc_context::kScratchWords * sizeof(uint32_t)) != cudaSuccess)
    return fail("cudaMalloc");
  if (cudaMallocHost((void**)&ctx->h_scratch, tc_context::kScratchWords * sizeof(uint32_t)) != cudaSuccess)
    return fail("cudaMallocHost");
  if (cudaMalloc((void**)&ctx->d_planes, (size_t)tc_context::kPlaneWords * tc_context::kPlaneStride *
                                             sizeof(uint32_t)) != cudaSuccess)
    return fail("cudaMalloc");
  if (tci_scratch_arm(ctx) != TC_OK) return fail("scratch init");
  // keep freed blocks cached in the stream-ordered pool: no cudaMalloc/cudaFree per call
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  // debug overrides for A/B runs of unmodified callers (tools/, bench.py)
  if (const char* e = std::getenv("TC_SEARCH_FLAGS")) tc_debug_set_search_flags(atoi(e));
  if (const char* e = std::getenv("TC_FINE_CAP")) tc_debug_set_fine_cap(atoi(e));
  if (const char* e = std::getenv("TC_ICP_KEEP")) tc_debug_set_icp_keep(atoi(e));
  *out = ctx;
  return TC_OK;
}

extern "C" void tc_context_destroy(tc_context* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < tc_context::kWsSlots; ++i)
    if (ctx->ws[i]) cudaFree(ctx->ws[i]);
  for (int i = 0; i < tc_context::kArenaSlots; ++i)
    if (ctx->arena_cache[i]) cudaFree(ctx->arena_cache[i]);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_planes) cudaFree(ctx->d_planes);
  if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* tc_last_error(const tc_context* ctx) { return ctx ? ctx->err.c_str() : ""; }
extern "C" void* tc_context_stream(tc_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t tc_launch_count(const tc_context* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int tc_stats_enable(tc_context* ctx, int on) {
  if (!ctx) return TC_INVALID_DATA;
  ctx->stats_on = on != 0;
  return TC_OK;
}
extern "C" int tc_last_stats(tc_context* ctx, tc_stats* out) {
  TC_ENTER(ctx);
  if (!out) return TC_INVALID_DATA;
  *out = ctx->icp_stats;  // ICP history (zero when no ICP call ran with statistics on)
  uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ctx->stats_on) {
    TC_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_scratch + 48, sizeof(h), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  out->queries = ctx->stats_queries;
  out->chain_queries = h[0];
  out->rounds = h[1];
  out->box_splits = h[2];
  out->retries = h[3];
  out->candidates_staged = h[4];
  out->merges = h[5];
  return TC_OK;
}

// Measured instruction-issue ceiling (the second roofline bench.py reports: exact kNN is bounded
// by issue slots, not by HBM).  Every thread runs 16 independent dependency chains, half FMNMX
// (ALU pipe) and half FMUL (FMA pipe) - the mix of the selection networks and the distance code -
// so the schedulers always have an eligible instruction: warp-instructions per second.
namespace {
__global__ void __launch_bounds__(256) k_issue_rate(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
  const float lim = seed * 1e30f, c = 1.0f + seed * 1e-7f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        a[i] = fminf(a[i], lim);      // FMNMX (ALU pipe); the FMUL in between keeps ptxas from
        a[i] = __fmul_rn(a[i], c);    // pairing two of them into one FMNMX3
      }
    }
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // keeps the loop alive
}
}  // namespace
extern "C" int tc_debug_issue_rate(tc_context* ctx, double* warp_inst_per_s) {
  TC_ENTER(ctx);
  if (!warp_inst_per_s) return TC_INVALID_DATA;
  float* d = nullptr;
  TC_TRY(tc_alloc(ctx, &d, 1));
  const int iters = 4096, blocks = ctx->sm_count * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    TC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    k_issue_rate<<<blocks, 256, 0, ctx->stream>>>(d, iters, 0.5f);
    TC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    TC_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.0f;
    TC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (rep > 0 && ms < best) best = ms;
  }
  tc_free(ctx, d);
  const double inst = (double)blocks * 8.0 /*warps*/ * (double)iters * 128.0 /*math per trip*/;
  *warp_inst_per_s = inst / ((double)best * 1e-3);
  return TC_OK;
}

extern "C" int tc_context_synchronize(tc_context* ctx) {
  TC_ENTER(ctx);
  TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TC_OK;
}
extern "C" int tc_timer_start(tc_context* ctx) {
  TC_ENTER(ctx);
  TC_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return TC_OK;
}
extern "C" int tc_timer_stop(tc_context* ctx, float* ms_out) {
  TC_ENTER(ctx);
  TC_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  TC_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  float ms = 0.0f;
  TC_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  if (ms_out) *ms_out = ms;
  return TC_OK;
}

// ------------------------------------------------------------------------------ memory helpers
extern "C" int tc_device_alloc(tc_context* ctx, uint64_t bytes, void** d_out) {
  TC_ENTER(ctx);
  if (!d_out) return TC_INVALID_DATA;
  uint8_t* p = nullptr;
  TC_TRY(tc_alloc(ctx, &p, bytes));
  *d_out = p;
  return TC_OK;
}
extern "C" int tc_device_free(tc_context* ctx, void* d_ptr) {
  TC_ENTER(ctx);
  tc_free(ctx, d_ptr);
  return TC_OK;
}
extern "C" int tc_copy_to_device(tc_context* ctx, void* d_dst, const void* h_src, uint64_t bytes) {
  TC_ENTER(ctx);
  if (bytes == 0) return TC_OK;
  TC_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return TC_OK;
}
extern "C" int tc_copy_to_host(tc_context* ctx, void* h_dst, const void* d_src, uint64_t bytes) {
  TC_ENTER(ctx);
  if (bytes == 0) return TC_OK;
  TC_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TC_OK;
}
extern "C" int tc_host_alloc_pinned(uint64_t bytes, void** h_out) {
  if (!h_out) return TC_INVALID_DATA;
  *h_out = nullptr;
  if (cudaMallocHost(h_out, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    return TC_GPU;
  }
  return TC_OK;
}
extern "C" int tc_host_free_pinned(void* h_ptr) {
  if (h_ptr && cudaFreeHost(h_ptr) != cudaSuccess) {
    cudaGetLastError();
    return TC_GPU;
  }
  return TC_OK;
}

// -------------------------------------------------------------------------------------- cloud
namespace {
__global__ void k_destride(const uint8_t* __restrict__ src, uint64_t n, uint32_t stride,
                           float* __restrict__ dst) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const float* p = reinterpret_cast<const float*>(src + i * stride);
    dst[3 * i + 0] = p[0];
    dst[3 * i + 1] = p[1];
    dst[3 * i + 2] = p[2];
  }
}
}  // namespace

extern "C" int tc_cloud_upload(tc_context* ctx, const float* xyz_aos, uint64_t n, tc_cloud** out) {
  TC_ENTER(ctx);
  if (!out || (n > 0 && !xyz_aos)) return TC_INVALID_DATA;
  *out = nullptr;
  if (n >= 0xFFFFFFFFull) return tc_fail(ctx, TC_INVALID_DATA, "cloud too large (N < 2^32-1)");
  tc_cloud* c = new tc_cloud();
  c->ctx = ctx;
  c->n = n;
  int st = tc_alloc(ctx, &c->d_xyz, 3 * n);
  if (st == TC_OK && n > 0) {
    if (cudaMemcpyAsync(c->d_xyz, xyz_aos, 3 * n * sizeof(float), cudaMemcpyHostToDevice,
                        ctx->stream) != cudaSuccess)
      st = tc_fail(ctx, TC_GPU, "cloud upload failed");
  }
  if (st != TC_OK) {
    tc_cloud_free(c);
    return st;
  }
  *out = c;
  return TC_OK;
}

extern "C" int tc_cloud_upload_strided(tc_context* ctx, const void* base, uint64_t n,
                                       uint32_t stride_bytes, tc_cloud** out) {
  TC_ENTER(ctx);
  if (!out || (n > 0 && !base)) return TC_INVALID_DATA;
  if (stride_bytes < 12 || (stride_bytes & 3u))
    return tc_fail(ctx, TC_INVALID_DATA, "stride must be a multiple of 4 and at least 12 bytes");
  if (stride_bytes == 12) return tc_cloud_upload(ctx, (const float*)base, n, out);
  *out = nullptr;
  if (n >= 0xFFFFFFFFull) return tc_fail(ctx, TC_INVALID_DATA, "cloud too large (N < 2^32-1)");
  tc_cloud* c = new tc_cloud();
  c->ctx = ctx;
  c->n = n;
  uint8_t* d_raw = nullptr;
  int st = tc_alloc(ctx, &c->d_xyz, 3 * n);
  if (st == TC_OK) st = tc_alloc(ctx, &d_raw, n * stride_bytes);
  if (st == TC_OK && n > 0) {
    // raw records go up in one copy and are de-interleaved on the device
    if (cudaMemcpyAsync(d_raw, base, n * stride_bytes, cudaMemcpyHostToDevice, ctx->stream) !=
        cudaSuccess)
      st = tc_fail(ctx, TC_GPU, "cloud upload failed");
    if (st == TC_OK) {
      const int blocks = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8);
      k_destride<<<blocks, 256, 0, ctx->stream>>>(d_raw, n, stride_bytes, c->d_xyz);
      ctx->launches++;
    }
  }
  tc_free(ctx, d_raw);
  if (st != TC_OK) {
    tc_cloud_free(c);
    return st;
  }
  *out = c;
  return TC_OK;
}

extern "C" int tc_cloud_from_device(tc_context* ctx, const float* d_xyz_aos, uint64_t n,
                                    tc_cloud** out) {
  TC_ENTER(ctx);
  if (!out || (n > 0 && !d_xyz_aos)) return TC_INVALID_DATA;
  *out = nullptr;
  if (n >= 0xFFFFFFFFull) return tc_fail(ctx, TC_INVALID_DATA, "cloud too large (N < 2^32-1)");
  tc_cloud* c = new tc_cloud();
  c->ctx = ctx;
  c->n = n;
  int st = tc_alloc(ctx, &c->d_xyz, 3 * n);
  if (st == TC_OK && n > 0 &&
      cudaMemcpyAsync(c->d_xyz, d_xyz_aos, 3 * n * sizeof(float), cudaMemcpyDeviceToDevice,
                      ctx->stream) != cudaSuccess)
    st = tc_fail(ctx, TC_GPU, "device copy failed");
  if (st != TC_OK) {
    tc_cloud_free(c);
    return st;
  }
  *out = c;
  return TC_OK;
}

extern "C" int tc_cloud_download(tc_context* ctx, const tc_cloud* cloud, float* xyz_aos_out) {
  TC_ENTER(ctx);
  if (!cloud || (cloud->n > 0 && !xyz_aos_out)) return TC_INVALID_DATA;
  if (cloud->n > 0)
    TC_CUDA(ctx, cudaMemcpyAsync(xyz_aos_out, cloud->d_xyz, 3 * cloud->n * sizeof(float),
                                 cudaMemcpyDeviceToHost, ctx->stream));
  TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return TC_OK;
}

extern "C" void tc_cloud_free(tc_cloud* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  tc_free(c->ctx, c->d_xyz);
  delete c;
}
extern "C" uint64_t tc_cloud_len(const tc_cloud* c) { return c ? c->n : 0; }

// ---------------------------------------------------------------------------------------- kNN
extern "C" int tc_knn_device(tc_context* ctx, const tc_index* ix, const float* d_queries_aos,
                             uint64_t nq, uint32_t k, int exclude_self, uint32_t* d_idx_out,
                             float* d_dist_out, uint32_t* d_count_out) {
  TC_ENTER(ctx);
  if (!ix) return TC_INVALID_DATA;
  if (ix->sharded)
    return tc_fail(ctx, TC_INVALID_DATA, "a slab-sharded index serves tc_estimate_normals_device only");
  const bool self_query = (d_queries_aos == nullptr);
  if (self_query && nq != ix->n)
    return tc_fail(ctx, TC_INVALID_DATA, "self query: nq must equal the indexed cloud's length");
  if (nq == 0) return TC_OK;
  if (k == 0 || ix->n == 0) {  // nearest_neighbor.rs:178-180
    if (d_count_out) TC_CUDA(ctx, cudaMemsetAsync(d_count_out, 0, nq * sizeof(uint32_t), ctx->stream));
    return TC_OK;
  }
  if (!d_idx_out) return TC_INVALID_DATA;
  if (self_query)
    return tci_knn_launch(ctx, ix, nullptr, 0, nq, k, exclude_self, true, d_idx_out, d_dist_out,
                          d_count_out);
  // external queries: validate + sort by the index's grid for warp coherence
  float mn[3], mx[3];
  TC_TRY(tci_bbox(ctx, d_queries_aos, nq, mn, mx));
  float4* d_sorted = nullptr;
  TC_TRY(tci_sort_by_grid(ctx, d_queries_aos, nq, ix->lv[0].g, &d_sorted));
  const int st = tci_knn_launch(ctx, ix, d_sorted, 0, nq, k, 0, false, d_idx_out, d_dist_out,
                                d_count_out);
  tc_free(ctx, d_sorted);
  return st;
}

extern "C" int tc_knn(tc_context* ctx, const tc_index* ix, const float* queries_aos, uint64_t nq,
                      uint32_t k, int exclude_self, uint32_t* idx_out, float* dist_out,
                      uint32_t* count_out) {
  TC_ENTER(ctx);
  if (!ix) return TC_INVALID_DATA;
  if (nq == 0) return TC_OK;
  if (k == 0 || ix->n == 0) {
    if (count_out) memset(count_out, 0, nq * sizeof(uint32_t));
    return TC_OK;
  }
  if (!idx_out) return TC_INVALID_DATA;
  float* d_q = nullptr;
  uint32_t *d_idx = nullptr, *d_cnt = nullptr;
  float* d_dist = nullptr;
  int st = TC_OK;
  if (queries_aos) {
    st = tc_alloc(ctx, &d_q, 3 * nq);
    if (st == TC_OK &&
        cudaMemcpyAsync(d_q, queries_aos, 3 * nq * sizeof(float), cudaMemcpyHostToDevice,
                        ctx->stream) != cudaSuccess)
      st = tc_fail(ctx, TC_GPU, "query upload failed");
  }
  if (st == TC_OK) st = tc_alloc(ctx, &d_idx, nq * k);
  if (st == TC_OK && dist_out) st = tc_alloc(ctx, &d_dist, nq * k);
  if (st == TC_OK) st = tc_alloc(ctx, &d_cnt, nq);
  if (st == TC_OK) st = tc_knn_device(ctx, ix, d_q, nq, k, exclude_self, d_idx, d_dist, d_cnt);
  if (st == TC_OK) {
    cudaError_t e = cudaMemcpyAsync(idx_out, d_idx, nq * k * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && dist_out)
      e = cudaMemcpyAsync(dist_out, d_dist, nq * k * sizeof(float), cudaMemcpyDeviceToHost,
                          ctx->stream);
    if (e == cudaSuccess && count_out)
      e = cudaMemcpyAsync(count_out, d_cnt, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                          ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = tc_fail(ctx, TC_GPU, std::string("kNN: ") + cudaGetErrorString(e));
  }
  tc_free(ctx, d_q);
  tc_free(ctx, d_idx);
  tc_free(ctx, d_dist);
  tc_free(ctx, d_cnt);
  return st;
}

extern "C" int tc_radius_search(tc_context* ctx, const tc_index* ix, const float query[3],
                                float radius, uint32_t* idx_out, float* dist_out, uint64_t capacity,
                                uint64_t* n_found) {
  TC_ENTER(ctx);
  if (!ix || !query || !n_found) return TC_INVALID_DATA;
  if (ix->sharded)
    return tc_fail(ctx, TC_INVALID_DATA, "a slab-sharded index serves tc_estimate_normals_device only");
  *n_found = 0;
  if (!(radius > 0.0f) || ix->n == 0) return TC_OK;  // nearest_neighbor.rs:255-257
  for (int a = 0; a < 3; ++a)
    if (!std::isfinite(query[a])) return tc_fail(ctx, TC_INVALID_DATA, "query must be finite");
  const uint32_t cap = (uint32_t)std::min<uint64_t>(ix->n, 0xFFFFFFFEull);
  uint32_t *d_idx = nullptr, *d_cnt = nullptr;
  float* d_d2 = nullptr;
  int st = tc_alloc(ctx, &d_idx, cap);
  if (st == TC_OK) st = tc_alloc(ctx, &d_d2, cap);
  if (st == TC_OK) st = tc_alloc(ctx, &d_cnt, 1);
  std::vector<uint32_t> hi;
  std::vector<float> hd;
  uint32_t found = 0;
  if (st == TC_OK) {
    cudaMemsetAsync(d_cnt, 0, sizeof(uint32_t), ctx->stream);
    st = tci_radius_search_launch(ctx, ix, query, radius, d_idx, d_d2, cap, d_cnt);
  }
  if (st == TC_OK) {
    cudaError_t e = cudaMemcpyAsync(&found, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && found > 0) {
      hi.resize(found);
      hd.resize(found);
      e = cudaMemcpyAsync(hi.data(), d_idx, found * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(hd.data(), d_d2, found * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) st = tc_fail(ctx, TC_GPU, std::string("radius search: ") + cudaGetErrorString(e));
  }
  tc_free(ctx, d_idx);
  tc_free(ctx, d_d2);
  tc_free(ctx, d_cnt);
  if (st != TC_OK) return st;
  // order the hits ascending by (d2, index) — host-side formatting of the device result
  std::vector<uint32_t> order(found);
  for (uint32_t i = 0; i < found; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    return hd[a] < hd[b] || (hd[a] == hd[b] && hi[a] < hi[b]);
  });
  *n_found = found;
  for (uint64_t i = 0; i < found && i < capacity; ++i) {
    if (idx_out) idx_out[i] = hi[order[i]];
    if (dist_out) dist_out[i] = std::sqrt(hd[order[i]]);
  }
  return TC_OK;
}

// ------------------------------------------------------------------------------------ normals
namespace {
// default viewpoint (normals.rs:275-303), f32 arithmetic in the reference's order
void default_viewpoint(const float mn[3], const float mx[3], float vp[3]) {
  const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
  volatile float e2 = ex * ex;
  volatile float t = ey * ey;
  e2 = e2 + t;
  t = ez * ez;
  e2 = e2 + t;
  const float extent = sqrtf(e2);
  vp[0] = (mn[0] + mx[0]) / 2.0f;
  vp[1] = (mn[1] + mx[1]) / 2.0f;
  vp[2] = (mn[2] + mx[2]) / 2.0f + extent;
}
}  // namespace

extern "C" int tc_estimate_normals_device(tc_context* ctx, const tc_index* ix, uint32_t k,
                                          float radius, int consistent_orientation,
                                          const float* viewpoint3, uint64_t shard_begin,
                                          uint64_t shard_end, float* d_out_aos) {
  TC_ENTER(ctx);
  if (!ix) return TC_INVALID_DATA;
  if (ix->n == 0) return TC_OK;  // empty -> Ok(empty), checked before k (normals.rs:261-263)
  if (k < 3) return tc_fail(ctx, TC_INVALID_DATA, "k_neighbors must be at least 3");
  if (!d_out_aos) return TC_INVALID_DATA;
  if (shard_end == UINT64_MAX && shard_begin == 0 && ix->shard_world > 1 && !ix->sharded) {
    shard_begin = ix->own_lo;  // complete index from tc_index_build_sharded: the rank's share
    shard_end = ix->own_hi;
  }
  if (shard_end > ix->n) shard_end = ix->n;
  float vp[3];
  if (viewpoint3) {
    vp[0] = viewpoint3[0];
    vp[1] = viewpoint3[1];
    vp[2] = viewpoint3[2];
  } else {
    default_viewpoint(ix->bbox_min, ix->bbox_max, vp);
  }
  if (radius > 0.0f) {  // Some(radius); radius <= 0 finds nothing and is the kNN rule for everyone
    if (ix->sharded)
      return tc_fail(ctx, TC_INVALID_DATA,
                     "radius-mode normals need a complete index (shard by sorted range instead)");
    return tci_normals_radius_launch(ctx, ix, radius, k, consistent_orientation ? 1 : 0, vp,
                                     shard_begin, shard_end, d_out_aos);
  }
  if (!ix->sharded)
    return tci_normals_launch(ctx, ix, k, consistent_orientation ? 1 : 0, vp, shard_begin, shard_end,
                              d_out_aos);
  // Slab-sharded index: exactly the rank's own rows.  A search that had to look beyond the halo
  // (counted on the device) may have missed points of the unbuilt part: the rows are then redone
  // on a complete index, whose sorted range of the same planes holds the same queries.
  uint32_t* d_unsafe = ctx->d_scratch + 42;
  TC_CUDA(ctx, cudaMemsetAsync(d_unsafe, 0, sizeof(uint32_t), ctx->stream));
  TC_TRY(tci_normals_launch(ctx, ix, k, consistent_orientation ? 1 : 0, vp, ix->own_lo, ix->own_hi,
                            d_out_aos, true));
  uint32_t* h = ctx->h_scratch + 47;
  TC_CUDA(ctx, cudaMemcpyAsync(h, d_unsafe, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  TC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*h == 0) return TC_OK;
  tc_index* full = nullptr;
  TC_TRY(tci_index_build(ctx, ix->cloud, ix->k_hint, 0.0f, 0, 1, &full, &ix->lv[0].g));
  int st = TC_OK;
  if (full->n_levels != 1 || full->lv[0].n_cells != ix->lv[0].n_cells) {
    st = tc_fail(ctx, TC_GPU, "sharded normals: the complete index chose a different grid");
  } else {
    uint32_t r[2];
    cudaError_t e = cudaMemcpyAsync(&r[0], full->lv[0].d_cell_start + ix->own_cell_lo, 4,
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(&r[1], full->lv[0].d_cell_start + ix->own_cell_hi, 4,
                          cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = tc_fail(ctx, TC_GPU, "sharded normals: range readback failed");
    else
      st = tci_normals_launch(ctx, full, k, consistent_orientation ? 1 : 0, vp, r[0], r[1], d_out_aos,
                              true);
  }
  tc_index_free(full);
  return st;
}

extern "C" int tc_estimate_normals_indexed(tc_context* ctx, const tc_index* ix, uint32_t k,
                                           float radius, int consistent_orientation,
                                           const float* viewpoint3, float* out_aos) {
  TC_ENTER(ctx);
  if (!ix) return TC_INVALID_DATA;
  if (ix->n == 0) return TC_OK;
  if (k < 3) return tc_fail(ctx, TC_INVALID_DATA, "k_neighbors must be at least 3");
  if (!out_aos) return TC_INVALID_DATA;
  float* d_out = nullptr;
  TC_TRY(tc_alloc(ctx, &d_out, 6 * ix->n));
  int st = tc_estimate_normals_device(ctx, ix, k, radius, consistent_orientation, viewpoint3, 0,
                                      ix->n, d_out);
  if (st == TC_OK) {
    cudaError_t e = cudaMemcpyAsync(out_aos, d_out, 6 * ix->n * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) st = tc_fail(ctx, TC_GPU, std::string("normals: ") + cudaGetErrorString(e));
  }
  tc_free(ctx, d_out);
  return st;
}

extern "C" int tc_estimate_normals(tc_context* ctx, const float* xyz_aos, uint64_t n, uint32_t k,
                                   float radius, int consistent_orientation,
                                   const float* viewpoint3, float* out_aos) {
  TC_ENTER(ctx);
  if (n == 0) return TC_OK;  // normals.rs:261-263
  if (k < 3) return tc_fail(ctx, TC_INVALID_DATA, "k_neighbors must be at least 3");
  if (!xyz_aos || !out_aos) return TC_INVALID_DATA;
  tc_cloud* cloud = nullptr;
  tc_index* ix = nullptr;
  int st = tc_cloud_upload(ctx, xyz_aos, n, &cloud);
  if (st == TC_OK) st = tc_index_build(ctx, cloud, k, 0.0f, &ix);
  if (st == TC_OK)
    st = tc_estimate_normals_indexed(ctx, ix, k, radius, consistent_orientation, viewpoint3,
                                     out_aos);
  tc_index_free(ix);
  tc_cloud_free(cloud);
  return st;
}

// ---------------------------------------------------------------------------------------- ICP
extern "C" int tc_icp_point_to_plane(tc_context* ctx, const float* src_aos, uint64_t ns,
                                     const float* tgt_aos, uint64_t nt,
                                     const float* tgt_normals_aos, uint64_t n_normals,
                                     const float init[7], uint32_t max_iters, float max_corr_dist,
                                     float conv_threshold, tc_icp_result* out,
                                     uint64_t* pairs_out) {
  TC_ENTER(ctx);
  // validation order of registration.rs:517-531
  if (ns == 0 || nt == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "Source or target point cloud is empty");
  if (n_normals != nt)
    return tc_fail(ctx, TC_INVALID_DATA,
                   "target_normals length must equal the number of target points");
  if (max_iters == 0) return tc_fail(ctx, TC_INVALID_DATA, "Max iterations must be positive");
  if (!src_aos || !tgt_aos || !tgt_normals_aos || !init || !out) return TC_INVALID_DATA;
  tc_cloud *src = nullptr, *tgt = nullptr;
  tc_index* ix = nullptr;
  float* d_nrm = nullptr;
  uint32_t* d_match = nullptr;
  int st = tc_cloud_upload(ctx, src_aos, ns, &src);
  if (st == TC_OK) st = tc_cloud_upload(ctx, tgt_aos, nt, &tgt);
  if (st == TC_OK) st = tc_alloc(ctx, &d_nrm, 3 * nt);
  if (st == TC_OK &&
      cudaMemcpyAsync(d_nrm, tgt_normals_aos, 3 * nt * sizeof(float), cudaMemcpyHostToDevice,
                      ctx->stream) != cudaSuccess)
    st = tc_fail(ctx, TC_GPU, "normals upload failed");
  if (st == TC_OK) st = tc_index_build(ctx, tgt, 1, 0.0f, &ix);  // KdTree::new(target), :536
  if (st == TC_OK && pairs_out) st = tc_alloc(ctx, &d_match, ns);
  if (st == TC_OK)
    st = tc_icp_point_to_plane_device(ctx, nullptr, src, ix, d_nrm, init, max_iters, max_corr_dist,
                                      conv_threshold, out, d_match);
  if (st == TC_OK && pairs_out) {
    std::vector<uint32_t> match(ns);
    cudaError_t e = cudaMemcpyAsync(match.data(), d_match, ns * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      st = tc_fail(ctx, TC_GPU, std::string("ICP pairs: ") + cudaGetErrorString(e));
    } else {
      uint64_t c = 0;
      for (uint64_t i = 0; i < ns; ++i)
        if (match[i] != TC_NO_INDEX) {
          pairs_out[2 * c] = i;
          pairs_out[2 * c + 1] = match[i];
          ++c;
        }
      out->n_correspondences = c;
    }
  }
  tc_free(ctx, d_match);
  tc_free(ctx, d_nrm);
  tc_index_free(ix);
  tc_cloud_free(tgt);
  tc_cloud_free(src);
  return st;
}

// icp_detailed on device-resident clouds: KdTree::new(target) (registration.rs:281), the device
// loop, and optionally the (source index, target index) pairs of the final iteration
static int p2p_on_clouds(tc_context* ctx, const tc_cloud* src, const tc_cloud* tgt,
                         const float init[7], uint32_t max_iters, float max_corr_dist,
                         float conv_threshold, tc_icp_result* out, uint64_t* pairs_out) {
  const uint64_t ns = src->n;
  tc_index* ix = nullptr;
  uint32_t* d_match = nullptr;
  int st = tc_index_build(ctx, tgt, 1, 0.0f, &ix);
  if (st == TC_OK && pairs_out) st = tc_alloc(ctx, &d_match, ns);
  if (st == TC_OK)
    st = tc_icp_point_to_point_device(ctx, nullptr, src, ix, init, max_iters, max_corr_dist,
                                      conv_threshold, out, d_match);
  if (st == TC_OK && pairs_out) {
    std::vector<uint32_t> match(ns);
    cudaError_t e = cudaMemcpyAsync(match.data(), d_match, ns * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      st = tc_fail(ctx, TC_GPU, std::string("ICP pairs: ") + cudaGetErrorString(e));
    } else {
      uint64_t c = 0;
      for (uint64_t i = 0; i < ns; ++i)
        if (match[i] != TC_NO_INDEX) {
          pairs_out[2 * c] = i;
          pairs_out[2 * c + 1] = match[i];
          ++c;
        }
      out->n_correspondences = c;
    }
  }
  tc_free(ctx, d_match);
  tc_index_free(ix);
  return st;
}

extern "C" int tc_icp_point_to_point(tc_context* ctx, const float* src_aos, uint64_t ns,
                                     const float* tgt_aos, uint64_t nt, const float init[7],
                                     uint32_t max_iters, float max_corr_dist, float conv_threshold,
                                     tc_icp_result* out, uint64_t* pairs_out) {
  TC_ENTER(ctx);
  // validation order of registration.rs:266-276
  if (ns == 0 || nt == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "Source or target point cloud is empty");
  if (max_iters == 0) return tc_fail(ctx, TC_INVALID_DATA, "Max iterations must be positive");
  if (!src_aos || !tgt_aos || !init || !out) return TC_INVALID_DATA;
  tc_cloud *src = nullptr, *tgt = nullptr;
  int st = tc_cloud_upload(ctx, src_aos, ns, &src);
  if (st == TC_OK) st = tc_cloud_upload(ctx, tgt_aos, nt, &tgt);
  if (st == TC_OK)
    st = p2p_on_clouds(ctx, src, tgt, init, max_iters, max_corr_dist, conv_threshold, out,
                       pairs_out);
  tc_cloud_free(tgt);
  tc_cloud_free(src);
  return st;
}

// multiscale_icp_point_to_point (registration.rs:704-789): per level, voxel-downsample both
// clouds and run icp_point_to_point from the previous level's transform; then refine on the
// full clouds.  Everything between the two uploads and the final read-back stays on the device.
extern "C" int tc_multiscale_icp_point_to_point(
    tc_context* ctx, const float* src_aos, uint64_t ns, const float* tgt_aos, uint64_t nt,
    const float init[7], const tc_icp_scale_level* levels, uint32_t n_levels,
    uint32_t final_refinement_iterations, float final_max_corr_dist, float conv_threshold,
    tc_icp_result* out, uint64_t* pairs_out) {
  TC_ENTER(ctx);
  // validation order of registration.rs:710-734
  if (ns == 0 || nt == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "Source or target point cloud is empty");
  if (n_levels == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "At least one ICP scale level is required");
  if (conv_threshold <= 0.0f)
    return tc_fail(ctx, TC_INVALID_DATA, "Convergence threshold must be positive");
  if (final_refinement_iterations == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "Final refinement iterations must be positive");
  if (!src_aos || !tgt_aos || !init || !levels || !out) return TC_INVALID_DATA;
  tc_cloud *src = nullptr, *tgt = nullptr;
  int st = tc_cloud_upload(ctx, src_aos, ns, &src);
  if (st == TC_OK) st = tc_cloud_upload(ctx, tgt_aos, nt, &tgt);
  float T[7];
  for (int i = 0; i < 7; ++i) T[i] = init[i];
  uint32_t total_iterations = 0;
  bool any = false;
  for (uint32_t l = 0; l < n_levels && st == TC_OK; ++l) {
    const tc_icp_scale_level& lv = levels[l];
    if (lv.voxel_size <= 0.0f) {
      st = tc_fail(ctx, TC_INVALID_DATA, "Scale voxel_size must be positive");
      break;
    }
    if (lv.max_iterations == 0) {
      st = tc_fail(ctx, TC_INVALID_DATA, "Scale max_iterations must be positive");
      break;
    }
    tc_cloud *sd = nullptr, *td = nullptr;
    st = tc_voxel_grid_filter(ctx, src, lv.voxel_size, &sd);
    if (st == TC_OK) st = tc_voxel_grid_filter(ctx, tgt, lv.voxel_size, &td);
    if (st == TC_OK && sd->n >= 3 && td->n >= 3) {
      tc_icp_result r{};
      st = p2p_on_clouds(ctx, sd, td, T, lv.max_iterations, lv.max_correspondence_distance,
                         conv_threshold, &r, nullptr);
      if (st == TC_OK) {
        for (int i = 0; i < 7; ++i) T[i] = r.transform[i];
        total_iterations += r.iterations;
        any = true;
      }
    }
    tc_cloud_free(sd);
    tc_cloud_free(td);
  }
  if (st == TC_OK && !any)
    st = tc_fail(ctx, TC_ALGORITHM, "No multiscale ICP level had enough downsampled points");
  if (st == TC_OK)
    st = p2p_on_clouds(ctx, src, tgt, T, final_refinement_iterations, final_max_corr_dist,
                       conv_threshold, out, pairs_out);
  if (st == TC_OK) out->iterations += total_iterations;
  tc_cloud_free(tgt);
  tc_cloud_free(src);
  return st;
}

// gicp (gicp.rs:117-312): validation in the reference's order, per-point covariances of both
// clouds (kNN(k) incl. self), then the Gauss-Newton loop on the device.
extern "C" int tc_gicp(tc_context* ctx, const float* src_aos, uint64_t ns, const float* tgt_aos,
                       uint64_t nt, const float init[7], uint32_t max_iterations,
                       float max_correspondence_distance, float convergence_threshold,
                       uint32_t k_correspondences, tc_icp_result* out, uint64_t* pairs_out) {
  TC_ENTER(ctx);
  if (ns == 0 || nt == 0)
    return tc_fail(ctx, TC_INVALID_DATA, "GICP: source or target point cloud is empty");
  if (max_iterations == 0) return tc_fail(ctx, TC_INVALID_DATA, "GICP: max_iterations must be > 0");
  const uint32_t min_k = std::max<uint32_t>(k_correspondences, 4);
  if (ns < min_k || nt < min_k)
    return tc_fail(ctx, TC_INVALID_DATA,
                   "GICP: clouds must have at least " + std::to_string(min_k) +
                       " points for reliable covariance estimation (k_correspondences=" +
                       std::to_string(k_correspondences) + "); got source=" + std::to_string(ns) +
                       ", target=" + std::to_string(nt));
  if (!src_aos || !tgt_aos || !init || !out) return TC_INVALID_DATA;
  tc_cloud *src = nullptr, *tgt = nullptr;
  tc_index* ix = nullptr;
  float4 *d_scov = nullptr, *d_tcov = nullptr;
  uint32_t* d_match = nullptr;
  int st = tc_cloud_upload(ctx, src_aos, ns, &src);
  if (st == TC_OK) st = tc_cloud_upload(ctx, tgt_aos, nt, &tgt);
  // coplanar / collinear clouds are rejected by their bounding box (gicp.rs:148-166)
  for (int which = 0; which < 2 && st == TC_OK; ++which) {
    const tc_cloud* c = which == 0 ? src : tgt;
    float mn[3], mx[3];
    st = tci_bbox(ctx, c->d_xyz, c->n, mn, mx);
    if (st != TC_OK) break;
    float min_extent = INFINITY;
    for (int a = 0; a < 3; ++a) min_extent = std::fmin(min_extent, mx[a] - mn[a]);
    if (min_extent < 1e-4f) {
      char buf[32];
      snprintf(buf, sizeof(buf), "%.2e", (double)min_extent);
      st = tc_fail(ctx, TC_INVALID_DATA,
                   std::string("GICP: ") + (which == 0 ? "source" : "target") +
                       " point cloud appears to be coplanar or collinear (smallest bounding-box "
                       "dimension = " + buf + "); GICP requires 3-D structure");
    }
  }
  // GicpConfig.max_correspondence_distance is a plain f32 (no Option): a negative value makes
  // `dist > max` true for every pair, i.e. the reference finds no correspondences (gicp.rs:207-213)
  if (st == TC_OK && max_correspondence_distance < 0.0f)
    st = tc_fail(ctx, TC_ALGORITHM, "GICP: insufficient correspondences (need >= 6)");
  if (st == TC_OK) st = tci_gicp_covariances(ctx, src, min_k, &d_scov);
  if (st == TC_OK) st = tci_gicp_covariances(ctx, tgt, min_k, &d_tcov);
  if (st == TC_OK) st = tc_index_build(ctx, tgt, 1, 0.0f, &ix);
  if (st == TC_OK && pairs_out) st = tc_alloc(ctx, &d_match, ns);
  if (st == TC_OK)
    st = tci_gicp_device(ctx, src, ix, d_scov, d_tcov, init, max_iterations,
                         max_correspondence_distance, convergence_threshold, out, d_match);
  if (st == TC_OK && pairs_out) {
    std::vector<uint32_t> match(ns);
    cudaError_t e = cudaMemcpyAsync(match.data(), d_match, ns * sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      st = tc_fail(ctx, TC_GPU, std::string("GICP pairs: ") + cudaGetErrorString(e));
    } else {
      uint64_t c = 0;
      for (uint64_t i = 0; i < ns; ++i)
        if (match[i] != TC_NO_INDEX) {
          pairs_out[2 * c] = i;
          pairs_out[2 * c + 1] = match[i];
          ++c;
        }
      out->n_correspondences = c;
    }
  }
  tc_free(ctx, d_match);
  tc_free(ctx, d_scov);
  tc_free(ctx, d_tcov);
  tc_index_free(ix);
  tc_cloud_free(tgt);
  tc_cloud_free(src);
  return st;
}
