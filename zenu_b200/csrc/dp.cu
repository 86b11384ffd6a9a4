// dp.cu — data-parallel gradient exchange: NCCL sum-allreduce over NVLink 5 / NVSwitch on a dedicated
// communication stream, fenced against the compute stream with events so that a bucket's allreduce overlaps
// the backward kernels that follow it.  New relative to the reference, which is single-GPU only
// (hook site = top of Optimizer::update, zenu-optimizer/src/sgd.rs:20-30).
// NCCL is resolved with dlopen at run time (the library already loaded by the host process wins, e.g. the
// copy bundled with PyTorch), so the C ABI itself has no link-time NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace zb {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.lib) return ZB_OK;
  const char* cands[] = {getenv("ZENU_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : cands) {
    if (!c) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_last_error("cannot dlopen libnccl (set ZENU_B200_NCCL_LIB): %s", dlerror());
    return ZB_ERR_NCCL;
  }
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.CommInitRankConfig = reinterpret_cast<decltype(g_nccl.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  g_nccl.CommAbort = reinterpret_cast<decltype(g_nccl.CommAbort)>(dlsym(h, "ncclCommAbort"));
  g_nccl.CommGetAsyncError = reinterpret_cast<decltype(g_nccl.CommGetAsyncError)>(dlsym(h, "ncclCommGetAsyncError"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_last_error("libnccl is missing required symbols");
    return ZB_ERR_NCCL;
  }
  g_nccl.lib = h;
  return ZB_OK;
}

#define ZB_CHECK_NCCL(expr)                                                                         \
  do {                                                                                              \
    ncclResult_t _r = (expr);                                                                       \
    if (_r != ncclSuccess) {                                                                        \
      zb::set_last_error("%s failed: %s", #expr, zb::g_nccl.GetErrorString ? zb::g_nccl.GetErrorString(_r) : "?"); \
      return ZB_ERR_NCCL;                                                                           \
    }                                                                                               \
  } while (0)

// Asynchronous NCCL failures (a peer died, a link went down) do not come back through the enqueue calls' return codes: they are
// polled from the communicator.  On error the communicator is aborted (so that kernels blocked in the collective are released
// instead of hanging the stream) and every later data-parallel call fails with ZB_ERR_NCCL.
int dp_poll_async_error(zb_ctx* ctx) {
  if (ctx->world <= 1 || ctx->nccl_comm == nullptr || g_nccl.CommGetAsyncError == nullptr) return ZB_OK;
  ncclResult_t async = ncclSuccess;
  const ncclResult_t r = g_nccl.CommGetAsyncError(static_cast<ncclComm_t>(ctx->nccl_comm), &async);
  if (r == ncclSuccess && (async == ncclSuccess || async == ncclInProgress)) return ZB_OK;
  const ncclResult_t bad = r != ncclSuccess ? r : async;
  set_last_error("NCCL asynchronous error on rank %d of %d: %s (communicator aborted)", ctx->rank, ctx->world,
                 g_nccl.GetErrorString ? g_nccl.GetErrorString(bad) : "?");
  if (g_nccl.CommAbort) g_nccl.CommAbort(static_cast<ncclComm_t>(ctx->nccl_comm));
  ctx->nccl_comm = nullptr;
  ctx->nccl_failed = true;
  return ZB_ERR_NCCL;
}

void dp_destroy(zb_ctx* ctx) {
  // The communicator itself is left to process teardown: ncclCommDestroy synchronises with the peers' destroy calls, and a rank
  // whose peers have already exited (the normal end of a torchrun job) would block in it.
  ctx->nccl_comm = nullptr;
  if (ctx->ev_ready) cudaEventDestroy(ctx->ev_ready);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  ctx->ev_ready = ctx->ev_done = nullptr;
}

}  // namespace zb

using namespace zb;

extern "C" {

int zb_dp_plan_buckets(const int64_t* numel, const int* kind, int n, int64_t bucket_bytes, int elem_size, int* bucket_out,
                       int64_t* offset_out, int* num_buckets, int64_t* total_elems, int64_t* buffer_elems) {
  // Pure host logic (no CUDA call): parameters are listed in forward order; gradients arrive in reverse order during
  // backward, so buckets are filled from the last parameter backwards (bucket 0 completes first).  Inside a bucket the
  // weights come before the biases (AdamW decays weights() only, zenu-optimizer/src/adamw.rs:28,61-65); every tensor
  // starts on a 16-byte boundary.  kind: 0 weight, 1 bias, 2 buffer (BN running statistics: no gradient, own area).
  ZB_REQUIRE(numel && kind && bucket_out && offset_out && n >= 0 && bucket_bytes > 0 && elem_size > 0, "zb_dp_plan_buckets: bad argument");
  auto align4 = [](int64_t v) { return (v + 3) & ~int64_t(3); };
  int nb = 0;
  int64_t acc = 0;
  for (int i = n - 1; i >= 0; --i) {
    bucket_out[i] = -1;
    if (kind[i] == 2) continue;
    if (acc > 0 && (acc + numel[i]) * elem_size > bucket_bytes) { ++nb; acc = 0; }
    bucket_out[i] = nb;
    acc += numel[i];
  }
  // The bucket that completes LAST (the front of the network: its final gradient is the stem's wgrad, the last kernel of backward)
  // has nothing left to overlap with except the optimizer: keep its exposed allreduce latency-sized by cutting a small tail bucket
  // (<= 1/16 of bucket_bytes, at most 2 MB) off its front.  The bulk of the old last bucket then starts its exchange earlier.
  {
    const int64_t tail_cap = std::min<int64_t>(bucket_bytes / 16, 2ll << 20);
    int64_t last_bytes = 0, tail_bytes = 0;
    for (int i = 0; i < n; ++i)
      if (bucket_out[i] == nb) last_bytes += numel[i] * elem_size;
    if (last_bytes > 2 * tail_cap) {
      int cut = -1;
      for (int i = 0; i < n; ++i) {
        if (bucket_out[i] != nb) continue;
        if (tail_bytes + numel[i] * elem_size > tail_cap && tail_bytes > 0) break;
        tail_bytes += numel[i] * elem_size;
        cut = i;
      }
      if (cut >= 0 && tail_bytes < last_bytes) {
        for (int i = 0; i <= cut; ++i)
          if (bucket_out[i] == nb) bucket_out[i] = nb + 1;
        ++nb;
      }
    }
  }
  const int buckets = nb + 1;
  int64_t total = 0, buf_total = 0;
  for (int b = 0; b < buckets; ++b)
    for (int k = 0; k < 2; ++k)
      for (int i = 0; i < n; ++i)
        if (bucket_out[i] == b && kind[i] == k) { offset_out[i] = total; total += align4(numel[i]); }
  for (int i = 0; i < n; ++i)
    if (kind[i] == 2) { offset_out[i] = buf_total; buf_total += align4(numel[i]); }
  if (num_buckets) *num_buckets = buckets;
  if (total_elems) *total_elems = total;
  if (buffer_elems) *buffer_elems = buf_total;
  return ZB_OK;
}

int zb_dp_unique_id(zb_ctx* ctx, void* host_id128) {
  ZB_API_RANGE();
  (void)ctx;
  int rc = load_nccl();
  if (rc != ZB_OK) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  ncclUniqueId id;
  ZB_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(host_id128, &id, sizeof(id));
  return ZB_OK;
}

int zb_dp_init(zb_ctx* ctx, const void* host_id128, int rank, int world) {
  ZB_API_RANGE();
  ZB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "zb_dp_init: bad rank/world %d/%d", rank, world);
  ctx->rank = rank;
  ctx->world = world;
  if (!ctx->ev_ready) {
    ZB_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_ready, cudaEventDisableTiming));
    ZB_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
  }
  if (world == 1) return ZB_OK;
  int rc = load_nccl();
  if (rc != ZB_OK) return rc;
  ncclUniqueId id;
  memcpy(&id, host_id128, sizeof(id));
  ncclComm_t comm;
  ZB_CHECK_CUDA(cudaSetDevice(ctx->device));
  // The compute kernels are persistent grids of one CTA per SM, so every SM a collective occupies delays a whole CTA's worth of tiles
  // of the tensor kernel beside it; ZENU_B200_NCCL_MAX_CTAS caps NCCL's CTA count per collective.  Measured at N = 2 on one box
  // (profiles/r2_scaling.md): NCCL's own choice 37.09 ms / step, cap 4 37.24, cap 16 37.12 (1 GPU: 36.56) - the cap buys nothing, so
  // the default is 0 = leave it to NCCL.
  static const int max_ctas = []() { const char* e = getenv("ZENU_B200_NCCL_MAX_CTAS"); return e ? atoi(e) : 0; }();
  if (g_nccl.CommInitRankConfig != nullptr && max_ctas > 0) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = max_ctas;
    cfg.minCTAs = 1;
    ZB_CHECK_NCCL(g_nccl.CommInitRankConfig(&comm, world, id, rank, &cfg));
  } else {
    ZB_CHECK_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  }
  ctx->nccl_comm = comm;
  return ZB_OK;
}

int zb_dp_allreduce_sum(zb_ctx* ctx, int dtype, void* buf, int64_t n) {
  ZB_API_RANGE();
  if (ctx->world <= 1 || n == 0) return ZB_OK;
  if (ctx->nccl_failed) { set_last_error("zb_dp_allreduce_sum: the NCCL communicator was aborted after an asynchronous error"); return ZB_ERR_NCCL; }
  ZB_REQUIRE(ctx->nccl_comm != nullptr, "zb_dp_allreduce_sum: zb_dp_init was not called");
  ZB_CHECK_CUDA(cudaEventRecord(ctx->ev_ready, ctx->stream));
  ZB_CHECK_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_ready, 0));
  ZB_CHECK_NCCL(g_nccl.AllReduce(buf, buf, static_cast<size_t>(n), dtype == ZB_F64 ? ncclFloat64 : ncclFloat32, ncclSum,
                                 static_cast<ncclComm_t>(ctx->nccl_comm), ctx->comm_stream));
  ZB_CHECK_CUDA(cudaEventRecord(ctx->ev_done, ctx->comm_stream));
  return ZB_OK;
}

int zb_dp_wait(zb_ctx* ctx) {
  ZB_API_RANGE();
  if (ctx->world <= 1 || !ctx->ev_done) return ZB_OK;
  if (ctx->nccl_failed) { set_last_error("zb_dp_wait: the NCCL communicator was aborted after an asynchronous error"); return ZB_ERR_NCCL; }
  ZB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_done, 0));
  // once per step (this is the optimizer's fence): has any collective enqueued so far failed asynchronously?
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) return dp_poll_async_error(ctx);
  return ZB_OK;
}

int zb_dp_check(zb_ctx* ctx) {
  ZB_API_RANGE();
  ZB_REQUIRE(ctx != nullptr, "zb_dp_check: ctx is NULL");
  if (ctx->nccl_failed) { set_last_error("the NCCL communicator was aborted after an asynchronous error"); return ZB_ERR_NCCL; }
  return dp_poll_async_error(ctx);
}

int zb_dp_rank(zb_ctx* ctx) { return ctx->rank; }
int zb_dp_world(zb_ctx* ctx) { return ctx->world; }

}  // extern "C"
