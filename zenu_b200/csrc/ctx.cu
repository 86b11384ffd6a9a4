// ctx.cu — context lifecycle, scratch arena, error reporting.
// Replaces the reference's process-global ZenuCudaState (zenu-cuda/src/lib.rs:18-112) and its
// cudaMallocAsync+synchronize allocation path (zenu-cuda/src/runtime/mod.rs:30-60).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <vector>

#include "common.cuh"

namespace zb {

struct ProfRecord { cudaEvent_t a, b; double work; };
struct ProfState {
  bool enabled = false;
  std::vector<ProfRecord> rec[PROF_NUM];
  cudaEvent_t open[PROF_NUM] = {nullptr, nullptr, nullptr};
};

bool prof_active(zb_ctx* ctx) { return ctx->prof && ctx->prof->enabled; }
void prof_begin(zb_ctx* ctx, int cls) {
  if (!ctx->prof || !ctx->prof->enabled) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, ctx->stream);
  ctx->prof->open[cls] = e;
}
void prof_end(zb_ctx* ctx, int cls, double work) {
  if (!ctx->prof || !ctx->prof->enabled || !ctx->prof->open[cls]) return;
  cudaEvent_t e;
  cudaEventCreate(&e);
  cudaEventRecord(e, ctx->stream);
  ctx->prof->rec[cls].push_back({ctx->prof->open[cls], e, work});
  ctx->prof->open[cls] = nullptr;
}
static void prof_clear(ProfState* p) {
  for (int c = 0; c < PROF_NUM; ++c) {
    for (auto& r : p->rec[c]) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    p->rec[c].clear();
    if (p->open[c]) { cudaEventDestroy(p->open[c]); p->open[c] = nullptr; }
  }
}

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

thread_local PlanTrace* tl_plan = nullptr;
void plan_note(const char* fmt, ...) {
  if (tl_plan == nullptr) return;
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  tl_plan->text += buf;
}

int ctx_workspace(zb_ctx* ctx, size_t bytes, void** out) {
  if (plan_dry()) {   // planning only: an aligned address that is never dereferenced
    *out = reinterpret_cast<void*>(uintptr_t(3) << 40);
    return ZB_OK;
  }
  if (bytes > ctx->ws_bytes) {
    // Never grow inside a stream capture: the free / malloc would become nodes of the graph (ctx->ws a graph-owned allocation that
    // later eager kernels dereference).  The capturing caller abandons the capture and runs the step eagerly (model_api.cu).
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
      ctx->ws_grow_refused = true;
      set_last_error("scratch arena would have to grow (%zu -> %zu bytes) during stream capture", ctx->ws_bytes, bytes);
      return ZB_ERR_UNSUPPORTED;
    }
    // stream-ordered: kernels already enqueued keep using the old block until they retire
    ++ctx->ws_generation;
    if (ctx->ws) ZB_CHECK_CUDA(cudaFreeAsync(ctx->ws, ctx->stream));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = bytes + (bytes >> 2);
    want = (want + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    ZB_CHECK_CUDA(cudaMallocAsync(&ctx->ws, want, ctx->stream));
    ctx->ws_bytes = want;
  }
  *out = ctx->ws;
  return ZB_OK;
}

}  // namespace zb

extern "C" {

const char* zb_last_error(void) { return zb::g_last_error; }
const char* zb_version(void) { return "zenu_b200 0.1 (sm_100a; tcgen05 kind::tf32 + FFMA/DFMA)"; }

int64_t zb_conv_out_size(int64_t in, int64_t k, int64_t pad, int64_t stride, int64_t dil) {
  // reference: zenu-matrix/src/nn/conv/utils.rs:52-60
  return ((in + 2 * pad - dil * (k - 1) - 1) / stride) + 1;
}

int zb_ctx_create(zb_ctx** out, int device, void* stream) {
  ZB_API_RANGE();
  ZB_REQUIRE(out != nullptr, "zb_ctx_create: out is NULL");
  ZB_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ZB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    zb::set_last_error("zenu_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    return ZB_ERR_UNSUPPORTED;
  }
  zb_ctx* ctx = new zb_ctx();
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ZB_CHECK_CUDA(cudaDriverGetVersion(&ctx->driver_version));
  if (stream) {
    ctx->stream = static_cast<cudaStream_t>(stream);
    ctx->owns_stream = false;
  } else {
    ZB_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->owns_stream = true;
  }
  ZB_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  {
    // Stream-ordered temporaries (the NCHW staging copies of api.cu, the scratch arena) come from the device's default memory pool.
    // Its release threshold is 0 by default: every synchronisation hands the freed blocks back to the driver and the next call pays
    // for mapping them again (measured: 2.7 ms instead of 0.3 ms for a staged 64 -> 64 3x3 conv at batch 256).  Keep them cached,
    // like the reference's own pool (zenu-cuda/src/runtime/mod.rs:208-214: release threshold 99 % of free memory).
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool != nullptr) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  ZB_CHECK_CUDA(cudaMalloc(&ctx->err_flag, sizeof(int)));
  ZB_CHECK_CUDA(cudaMemset(ctx->err_flag, 0, sizeof(int)));
  // Driver entry points for tensor-map encoding, resolved at run time so the library does not link libcuda.
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  ZB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  ZB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
  ctx->encode_tiled = reinterpret_cast<zb::EncodeTiledFn>(fn);
  fn = nullptr;
  ZB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
  ZB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col not available in this driver");
  ctx->encode_im2col = reinterpret_cast<zb::EncodeIm2colFn>(fn);
  ctx->default_math = ZB_MATH_TF32;
  ctx->bn_eps = 1e-10;
  ctx->rank = 0;
  ctx->world = 1;
  *out = ctx;
  return ZB_OK;
}

zb_ctx* zb_ctx_side(zb_ctx* ctx) {
  if (ctx == nullptr) return nullptr;
  if (ctx->parent != nullptr) return ctx;   // a side context has no side of its own
  if (ctx->side == nullptr) {
    zb_ctx* s = nullptr;
    if (zb_ctx_create(&s, ctx->device, nullptr) != ZB_OK) return nullptr;
    s->parent = ctx;
    s->default_math = ctx->default_math;
    s->bn_eps = ctx->bn_eps;
    if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      zb::set_last_error("zb_ctx_side: cannot create the fork / join events");
      zb_ctx_destroy(s);
      return nullptr;
    }
    ctx->side = s;
  }
  ctx->side->default_math = ctx->default_math;
  return ctx->side;
}

int zb_ctx_fork(zb_ctx* ctx) {
  ZB_REQUIRE(ctx != nullptr && ctx->side != nullptr, "zb_ctx_fork: no side context (call zb_ctx_side first)");
  ZB_CHECK_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  ZB_CHECK_CUDA(cudaStreamWaitEvent(ctx->side->stream, ctx->ev_fork, 0));
  return ZB_OK;
}

int zb_ctx_join(zb_ctx* ctx) {
  ZB_REQUIRE(ctx != nullptr, "zb_ctx_join: ctx is NULL");
  if (ctx->side == nullptr) return ZB_OK;
  ZB_CHECK_CUDA(cudaEventRecord(ctx->ev_join, ctx->side->stream));
  ZB_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return ZB_OK;
}

int zb_ctx_destroy(zb_ctx* ctx) {
  ZB_API_RANGE();
  if (!ctx) return ZB_OK;
  cudaSetDevice(ctx->device);
  if (ctx->side) {
    zb_ctx_destroy(ctx->side);
    ctx->side = nullptr;
    cudaEventDestroy(ctx->ev_fork);
    cudaEventDestroy(ctx->ev_join);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->comm_stream);
  zb::dp_destroy(ctx);
  if (ctx->prof) { zb::prof_clear(ctx->prof); delete ctx->prof; }
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->err_flag) cudaFree(ctx->err_flag);
  cudaStreamDestroy(ctx->comm_stream);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return ZB_OK;
}

int zb_ctx_profile_enable(zb_ctx* ctx, int enable) {
  ZB_API_RANGE();
  if (!ctx->prof) ctx->prof = new zb::ProfState();
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  zb::prof_clear(ctx->prof);
  ctx->prof->enabled = enable != 0;
  return ZB_OK;
}

int zb_ctx_profile_read(zb_ctx* ctx, int cls, int64_t* ops, double* total_ms, double* work) {
  ZB_API_RANGE();
  ZB_REQUIRE(cls >= 0 && cls < zb::PROF_NUM, "profile class out of range");
  if (ops) *ops = 0;
  if (total_ms) *total_ms = 0.0;
  if (work) *work = 0.0;
  if (!ctx->prof) return ZB_OK;
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  double ms = 0.0, w = 0.0;
  for (auto& r : ctx->prof->rec[cls]) {
    float t = 0.f;
    ZB_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    w += r.work;
  }
  if (ops) *ops = static_cast<int64_t>(ctx->prof->rec[cls].size());
  if (total_ms) *total_ms = ms;
  if (work) *work = w;
  return ZB_OK;
}

int zb_ctx_synchronize(zb_ctx* ctx) {
  ZB_API_RANGE();
  if (ctx->side) ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->side->stream));
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return ZB_OK;
}

int zb_ctx_set_math(zb_ctx* ctx, int math_mode) {
  ZB_API_RANGE();
  ZB_REQUIRE(math_mode == ZB_MATH_TF32 || math_mode == ZB_MATH_TF32X3 || math_mode == ZB_MATH_FP32, "unknown math mode %d", math_mode);
  ctx->default_math = math_mode;
  return ZB_OK;
}

int zb_ctx_set_bn_epsilon(zb_ctx* ctx, double eps) {
  ZB_REQUIRE(ctx != nullptr && eps >= 0.0, "zb_ctx_set_bn_epsilon: bad argument");
  ctx->bn_eps = eps;
  return ZB_OK;
}
double zb_ctx_bn_epsilon(zb_ctx* ctx) { return ctx ? ctx->bn_eps : 0.0; }

void* zb_ctx_stream(zb_ctx* ctx) { return ctx->stream; }

int zb_ctx_check(zb_ctx* ctx) {
  ZB_API_RANGE();
  if (ctx->side) {   // kernels on the side stream report through their own flag
    const int rc = zb_ctx_check(ctx->side);
    if (rc != ZB_OK) return rc;
  }
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  int flag = 0;
  ZB_CHECK_CUDA(cudaMemcpy(&flag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag != 0) {
    cudaMemset(ctx->err_flag, 0, sizeof(int));
    zb::set_last_error("device-side barrier wait timed out (tcgen05 pipeline protocol error)");
    return ZB_ERR_TIMEOUT;
  }
  if (ctx->nccl_failed) { zb::set_last_error("the NCCL communicator was aborted after an asynchronous error"); return ZB_ERR_NCCL; }
  return zb::dp_poll_async_error(ctx);
}

unsigned long long zb_ctx_launch_count(zb_ctx* ctx) { return ctx->launches + (ctx->side ? ctx->side->launches : 0ull); }

}  // extern "C"
