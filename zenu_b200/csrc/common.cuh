// common.cuh — context object, status/error plumbing and small device helpers shared by all kernels.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include "../../include/zenu_b200.h"

namespace zb {

void set_last_error(const char* fmt, ...);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace zb

namespace zb { struct ProfState; }

// The context: one per process/rank, single owner (mirrors the reference's single global
// ZENU_CUDA_STATE, zenu-cuda/src/lib.rs:18-25, minus the mutex and the library handles).
struct zb_ctx {
  int device;
  int sm_count;
  int driver_version;
  cudaStream_t stream;       // compute stream (all kernels of this ctx)
  cudaStream_t comm_stream;  // NCCL stream
  bool owns_stream;
  void* ws;                  // library-owned scratch (split-K partials, BN partial sums, layout staging)
  size_t ws_bytes;
  unsigned long long ws_generation;  // bumped whenever ws is reallocated: captured step graphs bake the address in
  bool ws_grow_refused;      // a call needed more scratch while the stream was being captured (the capture must be abandoned)
  int* err_flag;             // device word: set by a kernel whose mbarrier wait timed out
  zb::EncodeTiledFn encode_tiled;
  zb::EncodeIm2colFn encode_im2col;
  int default_math;          // zb_math_mode
  double bn_eps;             // BatchNorm epsilon (reference CPU path: 1e-10, zenu-matrix/src/nn/batch_norm.rs:296)
  // data-parallel state (dp.cu)
  void* nccl_lib;
  void* nccl_comm;
  int rank, world;
  bool nccl_failed;          // an asynchronous NCCL error was seen: the communicator has been aborted
  cudaEvent_t ev_ready, ev_done;  // compute->comm and comm->compute fences
  unsigned long long launches;  // number of kernels this ctx launched (bench.py's gpu_launches)
  // side context (zb_ctx_side): own stream + own scratch arena for work that may overlap this ctx's stream; forked / joined with events
  zb_ctx* side;
  zb_ctx* parent;            // non-NULL in a side context
  cudaEvent_t ev_fork, ev_join;
  zb::ProfState* prof;          // optional per-op CUDA-event timing (zb_ctx_profile_*)
};

namespace zb {

#define ZB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      zb::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ZB_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define ZB_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      zb::set_last_error(__VA_ARGS__);       \
      return ZB_ERR_INVALID;                 \
    }                                        \
  } while (0)

#define ZB_LAUNCH_CHECK(ctx)                                                                  \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t _e = cudaPeekAtLastError();                                                   \
    if (_e != cudaSuccess) {                                                                  \
      zb::set_last_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ZB_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

// NVTX range around every C-ABI entry point (tracing the reference does not have).  NVTX v3 is header-only: without
// a profiler attached a push / pop is one predictable branch, under nsys / ncu the ABI calls show up as named ranges on the timeline.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define ZB_API_RANGE() zb::NvtxRange zb_nvtx_range_(__func__)

// Plan trace (zb_conv2d_plan_describe / zb_ctx_plan_trace): while a PlanTrace is installed on the calling thread every launcher
// appends one "kernel<variant> key=value ...;" segment per launch it plans (size-dependent values carry a '~' prefix so that a
// test can compare the VARIANT chosen for a small batch with the one chosen for the benchmarked batch).  In dry mode the host
// planners run unchanged -- same dispatch, same tensor-map encodes, same heuristics -- but nothing is launched, allocated or
// written: scratch and temporaries are fake addresses.
struct PlanTrace {
  bool dry = false;
  std::string text;
};
extern thread_local PlanTrace* tl_plan;
static inline bool plan_dry() { return tl_plan != nullptr && tl_plan->dry; }
void plan_note(const char* fmt, ...);
// every kernel launch / async memset goes through this: skipped in dry mode, counted and checked otherwise
#define ZB_KLAUNCH(ctx, ...)            \
  do {                                  \
    if (!zb::plan_dry()) {              \
      __VA_ARGS__;                      \
      ZB_LAUNCH_CHECK(ctx);             \
    }                                   \
  } while (0)

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is per kernel AND per device, so the "already done" state is kept per
// device ordinal (a process may own contexts on several GPUs); each launcher holds one function-local static SmemOptIn.
struct SmemOptIn {
  static constexpr int kMaxDevices = 64;
  std::atomic<size_t> bytes[kMaxDevices];
};
template <typename K>
static inline int smem_opt_in(zb_ctx* ctx, SmemOptIn& st, K kernel, size_t bytes) {
  const int dev = ctx->device;
  if (dev >= 0 && dev < SmemOptIn::kMaxDevices && st.bytes[dev].load(std::memory_order_relaxed) >= bytes) return ZB_OK;
  ZB_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  if (dev >= 0 && dev < SmemOptIn::kMaxDevices) st.bytes[dev].store(bytes, std::memory_order_relaxed);
  return ZB_OK;
}

// Per-op timing with CUDA events on the launching stream, grouped by kernel class (bench.py roofline):
// class 0 = tcgen05 implicit-GEMM / GEMM launches (work = algorithmic FLOPs),
// class 1 = BatchNorm ops (work = algorithmic bytes), class 2 = other elementwise (bytes).
enum ProfClass { PROF_TENSOR = 0, PROF_BN = 1, PROF_EWISE = 2, PROF_NUM = 3 };
bool prof_active(zb_ctx* ctx);   // per-op timing is recording (events are being interleaved with the launches)
void prof_begin(zb_ctx* ctx, int cls);
void prof_end(zb_ctx* ctx, int cls, double work);

int dp_poll_async_error(zb_ctx* ctx);   // dp.cu: ncclCommGetAsyncError; aborts the communicator on failure
void dp_destroy(zb_ctx* ctx);

// Grow-only scratch; stream-ordered so earlier kernels that still read the old block stay valid.
int ctx_workspace(zb_ctx* ctx, size_t bytes, void** out);

// Environment switches select code paths for A/B experiments; they are read once per process, never per call.
#define ZB_ENV_FLAG(name) ([]() -> bool { static const bool v = getenv(name) != nullptr; return v; }())

static inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace zb
