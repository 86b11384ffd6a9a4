// input_stage.cu — library-owned input staging: pinned, multi-buffered uint8 batches + int32 labels on the host, an
// asynchronous host->device copy on a dedicated copy stream, and the on-device expansion into the model's NCHW float batch and
// one-hot targets on the compute stream.
//
// Replaces the reference's per-sample Vec<Variable> + CPU concat (zenu/src/dataset.rs:74-100) followed by a synchronous f32
// cudaMemcpy of the whole batch (zenu-matrix/src/matrix.rs:139-160,486; zenu-cuda/src/runtime/mod.rs:99-103): the batch crosses
// PCIe as bytes (a quarter of the f32 volume), the copy of batch i+1 overlaps the step of batch i, and nothing synchronises the host.
//
// Slot protocol (slot s of `slots` >= 2):
//   host_buffers(s)  pinned pointers the decoder fills (host_sync(s) first when the slot was submitted before: its copy must be done)
//   submit(s)        copy stream: waits until the previous batch of this slot was expanded, H2D images + labels, records `ready`
//   wait(s)          compute stream: waits for `ready`, expands u8 -> float NCHW (normalised) and labels -> one-hot, records `consumed`,
//                    returns the device pointers (stable per slot: step graphs keyed on them are captured once per slot).  The float
//                    batch of slot s is rewritten by the NEXT wait(s), which the compute stream orders after every step that read it.
#include <vector>

#include "common.cuh"

struct zb_input_stage {
  zb_ctx* ctx;
  int dtype, src_layout, slots;
  int64_t n, c, h, w, classes;
  bool has_norm;
  double mean[8], stdv[8];
  cudaStream_t copy_stream;
  struct Slot {
    void* host_img = nullptr;   // pinned [n*c*h*w] uint8
    void* host_lab = nullptr;   // pinned [n] int32
    void* dev_img = nullptr;
    void* dev_lab = nullptr;
    void* dev_x = nullptr;      // [n][c][h][w] dtype
    void* dev_t = nullptr;      // [n][classes] dtype
    cudaEvent_t ready = nullptr, consumed = nullptr;
    bool submitted = false, ever_consumed = false;
  };
  std::vector<Slot> slot;
};

extern "C" {

int zb_input_stage_destroy(zb_input_stage* st) {
  ZB_API_RANGE();
  if (!st) return ZB_OK;
  cudaStreamSynchronize(st->copy_stream);
  cudaStreamSynchronize(st->ctx->stream);
  for (auto& s : st->slot) {
    if (s.host_img) cudaFreeHost(s.host_img);
    if (s.host_lab) cudaFreeHost(s.host_lab);
    if (s.dev_img) cudaFree(s.dev_img);
    if (s.dev_lab) cudaFree(s.dev_lab);
    if (s.dev_x) cudaFree(s.dev_x);
    if (s.dev_t) cudaFree(s.dev_t);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.consumed) cudaEventDestroy(s.consumed);
  }
  cudaStreamDestroy(st->copy_stream);
  delete st;
  return ZB_OK;
}

int zb_input_stage_create(zb_ctx* ctx, int dtype, int src_layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t classes,
                          const double* host_mean, const double* host_std, int slots, zb_input_stage** out) {
  ZB_API_RANGE();
  ZB_REQUIRE(ctx != nullptr && out != nullptr, "input stage: NULL argument");
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "input stage: unknown dtype %d", dtype);
  ZB_REQUIRE(src_layout == ZB_NCHW || src_layout == ZB_NHWC, "input stage: unknown source layout %d", src_layout);
  ZB_REQUIRE(n > 0 && c >= 1 && c <= 8 && h > 0 && w > 0 && classes > 0, "input stage: bad batch geometry");
  ZB_REQUIRE(slots >= 2 && slots <= 8, "input stage: 2..8 slots");
  zb_input_stage* st = new zb_input_stage();
  st->ctx = ctx; st->dtype = dtype; st->src_layout = src_layout; st->slots = slots;
  st->n = n; st->c = c; st->h = h; st->w = w; st->classes = classes;
  st->has_norm = host_mean != nullptr || host_std != nullptr;
  for (int i = 0; i < 8; ++i) {
    st->mean[i] = host_mean && i < c ? host_mean[i] : 0.0;
    st->stdv[i] = host_std && i < c ? host_std[i] : 1.0;
  }
  st->copy_stream = nullptr;
  st->slot.resize(slots);
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  const size_t img_bytes = static_cast<size_t>(n * c * h * w), lab_bytes = static_cast<size_t>(n) * 4;
  cudaError_t e = cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < slots && e == cudaSuccess; ++i) {
    auto& s = st->slot[i];
    if ((e = cudaMallocHost(&s.host_img, img_bytes)) != cudaSuccess) break;
    if ((e = cudaMallocHost(&s.host_lab, lab_bytes)) != cudaSuccess) break;
    if ((e = cudaMalloc(&s.dev_img, img_bytes)) != cudaSuccess) break;
    if ((e = cudaMalloc(&s.dev_lab, lab_bytes)) != cudaSuccess) break;
    if ((e = cudaMalloc(&s.dev_x, img_bytes * esz)) != cudaSuccess) break;
    if ((e = cudaMalloc(&s.dev_t, static_cast<size_t>(n * classes) * esz)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming)) != cudaSuccess) break;
    if ((e = cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming)) != cudaSuccess) break;
  }
  if (e != cudaSuccess) {
    zb::set_last_error("input stage: %s", cudaGetErrorString(e));
    zb_input_stage_destroy(st);
    return ZB_ERR_CUDA;
  }
  *out = st;
  return ZB_OK;
}

int64_t zb_input_stage_h2d_bytes(const zb_input_stage* st) { return st ? st->n * st->c * st->h * st->w + st->n * 4 : 0; }

int zb_input_stage_host_buffers(zb_input_stage* st, int slot, void** images_u8, void** labels_i32) {
  ZB_API_RANGE();
  ZB_REQUIRE(st != nullptr && slot >= 0 && slot < st->slots, "input stage: slot out of range");
  if (images_u8) *images_u8 = st->slot[slot].host_img;
  if (labels_i32) *labels_i32 = st->slot[slot].host_lab;
  return ZB_OK;
}

int zb_input_stage_host_sync(zb_input_stage* st, int slot) {
  ZB_API_RANGE();
  ZB_REQUIRE(st != nullptr && slot >= 0 && slot < st->slots, "input stage: slot out of range");
  auto& s = st->slot[slot];
  if (s.submitted || s.ever_consumed) ZB_CHECK_CUDA(cudaEventSynchronize(s.ready));   // the H2D copy has left the pinned buffers
  return ZB_OK;
}

int zb_input_stage_submit(zb_input_stage* st, int slot) {
  ZB_API_RANGE();
  ZB_REQUIRE(st != nullptr && slot >= 0 && slot < st->slots, "input stage: slot out of range");
  auto& s = st->slot[slot];
  // the previous batch staged in this slot must have been expanded before its device staging is overwritten
  if (s.ever_consumed) ZB_CHECK_CUDA(cudaStreamWaitEvent(st->copy_stream, s.consumed, 0));
  ZB_CHECK_CUDA(cudaMemcpyAsync(s.dev_img, s.host_img, static_cast<size_t>(st->n * st->c * st->h * st->w), cudaMemcpyHostToDevice, st->copy_stream));
  ZB_CHECK_CUDA(cudaMemcpyAsync(s.dev_lab, s.host_lab, static_cast<size_t>(st->n) * 4, cudaMemcpyHostToDevice, st->copy_stream));
  ZB_CHECK_CUDA(cudaEventRecord(s.ready, st->copy_stream));
  s.submitted = true;
  return ZB_OK;
}

int zb_input_stage_wait(zb_input_stage* st, int slot, void** x_nchw, void** targets_onehot) {
  ZB_API_RANGE();
  ZB_REQUIRE(st != nullptr && slot >= 0 && slot < st->slots, "input stage: slot out of range");
  auto& s = st->slot[slot];
  ZB_REQUIRE(s.submitted, "input stage: wait on slot %d without a submit", slot);
  ZB_CHECK_CUDA(cudaStreamWaitEvent(st->ctx->stream, s.ready, 0));
  int rc = zb_input_u8_to_float(st->ctx, st->dtype, st->src_layout, s.dev_img, s.dev_x, st->n, st->c, st->h, st->w,
                                st->has_norm ? st->mean : nullptr, st->has_norm ? st->stdv : nullptr);
  if (rc != ZB_OK) return rc;
  rc = zb_onehot(st->ctx, st->dtype, s.dev_lab, s.dev_t, st->n, st->classes);
  if (rc != ZB_OK) return rc;
  // the expansion has consumed the device staging: the slot may be refilled while the step still reads dev_x / dev_t ...
  ZB_CHECK_CUDA(cudaEventRecord(s.consumed, st->ctx->stream));
  s.ever_consumed = true;
  s.submitted = false;
  if (x_nchw) *x_nchw = s.dev_x;
  if (targets_onehot) *targets_onehot = s.dev_t;
  return ZB_OK;
}

}  // extern "C"
