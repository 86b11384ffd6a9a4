// model_api.cu — extern "C" face of the host model (see include/zenu_b200.h "host model API").
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "autograd.h"

using namespace zb::host;

struct zb_model {
  zb_ctx* ctx;
  std::unique_ptr<Runtime> rt;
  std::shared_ptr<Model> model;
  ParamStore params;
  Optimizer opt;
  bool opt_ready = false;
  Variable last_loss;
  // CUDA-graph replay of the train step (zb_model_set_graph): one instantiated graph per distinct call signature
  // A captured step bakes in every address and every host-side decision: the call signature below, the activation allocator's
  // blocks (Allocator::generation), the ctx scratch arena (zb_ctx::ws_generation), train / eval mode, the math mode and the
  // data-parallel world (bucket allreduces are nodes of the graph).
  struct StepSig {
    const void* x; const void* t; void* loss_dev;
    int64_t b, c, h, w;
    bool train;
    int math, world;
    bool operator==(const StepSig& o) const {
      return x == o.x && t == o.t && loss_dev == o.loss_dev && b == o.b && c == o.c && h == o.h && w == o.w && train == o.train &&
             math == o.math && world == o.world;
    }
  };
  struct StepGraph {
    StepSig sig;
    uint64_t generation;              // Allocator::generation() at capture time
    unsigned long long ws_generation; // zb_ctx::ws_generation at capture time
    unsigned long long launches;      // kernels per replay (for zb_ctx_launch_count)
    cudaGraphExec_t exec;
  };
  struct WarmUp { StepSig sig; int eager; };   // eager steps run per signature before it is captured (two: allocator + arena settle)
  bool graph_enabled = false;
  std::vector<WarmUp> warm;
  std::vector<StepGraph> graphs;
  void* pinned_loss = nullptr;  // 8 bytes of pinned host memory: the loss read-back node of the graphs
  // deferred loss read (zb_model_train_step_async / zb_model_loss_wait): two pinned slots + the events that mark them written
  void* loss_ring = nullptr;
  cudaEvent_t loss_ev[2] = {nullptr, nullptr};
  unsigned long long loss_seq = 0;   // steps enqueued through zb_model_train_step_async
  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
    warm.clear();
  }
};

#define ZB_HOST_TRY(body)                                  \
  try {                                                    \
    body;                                                  \
    return ZB_OK;                                          \
  } catch (const HostError& e) {                           \
    zb::set_last_error("%s", e.what());                    \
    return ZB_ERR_INVALID;                                 \
  } catch (const std::exception& e) {                      \
    zb::set_last_error("host error: %s", e.what());        \
    return ZB_ERR_INVALID;                                 \
  }

namespace zb { namespace host {   // accessors for checkpoint.cu
ParamStore& model_params(zb_model* m) { return m->params; }
zb_ctx* model_ctx(zb_model* m) { return m->ctx; }
int model_dtype(zb_model* m) { return m->rt->dtype; }
Optimizer& model_optimizer(zb_model* m) { return m->opt; }
void model_drop_graphs(zb_model* m) { m->drop_graphs(); }
} }

static Variable run_forward(zb_model* m, const void* x_nchw, int64_t batch, int64_t c, int64_t h, int64_t w) {
  Runtime& rt = *m->rt;
  // The batch stays NCHW (the reference's input contract): the stem conv consumes it directly when it can, otherwise
  // conv2d() converts it.  The reference computes the input gradient of every conv, the first one included
  // (conv_without_bias.rs:110-121); requires_grad on the input keeps that work in the step.
  Variable xin = Variable::leaf(rt.borrow(const_cast<void*>(x_nchw), {batch, c, h, w}), /*requires_grad=*/true);
  xin->nchw = true;
  return m->model->call(rt, xin);
}

extern "C" {

int zb_model_create(zb_ctx* ctx, const char* arch, int dtype, int num_classes, int fused, uint64_t seed, int64_t bucket_bytes,
                    zb_model** out) {
  ZB_API_RANGE();
  ZB_REQUIRE(ctx && arch && out, "zb_model_create: NULL argument");
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "zb_model_create: unknown dtype %d", dtype);
  ZB_HOST_TRY({
    auto m = std::make_unique<zb_model>();
    m->ctx = ctx;
    m->rt = std::make_unique<Runtime>(ctx);
    m->rt->dtype = dtype;
    m->model = make_model(arch, num_classes, fused != 0);
    m->params.build(*m->rt, *m->model, seed, bucket_bytes > 0 ? bucket_bytes : (25ll << 20));
    *out = m.release();
  });
}

int zb_model_destroy(zb_model* m) {
  ZB_API_RANGE();
  if (!m) return ZB_OK;
  if (m->ctx->side) cudaStreamSynchronize(m->ctx->side->stream);
  cudaStreamSynchronize(m->ctx->stream);
  cudaStreamSynchronize(m->ctx->comm_stream);
  m->rt->side_hold.clear();
  m->rt->side_pending = false;
  if (m->last_loss.defined()) m->last_loss.clear_grad();
  m->drop_graphs();
  if (m->pinned_loss) cudaFreeHost(m->pinned_loss);
  if (m->loss_ring) cudaFreeHost(m->loss_ring);
  for (auto& e : m->loss_ev)
    if (e) cudaEventDestroy(e);
  delete m;
  return ZB_OK;
}

int zb_model_param_count(zb_model* m) { return static_cast<int>(m->params.entries.size()); }

int zb_model_param_info(zb_model* m, int index, char* name, int name_cap, int64_t* shape, int* ndim, int* kind, void** data,
                        void** grad) {
  ZB_API_RANGE();
  ZB_REQUIRE(index >= 0 && index < static_cast<int>(m->params.entries.size()), "param index out of range");
  const ParamEntry& e = m->params.entries[index];
  if (name && name_cap > 0) {
    strncpy(name, e.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  const auto& s = e.var->data.shape;
  if (ndim) *ndim = static_cast<int>(s.size());
  if (shape) for (size_t i = 0; i < s.size() && i < 4; ++i) shape[i] = s[i];
  if (kind) *kind = e.kind;
  if (data) *data = e.var->data.ptr;
  if (grad) *grad = e.var->grad_slot.ptr;
  return ZB_OK;
}

int zb_model_set_train(zb_model* m, int train) {
  ZB_API_RANGE();
  m->rt->train = train != 0;
  return ZB_OK;
}

int zb_model_set_optimizer(zb_model* m, int kind, double lr, double beta1, double beta2, double eps, double weight_decay) {
  ZB_API_RANGE();
  ZB_REQUIRE(kind >= 0 && kind <= 2, "unknown optimizer kind %d", kind);
  cudaStreamSynchronize(m->ctx->stream);
  m->drop_graphs();   // learning rate, betas ... are kernel arguments baked into captured steps
  ZB_HOST_TRY({
    m->opt.kind = kind;
    m->opt.lr = lr;
    m->opt.beta1 = beta1;
    m->opt.beta2 = beta2;
    m->opt.eps = eps;
    m->opt.weight_decay = weight_decay;
    m->opt.init(*m->rt, m->params);
    m->opt_ready = true;
  });
}

int zb_model_forward(zb_model* m, const void* x_nchw, int64_t batch, int64_t c, int64_t h, int64_t w, void* logits_out) {
  ZB_API_RANGE();
  ZB_HOST_TRY({
    Variable logits = run_forward(m, x_nchw, batch, c, h, w);
    check_rc(zb_copy(m->ctx, m->rt->dtype, logits->data.ptr, logits_out, logits->data.numel()), "copy logits");
    logits.clear_grad();
  });
}

int zb_model_forward_backward(zb_model* m, const void* x_nchw, const void* targets, int64_t batch, int64_t c, int64_t h,
                              int64_t w, void* loss_dev) {
  ZB_API_RANGE();
  ZB_HOST_TRY({
    Runtime& rt = *m->rt;
    rt.join_side();   // (only pending when an earlier step threw half way through its backward)
    if (m->last_loss.defined()) { m->last_loss.clear_grad(); m->last_loss = Variable(); }
    for (auto& e : m->params.entries) e.var->grad = Tensor();  // loss.clear_grad() of the previous step
    m->params.reset_pending();
    Variable logits = run_forward(m, x_nchw, batch, c, h, w);
    Tensor t = rt.borrow(const_cast<void*>(targets), {batch, logits.shape()[1]});
    Variable loss = softmax_cross_entropy(rt, logits, t);
    ParamStore& ps = m->params;
    zb_ctx* ctx = m->ctx;
    const size_t esz = rt.dtype == ZB_F64 ? 8 : 4;
    // hook: called with -1-bucket each time a parameter of that bucket received its gradient
    loss.backward(rt, [&](int code) {
      const int b = -1 - code;
      if (b < 0 || b >= static_cast<int>(ps.buckets.size())) return;
      if (--ps.buckets[b].pending == 0 && zb_dp_world(ctx) > 1) rt.join_side();   // the bucket's last wgrad may be on the side stream
      // ZENU_B200_DP_MODE (measurement only, profiles/r2_scaling.md): "skip" = no exchange at all (wrong gradients; what the step
      // costs without NCCL beside it), "tail" = every bucket exchanged after the backward pass (no overlap)
      static const int dp_mode = []() { const char* e = getenv("ZENU_B200_DP_MODE"); return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 't' ? 2 : 0)); }();
      if (ps.buckets[b].pending == 0 && zb_dp_world(ctx) > 1 && dp_mode == 0)
        check_rc(zb_dp_allreduce_sum(ctx, rt.dtype, static_cast<uint8_t*>(ps.flat_grads.ptr) + ps.buckets[b].offset * esz,
                                     ps.buckets[b].numel), "bucket allreduce");
    });
    rt.join_side();   // every gradient is complete on the compute stream from here on (optimizer, capture end)
    {
      const char* e = getenv("ZENU_B200_DP_MODE");
      if (e != nullptr && e[0] == 't' && zb_dp_world(ctx) > 1)
        for (auto& bk : ps.buckets)
          check_rc(zb_dp_allreduce_sum(ctx, rt.dtype, static_cast<uint8_t*>(ps.flat_grads.ptr) + bk.offset * esz, bk.numel), "bucket allreduce (tail)");
    }
    if (loss_dev) check_rc(zb_copy(ctx, rt.dtype, loss->data.ptr, loss_dev, 1), "copy loss");
    m->last_loss = loss;
  });
}

int zb_model_update(zb_model* m) {
  ZB_API_RANGE();
  ZB_REQUIRE(m->opt_ready, "zb_model_update: call zb_model_set_optimizer first");
  ZB_HOST_TRY({ m->opt.update(*m->rt, m->params); });
}

int zb_model_set_wgrad_overlap(zb_model* m, int enable) {
  ZB_API_RANGE();
  ZB_REQUIRE(m, "zb_model_set_wgrad_overlap: NULL model");
  ZB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  m->drop_graphs();   // the fork / join edges are part of a captured step
  m->rt->overlap_wgrad = enable != 0 && zb_ctx_side(m->ctx) != nullptr;
  return ZB_OK;
}

int zb_model_set_graph(zb_model* m, int enable) {
  ZB_API_RANGE();
  ZB_REQUIRE(m, "zb_model_set_graph: NULL model");
  ZB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  m->drop_graphs();
  m->graph_enabled = enable != 0;
  if (m->graph_enabled && !m->pinned_loss) ZB_CHECK_CUDA(cudaMallocHost(&m->pinned_loss, 8));
  return ZB_OK;
}

int zb_model_graph_count(zb_model* m) { return m ? static_cast<int>(m->graphs.size()) : 0; }

// Train step through a CUDA graph.  The tape is dynamic (rebuilt by the host every step, like the reference's), but for a fixed
// model, batch shape and buffer addresses it enqueues the same kernels with the same arguments: after two eager steps (the caching
// allocator and the scratch arena have reached their steady state, every kernel attribute is set) the step is captured once from
// the compute stream and replayed.  Data parallel: the bucket allreduces are captured with it (the comm stream forks from the
// compute stream at each bucket's ready event and joins at the optimizer's wait).  Nothing the host computes per step is a kernel
// argument: Adam's bias corrections come from a device table indexed by a device-side step counter (zb_adam_step_table).  Returns 1 when the step was run from a graph, 0 when the caller should run it eagerly, < 0 on error.
static int train_step_graph(zb_model* m, const void* x, const void* t, int64_t b, int64_t c, int64_t h, int64_t w, void* loss_dev,
                            double* host_loss) {
  zb_ctx* ctx = m->ctx;
  if (!m->graph_enabled || !m->opt_ready || m->rt->prof.enabled || zb::prof_active(ctx)) return 0;
  const size_t esz = m->rt->dtype == ZB_F64 ? 8 : 4;
  const uint64_t gen = m->rt->alloc.generation();
  const zb_model::StepSig sig{x, t, loss_dev, b, c, h, w, m->rt->train, ctx->default_math, zb_dp_world(ctx)};
  zb_model::StepGraph* hit = nullptr;
  for (auto& g : m->graphs)
    if (g.sig == sig) hit = &g;
  // any captured step whose buffers went away is unusable, and so are its siblings (they share the allocator and the arena)
  const unsigned long long ws_gen = ctx->ws_generation + (ctx->side ? ctx->side->ws_generation : 0ull);   // both arenas only ever count up
  if (!m->graphs.empty() && (m->graphs.front().generation != gen || m->graphs.front().ws_generation != ws_gen)) {
    for (auto& g : m->graphs) cudaGraphExecDestroy(g.exec);
    m->graphs.clear();
    for (auto& wu : m->warm) wu.eager = 0;   // every signature warms up again before it is re-captured
    hit = nullptr;
  }
  if (!hit) {
    zb_model::WarmUp* wu = nullptr;
    for (auto& e : m->warm)
      if (e.sig == sig) wu = &e;
    if (wu == nullptr) {
      if (m->warm.size() >= 16) m->warm.clear();
      m->warm.push_back({sig, 0});
      wu = &m->warm.back();
    }
    if (wu->eager < 2) { ++wu->eager; return 0; }   // per signature: a new batch shape runs eagerly twice before its capture
    if (m->graphs.size() >= 4) {   // callers rotating through many buffers: start over rather than grow
      for (auto& g : m->graphs) cudaGraphExecDestroy(g.exec);
      m->graphs.clear();
    }
    const unsigned long long l0 = zb_ctx_launch_count(ctx);
    ctx->ws_grow_refused = false;
    if (ctx->side) ctx->side->ws_grow_refused = false;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();            // e.g. the ctx runs on the legacy default stream, which cannot be captured
      m->graph_enabled = false;
      return 0;
    }
    int rc = zb_model_forward_backward(m, x, t, b, c, h, w, loss_dev);
    m->opt.in_replay_capture = true;   // the step count / Adam table is advanced below, once per replay, outside the graph
    if (rc == ZB_OK) rc = zb_model_update(m);
    m->opt.in_replay_capture = false;
    if (rc == ZB_OK && cudaMemcpyAsync(m->pinned_loss, m->last_loss->data.ptr, esz, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = ZB_ERR_CUDA;
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    const unsigned long long captured = zb_ctx_launch_count(ctx) - l0;
    ctx->launches = l0 - (ctx->side ? ctx->side->launches : 0ull);   // nothing ran during capture: the replay below counts it
    if (rc != ZB_OK || ce != cudaSuccess || graph == nullptr) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      if (ctx->ws_grow_refused || (ctx->side && ctx->side->ws_grow_refused)) {   // an arena had to grow: not capturable yet.  Nothing ran; this step goes eagerly and warms up again
        ctx->ws_grow_refused = false;
        if (ctx->side) ctx->side->ws_grow_refused = false;
        m->rt->side_hold.clear();
        m->rt->side_pending = false;
        wu->eager = 1;
        return 0;
      }
      if (rc != ZB_OK) return rc < 0 ? rc : -rc;   // the step's own error (message already recorded)
      m->graph_enabled = false;                    // this step cannot be captured: stay eager from now on
      return 0;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); m->graph_enabled = false; return 0; }
    m->graphs.push_back({sig, m->rt->alloc.generation(), ctx->ws_generation + (ctx->side ? ctx->side->ws_generation : 0ull), captured, exec});
    hit = &m->graphs.back();
  }
  try {
    m->opt.begin_step(*m->rt);   // host step count; Adam: (re)fill the bias-correction window when the step leaves it
  } catch (const std::exception& e) {
    zb::set_last_error("%s", e.what());
    return -ZB_ERR_CUDA;
  }
  if (cudaGraphLaunch(hit->exec, ctx->stream) != cudaSuccess) {
    zb::set_last_error("cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -ZB_ERR_CUDA;
  }
  ctx->launches += hit->launches;
  if (host_loss) {
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      zb::set_last_error("train step graph: %s", cudaGetErrorString(cudaGetLastError()));
      return -ZB_ERR_CUDA;
    }
    *host_loss = esz == 8 ? *static_cast<double*>(m->pinned_loss) : static_cast<double>(*static_cast<float*>(m->pinned_loss));
  }
  return 1;
}

int zb_model_train_step(zb_model* m, const void* x_nchw, const void* targets, int64_t batch, int64_t c, int64_t h, int64_t w,
                        void* loss_dev, double* host_loss) {
  ZB_API_RANGE();
  const int g = train_step_graph(m, x_nchw, targets, batch, c, h, w, loss_dev, host_loss);
  if (g == 1) return ZB_OK;
  if (g < 0) return -g;
  int rc = zb_model_forward_backward(m, x_nchw, targets, batch, c, h, w, loss_dev);
  if (rc != ZB_OK) return rc;
  rc = zb_model_update(m);
  if (rc != ZB_OK) return rc;
  if (host_loss) {
    ZB_REQUIRE(m->last_loss.defined(), "no loss");
    if (m->rt->dtype == ZB_F64) {
      ZB_CHECK_CUDA(cudaMemcpyAsync(host_loss, m->last_loss->data.ptr, 8, cudaMemcpyDeviceToHost, m->ctx->stream));
      ZB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
    } else {
      float f = 0.f;
      ZB_CHECK_CUDA(cudaMemcpyAsync(&f, m->last_loss->data.ptr, 4, cudaMemcpyDeviceToHost, m->ctx->stream));
      ZB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
      *host_loss = f;
    }
  }
  return ZB_OK;
}

// The step without the host waiting for it: the loss goes to a pinned ring slot behind the step, on the compute stream, and an event
// marks the slot written.  A training loop that logs every loss then blocks on the step BEFORE the one it has just enqueued
// (zb_model_loss_wait(m, 1, ..)): the device always has the next step queued while the host prepares the one after.
int zb_model_train_step_async(zb_model* m, const void* x_nchw, const void* targets, int64_t batch, int64_t c, int64_t h, int64_t w,
                              void* loss_dev) {
  ZB_API_RANGE();
  ZB_REQUIRE(m != nullptr && loss_dev != nullptr, "train_step_async: a device loss scalar is required");
  if (!m->loss_ring) {
    ZB_CHECK_CUDA(cudaMallocHost(&m->loss_ring, 16));
    for (auto& e : m->loss_ev) ZB_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int rc = zb_model_train_step(m, x_nchw, targets, batch, c, h, w, loss_dev, nullptr);
  if (rc != ZB_OK) return rc;
  const int slot = static_cast<int>(m->loss_seq & 1ull);
  const size_t esz = m->rt->dtype == ZB_F64 ? 8 : 4;
  ZB_CHECK_CUDA(cudaMemcpyAsync(static_cast<char*>(m->loss_ring) + 8 * slot, loss_dev, esz, cudaMemcpyDeviceToHost, m->ctx->stream));
  ZB_CHECK_CUDA(cudaEventRecord(m->loss_ev[slot], m->ctx->stream));
  ++m->loss_seq;
  return ZB_OK;
}

int zb_model_loss_wait(zb_model* m, int age, double* host_loss) {
  ZB_API_RANGE();
  ZB_REQUIRE(m != nullptr && host_loss != nullptr, "loss_wait: null argument");
  ZB_REQUIRE(age == 0 || age == 1, "loss_wait: age must be 0 (the step enqueued last) or 1 (the one before it)");
  ZB_REQUIRE(m->loss_seq > static_cast<unsigned long long>(age), "loss_wait: no such step has been enqueued");
  const int slot = static_cast<int>((m->loss_seq - 1ull - static_cast<unsigned long long>(age)) & 1ull);
  ZB_CHECK_CUDA(cudaEventSynchronize(m->loss_ev[slot]));
  const char* src = static_cast<const char*>(m->loss_ring) + 8 * slot;
  *host_loss = m->rt->dtype == ZB_F64 ? *reinterpret_cast<const double*>(src) : static_cast<double>(*reinterpret_cast<const float*>(src));
  return ZB_OK;
}

int zb_model_profile_enable(zb_model* m, int enable) {
  ZB_API_RANGE();
  ZB_CHECK_CUDA(cudaStreamSynchronize(m->ctx->stream));
  m->rt->prof.clear();
  m->rt->prof.enabled = enable != 0;
  return ZB_OK;
}

int64_t zb_model_profile_dump(zb_model* m, char* buf, int64_t cap) {
  ZB_API_RANGE();
  cudaStreamSynchronize(m->ctx->stream);
  struct Agg { int64_t n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : m->rt->prof.recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
    if (!agg.count(r.key)) order.push_back(r.key);
    Agg& a = agg[r.key];
    a.n++; a.ms += t; a.flops += r.flops; a.bytes += r.bytes;
  }
  std::string out;
  char line[512];
  for (auto& k : order) {
    const Agg& a = agg[k];
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.6e\t%.6e\n", k.c_str(), static_cast<long long>(a.n), a.ms, a.flops, a.bytes);
    out += line;
  }
  if (buf && cap > 0) {
    const size_t n = std::min<size_t>(out.size(), static_cast<size_t>(cap - 1));
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return static_cast<int64_t>(out.size() + 1);
}

int64_t zb_model_bytes_reserved(zb_model* m) { return static_cast<int64_t>(m->rt->alloc.bytes_reserved()); }

}  // extern "C"
