// autograd.cu — tape, differentiable functions, layers, parameter store and optimizers (see autograd.h).
#include "autograd.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>
#include <random>
#include <set>

namespace zb {
namespace host {

void check_rc(int rc, const char* what) {
  if (rc != ZB_OK) throw HostError(std::string(what) + ": " + zb_last_error());
}

// ------------------------------------------------------------------------------------------------ allocator
Allocator::~Allocator() { release_cached(); }
void* Allocator::alloc(size_t bytes) {
  bytes = std::max<size_t>((bytes + 511) & ~size_t(511), 512);
  auto it = free_.find(bytes);
  if (it != free_.end() && !it->second.empty()) {
    void* p = it->second.back();
    it->second.pop_back();
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    release_cached();
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) throw HostError(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
  }
  reserved_ += bytes;
  return p;
}
void Allocator::free(void* p, size_t bytes) {
  bytes = std::max<size_t>((bytes + 511) & ~size_t(511), 512);
  free_[bytes].push_back(p);
}
void Allocator::release_cached() {
  ++generation_;
  for (auto& kv : free_) {
    for (void* p : kv.second) { cudaFree(p); reserved_ -= kv.first; }
    kv.second.clear();
  }
}

Runtime::Runtime(zb_ctx* c) : ctx(c), alloc(c) {
  // Off by default (zb_model_set_wgrad_overlap / ZENU_B200_WGRAD_OVERLAP=1 turn it on).  Measured on ResNet-50, batch 256
  // (profiles/r2b_wgrad_overlap.md): 36.99 ms / step with the overlap, 36.68 without.  The tensor-core kernels hold ~225 registers
  // x 192 threads and up to 227 KB of shared memory per SM, so at most ONE 256-thread BatchNorm CTA fits beside a resident wgrad CTA
  // (the reduce kernels, which need 8 KB of shared memory, none): the HBM-bound kernel on the critical path loses more bandwidth
  // than the off-path wgrad gains.
  static const bool on = []() { const char* e = getenv("ZENU_B200_WGRAD_OVERLAP"); return e != nullptr && e[0] == '1'; }();
  overlap_wgrad = on && zb_ctx_side(c) != nullptr;
  lazy_mask = getenv("ZENU_B200_NO_LAZY_MASK") == nullptr;
  fuse_stem_pool = getenv("ZENU_B200_NO_STEM_POOL_FUSION") == nullptr;   // read per Runtime (not cached): tests build models both ways
}

void Runtime::join_side() {
  if (!side_pending) return;
  check_rc(zb_ctx_join(ctx), "join");
  side_hold.clear();
  side_pending = false;
}

Tensor Runtime::empty(std::vector<int64_t> shape) {
  Tensor t;
  t.shape = std::move(shape);
  t.dtype = dtype;
  auto st = std::make_shared<Storage>();
  st->alloc = &alloc;
  st->bytes = std::max<size_t>(t.bytes(), 4);
  st->ptr = alloc.alloc(st->bytes);
  t.storage = st;
  t.ptr = st->ptr;
  return t;
}
Tensor Runtime::zeros(std::vector<int64_t> shape) {
  Tensor t = empty(std::move(shape));
  check_rc(zb_fill(ctx, dtype, t.ptr, 0.0, t.numel()), "zeros");
  return t;
}
Tensor Runtime::borrow(void* p, std::vector<int64_t> shape) {
  Tensor t;
  t.shape = std::move(shape);
  t.dtype = dtype;
  t.ptr = p;
  return t;
}

// ------------------------------------------------------------------------------------------------ per-node timing
void OpProfiler::clear() {
  for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  recs.clear();
}
ProfScope::ProfScope(Runtime& r, std::string key, double flops, double bytes) : rt(r) {
  if (!rt.prof.enabled) return;
  OpProfiler::Rec rec{std::move(key), nullptr, nullptr, flops, bytes};
  cudaEventCreate(&rec.a);
  cudaEventCreate(&rec.b);
  cudaEventRecord(rec.a, rt.ctx->stream);
  idx = rt.prof.recs.size();
  rt.prof.recs.push_back(std::move(rec));
}
ProfScope::~ProfScope() {
  if (idx != static_cast<size_t>(-1)) cudaEventRecord(rt.prof.recs[idx].b, rt.ctx->stream);
}
static std::string shape_str(const std::vector<int64_t>& s) {
  std::string o;
  for (size_t i = 0; i < s.size(); ++i) o += (i ? "x" : "") + std::to_string(s[i]);
  return o;
}
static std::string conv_key(const char* pass, const zb_conv2d_desc& d) {
  return std::string("conv.") + pass + " n" + std::to_string(d.n) + " c" + std::to_string(d.c) + " hw" + std::to_string(d.h) + " k" +
         std::to_string(d.k) + " r" + std::to_string(d.kh) + " s" + std::to_string(d.stride_h);
}
static double conv_flops(const zb_conv2d_desc& d) {
  const double P = zb_conv_out_size(d.h, d.kh, d.pad_h, d.stride_h, d.dil_h), Q = zb_conv_out_size(d.w, d.kw, d.pad_w, d.stride_w, d.dil_w);
  return 2.0 * d.n * P * Q * d.k * d.c * d.kh * d.kw;
}
static double conv_bytes(const zb_conv2d_desc& d, size_t esz) {
  const double P = zb_conv_out_size(d.h, d.kh, d.pad_h, d.stride_h, d.dil_h), Q = zb_conv_out_size(d.w, d.kw, d.pad_w, d.stride_w, d.dil_w);
  return esz * (static_cast<double>(d.n) * d.h * d.w * d.c + static_cast<double>(d.k) * d.kh * d.kw * d.c + static_cast<double>(d.n) * P * Q * d.k);
}

// ------------------------------------------------------------------------------------------------ tape
Variable Variable::leaf(Tensor t, bool requires_grad, std::string name) {
  auto p = std::make_shared<VariableInner>();
  p->data = std::move(t);
  p->requires_grad = requires_grad;
  p->name = std::move(name);
  return Variable(p);
}

static Variable make_output(Tensor data, std::shared_ptr<Function> fn) {
  auto p = std::make_shared<VariableInner>();
  p->data = std::move(data);
  int gen = 0;
  bool req = false;
  for (auto& in : fn->inputs) {
    gen = std::max(gen, in->gen);
    req = req || in->requires_grad;
  }
  fn->gen = gen;
  p->gen = gen + 1;
  p->requires_grad = req;
  if (req) {
    p->creator = fn;
    fn->output = p;
  }
  return Variable(p);
}

Tensor grad_target(Runtime& rt, VariableInner& v) {
  if (v.is_param && v.grad_slot.defined() && !v.grad.defined()) return v.grad_slot;
  return rt.empty(v.data.shape);
}

void materialise_grad(Runtime& rt, VariableInner& v) {
  if (!v.grad_mask.defined()) return;
  ProfScope ps(rt, "grad.mask " + shape_str(v.data.shape), 0.0, (2.0 + 1.0 / 32.0) * v.grad.bytes());
  Tensor dst = (v.grad.storage && v.grad.storage.use_count() == 1) ? v.grad : rt.empty(v.data.shape);
  check_rc(zb_mask_apply(rt.ctx, v.grad.dtype, v.grad.ptr, v.grad_mask.ptr, dst.ptr, v.grad.numel()), "grad mask");
  v.grad = dst;
  v.grad_mask = Tensor();
}

void commit_grad_masked(Runtime& rt, VariableInner& v, const Tensor& g, const Tensor& bits) {
  if (!v.grad.defined() && !v.is_param) {
    v.grad = g;
    v.grad.shape = v.data.shape;
    v.grad_mask = bits;
    return;
  }
  Tensor m = rt.empty(v.data.shape);
  check_rc(zb_mask_apply(rt.ctx, g.dtype, g.ptr, bits.ptr, m.ptr, g.numel()), "grad mask");
  commit_grad(rt, v, m);
}

void commit_grad(Runtime& rt, VariableInner& v, const Tensor& g) {
  materialise_grad(rt, v);
  if (!v.grad.defined()) {
    if (v.is_param && v.grad_slot.defined() && g.ptr != v.grad_slot.ptr) {
      ProfScope ps(rt, "grad.copy", 0.0, 2.0 * g.bytes());
      check_rc(zb_copy(rt.ctx, g.dtype, g.ptr, v.grad_slot.ptr, g.numel()), "grad copy");
      v.grad = v.grad_slot;
    } else {
      v.grad = g;
      v.grad.shape = v.data.shape;
    }
    return;
  }
  // second arrival: grad + old (lib.rs:480-481)
  ProfScope ps(rt, "grad.accumulate " + shape_str(v.data.shape), 0.0, 3.0 * g.bytes());
  if (v.is_param || (v.grad.storage && v.grad.storage.use_count() == 1)) {
    check_rc(zb_binary(rt.ctx, g.dtype, ZB_OP_ADD, v.grad.ptr, g.ptr, v.grad.ptr, g.numel()), "grad accumulate");
  } else {
    Tensor s = rt.empty(v.data.shape);
    check_rc(zb_binary(rt.ctx, g.dtype, ZB_OP_ADD, v.grad.ptr, g.ptr, s.ptr, g.numel()), "grad accumulate");
    v.grad = s;
  }
}
void accumulate_grad(Runtime& rt, VariableInner& v, const Tensor& g) { commit_grad(rt, v, g); }

void Variable::backward(Runtime& rt, const std::function<void(int)>& on_bucket_ready) const {
  if (!p_->creator) return;
  if (!p_->grad.defined()) {
    Tensor one = rt.empty(p_->data.shape);
    check_rc(zb_fill(rt.ctx, one.dtype, one.ptr, 1.0, one.numel()), "seed grad");
    p_->grad = one;
  }
  struct Cmp {
    bool operator()(const std::shared_ptr<Function>& a, const std::shared_ptr<Function>& b) const {
      return a->gen < b->gen || (a->gen == b->gen && a->order_hint() > b->order_hint());
    }
  };
  std::priority_queue<std::shared_ptr<Function>, std::vector<std::shared_ptr<Function>>, Cmp> heap;
  std::set<Function*> seen;
  heap.push(p_->creator);
  seen.insert(p_->creator.get());
  while (!heap.empty()) {
    auto fn = heap.top();
    heap.pop();
    VarPtr out = fn->output.lock();
    if (!out || !out->grad.defined()) continue;
    if (out->grad_mask.defined() && !fn->takes_masked_grad()) materialise_grad(rt, *out);
    fn->gy_mask = out->grad_mask;
    Tensor gy = out->grad;
    fn->backward(rt, gy);
    fn->gy_mask = Tensor();
    out->grad = Tensor();  // intermediate gradients are not retained
    out->grad_mask = Tensor();
    for (auto& in : fn->inputs) {
      if (in->is_param && in->grad.defined() && in->bucket >= 0 && on_bucket_ready) on_bucket_ready(-1 - in->bucket);
      if (in->creator && !seen.count(in->creator.get())) {
        seen.insert(in->creator.get());
        heap.push(in->creator);
      }
    }
  }
}

void Variable::clear_grad() const {
  // iterative walk: drop grads and creators so activations return to the allocator
  std::vector<VarPtr> stack{p_};
  std::set<VariableInner*> seen;
  while (!stack.empty()) {
    VarPtr v = stack.back();
    stack.pop_back();
    if (!v || seen.count(v.get())) continue;
    seen.insert(v.get());
    v->grad = Tensor();
    v->grad_mask = Tensor();
    if (v->creator) {
      for (auto& in : v->creator->inputs) stack.push_back(in);
      v->creator.reset();
    }
  }
}

// ------------------------------------------------------------------------------------------------ functions
static int64_t out_size(int64_t in, int64_t k, int64_t pad, int64_t stride, int64_t dil) {
  return zb_conv_out_size(in, k, pad, stride, dil);
}

struct ConvFn : Function {
  Tensor x, w;
  zb_conv2d_desc d;
  bool has_bias, need_dx;
  int x_layout = ZB_NHWC;  // ZB_NCHW_X when the stem consumed the NCHW network input directly
  const char* name() const override { return "conv2d"; }
  int order_hint() const override {
    return (d.kh == 1 && d.kw == 1 && (d.stride_h > 1 || d.stride_w > 1) && !ZB_ENV_FLAG("ZENU_B200_NO_ORDER_HINT")) ? 1 : 0;
  }
  void backward(Runtime& rt, const Tensor& gy) override {
    VariableInner& xv = *inputs[0];
    VariableInner& wv = *inputs[1];
    // wgrad feeds nothing but the optimizer: with overlap enabled it runs on the side stream, after the dgrad below has been
    // enqueued on the main stream (which is the critical path: it is dispatched first), and overlaps the BatchNorm-backward kernels
    // of the next layer.  The previous layer's side work is joined first, so at most one wgrad is in flight and the tensors it reads
    // (held in rt.side_hold) are released before the main stream can be handed their memory again.
    const bool side_wgrad = wv.requires_grad && rt.overlap_wgrad && !rt.prof.enabled && !prof_active(rt.ctx) && x_layout == ZB_NHWC &&
                            !wv.grad.defined();
    if (rt.side_pending) rt.join_side();
    if (side_wgrad) check_rc(zb_ctx_fork(rt.ctx), "fork");   // gy (and everything before it) is complete for the side stream
    if (wv.requires_grad && !side_wgrad) {
      Tensor dw = grad_target(rt, wv);
      ProfScope ps(rt, conv_key("wgrad", d), conv_flops(d), conv_bytes(d, gy.elem_size()));
      check_rc(zb_conv2d_wgrad(rt.ctx, gy.dtype, x_layout, ZB_MATH_DEFAULT, &d, gy.ptr, x.ptr, dw.ptr), "conv wgrad");
      commit_grad(rt, wv, dw);
    }
    if (has_bias && inputs[2]->requires_grad) {
      VariableInner& bv = *inputs[2];
      Tensor db = grad_target(rt, bv);
      const int64_t P = out_size(d.h, d.kh, d.pad_h, d.stride_h, d.dil_h), Q = out_size(d.w, d.kw, d.pad_w, d.stride_w, d.dil_w);
      check_rc(zb_conv2d_bias_bwd(rt.ctx, gy.dtype, ZB_NHWC, gy.ptr, db.ptr, d.n, d.k, P, Q), "conv bias bwd");
      commit_grad(rt, bv, db);
    }
    if (need_dx && (xv.requires_grad || xv.creator)) {
      ProfScope ps(rt, conv_key("dgrad", d), conv_flops(d), conv_bytes(d, gy.elem_size()));
      if (xv.grad.defined() && !xv.is_param && xv.grad.storage && xv.grad.storage.use_count() == 1 && x_layout == ZB_NHWC) {
        // second arrival (residual fan-in): accumulate inside the dgrad epilogue instead of a separate add pass; a lazy masked
        // first arrival (the residual gradient of a fused BN + add + ReLU) is masked as the epilogue reads it
        if (xv.grad_mask.defined()) {
          check_rc(zb_conv2d_dgrad_acc_masked(rt.ctx, gy.dtype, ZB_NHWC, ZB_MATH_DEFAULT, &d, gy.ptr, w.ptr, xv.grad.ptr, xv.grad_mask.ptr),
                   "conv dgrad (masked accumulate)");
          xv.grad_mask = Tensor();
        } else {
          check_rc(zb_conv2d_dgrad_acc(rt.ctx, gy.dtype, ZB_NHWC, ZB_MATH_DEFAULT, &d, gy.ptr, w.ptr, xv.grad.ptr), "conv dgrad (accumulate)");
        }
      } else {
        // (for an NCHW network input the gradient is produced in NHWC order; nothing consumes it)
        Tensor dx = grad_target(rt, xv);
        check_rc(zb_conv2d_dgrad(rt.ctx, gy.dtype, ZB_NHWC, ZB_MATH_DEFAULT, &d, gy.ptr, w.ptr, dx.ptr), "conv dgrad");
        commit_grad(rt, xv, dx);
      }
    }
    if (side_wgrad) {
      Tensor dw = grad_target(rt, wv);
      check_rc(zb_conv2d_wgrad(zb_ctx_side(rt.ctx), gy.dtype, x_layout, ZB_MATH_DEFAULT, &d, gy.ptr, x.ptr, dw.ptr), "conv wgrad (side stream)");
      rt.side_hold.push_back(x);
      rt.side_hold.push_back(gy);
      rt.side_pending = true;
      commit_grad(rt, wv, dw);
    }
    x = Tensor();
  }
};

Variable conv2d(Runtime& rt, const Variable& x_in, const Variable& w, const Variable* bias, const ConvArgs& a, bool need_dx,
                const Variable* bn_shift) {
  Variable x = x_in;
  // fused BatchNorm statistics: only while training, f32 (the kernel reports 0 rows when the shape cannot fuse them)
  Tensor stats;
  int64_t stat_rows = 0;
  const bool want_stats = bn_shift != nullptr && rt.train && rt.dtype == ZB_F32 && w.shape().size() == 4 &&
                          (*bn_shift)->data.numel() == w.shape()[0] && !ZB_ENV_FLAG("ZENU_B200_NO_BNSTATS");
  if (want_stats) stats = rt.empty({static_cast<int64_t>(zb_conv2d_bnstats_rows(rt.ctx)), 2, w.shape()[0]});
  auto run_fprop = [&](int layout, const zb_conv2d_desc* d, const void* xp, void* yp, int dtype) {
    const void* bp = bias ? (*bias)->data.ptr : nullptr;
    if (want_stats)
      return zb_conv2d_fprop_bnstats(rt.ctx, dtype, layout, ZB_MATH_DEFAULT, d, xp, w->data.ptr, bp, yp, (*bn_shift)->data.ptr, stats.ptr,
                                     &stat_rows);
    return zb_conv2d_fprop(rt.ctx, dtype, layout, ZB_MATH_DEFAULT, d, xp, w->data.ptr, bp, yp);
  };
  const auto& ws = w.shape();
  if (x.shape().size() != 4 || ws.size() != 4) throw HostError("conv2d: bad shapes (NHWC input, KRSC filter)");
  auto fn = std::make_shared<ConvFn>();
  Tensor y;
  bool done = false;
  if (x->nchw) {
    // stem: try to consume the NCHW batch directly (C <= 4 on the TF32 path repacks inside the conv's staging pass)
    const auto& s = x.shape();  // [N, C, H, W]
    if (s[1] != ws[3]) throw HostError("conv2d: channel mismatch");
    fn->d = zb_conv2d_desc{s[0], s[1], s[2], s[3], ws[0], ws[1], ws[2], a.pad_h, a.pad_w, a.stride_h, a.stride_w, a.dil_h, a.dil_w};
    const int64_t P = out_size(s[2], ws[1], a.pad_h, a.stride_h, a.dil_h), Q = out_size(s[3], ws[2], a.pad_w, a.stride_w, a.dil_w);
    y = rt.empty({s[0], P, Q, ws[0]});
    int rc;
    {
      ProfScope ps(rt, conv_key("fprop", fn->d), conv_flops(fn->d), conv_bytes(fn->d, y.elem_size()));
      rc = run_fprop(ZB_NCHW_X, &fn->d, x->data.ptr, y.ptr, y.dtype);
    }
    if (rc == ZB_OK) {
      fn->x_layout = ZB_NCHW_X;
      done = true;
    } else if (rc == ZB_ERR_UNSUPPORTED) {
      x = nchw_to_nhwc(rt, x);
    } else {
      check_rc(rc, "conv fprop");
    }
  }
  if (!done) {
    const auto& xs = x.shape();
    if (xs[3] != ws[3]) throw HostError("conv2d: bad shapes (NHWC input, KRSC filter)");
    fn->d = zb_conv2d_desc{xs[0], xs[3], xs[1], xs[2], ws[0], ws[1], ws[2], a.pad_h, a.pad_w, a.stride_h, a.stride_w, a.dil_h, a.dil_w};
    const int64_t P = out_size(xs[1], ws[1], a.pad_h, a.stride_h, a.dil_h), Q = out_size(xs[2], ws[2], a.pad_w, a.stride_w, a.dil_w);
    y = rt.empty({xs[0], P, Q, ws[0]});
    ProfScope ps(rt, conv_key("fprop", fn->d), conv_flops(fn->d), conv_bytes(fn->d, y.elem_size()));
    check_rc(run_fprop(ZB_NHWC, &fn->d, x->data.ptr, y.ptr, y.dtype), "conv fprop");
  }
  fn->inputs = {x.ptr(), w.ptr()};
  if (bias) fn->inputs.push_back(bias->ptr());
  fn->has_bias = bias != nullptr;
  fn->need_dx = need_dx;
  fn->x = x->data;
  fn->w = w->data;
  Variable out = make_output(y, fn);
  if (want_stats && stat_rows > 0) {
    out->bn_stats = stats;
    out->bn_stat_rows = stat_rows;
    out->bn_shift = (*bn_shift)->data.ptr;
  }
  return out;
}

struct BnFn : Function {
  Tensor x, y, scale, bias, saved_mean, saved_inv;
  Tensor relu_mask;   // fused BN+add+ReLU: 1 bit per output element written by the forward (replaces keeping / re-reading y)
  int64_t n, c, h, w;
  bool relu, has_res;
  const char* name() const override { return "batch_norm_2d"; }
  // the plain BN of a downsample branch reads its (lazy) output gradient through the bit mask of the block's closing ReLU
  bool takes_masked_grad() const override { return !relu && !has_res && x.dtype == ZB_F32 && c % 32 == 0; }
  void backward(Runtime& rt, const Tensor& gy) override {
    VariableInner& xv = *inputs[0];
    Tensor dx = grad_target(rt, xv);
    Tensor ds = grad_target(rt, *inputs[1]);
    Tensor db = grad_target(rt, *inputs[2]);
    Tensor dres;
    void* dres_ptr = nullptr;
    // fused BN + add + ReLU with a bit mask: the residual branch receives (gy, bits) lazily, the masked product is never stored
    const bool lazy_res = has_res && relu && relu_mask.defined() && !inputs[3]->grad.defined() && !inputs[3]->is_param && rt.lazy_mask;
    if (has_res && relu && !lazy_res) {
      dres = grad_target(rt, *inputs[3]);
      // a second consumer may already have written the slot: then produce into fresh memory and add
      if (inputs[3]->grad.defined()) dres = rt.empty(inputs[3]->data.shape);
      dres_ptr = dres.ptr;
    }
    ProfScope ps(rt, std::string("bn.bwd") + (relu ? "+relu" : "") + (has_res ? "+res" : "") + " " + shape_str(x.shape), 0.0,
                 static_cast<double>(x.bytes()) * (5.0 + (gy_mask.defined() ? 2.0 / 32.0 : 0.0) +
                     (relu && has_res ? (lazy_res ? 2.0 / 32.0 : relu_mask.defined() ? 1.0 + 1.0 / 32.0 : 2.0) : 0.0)));
    if (gy_mask.defined())
      check_rc(zb_bn2d_bwd_mask(rt.ctx, gy.dtype, ZB_NHWC, n, c, h, w, x.ptr, gy.ptr, scale.ptr, saved_mean.ptr, saved_inv.ptr, dx.ptr,
                                ds.ptr, db.ptr, gy_mask.ptr, nullptr), "bn bwd (lazy masked gradient)");
    else if (relu_mask.defined())
      check_rc(zb_bn2d_bwd_mask(rt.ctx, gy.dtype, ZB_NHWC, n, c, h, w, x.ptr, gy.ptr, scale.ptr, saved_mean.ptr, saved_inv.ptr, dx.ptr,
                                ds.ptr, db.ptr, relu_mask.ptr, dres_ptr), "bn bwd (bit mask)");
    else if (relu && !has_res)  // mask recomputed from x: the forward output is neither kept nor read
      check_rc(zb_bn2d_relu_bwd(rt.ctx, gy.dtype, ZB_NHWC, n, c, h, w, x.ptr, gy.ptr, scale.ptr, bias.ptr, saved_mean.ptr,
                                saved_inv.ptr, dx.ptr, ds.ptr, db.ptr), "bn relu bwd");
    else
      check_rc(zb_bn2d_bwd(rt.ctx, gy.dtype, ZB_NHWC, n, c, h, w, x.ptr, gy.ptr, scale.ptr, saved_mean.ptr, saved_inv.ptr, dx.ptr,
                           ds.ptr, db.ptr, relu ? y.ptr : nullptr, dres_ptr), "bn bwd");
    commit_grad(rt, xv, dx);
    commit_grad(rt, *inputs[1], ds);
    commit_grad(rt, *inputs[2], db);
    if (has_res) {
      if (lazy_res) commit_grad_masked(rt, *inputs[3], gy, relu_mask);
      else commit_grad(rt, *inputs[3], relu ? dres : gy);
    }
    x = Tensor();
    y = Tensor();
    relu_mask = Tensor();
  }
};

Variable batch_norm_2d(Runtime& rt, const Variable& x, const Variable& scale, const Variable& bias, const Variable& mean,
                       const Variable& variance, double momentum, const Variable* residual, bool relu) {
  const auto& s = x.shape();
  if (s.size() != 4) throw HostError("batch_norm_2d: expected NHWC input");
  const int64_t n = s[0], h = s[1], w = s[2], c = s[3];
  Tensor y = rt.empty(s);
  if (!rt.train) {
    // inference branch (zenu-autograd/src/nn/batch_norm.rs:107-119); fused residual/relu applied separately
    check_rc(zb_bn2d_fwd_infer(rt.ctx, y.dtype, ZB_NHWC, n, c, h, w, x->data.ptr, scale->data.ptr, bias->data.ptr, mean->data.ptr,
                               variance->data.ptr, y.ptr), "bn infer");
    if (residual) check_rc(zb_binary(rt.ctx, y.dtype, ZB_OP_ADD, y.ptr, (*residual)->data.ptr, y.ptr, y.numel()), "bn infer add");
    if (relu) check_rc(zb_relu(rt.ctx, y.dtype, y.ptr, y.ptr, 0.0, y.numel()), "bn infer relu");
    return Variable::leaf(y);
  }
  auto fn = std::make_shared<BnFn>();
  fn->saved_mean = rt.empty({c});
  fn->saved_inv = rt.empty({c});
  const bool have_stats = x->bn_stat_rows > 0 && x->bn_shift == mean->data.ptr;   // statistics came with the producing conv
  const bool want_mask = relu && residual != nullptr && y.dtype == ZB_F32 && c % 32 == 0 && !ZB_ENV_FLAG("ZENU_B200_NO_RELU_MASK");
  // algorithmic bytes: x read twice (once when the statistics came with the conv), y written (+ residual read)
  ProfScope ps(rt, std::string("bn.fwd") + (relu ? "+relu" : "") + (residual ? "+res" : "") + " " + shape_str(s), 0.0,
               static_cast<double>(y.bytes()) * ((residual ? 4.0 : 3.0) - (have_stats ? 1.0 : 0.0)));
  if (want_mask) fn->relu_mask = rt.empty({zb_bn2d_relu_mask_words(n, c, h, w)});
  if (have_stats || want_mask) {
    check_rc(zb_bn2d_fwd_train_fused(rt.ctx, y.dtype, ZB_NHWC, n, c, h, w, momentum, x->data.ptr, scale->data.ptr, bias->data.ptr,
                                     mean->data.ptr, variance->data.ptr, fn->saved_mean.ptr, fn->saved_inv.ptr, y.ptr,
                                     residual ? (*residual)->data.ptr : nullptr, relu ? 1 : 0, have_stats ? x->bn_stats.ptr : nullptr,
                                     have_stats ? x->bn_stat_rows : 0, have_stats ? x->bn_shift : nullptr,
                                     want_mask ? fn->relu_mask.ptr : nullptr), "bn fwd (fused)");
    x->bn_stats = Tensor();
    x->bn_stat_rows = 0;
  } else {
    check_rc(zb_bn2d_fwd_train(rt.ctx, y.dtype, ZB_NHWC, n, c, h, w, momentum, x->data.ptr, scale->data.ptr, bias->data.ptr,
                               mean->data.ptr, variance->data.ptr, fn->saved_mean.ptr, fn->saved_inv.ptr, y.ptr,
                               residual ? (*residual)->data.ptr : nullptr, relu ? 1 : 0), "bn fwd");
  }
  fn->inputs = {x.ptr(), scale.ptr(), bias.ptr()};
  if (residual) fn->inputs.push_back(residual->ptr());
  fn->x = x->data;
  fn->scale = scale->data;
  fn->bias = bias->data;
  fn->relu = relu;
  fn->has_res = residual != nullptr;
  if (relu && residual && !fn->relu_mask.defined()) fn->y = y;
  fn->n = n; fn->c = c; fn->h = h; fn->w = w;
  return make_output(y, fn);
}

struct ReluFn : Function {
  Tensor x;
  const char* name() const override { return "relu"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    Tensor dx = grad_target(rt, *inputs[0]);
    ProfScope ps(rt, "relu.bwd " + shape_str(x.shape), 0.0, 3.0 * x.bytes());
    check_rc(zb_relu_bwd(rt.ctx, gy.dtype, x.ptr, gy.ptr, dx.ptr, 0.0, gy.numel()), "relu bwd");
    commit_grad(rt, *inputs[0], dx);
    x = Tensor();
  }
};
Variable relu(Runtime& rt, const Variable& x) {
  Tensor y = rt.empty(x.shape());
  ProfScope ps(rt, "relu.fwd " + shape_str(x.shape()), 0.0, 2.0 * y.bytes());
  check_rc(zb_relu(rt.ctx, y.dtype, x->data.ptr, y.ptr, 0.0, y.numel()), "relu");
  auto fn = std::make_shared<ReluFn>();
  fn->inputs = {x.ptr()};
  fn->x = x->data;
  return make_output(y, fn);
}

struct AddFn : Function {
  const char* name() const override { return "add"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    commit_grad(rt, *inputs[0], gy);
    commit_grad(rt, *inputs[1], gy);
  }
};
Variable add(Runtime& rt, const Variable& a, const Variable& b) {
  if (a.shape() != b.shape()) throw HostError("add: shape mismatch");
  Tensor y = rt.empty(a.shape());
  ProfScope ps(rt, "add.fwd " + shape_str(a.shape()), 0.0, 3.0 * y.bytes());
  check_rc(zb_binary(rt.ctx, y.dtype, ZB_OP_ADD, a->data.ptr, b->data.ptr, y.ptr, y.numel()), "add");
  auto fn = std::make_shared<AddFn>();
  fn->inputs = {a.ptr(), b.ptr()};
  return make_output(y, fn);
}

struct LinearFn : Function {
  Tensor x, w;
  int64_t b, in_f, out_f;
  bool has_bias;
  const char* name() const override { return "linear"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    VariableInner& xv = *inputs[0];
    Tensor dw = grad_target(rt, *inputs[1]);
    Tensor db;
    if (has_bias) db = grad_target(rt, *inputs[2]);
    Tensor dx;
    const bool want_dx = xv.requires_grad || xv.creator;
    if (want_dx) dx = grad_target(rt, xv);
    ProfScope ps(rt, "linear.bwd " + std::to_string(b) + "x" + std::to_string(in_f) + "x" + std::to_string(out_f),
                 (want_dx ? 4.0 : 2.0) * b * in_f * out_f, 0.0);
    check_rc(zb_linear_bwd(rt.ctx, gy.dtype, ZB_MATH_DEFAULT, x.ptr, w.ptr, gy.ptr, want_dx ? dx.ptr : nullptr, dw.ptr,
                           has_bias ? db.ptr : nullptr, b, in_f, out_f), "linear bwd");
    commit_grad(rt, *inputs[1], dw);
    if (has_bias) commit_grad(rt, *inputs[2], db);
    if (want_dx) commit_grad(rt, xv, dx);
    x = Tensor();
  }
};
Variable linear(Runtime& rt, const Variable& x, const Variable& w, const Variable* bias) {
  const auto& xs = x.shape();
  const auto& ws = w.shape();
  if (xs.size() != 2 || ws.size() != 2 || xs[1] != ws[1]) throw HostError("linear: bad shapes");
  Tensor y = rt.empty({xs[0], ws[0]});
  ProfScope ps(rt, "linear.fwd " + std::to_string(xs[0]) + "x" + std::to_string(xs[1]) + "x" + std::to_string(ws[0]),
               2.0 * xs[0] * xs[1] * ws[0], 0.0);
  check_rc(zb_linear_fwd(rt.ctx, y.dtype, ZB_MATH_DEFAULT, x->data.ptr, w->data.ptr, bias ? (*bias)->data.ptr : nullptr, y.ptr,
                         xs[0], xs[1], ws[0]), "linear fwd");
  auto fn = std::make_shared<LinearFn>();
  fn->inputs = {x.ptr(), w.ptr()};
  if (bias) fn->inputs.push_back(bias->ptr());
  fn->x = x->data;
  fn->w = w->data;
  fn->b = xs[0]; fn->in_f = xs[1]; fn->out_f = ws[0];
  fn->has_bias = bias != nullptr;
  return make_output(y, fn);
}

struct MaxPoolFn : Function {
  Tensor x, idx;  // idx (uint8 winning tap per output element) when the indexed NHWC kernels apply, else x is kept
  int64_t n, c, h, w, k, stride, pad;
  const char* name() const override { return "max_pool_2d"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    Tensor dx = grad_target(rt, *inputs[0]);
    ProfScope ps(rt, "maxpool.bwd " + shape_str(inputs[0]->data.shape), 0.0, static_cast<double>(dx.bytes() + gy.bytes()) + gy.numel());
    if (idx.defined())
      check_rc(zb_maxpool2d_bwd_idx(rt.ctx, gy.dtype, ZB_NHWC, gy.ptr, idx.ptr, dx.ptr, n, c, h, w, k, k, stride, stride, pad, pad), "maxpool bwd");
    else
      check_rc(zb_maxpool2d_bwd(rt.ctx, gy.dtype, ZB_NHWC, x.ptr, gy.ptr, dx.ptr, n, c, h, w, k, k, stride, stride, pad, pad), "maxpool bwd");
    commit_grad(rt, *inputs[0], dx);
    x = Tensor();
    idx = Tensor();
  }
};
Variable max_pool_2d(Runtime& rt, const Variable& x, int64_t k, int64_t stride, int64_t pad) {
  const auto& s = x.shape();
  const int64_t P = (s[1] + 2 * pad - k) / stride + 1, Q = (s[2] + 2 * pad - k) / stride + 1;
  Tensor y = rt.empty({s[0], P, Q, s[3]});
  auto fn = std::make_shared<MaxPoolFn>();
  ProfScope ps(rt, "maxpool.fwd " + shape_str(s), 0.0, static_cast<double>(x->data.bytes() + y.bytes()));
  if (s[3] % 4 == 0 && k * k < 255) {
    fn->idx = rt.empty({(y.numel() + 3) / 4});  // one byte per output element
    fn->idx.dtype = ZB_F32;
    check_rc(zb_maxpool2d_fwd_idx(rt.ctx, y.dtype, ZB_NHWC, x->data.ptr, y.ptr, fn->idx.ptr, s[0], s[3], s[1], s[2], k, k, stride, stride, pad, pad), "maxpool");
  } else {
    check_rc(zb_maxpool2d_fwd(rt.ctx, y.dtype, ZB_NHWC, x->data.ptr, y.ptr, s[0], s[3], s[1], s[2], k, k, stride, stride, pad, pad), "maxpool");
    fn->x = x->data;
  }
  fn->inputs = {x.ptr()};
  fn->n = s[0]; fn->c = s[3]; fn->h = s[1]; fn->w = s[2]; fn->k = k; fn->stride = stride; fn->pad = pad;
  return make_output(y, fn);
}

// BatchNorm2d(train) + ReLU + max_pool_2d(3, 2, 1) as one node (the ResNet stem): the BN output, the largest activation of the
// network, is neither written nor read back; backward gathers the pooled gradient inside both BN-backward passes
// (zb_bn2d_relu_maxpool_fwd_train / _bwd).  Same results as the three separate nodes.
struct BnReluPoolFn : Function {
  Tensor x, idx, scale, bias, saved_mean, saved_inv;
  int64_t n, c, h, w;
  const char* name() const override { return "batch_norm_2d+relu+max_pool_2d"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    VariableInner& xv = *inputs[0];
    Tensor dx = grad_target(rt, xv);
    Tensor ds = grad_target(rt, *inputs[1]);
    Tensor db = grad_target(rt, *inputs[2]);
    ProfScope ps(rt, "bn.bwd+relu+maxpool " + shape_str(x.shape), 0.0, 3.0 * x.bytes() + 2.0 * (gy.bytes() + gy.numel()));
    check_rc(zb_bn2d_relu_maxpool_bwd(rt.ctx, gy.dtype, ZB_NHWC, n, c, h, w, 3, 2, 1, x.ptr, gy.ptr, idx.ptr, scale.ptr, bias.ptr,
                                      saved_mean.ptr, saved_inv.ptr, dx.ptr, ds.ptr, db.ptr), "bn + relu + maxpool bwd");
    commit_grad(rt, xv, dx);
    commit_grad(rt, *inputs[1], ds);
    commit_grad(rt, *inputs[2], db);
    x = Tensor();
    idx = Tensor();
  }
};

bool batch_norm_relu_max_pool_fusable(Runtime& rt, const Variable& x, int64_t k, int64_t stride, int64_t pad) {
  const auto& s = x.shape();
  if (!rt.train || !rt.fuse_stem_pool || rt.dtype != ZB_F32 || s.size() != 4 || k != 3 || stride != 2 || pad != 1) return false;
  const int64_t c4 = s[3] / 4;
  return s[3] % 4 == 0 && c4 >= 1 && c4 <= 256 && (256 % c4) == 0 && s[1] >= 2 && s[2] >= 2 &&
         s[0] * (s[1] + 1) * (s[2] + 1) * s[3] < (1ll << 31) - (1ll << 24);
}

Variable batch_norm_relu_max_pool(Runtime& rt, const Variable& x, const Variable& scale, const Variable& bias, const Variable& mean,
                                  const Variable& variance, double momentum) {
  const auto& s = x.shape();
  const int64_t n = s[0], h = s[1], w = s[2], c = s[3];
  const int64_t P = (h + 2 - 3) / 2 + 1, Q = (w + 2 - 3) / 2 + 1;
  Tensor y = rt.empty({n, P, Q, c});
  auto fn = std::make_shared<BnReluPoolFn>();
  fn->saved_mean = rt.empty({c});
  fn->saved_inv = rt.empty({c});
  fn->idx = rt.empty({(y.numel() + 3) / 4});   // one byte per pooled element
  const bool have_stats = x->bn_stat_rows > 0 && x->bn_shift == mean->data.ptr;
  {
    ProfScope ps(rt, "bn.fwd+relu+maxpool " + shape_str(s), 0.0,
                 static_cast<double>(x->data.bytes()) * (have_stats ? 1.0 : 2.0) + static_cast<double>(y.bytes()) + y.numel());
    check_rc(zb_bn2d_relu_maxpool_fwd_train(rt.ctx, y.dtype, ZB_NHWC, n, c, h, w, 3, 2, 1, momentum, x->data.ptr, scale->data.ptr,
                                            bias->data.ptr, mean->data.ptr, variance->data.ptr, fn->saved_mean.ptr, fn->saved_inv.ptr,
                                            y.ptr, fn->idx.ptr, have_stats ? x->bn_stats.ptr : nullptr, have_stats ? x->bn_stat_rows : 0,
                                            have_stats ? x->bn_shift : nullptr), "bn + relu + maxpool fwd");
  }
  x->bn_stats = Tensor();
  x->bn_stat_rows = 0;
  fn->inputs = {x.ptr(), scale.ptr(), bias.ptr()};
  fn->x = x->data;
  fn->scale = scale->data;
  fn->bias = bias->data;
  fn->n = n; fn->c = c; fn->h = h; fn->w = w;
  return make_output(y, fn);
}

struct GapFn : Function {
  int64_t n, c, hw;
  const char* name() const override { return "global_avg_pool"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    Tensor dx = grad_target(rt, *inputs[0]);
    ProfScope ps(rt, "gap.bwd", 0.0, static_cast<double>(dx.bytes()));
    check_rc(zb_gap_bwd(rt.ctx, gy.dtype, ZB_NHWC, gy.ptr, dx.ptr, n, c, hw), "gap bwd");
    commit_grad(rt, *inputs[0], dx);
  }
};
Variable global_avg_pool(Runtime& rt, const Variable& x) {
  const auto& s = x.shape();
  Tensor y = rt.empty({s[0], s[3]});
  ProfScope ps(rt, "gap.fwd", 0.0, static_cast<double>(x->data.bytes()));
  check_rc(zb_gap_fwd(rt.ctx, y.dtype, ZB_NHWC, x->data.ptr, y.ptr, s[0], s[3], s[1] * s[2]), "gap");
  auto fn = std::make_shared<GapFn>();
  fn->inputs = {x.ptr()};
  fn->n = s[0]; fn->c = s[3]; fn->hw = s[1] * s[2];
  return make_output(y, fn);
}

struct ViewFn : Function {
  const char* name() const override { return "flatten"; }
  void backward(Runtime& rt, const Tensor& gy) override { commit_grad(rt, *inputs[0], gy.view(inputs[0]->data.shape)); }
};
Variable flatten(Runtime& rt, const Variable& x) {
  (void)rt;
  const auto& s = x.shape();
  int64_t rest = 1;
  for (size_t i = 1; i < s.size(); ++i) rest *= s[i];
  auto fn = std::make_shared<ViewFn>();
  fn->inputs = {x.ptr()};
  return make_output(x->data.view({s[0], rest}), fn);  // no copy (the reference copies, functions/flatten.rs:22-31)
}

// NHWC -> NCHW with gradient: used in front of flatten so that Linear sees the reference's C*H*W feature order
struct ToNchwFn : Function {
  const char* name() const override { return "nhwc_to_nchw"; }
  void backward(Runtime& rt, const Tensor& gy) override {
    const auto& s = inputs[0]->data.shape;  // NHWC
    Tensor dx = grad_target(rt, *inputs[0]);
    check_rc(zb_nchw_to_nhwc(rt.ctx, gy.dtype, gy.ptr, dx.ptr, s[0], s[3], s[1], s[2]), "to_nchw bwd");
    commit_grad(rt, *inputs[0], dx);
  }
};
Variable nhwc_to_nchw(Runtime& rt, const Variable& x) {
  const auto& s = x.shape();
  Tensor y = rt.empty({s[0], s[3], s[1], s[2]});
  check_rc(zb_nhwc_to_nchw(rt.ctx, y.dtype, x->data.ptr, y.ptr, s[0], s[3], s[1], s[2]), "nhwc_to_nchw");
  auto fn = std::make_shared<ToNchwFn>();
  fn->inputs = {x.ptr()};
  return make_output(y, fn);
}

Variable nchw_to_nhwc(Runtime& rt, const Variable& x) {
  const auto& s = x.shape();
  Tensor y = rt.empty({s[0], s[2], s[3], s[1]});
  ProfScope ps(rt, "input.nchw_to_nhwc " + shape_str(s), 0.0, 2.0 * y.bytes());
  check_rc(zb_nchw_to_nhwc(rt.ctx, y.dtype, x->data.ptr, y.ptr, s[0], s[1], s[2], s[3]), "nchw_to_nhwc");
  // The reference computes the input gradient of every conv, the first one included (conv_without_bias.rs:110-121);
  // marking the network input as requiring a gradient keeps that work in the step.
  return Variable::leaf(y, /*requires_grad=*/true);
}

struct XentFn : Function {
  Tensor dz;
  const char* name() const override { return "softmax_cross_entropy"; }
  void backward(Runtime& rt, const Tensor&) override {
    commit_grad(rt, *inputs[0], dz);  // seed is 1 for the scalar loss
    dz = Tensor();
  }
};
Variable softmax_cross_entropy(Runtime& rt, const Variable& logits, const Tensor& targets) {
  const auto& s = logits.shape();
  Tensor loss = rt.empty({1});
  auto fn = std::make_shared<XentFn>();
  const bool need = logits->requires_grad;
  if (need) fn->dz = rt.empty(s);
  ProfScope ps(rt, "softmax_xent " + shape_str(s), 0.0, 3.0 * logits->data.bytes());
  check_rc(zb_softmax_xent(rt.ctx, loss.dtype, logits->data.ptr, targets.ptr, loss.ptr, need ? fn->dz.ptr : nullptr, s[0], s[1]), "softmax_xent");
  fn->inputs = {logits.ptr()};
  return make_output(loss, fn);
}

// ------------------------------------------------------------------------------------------------ layers
ParamMap Module::parameters() const {
  ParamMap m;
  weights("", m);
  biases("", m);
  buffers("", m);
  return m;
}

static std::string join(const std::string& p, const char* leaf) { return p.empty() ? std::string(leaf) : p + "." + leaf; }

Conv2d::Conv2d(int64_t ci, int64_t co, int64_t k, int64_t stride, int64_t pad, int64_t dil, bool b)
    : args{pad, pad, stride, stride, dil, dil}, has_bias(b), cin(ci), cout(co), kh(k), kw(k) {}
Variable Conv2d::call(Runtime& rt, const Variable& x) {
  return conv2d(rt, x, filter, has_bias ? &bias : nullptr, args, need_dx, next_bn ? &next_bn->mean : nullptr);
}
void Conv2d::weights(const std::string& p, ParamMap& o) const { o[join(p, "conv2d.filter")] = filter; }
void Conv2d::biases(const std::string& p, ParamMap& o) const { if (has_bias) o[join(p, "conv2d.bias")] = bias; }

BatchNorm2d::BatchNorm2d(int64_t ch, double mom) : momentum(mom), channels(ch) {}
Variable BatchNorm2d::call(Runtime& rt, const Variable& x) { return batch_norm_2d(rt, x, scale, bias, mean, variance, momentum, nullptr, false); }
Variable BatchNorm2d::call_fused(Runtime& rt, const Variable& x, const Variable* residual, bool relu) {
  return batch_norm_2d(rt, x, scale, bias, mean, variance, momentum, residual, relu);
}
void BatchNorm2d::weights(const std::string& p, ParamMap& o) const { o[join(p, "batch_norm_2d.scale")] = scale; }
void BatchNorm2d::biases(const std::string& p, ParamMap& o) const { o[join(p, "batch_norm_2d.bias")] = bias; }
void BatchNorm2d::buffers(const std::string& p, ParamMap& o) const {
  o[join(p, "batch_norm_2d.mean")] = mean;
  o[join(p, "batch_norm_2d.variance")] = variance;
}

Linear::Linear(int64_t i, int64_t o, bool b) : has_bias(b), in_f(i), out_f(o) {}
Variable Linear::call(Runtime& rt, const Variable& x) { return linear(rt, x, weight, has_bias ? &bias : nullptr); }
void Linear::weights(const std::string& p, ParamMap& o) const { o[join(p, "linear.weight")] = weight; }
void Linear::biases(const std::string& p, ParamMap& o) const { if (has_bias) o[join(p, "linear.bias")] = bias; }

void Model::weights(const std::string& p, ParamMap& o) const { for (auto& c : children) c.second->weights(join(p, c.first.c_str()), o); }
void Model::biases(const std::string& p, ParamMap& o) const { for (auto& c : children) c.second->biases(join(p, c.first.c_str()), o); }
void Model::buffers(const std::string& p, ParamMap& o) const { for (auto& c : children) c.second->buffers(join(p, c.first.c_str()), o); }

// ---- parameter initialisers (identical on every rank: seeded per tensor) --------------------------------
static void spec_conv(std::vector<ParamSpec>& specs, const std::string& prefix, Conv2d& c) {
  const double he = std::sqrt(2.0 / static_cast<double>(c.cin * c.kh * c.kw));
  // reference init is unscaled N(0,1) (conv2d.rs:106), which overflows in a 50-layer net; He-normal instead
  specs.push_back({join(prefix, "conv2d.filter"), {c.cout, c.kh, c.kw, c.cin}, 0,
                   [he](float* p, int64_t n) { for (int64_t i = 0; i < n; ++i) p[i] *= static_cast<float>(he); }, &c.filter});
  if (c.has_bias) specs.push_back({join(prefix, "conv2d.bias"), {c.cout}, 1, nullptr, &c.bias});
}
static void spec_bn(std::vector<ParamSpec>& specs, const std::string& prefix, BatchNorm2d& b) {
  auto ones = [](float* p, int64_t n) { for (int64_t i = 0; i < n; ++i) p[i] = 1.f; };
  specs.push_back({join(prefix, "batch_norm_2d.scale"), {b.channels}, 0, ones, &b.scale});   // scale is a weight (adamw.rs:28)
  specs.push_back({join(prefix, "batch_norm_2d.bias"), {b.channels}, 1, nullptr, &b.bias});
  specs.push_back({join(prefix, "batch_norm_2d.mean"), {b.channels}, 2, nullptr, &b.mean});
  specs.push_back({join(prefix, "batch_norm_2d.variance"), {b.channels}, 2, ones, &b.variance});
}
static void spec_linear(std::vector<ParamSpec>& specs, const std::string& prefix, Linear& l) {
  const double s = 1.0 / std::sqrt(static_cast<double>(l.in_f));  // linear.rs:56-60
  specs.push_back({join(prefix, "linear.weight"), {l.out_f, l.in_f}, 0,
                   [s](float* p, int64_t n) { for (int64_t i = 0; i < n; ++i) p[i] *= static_cast<float>(s); }, &l.weight});
  if (l.has_bias) specs.push_back({join(prefix, "linear.bias"), {l.out_f}, 1, nullptr, &l.bias});
}

// ---- cfg1: the small CIFAR CNN of zenu/examples/cifar10.rs:29-69 ----------------------------------------
struct SmallCnn : Model {
  std::shared_ptr<Conv2d> conv1, conv2;
  std::shared_ptr<BatchNorm2d> bn1, bn2;
  std::shared_ptr<Linear> linear1, linear2;
  bool fused;
  SmallCnn(int num_classes, bool f) : fused(f) {
    arch = "small_cnn";
    conv1 = std::make_shared<Conv2d>(3, 32, 3, 1, 1, 1, true);
    bn1 = std::make_shared<BatchNorm2d>(32, 0.9);
    conv2 = std::make_shared<Conv2d>(32, 64, 3, 1, 1, 1, true);
    bn2 = std::make_shared<BatchNorm2d>(64, 0.9);
    linear1 = std::make_shared<Linear>(64 * 32 * 32, 512, true);
    linear2 = std::make_shared<Linear>(512, num_classes, true);
    children = {{"conv1", conv1}, {"batch_norm1", bn1}, {"conv2", conv2}, {"batch_norm2", bn2}, {"linear1", linear1}, {"linear2", linear2}};
    if (fused) { conv1->next_bn = bn1.get(); conv2->next_bn = bn2.get(); }
  }
  void collect(std::vector<ParamSpec>& s) override {
    spec_conv(s, "conv1", *conv1); spec_bn(s, "batch_norm1", *bn1);
    spec_conv(s, "conv2", *conv2); spec_bn(s, "batch_norm2", *bn2);
    spec_linear(s, "linear1", *linear1); spec_linear(s, "linear2", *linear2);
  }
  Variable call(Runtime& rt, const Variable& x) override {
    Variable h = conv1->call(rt, x);
    h = fused ? bn1->call_fused(rt, h, nullptr, true) : relu(rt, bn1->call(rt, h));
    h = conv2->call(rt, h);
    h = fused ? bn2->call_fused(rt, h, nullptr, true) : relu(rt, bn2->call(rt, h));
    h = flatten(rt, nhwc_to_nchw(rt, h));  // reference feature order: [N, C*H*W] (functions/flatten.rs)
    h = relu(rt, linear1->call(rt, h));
    return linear2->call(rt, h);
  }
};

// ---- ResNet-18 / ResNet-50, torchvision v1.5 topology built from the reference's layer API --------------
// block output = ReLU(BN(conv(..)) + shortcut), as the reference's ResBlock sketch (zenu/examples/resnet.rs:18-28)
struct ResBlock : Module {
  std::vector<std::shared_ptr<Conv2d>> convs;
  std::vector<std::shared_ptr<BatchNorm2d>> bns;
  std::shared_ptr<Conv2d> down_conv;
  std::shared_ptr<BatchNorm2d> down_bn;
  bool fused;
  ResBlock(bool bottleneck, int64_t cin, int64_t width, int64_t stride, bool f) : fused(f) {
    const int64_t cout = bottleneck ? width * 4 : width;
    if (bottleneck) {
      convs = {std::make_shared<Conv2d>(cin, width, 1, 1, 0, 1, false), std::make_shared<Conv2d>(width, width, 3, stride, 1, 1, false),
               std::make_shared<Conv2d>(width, cout, 1, 1, 0, 1, false)};
      bns = {std::make_shared<BatchNorm2d>(width, 0.9), std::make_shared<BatchNorm2d>(width, 0.9), std::make_shared<BatchNorm2d>(cout, 0.9)};
    } else {
      convs = {std::make_shared<Conv2d>(cin, width, 3, stride, 1, 1, false), std::make_shared<Conv2d>(width, cout, 3, 1, 1, 1, false)};
      bns = {std::make_shared<BatchNorm2d>(width, 0.9), std::make_shared<BatchNorm2d>(cout, 0.9)};
    }
    if (stride != 1 || cin != cout) {
      down_conv = std::make_shared<Conv2d>(cin, cout, 1, stride, 0, 1, false);
      down_bn = std::make_shared<BatchNorm2d>(cout, 0.9);
    }
    if (fused) {   // every conv feeds a BatchNorm: its statistics pass runs inside the conv epilogue
      for (size_t i = 0; i < convs.size(); ++i) convs[i]->next_bn = bns[i].get();
      if (down_conv) down_conv->next_bn = down_bn.get();
    }
  }
  void collect(std::vector<ParamSpec>& s, const std::string& p) {
    for (size_t i = 0; i < convs.size(); ++i) {
      spec_conv(s, p + ".conv" + std::to_string(i + 1), *convs[i]);
      spec_bn(s, p + ".bn" + std::to_string(i + 1), *bns[i]);
    }
    if (down_conv) { spec_conv(s, p + ".downsample_conv", *down_conv); spec_bn(s, p + ".downsample_bn", *down_bn); }
  }
  void each(const std::string& p, const std::function<void(const std::string&, const Module&)>& f) const {
    for (size_t i = 0; i < convs.size(); ++i) {
      f(join(p, ("conv" + std::to_string(i + 1)).c_str()), *convs[i]);
      f(join(p, ("bn" + std::to_string(i + 1)).c_str()), *bns[i]);
    }
    if (down_conv) { f(join(p, "downsample_conv"), *down_conv); f(join(p, "downsample_bn"), *down_bn); }
  }
  void weights(const std::string& p, ParamMap& o) const override { each(p, [&](const std::string& q, const Module& m) { m.weights(q, o); }); }
  void biases(const std::string& p, ParamMap& o) const override { each(p, [&](const std::string& q, const Module& m) { m.biases(q, o); }); }
  void buffers(const std::string& p, ParamMap& o) const override { each(p, [&](const std::string& q, const Module& m) { m.buffers(q, o); }); }
  Variable call(Runtime& rt, const Variable& x) override {
    Variable shortcut = x;
    if (down_conv) shortcut = down_bn->call(rt, down_conv->call(rt, x));
    Variable h = x;
    const size_t last = convs.size() - 1;
    for (size_t i = 0; i < last; ++i) {
      h = convs[i]->call(rt, h);
      h = fused ? bns[i]->call_fused(rt, h, nullptr, true) : relu(rt, bns[i]->call(rt, h));
    }
    h = convs[last]->call(rt, h);
    if (fused) return bns[last]->call_fused(rt, h, &shortcut, true);
    return relu(rt, add(rt, bns[last]->call(rt, h), shortcut));
  }
};

struct ResNet : Model {
  std::shared_ptr<Conv2d> conv1;
  std::shared_ptr<BatchNorm2d> bn1;
  std::vector<std::pair<std::string, std::shared_ptr<ResBlock>>> blocks;
  std::shared_ptr<Linear> fc;
  bool fused;
  ResNet(int depth, int num_classes, bool f) : fused(f) {
    arch = depth == 18 ? "resnet18" : "resnet50";
    const bool bottleneck = depth == 50;
    const int counts18[4] = {2, 2, 2, 2}, counts50[4] = {3, 4, 6, 3};
    const int* counts = bottleneck ? counts50 : counts18;
    conv1 = std::make_shared<Conv2d>(3, 64, 7, 2, 3, 1, false);
    conv1->need_dx = true;  // the reference computes the input gradient of every conv (conv_without_bias.rs:110-121)
    bn1 = std::make_shared<BatchNorm2d>(64, 0.9);
    if (fused) conv1->next_bn = bn1.get();
    children = {{"conv1", conv1}, {"bn1", bn1}};
    int64_t cin = 64;
    for (int stage = 0; stage < 4; ++stage) {
      const int64_t width = 64ll << stage;
      for (int b = 0; b < counts[stage]; ++b) {
        const int64_t stride = (b == 0 && stage > 0) ? 2 : 1;
        auto blk = std::make_shared<ResBlock>(bottleneck, cin, width, stride, fused);
        const std::string name = "layer" + std::to_string(stage + 1) + "." + std::to_string(b);
        blocks.push_back({name, blk});
        children.push_back({name, blk});
        cin = bottleneck ? width * 4 : width;
      }
    }
    fc = std::make_shared<Linear>(cin, num_classes, true);
    children.push_back({"fc", fc});
  }
  void collect(std::vector<ParamSpec>& s) override {
    spec_conv(s, "conv1", *conv1);
    spec_bn(s, "bn1", *bn1);
    for (auto& b : blocks) b.second->collect(s, b.first);
    spec_linear(s, "fc", *fc);
  }
  Variable call(Runtime& rt, const Variable& x) override {
    Variable h = conv1->call(rt, x);
    if (fused && batch_norm_relu_max_pool_fusable(rt, h, 3, 2, 1)) {
      h = batch_norm_relu_max_pool(rt, h, bn1->scale, bn1->bias, bn1->mean, bn1->variance, bn1->momentum);
    } else {
      h = fused ? bn1->call_fused(rt, h, nullptr, true) : relu(rt, bn1->call(rt, h));
      h = max_pool_2d(rt, h, 3, 2, 1);
    }
    for (auto& b : blocks) h = b.second->call(rt, h);
    h = global_avg_pool(rt, h);
    return fc->call(rt, h);
  }
};

std::shared_ptr<Model> make_model(const std::string& arch, int num_classes, bool fused) {
  if (arch == "small_cnn") return std::make_shared<SmallCnn>(num_classes, fused);
  if (arch == "resnet18") return std::make_shared<ResNet>(18, num_classes, fused);
  if (arch == "resnet50") return std::make_shared<ResNet>(50, num_classes, fused);
  throw HostError("unknown arch '" + arch + "' (small_cnn | resnet18 | resnet50)");
}

// ------------------------------------------------------------------------------------------------ parameter store
void ParamStore::build(Runtime& r, Model& model, uint64_t seed, int64_t bucket_bytes) {
  rt = &r;
  std::vector<ParamSpec> specs;
  model.collect(specs);
  // bucket assignment + flat layout: zb_dp_plan_buckets (pure host logic, shared with the CPU tests)
  const size_t esz = r.dtype == ZB_F64 ? 8 : 4;
  std::vector<int64_t> numels(specs.size());
  std::vector<int> kinds(specs.size()), spec_bucket(specs.size(), -1);
  std::vector<int64_t> offsets(specs.size(), 0);
  for (size_t i = 0; i < specs.size(); ++i) {
    int64_t n = 1;
    for (auto sdim : specs[i].shape) n *= sdim;
    numels[i] = n;
    kinds[i] = specs[i].kind;
  }
  int num_buckets = 0;
  int64_t total = 0, buf_total = 0;
  check_rc(zb_dp_plan_buckets(numels.data(), kinds.data(), static_cast<int>(specs.size()), bucket_bytes, static_cast<int>(esz),
                              spec_bucket.data(), offsets.data(), &num_buckets, &total, &buf_total), "bucket plan");
  buckets.assign(num_buckets, Bucket{0, 0, 0, 0});
  for (int b = 0; b < num_buckets; ++b) {
    int64_t lo = -1, hi = 0;
    for (size_t i = 0; i < specs.size(); ++i)
      if (spec_bucket[i] == b) {
        if (lo < 0 || offsets[i] < lo) lo = offsets[i];
        hi = std::max(hi, offsets[i] + ((numels[i] + 3) & ~int64_t(3)));
        buckets[b].total++;
      }
    buckets[b].offset = lo < 0 ? 0 : lo;
    buckets[b].numel = hi - buckets[b].offset;
  }
  flat_params = r.zeros({total});
  flat_grads = r.zeros({total});
  flat_buffers = r.zeros({std::max<int64_t>(buf_total, 4)});
  // host init, then one upload
  std::vector<float> hp(total, 0.f), hb(std::max<int64_t>(buf_total, 4), 0.f);
  entries.clear();
  for (size_t i = 0; i < specs.size(); ++i) {
    int64_t n = 1;
    for (auto s : specs[i].shape) n *= s;
    float* dst = (specs[i].kind == 2 ? hb.data() : hp.data()) + offsets[i];
    if (specs[i].kind == 0 && specs[i].init) {
      std::mt19937_64 gen(seed * 1000003ull + i);
      std::normal_distribution<float> nd(0.f, 1.f);
      for (int64_t j = 0; j < n; ++j) dst[j] = nd(gen);
    }
    if (specs[i].init) specs[i].init(dst, n);
    uint8_t* base = static_cast<uint8_t*>(specs[i].kind == 2 ? flat_buffers.ptr : flat_params.ptr);
    Tensor data = r.borrow(base + offsets[i] * esz, specs[i].shape);
    Variable v = Variable::leaf(data, specs[i].kind != 2, specs[i].name);
    v->is_param = specs[i].kind != 2;
    if (specs[i].kind != 2) {
      v->grad_slot = r.borrow(static_cast<uint8_t*>(flat_grads.ptr) + offsets[i] * esz, specs[i].shape);
      v->bucket = spec_bucket[i];
    }
    *specs[i].target = v;
    entries.push_back({specs[i].name, v, specs[i].kind, offsets[i], n, spec_bucket[i]});
  }
  if (r.dtype == ZB_F32) {
    cudaMemcpyAsync(flat_params.ptr, hp.data(), sizeof(float) * total, cudaMemcpyHostToDevice, r.ctx->stream);
    cudaMemcpyAsync(flat_buffers.ptr, hb.data(), sizeof(float) * hb.size(), cudaMemcpyHostToDevice, r.ctx->stream);
    cudaStreamSynchronize(r.ctx->stream);
  } else {
    std::vector<double> dp(hp.begin(), hp.end()), db(hb.begin(), hb.end());
    cudaMemcpyAsync(flat_params.ptr, dp.data(), sizeof(double) * total, cudaMemcpyHostToDevice, r.ctx->stream);
    cudaMemcpyAsync(flat_buffers.ptr, db.data(), sizeof(double) * db.size(), cudaMemcpyHostToDevice, r.ctx->stream);
    cudaStreamSynchronize(r.ctx->stream);
  }
  reset_pending();
}

void ParamStore::reset_pending() {
  for (auto& b : buckets) b.pending = b.total;
}

// ------------------------------------------------------------------------------------------------ optimizers
void Optimizer::init(Runtime& rt, ParamStore& ps) {
  if (kind != OPT_SGD) {
    m = rt.zeros({ps.flat_params.numel()});
    v = rt.zeros({ps.flat_params.numel()});
  }
  step = 0;
  tbl_first = 0;
}

void Optimizer::begin_step(Runtime& rt) {
  ++step;
  if (kind == OPT_SGD) return;
  if (!bc_table.defined()) {
    bc_table = rt.empty({kTblSteps, 2});
    step_index = rt.empty({2});   // (one int32 used; sized in elements of the model dtype)
  }
  if (tbl_first == 0 || step >= tbl_first + kTblSteps) {   // (re)fill the window starting at this step, index back to 0
    tbl_first = step;
    check_rc(zb_adam_table_fill(rt.ctx, rt.dtype, beta1, beta2, tbl_first, kTblSteps, bc_table.ptr), "adam table");
    if (cudaMemsetAsync(step_index.ptr, 0, 4, rt.ctx->stream) != cudaSuccess) throw HostError("adam step index reset failed");
  }
}

static double ps_bytes(const ParamStore& ps) { return static_cast<double>(ps.flat_params.bytes()); }
void Optimizer::update(Runtime& rt, ParamStore& ps) {
  check_rc(zb_dp_wait(rt.ctx), "dp wait");
  const double gscale = 1.0 / static_cast<double>(std::max(1, zb_dp_world(rt.ctx)));
  const size_t esz = rt.dtype == ZB_F64 ? 8 : 4;
  if (!in_replay_capture) begin_step(rt);
  const int32_t* sidx = kind == OPT_SGD ? nullptr : static_cast<const int32_t*>(step_index.ptr);
  ProfScope scope(rt, "optimizer.update", 0.0, 3.0 * ps_bytes(ps));
  auto at = [&](const Tensor& t, int64_t off) { return static_cast<void*>(static_cast<uint8_t*>(t.ptr) + off * esz); };
  if (kind == OPT_SGD) {
    // p -= lr * g over the whole flat buffer (sgd.rs:20-30), gradient averaging folded in
    check_rc(zb_sgd_step(rt.ctx, rt.dtype, ps.flat_params.ptr, ps.flat_grads.ptr, lr, gscale, ps.flat_params.numel()), "sgd step");
  } else if (kind == OPT_ADAM) {
    check_rc(zb_adam_step_table(rt.ctx, rt.dtype, ps.flat_params.ptr, ps.flat_grads.ptr, m.ptr, v.ptr, lr, beta1, beta2, eps, 0.0, 0,
                                bc_table.ptr, sidx, gscale, ps.flat_params.numel()), "adam step");
  } else {
    // AdamW: decoupled decay on weights() only (adamw.rs:28,61-65): per bucket, the weight run then the bias run
    for (auto& b : ps.buckets) {
      int64_t w_end = b.offset;
      for (auto& e : ps.entries)
        if (e.kind == 0 && e.bucket == static_cast<int>(&b - &ps.buckets[0])) w_end = std::max(w_end, e.offset + ((e.numel + 3) & ~int64_t(3)));
      const int64_t nw = w_end - b.offset, nbias = b.numel - nw;
      if (nw > 0)
        check_rc(zb_adam_step_table(rt.ctx, rt.dtype, at(ps.flat_params, b.offset), at(ps.flat_grads, b.offset), at(m, b.offset),
                                    at(v, b.offset), lr, beta1, beta2, eps, weight_decay, 1, bc_table.ptr, sidx, gscale, nw), "adamw step");
      if (nbias > 0)
        check_rc(zb_adam_step_table(rt.ctx, rt.dtype, at(ps.flat_params, w_end), at(ps.flat_grads, w_end), at(m, w_end), at(v, w_end), lr,
                                    beta1, beta2, eps, weight_decay, 0, bc_table.ptr, sidx, gscale, nbias), "adamw step");
    }
  }
  if (kind != OPT_SGD) check_rc(zb_adam_advance(rt.ctx, static_cast<int32_t*>(step_index.ptr)), "adam advance");
}

}  // namespace host
}  // namespace zb
