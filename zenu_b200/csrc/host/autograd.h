// autograd.h — host side above the C ABI: device tensors, the define-by-run tape, layers, optimizers.
//
// Mirrors the reference's host interface for this path (it is compiled Rust there, compiled C++ here):
//   Variable / Function / backward / set_grad / clear_grad   zenu-autograd/src/lib.rs:48-56,126-135,220-237,413-486
//   Module::call, Parameters::{weights,biases,parameters}     zenu-layer/src/lib.rs:21-51
//   Conv2d / BatchNorm2d / Linear / MaxPool2d                 zenu-layer/src/layers/*.rs
//   Optimizer::update, SGD / Adam / AdamW                     zenu-optimizer/src/{lib,sgd,adam,adamw}.rs
// Differences by design (B200-first, documented in DESIGN.md):
//   * activations are NHWC and filters KRSC on the device; NCHW/KCRS only at the model boundary;
//   * backward is first order only (the reference can build higher-order graphs; its BN double-backward panics);
//   * parameters and their gradients live in two flat buffers so the optimizer step is one fused kernel per
//     bucket and the data-parallel allreduce needs no gather copy;
//   * BN+ReLU(+residual) and conv+bias are single nodes (the unfused nodes exist too).
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../common.cuh"

namespace zb {
namespace host {

class HostError : public std::exception {
 public:
  explicit HostError(std::string m) : msg_(std::move(m)) {}
  const char* what() const noexcept override { return msg_.c_str(); }
 private:
  std::string msg_;
};
void check_rc(int rc, const char* what);

// ---- caching device allocator (reference: zenu-matrix/src/memory_pool/mod.rs:80-146) -------------------
// Exact-size free lists: a training loop asks for the same sizes every step, so after the first step no
// cudaMalloc is issued.  Single stream, so reuse needs no event tracking.
class Allocator {
 public:
  explicit Allocator(zb_ctx* ctx) : ctx_(ctx) {}
  ~Allocator();
  void* alloc(size_t bytes);
  void free(void* p, size_t bytes);
  size_t bytes_reserved() const { return reserved_; }
  void release_cached();
  // bumped whenever cached blocks went back to the driver: captured CUDA graphs that address them are stale
  uint64_t generation() const { return generation_; }
 private:
  zb_ctx* ctx_;
  std::unordered_map<size_t, std::vector<void*>> free_;
  size_t reserved_ = 0;
  uint64_t generation_ = 0;
};

struct Storage {
  Allocator* alloc = nullptr;  // nullptr: not owned (view into a flat buffer / caller memory)
  void* ptr = nullptr;
  size_t bytes = 0;
  ~Storage() { if (alloc && ptr) alloc->free(ptr, bytes); }
};

struct Tensor {
  std::shared_ptr<Storage> storage;
  void* ptr = nullptr;
  std::vector<int64_t> shape;
  int dtype = ZB_F32;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
  size_t elem_size() const { return dtype == ZB_F64 ? 8 : 4; }
  size_t bytes() const { return static_cast<size_t>(numel()) * elem_size(); }
  bool defined() const { return ptr != nullptr; }
  Tensor view(std::vector<int64_t> new_shape) const { Tensor t = *this; t.shape = std::move(new_shape); return t; }
};

// Optional per-node timing (CUDA events on the compute stream), keyed by "op.pass shape"; bench.py's per-layer table.
struct OpProfiler {
  struct Rec { std::string key; cudaEvent_t a, b; double flops, bytes; };
  bool enabled = false;
  std::vector<Rec> recs;
  void clear();
  ~OpProfiler() { clear(); }
};

struct Runtime {  // one per model: ctx + allocator + train flag (reference: global is_train, lib.rs:58-79)
  zb_ctx* ctx;
  Allocator alloc;
  bool train = true;
  int dtype = ZB_F32;
  OpProfiler prof;
  // conv wgrad on the ctx's side stream (zb_ctx_side): overlaps the BatchNorm-backward / dgrad chain of the following layers.
  // Tensors the side kernels read are held here until join_side() has made the main stream wait for them.
  bool overlap_wgrad = false;
  // fused BN + add + ReLU backward hands its residual gradient on as (gy, ReLU bits) instead of writing gy (.) bits out; the consumer
  // (the next dgrad's accumulate epilogue, or a downsample BN's backward) masks as it reads.  ZENU_B200_NO_LAZY_MASK=1 turns it off.
  bool lazy_mask = true;
  // the ResNet stem's BatchNorm + ReLU + max-pool run as one node (ZENU_B200_NO_STEM_POOL_FUSION=1: three nodes)
  bool fuse_stem_pool = true;
  bool side_pending = false;
  std::vector<Tensor> side_hold;
  void join_side();
  explicit Runtime(zb_ctx* c);
  Tensor empty(std::vector<int64_t> shape);
  Tensor zeros(std::vector<int64_t> shape);
  Tensor borrow(void* p, std::vector<int64_t> shape);  // non-owning
};

struct ProfScope {  // RAII: records an event pair around the kernels a node enqueues
  Runtime& rt;
  size_t idx = static_cast<size_t>(-1);
  ProfScope(Runtime& r, std::string key, double flops = 0.0, double bytes = 0.0);
  ~ProfScope();
};

// ---- tape ------------------------------------------------------------------------------------------
struct VariableInner;
using VarPtr = std::shared_ptr<VariableInner>;

struct Function {
  virtual ~Function() = default;
  std::vector<VarPtr> inputs;
  std::weak_ptr<VariableInner> output;
  int gen = 0;
  // consume the gradient of the output, produce/accumulate gradients of the inputs
  virtual void backward(Runtime& rt, const Tensor& gy) = 0;
  virtual const char* name() const = 0;
  // A "lazy masked" output gradient is the pair (gy, 1 bit per element) whose value is gy where the bit is set, else 0 (what the
  // fused BN + add + ReLU backward hands to its residual branch without writing the product out).  A function that can apply the
  // mask while it reads gy says so here and finds the bits in gy_mask; for all others the sweep materialises the product first.
  virtual bool takes_masked_grad() const { return false; }
  // Functions of one generation may run in any order (their input gradients are summed, and a + b == b + a bit for bit); the sweep
  // runs lower hints first.  A strided pointwise conv asks to go last: as the FIRST arrival its dgrad must zero-fill the whole input
  // gradient and scatter into it, and the sibling then re-reads all of it to accumulate; as the second arrival it only touches the
  // pixels it reaches, on top of the sibling's plain write.
  virtual int order_hint() const { return 0; }
  Tensor gy_mask;
};

struct VariableInner {
  Tensor data;
  Tensor grad;                       // undefined until a gradient arrives
  Tensor grad_mask;                  // defined: grad is lazy, its value is grad where the bit is set else 0 (see Function::gy_mask)
  Tensor grad_slot;                  // parameters: pre-assigned view into the flat gradient buffer
  std::shared_ptr<Function> creator;
  int gen = 0;
  bool requires_grad = false;        // reference: is_train on parameters; inputs always get grads there
  bool is_param = false;
  int bucket = -1;                   // data-parallel bucket this parameter's gradient belongs to
  bool nchw = false;                 // network input still in the reference's NCHW layout (consumed by the stem conv)
  // conv output followed by a BatchNorm: per-channel statistics already accumulated in the conv epilogue
  Tensor bn_stats;                   // [rows][2][K] partial sums (zb_conv2d_fprop_bnstats)
  int64_t bn_stat_rows = 0;
  const void* bn_shift = nullptr;    // the shift vector they were taken against (the BN's running mean)
  std::string name;
};

class Variable {
 public:
  Variable() = default;
  explicit Variable(VarPtr p) : p_(std::move(p)) {}
  static Variable leaf(Tensor t, bool requires_grad = false, std::string name = "");
  VariableInner* operator->() const { return p_.get(); }
  const VarPtr& ptr() const { return p_; }
  bool defined() const { return static_cast<bool>(p_); }
  const std::vector<int64_t>& shape() const { return p_->data.shape; }
  // reverse sweep in generation order (lib.rs:220-237); seeds d(out)/d(out) = 1 for a scalar loss
  void backward(Runtime& rt, const std::function<void(int bucket)>& on_bucket_ready = nullptr) const;
  void clear_grad() const;  // drops the graph reachable from this variable (lib.rs:434-440)
 private:
  VarPtr p_;
};

// accumulate a gradient into v: first arrival is stored (directly in the parameter's slot), later arrivals
// are added — `grad + old` like set_grad (lib.rs:466-486) but in place
void accumulate_grad(Runtime& rt, VariableInner& v, const Tensor& g);
// Returns where a first-arriving gradient of v should be written (the flat slot for parameters, else fresh memory)
Tensor grad_target(Runtime& rt, VariableInner& v);
void commit_grad(Runtime& rt, VariableInner& v, const Tensor& g);  // g was produced in grad_target() or elsewhere
// first arrival: keeps (g, bits) lazily; otherwise the product is materialised and added
void commit_grad_masked(Runtime& rt, VariableInner& v, const Tensor& g, const Tensor& bits);
void materialise_grad(Runtime& rt, VariableInner& v);  // v.grad <- v.grad (.) bits, mask dropped (no-op when not lazy)

// ---- differentiable functions (NHWC activations, KRSC filters) ------------------------------------------
struct ConvArgs { int64_t pad_h, pad_w, stride_h, stride_w, dil_h, dil_w; };
// bn_shift: running mean of the BatchNorm2d that consumes the output (training): its statistics pass is fused into the conv
Variable conv2d(Runtime& rt, const Variable& x, const Variable& w, const Variable* bias, const ConvArgs& a, bool need_dx = true,
                const Variable* bn_shift = nullptr);
// BatchNorm2d with optional fused residual add and ReLU; running stats updated in place when training
Variable batch_norm_2d(Runtime& rt, const Variable& x, const Variable& scale, const Variable& bias, const Variable& mean,
                       const Variable& variance, double momentum, const Variable* residual, bool relu);
// BatchNorm2d(train) + ReLU + max_pool_2d(k, stride, pad) as one node where zb_bn2d_relu_maxpool_* serves the geometry
bool batch_norm_relu_max_pool_fusable(Runtime& rt, const Variable& x, int64_t k, int64_t stride, int64_t pad);
Variable batch_norm_relu_max_pool(Runtime& rt, const Variable& x, const Variable& scale, const Variable& bias, const Variable& mean,
                                  const Variable& variance, double momentum);
Variable relu(Runtime& rt, const Variable& x);
Variable add(Runtime& rt, const Variable& a, const Variable& b);
Variable linear(Runtime& rt, const Variable& x, const Variable& w, const Variable* bias);
Variable max_pool_2d(Runtime& rt, const Variable& x, int64_t k, int64_t stride, int64_t pad);
Variable global_avg_pool(Runtime& rt, const Variable& x);
Variable flatten(Runtime& rt, const Variable& x);  // [N,H,W,C] -> [N, H*W*C] view
Variable nchw_to_nhwc(Runtime& rt, const Variable& x);
Variable nhwc_to_nchw(Runtime& rt, const Variable& x);  // differentiable
Variable softmax_cross_entropy(Runtime& rt, const Variable& logits, const Tensor& targets);  // scalar loss

// ---- layers ----------------------------------------------------------------------------------------------
using ParamMap = std::map<std::string, Variable>;  // ordered: deterministic iteration (the reference's HashMap is not)

struct Module {
  virtual ~Module() = default;
  virtual Variable call(Runtime& rt, const Variable& x) = 0;
  virtual void weights(const std::string& prefix, ParamMap& out) const = 0;
  virtual void biases(const std::string& prefix, ParamMap& out) const = 0;
  virtual void buffers(const std::string& prefix, ParamMap& out) const {}  // BN running stats (parameters() in the reference)
  ParamMap parameters() const;  // weights + biases + buffers, reference naming
};

struct ParamSpec {  // collected first, materialised into the flat buffers by ParamStore
  std::string name;
  std::vector<int64_t> shape;
  int kind;  // 0 weight, 1 bias, 2 buffer (no gradient)
  std::function<void(float* host, int64_t n)> init;
  Variable* target;
};

struct Conv2d : Module {
  Variable filter, bias;  // filter [K,R,S,C] (KRSC); bias [K] (reference shape [1,K,1,1])
  ConvArgs args;
  bool has_bias, need_dx = true;
  int64_t cin, cout, kh, kw;
  const struct BatchNorm2d* next_bn = nullptr;   // set by fused model builders: the BatchNorm2d applied to this conv's output
  Conv2d(int64_t cin, int64_t cout, int64_t k, int64_t stride, int64_t pad, int64_t dil, bool bias);
  Variable call(Runtime& rt, const Variable& x) override;
  void weights(const std::string& p, ParamMap& o) const override;
  void biases(const std::string& p, ParamMap& o) const override;
};
struct BatchNorm2d : Module {
  Variable scale, bias, mean, variance;
  double momentum;
  int64_t channels;
  BatchNorm2d(int64_t channels, double momentum);
  Variable call(Runtime& rt, const Variable& x) override;
  Variable call_fused(Runtime& rt, const Variable& x, const Variable* residual, bool relu);
  void weights(const std::string& p, ParamMap& o) const override;
  void biases(const std::string& p, ParamMap& o) const override;
  void buffers(const std::string& p, ParamMap& o) const override;
};
struct Linear : Module {
  Variable weight, bias;  // weight [out, in]
  bool has_bias;
  int64_t in_f, out_f;
  Linear(int64_t in_f, int64_t out_f, bool bias);
  Variable call(Runtime& rt, const Variable& x) override;
  void weights(const std::string& p, ParamMap& o) const override;
  void biases(const std::string& p, ParamMap& o) const override;
};

// ---- models ----------------------------------------------------------------------------------------------
struct Model : Module {
  std::string arch;
  // (name, module) in definition order; names follow the derive(Parameters) field-prefix convention
  std::vector<std::pair<std::string, std::shared_ptr<Module>>> children;
  void weights(const std::string& p, ParamMap& o) const override;
  void biases(const std::string& p, ParamMap& o) const override;
  void buffers(const std::string& p, ParamMap& o) const override;
  // reference-order list used to place parameters in the flat buffers (forward order)
  virtual void collect(std::vector<ParamSpec>& specs) = 0;
};
std::shared_ptr<Model> make_model(const std::string& arch, int num_classes, bool fused);

// ---- parameters in flat buffers + optimizers ----------------------------------------------------------
struct ParamEntry {
  std::string name;
  Variable var;
  int kind;
  int64_t offset, numel;  // offset in the flat param buffer (kinds 0,1) or buffer area (kind 2)
  int bucket;
};

struct ParamStore {
  Runtime* rt = nullptr;
  Tensor flat_params, flat_grads, flat_buffers;
  std::vector<ParamEntry> entries;
  std::vector<uint8_t> decay_mask_host;  // per entry: 1 if weights() (AdamW decays those only)
  struct Bucket { int64_t offset, numel; int pending, total; };
  std::vector<Bucket> buckets;           // buckets[0] = parameters closest to the loss (ready first)
  void build(Runtime& rt, Model& model, uint64_t seed, int64_t bucket_bytes);
  void reset_pending();
  int64_t trainable_numel() const { return flat_params.numel(); }
};

enum OptimKind { OPT_SGD = 0, OPT_ADAM = 1, OPT_ADAMW = 2 };
struct Optimizer {
  int kind = OPT_SGD;
  double lr = 0.01, beta1 = 0.9, beta2 = 0.999, eps = 1e-8, weight_decay = 0.0;
  int64_t step = 0;
  Tensor m, v;  // Adam state over the flat parameter buffer
  // Adam bias corrections for steps [tbl_first, tbl_first + kTblSteps) in device memory + the device-side index of the current
  // step: the kernels take nothing that changes from step to step as an argument, so a captured step can be replayed
  static constexpr int64_t kTblSteps = 4096;
  Tensor bc_table, step_index;
  int64_t tbl_first = 0;
  void init(Runtime& rt, ParamStore& ps);
  // advances the host step count and makes the device table cover it (called by update(), and by the graph replay path instead
  // of update())
  void begin_step(Runtime& rt);
  bool in_replay_capture = false;   // set by the step-graph capture, which calls begin_step itself (outside the captured region)
  // Optimizer::update: waits for the bucket allreduces (data parallel), then one fused kernel per bucket
  void update(Runtime& rt, ParamStore& ps);
};

}  // namespace host
}  // namespace zb
