// checkpoint.cu — reader / writer of the reference's model files and zb_model_save / zb_model_load.
//
// Format (reference zenu/src/lib.rs:26-67: `bincode::serialize(&model.parameters())`, bincode 1.3.3 default options =
// little endian, fixed-width integers, u64 lengths; HashMap<String, Variable<T, D>>; a Variable serialises as its data
// Matrix, zenu-autograd/src/lib.rs:149-160,320-331; Matrix fields in order, zenu-matrix/src/impl_serde.rs:11-40):
//   u64 n_entries
//   n_entries x { u64 len, utf-8 key;
//                 u64 ndim, ndim x u64 shape;   u64 ndim, ndim x u64 stride (elements);
//                 u64 numel, numel x T data (f32 / f64 little endian, the tensor linearised row-major);
//                 u64 len, utf-8 data_type ("f32" | "f64" = std::any::type_name::<T>());   u64 ptr_offset }
// Reference layouts: conv filter [K,C,R,S], conv bias [1,K,1,1], BatchNorm scale/bias/mean/variance [C], Linear weight
// [out,in], bias [out] (zenu-layer/src/layers/*.rs); this library keeps filters KRSC and conv biases [K] on the device, the
// conversion happens here.  load mirrors load_model: every key of the file must exist in the model (else an error), model
// parameters the file does not name keep their values.
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "autograd.h"

using namespace zb::host;

struct zb_ckpt_entry_t {
  std::string name;
  std::vector<int64_t> shape, stride;
  int dtype;
  int64_t ptr_offset;
  std::vector<uint8_t> data;   // dense row-major (strides already resolved)
};
struct zb_ckpt {
  std::vector<zb_ckpt_entry_t> entries;
};

namespace {

struct Writer {
  std::vector<uint8_t> buf;
  void u64(uint64_t v) { for (int i = 0; i < 8; ++i) buf.push_back(static_cast<uint8_t>(v >> (8 * i))); }
  void str(const std::string& s) { u64(s.size()); buf.insert(buf.end(), s.begin(), s.end()); }
  void raw(const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); buf.insert(buf.end(), b, b + n); }
};

struct Reader {
  const uint8_t* p;
  size_t n, pos = 0;
  bool ok = true;
  uint64_t u64() {
    if (pos + 8 > n) { ok = false; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= static_cast<uint64_t>(p[pos + i]) << (8 * i);
    pos += 8;
    return v;
  }
  std::string str() {
    const uint64_t len = u64();
    if (!ok || len > n - pos) { ok = false; return std::string(); }
    std::string s(reinterpret_cast<const char*>(p + pos), len);
    pos += len;
    return s;
  }
};

void append_entry(Writer& w, const std::string& name, const std::vector<int64_t>& shape, int dtype, const void* data) {
  w.str(name);
  w.u64(shape.size());
  int64_t numel = 1;
  for (int64_t d : shape) { w.u64(static_cast<uint64_t>(d)); numel *= d; }
  w.u64(shape.size());
  for (size_t i = 0; i < shape.size(); ++i) {   // default (row-major) strides, as Matrix::from_vec produces
    int64_t st = 1;
    for (size_t j = i + 1; j < shape.size(); ++j) st *= shape[j];
    w.u64(static_cast<uint64_t>(st));
  }
  w.u64(static_cast<uint64_t>(numel));
  w.raw(data, static_cast<size_t>(numel) * (dtype == ZB_F64 ? 8 : 4));
  w.str(dtype == ZB_F64 ? "f64" : "f32");
  w.u64(0);
}

int write_file(const char* path, const std::vector<uint8_t>& buf) {
  FILE* f = fopen(path, "wb");
  if (!f) { zb::set_last_error("Failed to save model: cannot open %s", path); return ZB_ERR_INVALID; }
  const size_t n = fwrite(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (n != buf.size()) { zb::set_last_error("Failed to save model: short write to %s", path); return ZB_ERR_INVALID; }
  return ZB_OK;
}

}  // namespace

extern "C" {

int zb_ckpt_write(const char* path, int dtype, int n, const char* const* names, const int* ndims, const int64_t* const* shapes,
                  const void* const* host_data) {
  ZB_API_RANGE();
  ZB_REQUIRE(path && (n == 0 || (names && ndims && shapes && host_data)), "zb_ckpt_write: NULL argument");
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "zb_ckpt_write: unknown dtype %d", dtype);
  Writer w;
  w.u64(static_cast<uint64_t>(n));
  for (int i = 0; i < n; ++i) {
    ZB_REQUIRE(ndims[i] >= 0 && ndims[i] <= 8, "zb_ckpt_write: bad rank");
    append_entry(w, names[i], std::vector<int64_t>(shapes[i], shapes[i] + ndims[i]), dtype, host_data[i]);
  }
  return write_file(path, w.buf);
}

static int ckpt_open_impl(const char* path, zb_ckpt** out);
int zb_ckpt_open(const char* path, zb_ckpt** out) {
  ZB_API_RANGE();
  ZB_REQUIRE(path && out, "zb_ckpt_open: NULL argument");
  try {   // nothing thrown by the parser (bad_alloc, length_error on hostile sizes) may cross the C boundary
    return ckpt_open_impl(path, out);
  } catch (const std::exception& e) {
    zb::set_last_error("Failed to load model: %s", e.what());
    return ZB_ERR_INVALID;
  }
}
static int ckpt_open_impl(const char* path, zb_ckpt** out) {
  FILE* f = fopen(path, "rb");
  if (!f) { zb::set_last_error("Failed to load model: cannot open %s", path); return ZB_ERR_INVALID; }
  std::vector<uint8_t> bin;
  uint8_t chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) bin.insert(bin.end(), chunk, chunk + got);
  fclose(f);
  Reader r{bin.data(), bin.size()};
  auto ck = std::make_unique<zb_ckpt>();
  const uint64_t n = r.u64();
  ZB_REQUIRE(r.ok && n < (1ull << 32), "Failed to load model: truncated header");
  for (uint64_t i = 0; i < n; ++i) {
    zb_ckpt_entry_t e;
    e.name = r.str();
    const uint64_t nd = r.u64();
    ZB_REQUIRE(r.ok && nd <= 8, "Failed to load model: bad rank in entry %llu", static_cast<unsigned long long>(i));
    for (uint64_t k = 0; k < nd; ++k) e.shape.push_back(static_cast<int64_t>(r.u64()));
    const uint64_t ns = r.u64();
    ZB_REQUIRE(r.ok && ns == nd, "Failed to load model: shape / stride rank mismatch in '%s'", e.name.c_str());
    for (uint64_t k = 0; k < ns; ++k) e.stride.push_back(static_cast<int64_t>(r.u64()));
    const uint64_t len = r.u64();
    ZB_REQUIRE(r.ok, "Failed to load model: truncated entry '%s'", e.name.c_str());
    const size_t data_pos = r.pos;
    // element size is only known after data_type: find it by trying both widths (the type string follows the data)
    int dtype = -1;
    for (int cand = 0; cand < 2 && dtype < 0; ++cand) {
      const size_t esz = cand == 0 ? 4 : 8;
      if (len > (bin.size() - data_pos) / esz) continue;
      Reader t{bin.data(), bin.size()};
      t.pos = data_pos + len * esz;
      const std::string ty = t.str();
      if (t.ok && ty == (cand == 0 ? "f32" : "f64")) { dtype = cand == 0 ? ZB_F32 : ZB_F64; r.pos = t.pos; }
    }
    ZB_REQUIRE(dtype >= 0, "Failed to load model: Data type mismatch in '%s' (f32 / f64 expected)", e.name.c_str());
    e.dtype = dtype;
    e.ptr_offset = static_cast<int64_t>(r.u64());
    ZB_REQUIRE(r.ok, "Failed to load model: truncated entry '%s'", e.name.c_str());
    const size_t esz = dtype == ZB_F64 ? 8 : 4;
    // the shape is file content: every extent must be non-negative and the element count (overflow-checked) can never exceed
    // the elements the entry actually carries (a broadcast view, stride 0, may be smaller than len but never larger)
    int64_t numel = 1;
    for (int64_t d : e.shape) {
      ZB_REQUIRE(d >= 0, "Failed to load model: negative extent in the shape of '%s'", e.name.c_str());
      ZB_REQUIRE(d == 0 || numel <= static_cast<int64_t>(bin.size()) / d, "Failed to load model: shape of '%s' is larger than the file",
                 e.name.c_str());
      numel *= d;
    }
    ZB_REQUIRE(static_cast<uint64_t>(numel) * esz <= bin.size(), "Failed to load model: shape of '%s' is larger than the file", e.name.c_str());
    // resolve (shape, stride, ptr_offset) into a dense row-major copy (Matrix::new(ptr, shape, stride), impl_serde.rs:160-168)
    e.data.resize(static_cast<size_t>(numel) * esz);
    std::vector<int64_t> idx(e.shape.size(), 0);
    for (int64_t lin = 0; lin < numel; ++lin) {
      int64_t src = e.ptr_offset;
      for (size_t k = 0; k < idx.size(); ++k) src += idx[k] * e.stride[k];
      ZB_REQUIRE(src >= 0 && static_cast<uint64_t>(src) < len, "Failed to load model: stride walks outside the data of '%s'", e.name.c_str());
      memcpy(e.data.data() + lin * esz, bin.data() + data_pos + src * esz, esz);
      for (int k = static_cast<int>(idx.size()) - 1; k >= 0; --k) {
        if (++idx[k] < e.shape[k]) break;
        idx[k] = 0;
      }
    }
    ck->entries.push_back(std::move(e));
  }
  *out = ck.release();
  return ZB_OK;
}

int zb_ckpt_count(const zb_ckpt* ck) { return ck ? static_cast<int>(ck->entries.size()) : 0; }

int zb_ckpt_entry(const zb_ckpt* ck, int index, char* name, int name_cap, int64_t* shape, int* ndim, int* dtype, const void** host_data,
                  int64_t* numel) {
  ZB_API_RANGE();
  ZB_REQUIRE(ck && index >= 0 && index < static_cast<int>(ck->entries.size()), "zb_ckpt_entry: index out of range");
  const zb_ckpt_entry_t& e = ck->entries[index];
  if (name && name_cap > 0) { strncpy(name, e.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ndim) *ndim = static_cast<int>(e.shape.size());
  if (shape) for (size_t i = 0; i < e.shape.size(); ++i) shape[i] = e.shape[i];
  if (dtype) *dtype = e.dtype;
  if (host_data) *host_data = e.data.data();
  if (numel) *numel = static_cast<int64_t>(e.data.size() / (e.dtype == ZB_F64 ? 8 : 4));
  return ZB_OK;
}

int zb_ckpt_close(zb_ckpt* ck) {
  ZB_API_RANGE();
  delete ck;
  return ZB_OK;
}

}  // extern "C"

// ---- model <-> file ----------------------------------------------------------------------------------------------------
struct zb_model;
namespace zb { namespace host {
ParamStore& model_params(zb_model* m);
zb_ctx* model_ctx(zb_model* m);
int model_dtype(zb_model* m);
Optimizer& model_optimizer(zb_model* m);
void model_drop_graphs(zb_model* m);
} }

static bool ends_with(const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

// KRSC (device) <-> KCRS (reference) on the host
template <typename T>
static void krsc_to_kcrs(const T* src, T* dst, int64_t K, int64_t R, int64_t S, int64_t C) {
  for (int64_t k = 0; k < K; ++k)
    for (int64_t r = 0; r < R; ++r)
      for (int64_t s = 0; s < S; ++s)
        for (int64_t c = 0; c < C; ++c) dst[((k * C + c) * R + r) * S + s] = src[((k * R + r) * S + s) * C + c];
}
template <typename T>
static void kcrs_to_krsc(const T* src, T* dst, int64_t K, int64_t C, int64_t R, int64_t S) {
  for (int64_t k = 0; k < K; ++k)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t r = 0; r < R; ++r)
        for (int64_t s = 0; s < S; ++s) dst[((k * R + r) * S + s) * C + c] = src[((k * C + c) * R + r) * S + s];
}

// one tensor of the model (a parameter, or the slice of an optimizer-state buffer that belongs to it) -> a file entry in the
// reference's layout: filters KCRS, conv bias [1,K,1,1]
static int append_device_tensor(Writer& w, const std::string& key, const std::string& param_name, const void* dev, const std::vector<int64_t>& dev_shape,
                                int dtype, std::vector<uint8_t>& host, std::vector<uint8_t>& conv) {
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  int64_t numel = 1;
  for (int64_t d : dev_shape) numel *= d;
  host.resize(static_cast<size_t>(numel) * esz);
  ZB_CHECK_CUDA(cudaMemcpy(host.data(), dev, host.size(), cudaMemcpyDeviceToHost));
  std::vector<int64_t> shape = dev_shape;
  const void* data = host.data();
  if (ends_with(param_name, "conv2d.filter") && shape.size() == 4) {
    conv.resize(host.size());
    const int64_t K = shape[0], R = shape[1], S = shape[2], C = shape[3];
    if (dtype == ZB_F64) krsc_to_kcrs(reinterpret_cast<const double*>(host.data()), reinterpret_cast<double*>(conv.data()), K, R, S, C);
    else krsc_to_kcrs(reinterpret_cast<const float*>(host.data()), reinterpret_cast<float*>(conv.data()), K, R, S, C);
    shape = {K, C, R, S};
    data = conv.data();
  } else if (ends_with(param_name, "conv2d.bias") && shape.size() == 1) {
    shape = {1, shape[0], 1, 1};   // zenu-layer/src/layers/conv2d.rs:99
  }
  append_entry(w, key, shape, dtype, data);
  return ZB_OK;
}

// a file entry -> the device tensor of parameter `param_name` (or its optimizer-state slice); shapes validated by the caller
static int load_device_tensor(const zb_ckpt_entry_t& e, const std::string& param_name, void* dev, const std::vector<int64_t>& dev_shape, int dtype,
                              std::vector<uint8_t>& conv) {
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  const void* src = e.data.data();
  if (ends_with(param_name, "conv2d.filter") && dev_shape.size() == 4) {
    conv.resize(e.data.size());
    const int64_t K = e.shape[0], C = e.shape[1], R = e.shape[2], S = e.shape[3];
    if (dtype == ZB_F64) kcrs_to_krsc(reinterpret_cast<const double*>(e.data.data()), reinterpret_cast<double*>(conv.data()), K, C, R, S);
    else kcrs_to_krsc(reinterpret_cast<const float*>(e.data.data()), reinterpret_cast<float*>(conv.data()), K, C, R, S);
    src = conv.data();
  }
  int64_t numel = 1;
  for (int64_t d : dev_shape) numel *= d;
  ZB_CHECK_CUDA(cudaMemcpy(dev, src, static_cast<size_t>(numel) * esz, cudaMemcpyHostToDevice));
  return ZB_OK;
}

static int check_entry_against(const zb_ckpt_entry_t& e, const std::string& param_name, const Tensor& t, int dtype) {
  ZB_REQUIRE(e.dtype == dtype, "Failed to load model: Data type mismatch in '%s'", e.name.c_str());
  int64_t numel = 1;
  for (int64_t d : e.shape) numel *= d;
  ZB_REQUIRE(numel == t.numel(), "Failed to load model: '%s' has %lld elements, the model expects %lld", e.name.c_str(),
             static_cast<long long>(numel), static_cast<long long>(t.numel()));
  if (ends_with(param_name, "conv2d.filter") && t.shape.size() == 4)
    ZB_REQUIRE(e.shape.size() == 4 && e.shape[0] == t.shape[0] && e.shape[1] == t.shape[3] && e.shape[2] == t.shape[1] &&
                   e.shape[3] == t.shape[2], "Failed to load model: filter '%s' shape mismatch (KCRS expected)", e.name.c_str());
  return ZB_OK;
}

// Training-state files (zb_model_save_state / zb_model_load_state): the same container with, next to the parameters, the optimizer
// state the reference keeps in memory only (zenu-optimizer/src/adam.rs:9-17: step, m and v as HashMap<String, Variable> keyed by the
// parameter names): "optimizer.step" (scalar, completed updates), and for Adam / AdamW "optimizer.m.<param>" / "optimizer.v.<param>"
// in the parameter's own reference layout.  A plain model file is a valid state file without optimizer entries.
static const char kOptStep[] = "optimizer.step";
static const char kOptM[] = "optimizer.m.";
static const char kOptV[] = "optimizer.v.";

static int save_impl(zb_model* m, const char* path, bool with_state) {
  ParamStore& ps = model_params(m);
  zb_ctx* ctx = model_ctx(m);
  const int dtype = model_dtype(m);
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  Optimizer& opt = model_optimizer(m);
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  const bool adam = with_state && opt.kind != OPT_SGD && opt.m.defined() && opt.v.defined();
  size_t count = ps.entries.size();
  if (with_state) {
    ++count;
    if (adam)
      for (const ParamEntry& e : ps.entries) count += e.kind != 2 ? 2 : 0;
  }
  Writer w;
  w.u64(count);
  std::vector<uint8_t> host, conv;
  int rc;
  for (const ParamEntry& e : ps.entries)
    if ((rc = append_device_tensor(w, e.name, e.name, e.var->data.ptr, e.var->data.shape, dtype, host, conv)) != ZB_OK) return rc;
  if (with_state) {
    const double step_d = static_cast<double>(opt.step);
    const float step_f = static_cast<float>(opt.step);
    append_entry(w, kOptStep, {}, dtype, dtype == ZB_F64 ? static_cast<const void*>(&step_d) : static_cast<const void*>(&step_f));
    if (adam)
      for (const ParamEntry& e : ps.entries) {
        if (e.kind == 2) continue;   // buffers (BN running statistics) have no gradient and no optimizer state
        const uint8_t* mp = static_cast<const uint8_t*>(opt.m.ptr) + e.offset * esz;
        const uint8_t* vp = static_cast<const uint8_t*>(opt.v.ptr) + e.offset * esz;
        if ((rc = append_device_tensor(w, kOptM + e.name, e.name, mp, e.var->data.shape, dtype, host, conv)) != ZB_OK) return rc;
        if ((rc = append_device_tensor(w, kOptV + e.name, e.name, vp, e.var->data.shape, dtype, host, conv)) != ZB_OK) return rc;
      }
  }
  return write_file(path, w.buf);
}

static int load_impl(zb_model* m, const char* path, bool with_state) {
  zb_ckpt* ck = nullptr;
  int rc = zb_ckpt_open(path, &ck);
  if (rc != ZB_OK) return rc;
  std::unique_ptr<zb_ckpt> guard(ck);
  ParamStore& ps = model_params(m);
  zb_ctx* ctx = model_ctx(m);
  const int dtype = model_dtype(m);
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  Optimizer& opt = model_optimizer(m);
  std::map<std::string, const ParamEntry*> by_name;
  for (const ParamEntry& e : ps.entries) by_name[e.name] = &e;
  // validate everything before touching the model
  for (const zb_ckpt_entry_t& e : ck->entries) {
    std::string pname = e.name;
    if (with_state && e.name == kOptStep) {
      ZB_REQUIRE(e.dtype == dtype && e.data.size() == esz, "Failed to load model: bad '%s' entry", kOptStep);
      continue;
    }
    const bool is_m = with_state && e.name.compare(0, sizeof(kOptM) - 1, kOptM) == 0;
    const bool is_v = with_state && e.name.compare(0, sizeof(kOptV) - 1, kOptV) == 0;
    if (is_m || is_v) {
      pname = e.name.substr(sizeof(kOptM) - 1);
      ZB_REQUIRE(opt.kind != OPT_SGD && opt.m.defined() && opt.v.defined(),
                 "Failed to load model: the file carries Adam state ('%s') but the model's optimizer is not Adam / AdamW "
                 "(call zb_model_set_optimizer first)", e.name.c_str());
    }
    auto it = by_name.find(pname);
    ZB_REQUIRE(it != by_name.end(), "Failed to load model: the model has no parameter '%s'", e.name.c_str());
    ZB_REQUIRE(!(is_m || is_v) || it->second->kind != 2, "Failed to load model: '%s' names a buffer, which has no optimizer state", e.name.c_str());
    if ((rc = check_entry_against(e, pname, it->second->var->data, dtype)) != ZB_OK) return rc;
  }
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  std::vector<uint8_t> conv;
  for (const zb_ckpt_entry_t& e : ck->entries) {
    if (with_state && e.name == kOptStep) {
      const double v = dtype == ZB_F64 ? *reinterpret_cast<const double*>(e.data.data()) : static_cast<double>(*reinterpret_cast<const float*>(e.data.data()));
      ZB_REQUIRE(v >= 0.0 && v < 9.0e15, "Failed to load model: bad optimizer step %g", v);
      opt.step = static_cast<int64_t>(v);
      opt.tbl_first = 0;   // the device table of Adam bias corrections is refilled from the restored step on the next update
      continue;
    }
    const bool is_m = with_state && e.name.compare(0, sizeof(kOptM) - 1, kOptM) == 0;
    const bool is_v = with_state && e.name.compare(0, sizeof(kOptV) - 1, kOptV) == 0;
    if (is_m || is_v) {
      const ParamEntry* pe = by_name[e.name.substr(sizeof(kOptM) - 1)];
      uint8_t* base = static_cast<uint8_t*>(is_m ? opt.m.ptr : opt.v.ptr) + pe->offset * esz;
      if ((rc = load_device_tensor(e, pe->name, base, pe->var->data.shape, dtype, conv)) != ZB_OK) return rc;
    } else {
      const ParamEntry* pe = by_name[e.name];
      if ((rc = load_device_tensor(e, pe->name, pe->var->data.ptr, pe->var->data.shape, dtype, conv)) != ZB_OK) return rc;
    }
  }
  model_drop_graphs(m);   // captured steps stay valid address-wise, but a restored step count re-bases the Adam table window
  return ZB_OK;
}

extern "C" {

int zb_model_save(zb_model* m, const char* path) {
  ZB_API_RANGE();
  ZB_REQUIRE(m && path, "zb_model_save: NULL argument");
  return save_impl(m, path, false);
}

int zb_model_load(zb_model* m, const char* path) {
  ZB_API_RANGE();
  ZB_REQUIRE(m && path, "zb_model_load: NULL argument");
  return load_impl(m, path, false);
}

int zb_model_save_state(zb_model* m, const char* path) {
  ZB_API_RANGE();
  ZB_REQUIRE(m && path, "zb_model_save_state: NULL argument");
  return save_impl(m, path, true);
}

int zb_model_load_state(zb_model* m, const char* path) {
  ZB_API_RANGE();
  ZB_REQUIRE(m && path, "zb_model_load_state: NULL argument");
  return load_impl(m, path, true);
}

}  // extern "C"
