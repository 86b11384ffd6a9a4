// api.cu — C-ABI entry points for convolution, GEMM and Linear: argument validation and dispatch over
// dtype x layout x math mode onto the tcgen05 path (umma_gemm.cu) or the FFMA/DFMA path (conv_simt.cu).
// There is no CPU fallback anywhere: a request neither GPU path can serve returns ZB_ERR_UNSUPPORTED.
#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"

namespace zb {
// umma_gemm.cu
int umma_gemm(zb_ctx*, bool, bool, long long, long long, long long, float, const float*, long long, const float*, long long,
              float, float*, long long, const float*, const uint32_t* old_bits = nullptr);
void umma_set_chain_limit(int);
bool umma_conv_supported(const zb_conv2d_desc*);
int umma_conv_fprop_nhwc(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, const float*, float*, float, const float*, float*, int*);
int umma_conv_dgrad_nhwc(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float, const uint32_t* old_bits = nullptr);
int umma_conv_wgrad_nhwc(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);
bool umma_conv1x1_nchw_supported(const zb_conv2d_desc*, int pass);
int umma_conv1x1_nchw_fprop(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);
int umma_conv1x1_nchw_dgrad(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);
int umma_conv1x1_nchw_wgrad(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);
bool umma_conv_smallc_supported(const zb_conv2d_desc*);
int umma_conv_smallc_fprop(zb_ctx*, const zb_conv2d_desc*, const float*, int, const float*, const float*, float*, float, const float*, float*, int*);
int umma_conv_smallc_wgrad(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, int, float*, float);
bool umma_conv_smallc_dgrad_supported(const zb_conv2d_desc*);
int umma_conv_smallc_dgrad(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);
// conv_simt.cu
template <typename T> int simt_conv_fprop(zb_ctx*, int, const zb_conv2d_desc*, const T*, const T*, const T*, T*);
template <typename T> int simt_conv_dgrad(zb_ctx*, int, const zb_conv2d_desc*, const T*, const T*, T*);
template <typename T> int simt_conv_wgrad(zb_ctx*, int, const zb_conv2d_desc*, const T*, const T*, T*);
template <typename T> int simt_gemm(zb_ctx*, bool, bool, long long, long long, long long, T, const T*, long long, const T*, long long, T, T*, long long, const T*);
// elementwise.cu / batchnorm.cu
template <typename T> int transpose_batched(zb_ctx*, const T*, T*, long long, long long, long long);
template <typename T> int bias_add_nchw(zb_ctx*, const T*, const T*, T*, long long, long long, long long);
template <typename T> int channel_sum(zb_ctx*, int, long long, long long, long long, const T*, T*);

static int check_desc(const zb_conv2d_desc* d, long long* P, long long* Q) {
  ZB_REQUIRE(d != nullptr, "conv: desc is NULL");
  ZB_REQUIRE(d->n > 0 && d->c > 0 && d->h > 0 && d->w > 0 && d->k > 0 && d->kh > 0 && d->kw > 0, "conv: non-positive extent");
  ZB_REQUIRE(d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0 && d->pad_h >= 0 && d->pad_w >= 0,
             "conv: bad stride/dilation/padding");
  ZB_REQUIRE(d->h + 2 * d->pad_h >= d->dil_h * (d->kh - 1) + 1 && d->w + 2 * d->pad_w >= d->dil_w * (d->kw - 1) + 1,
             "conv: filter larger than padded input");
  *P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  *Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  ZB_REQUIRE(d->n * d->c * d->h * d->w < (1ll << 40) && d->n * d->k * (*P) * (*Q) < (1ll << 40), "conv: tensor too large");
  return ZB_OK;
}

static int resolve_math(zb_ctx* ctx, int dtype, int math, int* out) {
  int m = (math == ZB_MATH_DEFAULT) ? ctx->default_math : math;
  ZB_REQUIRE(m == ZB_MATH_TF32 || m == ZB_MATH_TF32X3 || m == ZB_MATH_FP32, "unknown math mode %d", math);
  if (dtype == ZB_F64) m = ZB_MATH_FP32;  // f64 always runs DFMA
  *out = m;
  return ZB_OK;
}

// Scoped stream-ordered temporaries for the NCHW (reference-contract) staging path.
struct Temp {
  zb_ctx* ctx;
  void* p = nullptr;
  explicit Temp(zb_ctx* c) : ctx(c) {}
  bool fake = false;   // plan dry run (zb_conv2d_plan_describe): an address nothing dereferences
  int alloc(size_t bytes) {
    if (plan_dry()) {
      fake = true;
      p = reinterpret_cast<void*>(uintptr_t(1) << 40);
      return ZB_OK;
    }
    ZB_CHECK_CUDA(cudaMallocAsync(&p, std::max<size_t>(bytes, 16), ctx->stream));
    return ZB_OK;
  }
  ~Temp() { if (p && !fake) cudaFreeAsync(p, ctx->stream); }
};

// ---------------------------------------------------------------------------------------------- 3xTF32
// ZB_MATH_TF32X3: every f32 operand is split once into hi = tf32(x) (round to nearest) and lo = x - hi (exact in fp32, at most
// 13 significant bits), and the product is accumulated on the tensor cores as lo*hi + hi*lo + hi*hi (the lo*lo term, ~2^-22
// relative, is dropped) by three launches of the same tcgen05 kernels, the 2nd and 3rd with an accumulating (beta = 1) epilogue.
// Inside each launch the TMEM accumulation chain is limited to 64 MMA steps (UmmaParams::chain_kb).
// Result: f32-level accuracy (rel. 1e-5 contract, ~1e-6 measured) at roughly a third of the TF32 rate, with no SIMT fallback.
__global__ void split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const float v = x[i];
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    const float hf = __uint_as_float(h);
    hi[i] = hf;
    lo[i] = v - hf;
  }
}

struct SplitOperand {
  Temp t;
  const float* hi = nullptr;
  const float* lo = nullptr;
  explicit SplitOperand(zb_ctx* c) : t(c) {}
  int make(const float* x, long long n) {
    int rc = t.alloc(sizeof(float) * 2 * static_cast<size_t>(n));
    if (rc != ZB_OK) return rc;
    float* h = static_cast<float*>(t.p);
    float* l = h + n;
    const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, t.ctx->sm_count * 16ll));
    plan_note("split_tf32;");
    ZB_KLAUNCH(t.ctx, split_tf32_kernel<<<std::max(grid, 1), 256, 0, t.ctx->stream>>>(x, h, l, n));
    hi = h;
    lo = l;
    return ZB_OK;
  }
};

// call(a_part, b_part, beta, last): `last` carries the bias; the first pass keeps the caller's beta, the others accumulate.
template <typename F>
static int run_tf32x3(zb_ctx* ctx, const float* a, long long na, const float* b, long long nb, float beta, F&& call) {
  SplitOperand sa(ctx), sb(ctx);
  int rc;
  // The tensor core accumulates into TMEM with truncation: over thousands of MMA steps that bias alone reaches 1e-5.
  // Chains are cut every 16 K blocks (64 MMA steps) and joined by round-to-nearest fp32 adds in the epilogue.
  struct ChainScope { ChainScope() { umma_set_chain_limit(16); } ~ChainScope() { umma_set_chain_limit(0); } } chain_scope;
  if ((rc = sa.make(a, na)) != ZB_OK) return rc;
  if ((rc = sb.make(b, nb)) != ZB_OK) return rc;
  if ((rc = call(sa.lo, sb.hi, beta, false)) != ZB_OK) return rc;   // ZB_ERR_UNSUPPORTED surfaces before anything is written
  if ((rc = call(sa.hi, sb.lo, 1.f, false)) != ZB_OK) return rc;
  return call(sa.hi, sb.hi, 1.f, true);
}

static int tc_gemm(zb_ctx* ctx, int mm, bool ta, bool tb, long long m, long long n, long long k, float alpha, const float* a,
                   long long lda, const float* b, long long ldb, float beta, float* c, long long ldc, const float* bias) {
  if (mm != ZB_MATH_TF32X3) return umma_gemm(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, bias);
  const long long na = ((ta ? k : m) - 1) * lda + (ta ? m : k), nb = ((tb ? n : k) - 1) * ldb + (tb ? k : n);
  return run_tf32x3(ctx, a, na, b, nb, beta, [&](const float* ap, const float* bp, float bt, bool last) {
    return umma_gemm(ctx, ta, tb, m, n, k, alpha, ap, lda, bp, ldb, bt, c, ldc, last ? bias : nullptr);
  });
}

struct BnStats {   // fused BatchNorm statistics request (zb_conv2d_fprop_bnstats); rows stays 0 when the path cannot fuse them
  const float* shift = nullptr;
  float* partial = nullptr;
  int rows = 0;
};

static int tc_fprop_nhwc(zb_ctx* ctx, int mm, const zb_conv2d_desc* d, const float* x, const float* w, const float* bias, float* y,
                         BnStats* bs = nullptr) {
  if (mm != ZB_MATH_TF32X3)
    return umma_conv_fprop_nhwc(ctx, d, x, w, bias, y, 0.f, bs ? bs->shift : nullptr, bs ? bs->partial : nullptr, bs ? &bs->rows : nullptr);
  return run_tf32x3(ctx, x, d->n * d->h * d->w * d->c, w, d->k * d->kh * d->kw * d->c, 0.f,
                    [&](const float* xp, const float* wp, float bt, bool last) {
                      return umma_conv_fprop_nhwc(ctx, d, xp, wp, last ? bias : nullptr, y, bt, nullptr, nullptr, nullptr);
                    });
}

static int tc_smallc_fprop(zb_ctx* ctx, int mm, const zb_conv2d_desc* d, const float* x, int x_nchw, const float* w, const float* bias,
                           float* y, BnStats* bs = nullptr) {
  if (mm != ZB_MATH_TF32X3)
    return umma_conv_smallc_fprop(ctx, d, x, x_nchw, w, bias, y, 0.f, bs ? bs->shift : nullptr, bs ? bs->partial : nullptr,
                                  bs ? &bs->rows : nullptr);
  return run_tf32x3(ctx, x, d->n * d->h * d->w * d->c, w, d->k * d->kh * d->kw * d->c, 0.f,
                    [&](const float* xp, const float* wp, float bt, bool last) {
                      return umma_conv_smallc_fprop(ctx, d, xp, x_nchw, wp, last ? bias : nullptr, y, bt, nullptr, nullptr, nullptr);
                    });
}

static int tc_dgrad_nhwc(zb_ctx* ctx, int mm, const zb_conv2d_desc* d, long long P, long long Q, const float* dy, const float* w,
                         float* dx, float beta, bool smallc, const uint32_t* old_bits = nullptr) {
  if (smallc && old_bits != nullptr) { set_last_error("dgrad: masked accumulate is not served by the small-C kernels"); return ZB_ERR_UNSUPPORTED; }
  if (mm != ZB_MATH_TF32X3)
    return smallc ? umma_conv_smallc_dgrad(ctx, d, dy, w, dx, beta) : umma_conv_dgrad_nhwc(ctx, d, dy, w, dx, beta, old_bits);
  int pass = 0;   // the mask belongs to the caller's old dx: first pass only, the other two accumulate onto complete values
  return run_tf32x3(ctx, dy, d->n * P * Q * d->k, w, d->k * d->kh * d->kw * d->c, beta,
                    [&](const float* gp, const float* wp, float bt, bool) {
                      const uint32_t* ob = pass++ == 0 ? old_bits : nullptr;
                      return smallc ? umma_conv_smallc_dgrad(ctx, d, gp, wp, dx, bt) : umma_conv_dgrad_nhwc(ctx, d, gp, wp, dx, bt, ob);
                    });
}

static int tc_wgrad_nhwc(zb_ctx* ctx, int mm, const zb_conv2d_desc* d, long long P, long long Q, const float* dy, const float* x,
                         int x_nchw, float* dw, bool smallc) {
  if (mm != ZB_MATH_TF32X3)
    return smallc ? umma_conv_smallc_wgrad(ctx, d, dy, x, x_nchw, dw, 0.f) : umma_conv_wgrad_nhwc(ctx, d, dy, x, dw, 0.f);
  return run_tf32x3(ctx, dy, d->n * P * Q * d->k, x, d->n * d->h * d->w * d->c, 0.f,
                    [&](const float* gp, const float* xp, float bt, bool) {
                      return smallc ? umma_conv_smallc_wgrad(ctx, d, gp, xp, x_nchw, dw, bt) : umma_conv_wgrad_nhwc(ctx, d, gp, xp, dw, bt);
                    });
}

}  // namespace zb

using namespace zb;

extern "C" {

static int fprop_impl(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* x, const void* w,
                      const void* bias, void* y, BnStats* bs);

int zb_conv2d_fprop(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* x, const void* w,
                    const void* bias, void* y) {
  ZB_API_RANGE();
  return fprop_impl(ctx, dtype, layout, math, d, x, w, bias, y, nullptr);
}

int zb_conv2d_bnstats_rows(zb_ctx* ctx) { return ctx->sm_count * 4; }

int zb_conv2d_fprop_bnstats(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* x, const void* w,
                            const void* bias, void* y, const void* shift, void* stat_partial, int64_t* stat_rows) {
  ZB_API_RANGE();
  ZB_REQUIRE(shift != nullptr && stat_partial != nullptr && stat_rows != nullptr, "conv fprop + bn stats: NULL statistics argument");
  BnStats bs;
  bs.shift = static_cast<const float*>(shift);
  bs.partial = static_cast<float*>(stat_partial);
  const int rc = fprop_impl(ctx, dtype, layout, math, d, x, w, bias, y, dtype == ZB_F32 ? &bs : nullptr);
  *stat_rows = bs.rows;
  return rc;
}

static int fprop_impl(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* x, const void* w,
                      const void* bias, void* y, BnStats* bs) {
  long long P, Q;
  int rc = check_desc(d, &P, &Q);
  if (rc != ZB_OK) return rc;
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC || layout == ZB_NCHW_X, "conv: unknown layout %d", layout);
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "conv: unknown dtype %d", dtype);
  int m;
  rc = resolve_math(ctx, dtype, math, &m);
  if (rc != ZB_OK) return rc;
  if (layout == ZB_NCHW_X) {
    if (dtype == ZB_F32 && m != ZB_MATH_FP32 && umma_conv_smallc_supported(d))
      return tc_smallc_fprop(ctx, m, d, static_cast<const float*>(x), 1, static_cast<const float*>(w), static_cast<const float*>(bias),
                             static_cast<float*>(y), bs);
    set_last_error("conv fprop: ZB_NCHW_X is served for C <= 4 on the TF32 / 3xTF32 paths only");
    return ZB_ERR_UNSUPPORTED;
  }
  if (dtype == ZB_F64)
    return simt_conv_fprop<double>(ctx, layout, d, static_cast<const double*>(x), static_cast<const double*>(w),
                                   static_cast<const double*>(bias), static_cast<double*>(y));
  const float* xf = static_cast<const float*>(x);
  const float* wf = static_cast<const float*>(w);
  const float* bf = static_cast<const float*>(bias);
  float* yf = static_cast<float*>(y);
  if (m != ZB_MATH_FP32 && layout == ZB_NHWC && umma_conv_smallc_supported(d))  // C <= 4 (network stems): sliding-window path
    return tc_smallc_fprop(ctx, m, d, xf, 0, wf, bf, yf, bs);
  if (m != ZB_MATH_FP32 && layout == ZB_NCHW && umma_conv1x1_nchw_supported(d, 0)) {
    // reference contract, pointwise conv: batched per-image GEMMs straight on the NCHW tensors (no layout staging)
    if (m == ZB_MATH_TF32X3)
      rc = run_tf32x3(ctx, xf, d->n * d->c * d->h * d->w, wf, d->k * d->c, 0.f,
                      [&](const float* xp, const float* wp, float bt, bool) { return umma_conv1x1_nchw_fprop(ctx, d, xp, wp, yf, bt); });
    else
      rc = umma_conv1x1_nchw_fprop(ctx, d, xf, wf, yf, 0.f);
    if (rc != ZB_OK || bf == nullptr) return rc;
    return bias_add_nchw<float>(ctx, yf, bf, yf, d->n, d->k, P * Q);
  }
  if (m != ZB_MATH_FP32 && layout == ZB_NCHW && umma_conv_smallc_supported(d)) {
    // reference contract with C <= 4 (a network's first layer): the sliding-window kernels read the NCHW batch as it is; only the
    // (tiny) filter and the output are staged
    Temp tw(ctx), ty(ctx);
    if ((rc = tw.alloc(sizeof(float) * d->k * d->c * d->kh * d->kw)) != ZB_OK) return rc;
    if ((rc = ty.alloc(sizeof(float) * d->n * d->k * P * Q)) != ZB_OK) return rc;
    if ((rc = transpose_batched<float>(ctx, wf, static_cast<float*>(tw.p), d->k, d->c, d->kh * d->kw)) != ZB_OK) return rc;
    if ((rc = tc_smallc_fprop(ctx, m, d, xf, 1, static_cast<float*>(tw.p), bf, static_cast<float*>(ty.p))) != ZB_OK) return rc;
    return transpose_batched<float>(ctx, static_cast<float*>(ty.p), yf, d->n, P * Q, d->k);
  }
  if (m == ZB_MATH_FP32 || !umma_conv_supported(d)) return simt_conv_fprop<float>(ctx, layout, d, xf, wf, bf, yf);
  if (layout == ZB_NHWC) return tc_fprop_nhwc(ctx, m, d, xf, wf, bf, yf, bs);
  // NCHW contract: stage through NHWC / KRSC
  Temp tx(ctx), tw(ctx), ty(ctx);
  if ((rc = tx.alloc(sizeof(float) * d->n * d->c * d->h * d->w)) != ZB_OK) return rc;
  if ((rc = tw.alloc(sizeof(float) * d->k * d->c * d->kh * d->kw)) != ZB_OK) return rc;
  if ((rc = ty.alloc(sizeof(float) * d->n * d->k * P * Q)) != ZB_OK) return rc;
  if ((rc = transpose_batched<float>(ctx, xf, static_cast<float*>(tx.p), d->n, d->c, d->h * d->w)) != ZB_OK) return rc;
  if ((rc = transpose_batched<float>(ctx, wf, static_cast<float*>(tw.p), d->k, d->c, d->kh * d->kw)) != ZB_OK) return rc;
  if ((rc = tc_fprop_nhwc(ctx, m, d, static_cast<float*>(tx.p), static_cast<float*>(tw.p), bf, static_cast<float*>(ty.p))) != ZB_OK) return rc;
  return transpose_batched<float>(ctx, static_cast<float*>(ty.p), yf, d->n, P * Q, d->k);
}

static int dgrad_impl(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy, const void* w,
                      void* dx, float beta, const uint32_t* old_bits = nullptr);

int zb_conv2d_dgrad(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy, const void* w,
                    void* dx) {
  ZB_API_RANGE();
  return dgrad_impl(ctx, dtype, layout, math, d, dy, w, dx, 0.f);
}

int zb_mask_apply(zb_ctx* ctx, int dtype, const void* x, const void* bits, void* out, int64_t n);

int zb_conv2d_dgrad_acc_masked(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy, const void* w,
                               void* dx, const void* old_bits) {
  ZB_API_RANGE();
  if (old_bits == nullptr) return zb_conv2d_dgrad_acc(ctx, dtype, layout, math, d, dy, w, dx);
  long long P, Q;
  int rc = check_desc(d, &P, &Q);
  if (rc != ZB_OK) return rc;
  int m;
  rc = resolve_math(ctx, dtype, math, &m);
  if (rc != ZB_OK) return rc;
  const bool fused = dtype == ZB_F32 && layout == ZB_NHWC && m != ZB_MATH_FP32 && (d->k % 32 == 0) && (d->c % 32 == 0) && d->kh * d->kw <= 64 &&
                     !ZB_ENV_FLAG("ZENU_B200_NO_MASKED_ACC");
  if (fused) {
    rc = dgrad_impl(ctx, dtype, layout, math, d, dy, w, dx, 1.f, static_cast<const uint32_t*>(old_bits));
    if (rc != ZB_ERR_UNSUPPORTED) return rc;
  }
  // paths without the masking epilogue: materialise the masked old value in place, then the plain accumulate
  if ((rc = zb_mask_apply(ctx, dtype, dx, old_bits, dx, d->n * d->c * d->h * d->w)) != ZB_OK) return rc;
  return zb_conv2d_dgrad_acc(ctx, dtype, layout, math, d, dy, w, dx);
}

int zb_conv2d_dgrad_acc(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy, const void* w,
                        void* dx) {
  ZB_API_RANGE();
  long long P, Q;
  int rc = check_desc(d, &P, &Q);
  if (rc != ZB_OK) return rc;
  int m;
  rc = resolve_math(ctx, dtype, math, &m);
  if (rc != ZB_OK) return rc;
  const bool fused = dtype == ZB_F32 && layout == ZB_NHWC && m != ZB_MATH_FP32 && (d->k % 32 == 0) && d->kh * d->kw <= 64;
  if (fused) {
    rc = dgrad_impl(ctx, dtype, layout, math, d, dy, w, dx, 1.f);
    if (rc != ZB_ERR_UNSUPPORTED) return rc;
  }
  // paths without an accumulating epilogue: compute into a temporary, then add
  Temp t(ctx);
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  const long long n = d->n * d->c * d->h * d->w;
  if ((rc = t.alloc(esz * n)) != ZB_OK) return rc;
  if ((rc = dgrad_impl(ctx, dtype, layout, math, d, dy, w, t.p, 0.f)) != ZB_OK) return rc;
  return zb_binary(ctx, dtype, ZB_OP_ADD, dx, t.p, dx, n);
}

static int dgrad_impl(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy, const void* w,
                      void* dx, float beta, const uint32_t* old_bits) {
  long long P, Q;
  int rc = check_desc(d, &P, &Q);
  if (rc != ZB_OK) return rc;
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC, "conv: unknown layout %d", layout);
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "conv: unknown dtype %d", dtype);
  int m;
  rc = resolve_math(ctx, dtype, math, &m);
  if (rc != ZB_OK) return rc;
  if (dtype == ZB_F64)
    return simt_conv_dgrad<double>(ctx, layout, d, static_cast<const double*>(dy), static_cast<const double*>(w), static_cast<double*>(dx));
  const float* gf = static_cast<const float*>(dy);
  const float* wf = static_cast<const float*>(w);
  float* df = static_cast<float*>(dx);
  const bool tc_ok = (d->k % 32 == 0) && d->kh * d->kw <= 64 && (layout == ZB_NHWC || d->c % 4 == 0 || umma_conv_smallc_dgrad_supported(d));
  if (beta != 0.f && !(layout == ZB_NHWC && m != ZB_MATH_FP32 && tc_ok)) {
    set_last_error("dgrad accumulate: not available on this path");
    return ZB_ERR_UNSUPPORTED;
  }
  if (m == ZB_MATH_FP32 || !tc_ok) return simt_conv_dgrad<float>(ctx, layout, d, gf, wf, df);
  if (layout == ZB_NHWC && beta == 0.f && umma_conv_smallc_dgrad_supported(d)) return tc_dgrad_nhwc(ctx, m, d, P, Q, gf, wf, df, 0.f, true);
  if (layout == ZB_NHWC) {
    rc = tc_dgrad_nhwc(ctx, m, d, P, Q, gf, wf, df, beta, false, old_bits);
    if (rc == ZB_ERR_UNSUPPORTED && beta == 0.f) return simt_conv_dgrad<float>(ctx, layout, d, gf, wf, df);
    return rc;
  }
  if (umma_conv1x1_nchw_supported(d, 1)) {   // pointwise conv on the reference's NCHW tensors: no layout staging
    if (m == ZB_MATH_TF32X3)
      return run_tf32x3(ctx, gf, d->n * d->k * P * Q, wf, d->k * d->c, 0.f,
                        [&](const float* gp, const float* wp, float bt, bool) { return umma_conv1x1_nchw_dgrad(ctx, d, gp, wp, df, bt); });
    return umma_conv1x1_nchw_dgrad(ctx, d, gf, wf, df, 0.f);
  }
  Temp tg(ctx), tw(ctx), td(ctx);
  if ((rc = tg.alloc(sizeof(float) * d->n * d->k * P * Q)) != ZB_OK) return rc;
  if ((rc = tw.alloc(sizeof(float) * d->k * d->c * d->kh * d->kw)) != ZB_OK) return rc;
  if ((rc = td.alloc(sizeof(float) * d->n * d->c * d->h * d->w)) != ZB_OK) return rc;
  if ((rc = transpose_batched<float>(ctx, gf, static_cast<float*>(tg.p), d->n, d->k, P * Q)) != ZB_OK) return rc;
  if ((rc = transpose_batched<float>(ctx, wf, static_cast<float*>(tw.p), d->k, d->c, d->kh * d->kw)) != ZB_OK) return rc;
  rc = tc_dgrad_nhwc(ctx, m, d, P, Q, static_cast<float*>(tg.p), static_cast<float*>(tw.p), static_cast<float*>(td.p), 0.f,
                     umma_conv_smallc_dgrad_supported(d));
  if (rc == ZB_ERR_UNSUPPORTED) return simt_conv_dgrad<float>(ctx, layout, d, gf, wf, df);
  if (rc != ZB_OK) return rc;
  return transpose_batched<float>(ctx, static_cast<float*>(td.p), df, d->n, d->h * d->w, d->c);
}

int zb_conv2d_wgrad(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy, const void* x,
                    void* dw) {
  ZB_API_RANGE();
  long long P, Q;
  int rc = check_desc(d, &P, &Q);
  if (rc != ZB_OK) return rc;
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC || layout == ZB_NCHW_X, "conv: unknown layout %d", layout);
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "conv: unknown dtype %d", dtype);
  int m;
  rc = resolve_math(ctx, dtype, math, &m);
  if (rc != ZB_OK) return rc;
  if (layout == ZB_NCHW_X) {
    if (dtype == ZB_F32 && m != ZB_MATH_FP32 && umma_conv_smallc_supported(d))
      return tc_wgrad_nhwc(ctx, m, d, P, Q, static_cast<const float*>(dy), static_cast<const float*>(x), 1, static_cast<float*>(dw), true);
    set_last_error("conv wgrad: ZB_NCHW_X is served for C <= 4 on the TF32 / 3xTF32 paths only");
    return ZB_ERR_UNSUPPORTED;
  }
  if (dtype == ZB_F64)
    return simt_conv_wgrad<double>(ctx, layout, d, static_cast<const double*>(dy), static_cast<const double*>(x), static_cast<double*>(dw));
  const float* gf = static_cast<const float*>(dy);
  const float* xf = static_cast<const float*>(x);
  float* wf = static_cast<float*>(dw);
  if (m != ZB_MATH_FP32 && layout == ZB_NHWC && umma_conv_smallc_supported(d)) return tc_wgrad_nhwc(ctx, m, d, P, Q, gf, xf, 0, wf, true);
  if (m != ZB_MATH_FP32 && layout == ZB_NCHW && umma_conv1x1_nchw_supported(d, 2)) {   // pointwise conv, NCHW tensors as they are
    if (m == ZB_MATH_TF32X3)
      return run_tf32x3(ctx, gf, d->n * d->k * P * Q, xf, d->n * d->c * d->h * d->w, 0.f,
                        [&](const float* gp, const float* xp, float bt, bool) { return umma_conv1x1_nchw_wgrad(ctx, d, gp, xp, wf, bt); });
    return umma_conv1x1_nchw_wgrad(ctx, d, gf, xf, wf, 0.f);
  }
  if (m != ZB_MATH_FP32 && layout == ZB_NCHW && umma_conv_smallc_supported(d)) {   // x stays NCHW; dy and the (tiny) dw are staged
    Temp tg(ctx), tw(ctx);
    if ((rc = tg.alloc(sizeof(float) * d->n * d->k * P * Q)) != ZB_OK) return rc;
    if ((rc = tw.alloc(sizeof(float) * d->k * d->c * d->kh * d->kw)) != ZB_OK) return rc;
    if ((rc = transpose_batched<float>(ctx, gf, static_cast<float*>(tg.p), d->n, d->k, P * Q)) != ZB_OK) return rc;
    if ((rc = tc_wgrad_nhwc(ctx, m, d, P, Q, static_cast<float*>(tg.p), xf, 1, static_cast<float*>(tw.p), true)) != ZB_OK) return rc;
    return transpose_batched<float>(ctx, static_cast<float*>(tw.p), wf, d->k, d->kh * d->kw, d->c);  // KRSC -> KCRS
  }
  if (m == ZB_MATH_FP32 || !umma_conv_supported(d)) return simt_conv_wgrad<float>(ctx, layout, d, gf, xf, wf);
  if (layout == ZB_NHWC) return tc_wgrad_nhwc(ctx, m, d, P, Q, gf, xf, 0, wf, false);
  Temp tg(ctx), tx(ctx), tw(ctx);
  if ((rc = tg.alloc(sizeof(float) * d->n * d->k * P * Q)) != ZB_OK) return rc;
  if ((rc = tx.alloc(sizeof(float) * d->n * d->c * d->h * d->w)) != ZB_OK) return rc;
  if ((rc = tw.alloc(sizeof(float) * d->k * d->c * d->kh * d->kw)) != ZB_OK) return rc;
  if ((rc = transpose_batched<float>(ctx, gf, static_cast<float*>(tg.p), d->n, d->k, P * Q)) != ZB_OK) return rc;
  if ((rc = transpose_batched<float>(ctx, xf, static_cast<float*>(tx.p), d->n, d->c, d->h * d->w)) != ZB_OK) return rc;
  if ((rc = tc_wgrad_nhwc(ctx, m, d, P, Q, static_cast<float*>(tg.p), static_cast<float*>(tx.p), 0, static_cast<float*>(tw.p), false)) != ZB_OK) return rc;
  return transpose_batched<float>(ctx, static_cast<float*>(tw.p), wf, d->k, d->kh * d->kw, d->c);  // KRSC -> KCRS
}

// ---- plan trace ---------------------------------------------------------------------------------------------------------
static int64_t copy_trace(const std::string& text, char* buf, int64_t cap) {
  if (buf != nullptr && cap > 0) {
    const size_t n = std::min<size_t>(text.size(), static_cast<size_t>(cap - 1));
    memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return static_cast<int64_t>(text.size()) + 1;
}

int64_t zb_conv2d_plan_describe(zb_ctx* ctx, int op, int dtype, int layout, int math, const zb_conv2d_desc* d, int flags, char* buf,
                                int64_t cap) {
  if (ctx == nullptr || d == nullptr) { zb::set_last_error("plan_describe: NULL argument"); return -1; }
  PlanTrace tr;
  tr.dry = true;
  PlanTrace* prev = tl_plan;
  tl_plan = &tr;
  // Fake operand addresses (16-byte aligned, never dereferenced: nothing is launched in dry mode).  The real entry points run, so the
  // description cannot drift from what a call does.
  void* const A = reinterpret_cast<void*>(uintptr_t(5) << 40);
  void* const B = reinterpret_cast<void*>(uintptr_t(6) << 40);
  void* const C = reinterpret_cast<void*>(uintptr_t(7) << 40);
  void* const S = reinterpret_cast<void*>(uintptr_t(9) << 40);
  void* const T = reinterpret_cast<void*>(uintptr_t(10) << 40);
  int rc;
  int64_t rows = 0;
  switch (op) {
    case ZB_PLAN_FPROP:
      if (flags & ZB_PLAN_BNSTATS) rc = zb_conv2d_fprop_bnstats(ctx, dtype, layout, math, d, A, B, (flags & ZB_PLAN_BIAS) ? S : nullptr, C, S, T, &rows);
      else rc = zb_conv2d_fprop(ctx, dtype, layout, math, d, A, B, (flags & ZB_PLAN_BIAS) ? S : nullptr, C);
      break;
    case ZB_PLAN_DGRAD:
      rc = (flags & ZB_PLAN_ACCUMULATE) ? zb_conv2d_dgrad_acc(ctx, dtype, layout, math, d, A, B, C) : zb_conv2d_dgrad(ctx, dtype, layout, math, d, A, B, C);
      break;
    case ZB_PLAN_WGRAD: rc = zb_conv2d_wgrad(ctx, dtype, layout, math, d, A, B, C); break;
    default: zb::set_last_error("plan_describe: unknown op %d", op); rc = ZB_ERR_INVALID; break;
  }
  tl_plan = prev;
  if (rc != ZB_OK) return -static_cast<int64_t>(rc);
  return copy_trace(tr.text, buf, cap);
}

int zb_ctx_plan_trace(zb_ctx* ctx, int enable) {
  ZB_REQUIRE(ctx != nullptr, "plan_trace: ctx is NULL");
  if (enable) {
    if (tl_plan == nullptr) tl_plan = new PlanTrace();
    tl_plan->dry = false;
    tl_plan->text.clear();
  } else if (tl_plan != nullptr && !tl_plan->dry) {
    delete tl_plan;
    tl_plan = nullptr;
  }
  return ZB_OK;
}

int64_t zb_ctx_plan_trace_read(zb_ctx* ctx, char* buf, int64_t cap) {
  if (ctx == nullptr || tl_plan == nullptr) { if (buf && cap > 0) buf[0] = 0; return 1; }
  const int64_t need = copy_trace(tl_plan->text, buf, cap);
  if (buf != nullptr) tl_plan->text.clear();
  return need;
}

int zb_conv2d_bias_add(zb_ctx* ctx, int dtype, int layout, const void* x, const void* bias, void* y, int64_t n, int64_t k,
                       int64_t h, int64_t w) {
  ZB_API_RANGE();
  if (layout == ZB_NHWC) return zb_binary_bcast_rows(ctx, dtype, ZB_OP_ADD, x, bias, y, n * h * w, k);
  ZB_REQUIRE(layout == ZB_NCHW, "bias_add: unknown layout %d", layout);
  if (dtype == ZB_F32) return bias_add_nchw<float>(ctx, static_cast<const float*>(x), static_cast<const float*>(bias), static_cast<float*>(y), n, k, h * w);
  if (dtype == ZB_F64) return bias_add_nchw<double>(ctx, static_cast<const double*>(x), static_cast<const double*>(bias), static_cast<double*>(y), n, k, h * w);
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

int zb_gemm(zb_ctx* ctx, int dtype, int math, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, double alpha,
            const void* a, int64_t lda, const void* b, int64_t ldb, double beta, void* c, int64_t ldc) {
  ZB_API_RANGE();
  ZB_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm: negative extent");
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "gemm: unknown dtype %d", dtype);
  if (m == 0 || n == 0) return ZB_OK;
  ZB_REQUIRE(lda >= (trans_a ? m : k) && ldb >= (trans_b ? k : n) && ldc >= n, "gemm: leading dimension too small");
  int mm;
  int rc = resolve_math(ctx, dtype, math, &mm);
  if (rc != ZB_OK) return rc;
  if (dtype == ZB_F64)
    return simt_gemm<double>(ctx, trans_a != 0, trans_b != 0, m, n, k, alpha, static_cast<const double*>(a), lda,
                             static_cast<const double*>(b), ldb, beta, static_cast<double*>(c), ldc, static_cast<const double*>(nullptr));
  if (mm != ZB_MATH_FP32 && k > 0) {
    rc = tc_gemm(ctx, mm, trans_a != 0, trans_b != 0, m, n, k, static_cast<float>(alpha), static_cast<const float*>(a), lda,
                 static_cast<const float*>(b), ldb, static_cast<float>(beta), static_cast<float*>(c), ldc, nullptr);
    if (rc != ZB_ERR_UNSUPPORTED) return rc;
  }
  return simt_gemm<float>(ctx, trans_a != 0, trans_b != 0, m, n, k, static_cast<float>(alpha), static_cast<const float*>(a), lda,
                          static_cast<const float*>(b), ldb, static_cast<float>(beta), static_cast<float*>(c), ldc, static_cast<const float*>(nullptr));
}

int zb_linear_fwd(zb_ctx* ctx, int dtype, int math, const void* x, const void* w, const void* bias, void* y, int64_t batch,
                  int64_t in_f, int64_t out_f) {
  ZB_API_RANGE();
  ZB_REQUIRE(batch > 0 && in_f > 0 && out_f > 0, "linear: non-positive extent");
  int mm;
  int rc = resolve_math(ctx, dtype, math, &mm);
  if (rc != ZB_OK) return rc;
  if (dtype == ZB_F64)
    return simt_gemm<double>(ctx, false, true, batch, out_f, in_f, 1.0, static_cast<const double*>(x), in_f,
                             static_cast<const double*>(w), in_f, 0.0, static_cast<double*>(y), out_f, static_cast<const double*>(bias));
  if (mm != ZB_MATH_FP32) {
    rc = tc_gemm(ctx, mm, false, true, batch, out_f, in_f, 1.f, static_cast<const float*>(x), in_f, static_cast<const float*>(w), in_f,
                 0.f, static_cast<float*>(y), out_f, static_cast<const float*>(bias));
    if (rc != ZB_ERR_UNSUPPORTED) return rc;
  }
  return simt_gemm<float>(ctx, false, true, batch, out_f, in_f, 1.f, static_cast<const float*>(x), in_f, static_cast<const float*>(w), in_f,
                          0.f, static_cast<float*>(y), out_f, static_cast<const float*>(bias));
}

int zb_linear_bwd(zb_ctx* ctx, int dtype, int math, const void* x, const void* w, const void* dy, void* dx, void* dw,
                  void* dbias, int64_t batch, int64_t in_f, int64_t out_f) {
  ZB_API_RANGE();
  ZB_REQUIRE(batch > 0 && in_f > 0 && out_f > 0, "linear: non-positive extent");
  int rc;
  if (dx) {  // dX[b,in] = dY[b,out] * W[out,in]
    rc = zb_gemm(ctx, dtype, math, 0, 0, batch, in_f, out_f, 1.0, dy, out_f, w, in_f, 0.0, dx, in_f);
    if (rc != ZB_OK) return rc;
  }
  if (dw) {  // dW[out,in] = dY^T[out,b] * X[b,in]
    rc = zb_gemm(ctx, dtype, math, 1, 0, out_f, in_f, batch, 1.0, dy, out_f, x, in_f, 0.0, dw, in_f);
    if (rc != ZB_OK) return rc;
  }
  if (dbias) {
    rc = zb_sum_rows(ctx, dtype, dy, dbias, batch, out_f);
    if (rc != ZB_OK) return rc;
  }
  return ZB_OK;
}

}  // extern "C"
