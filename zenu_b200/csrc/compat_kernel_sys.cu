// compat_kernel_sys.cu — drop-in symbols of zenu-cuda-kernel-sys (see include/zenu_kernel_compat.h).
// Unit-stride calls are forwarded to the vectorised native kernels through a process-wide context bound to
// the legacy default stream (so un-ported reference ops interleave correctly); other strides use the strided
// kernels below.  Reference: zenu-cuda-kernel-sys/kernel/*.cu.
#include <cmath>

#include "../../include/zenu_kernel_compat.h"
#include "common.cuh"

namespace zb {

zb_ctx* compat_ctx() {
  static zb_ctx* ctx = nullptr;
  if (!ctx) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (zb_ctx_create(&ctx, dev, cudaStreamLegacy) != ZB_OK) {
      fprintf(stderr, "zenu_b200 compat: cannot create context: %s\n", zb_last_error());
      ctx = nullptr;
    }
  }
  return ctx;
}

template <typename T, typename F>
__global__ void strided_unary(const T* a, int sa, T* out, int so, int n, F f) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[static_cast<long long>(i) * so] = f(a[static_cast<long long>(i) * sa]);
}
template <typename T, typename F>
__global__ void strided_binary(const T* a, int sa, const T* b, int sb, T* c, int sc, int n, F f) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    c[static_cast<long long>(i) * sc] = f(a[static_cast<long long>(i) * sa], b[static_cast<long long>(i) * sb]);
}
template <typename T, typename F>
__global__ void strided_scalar_ptr(const T* a, int sa, const T* s, T* out, int so, int n, F f) {
  const T sv = *s;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[static_cast<long long>(i) * so] = f(a[static_cast<long long>(i) * sa], sv);
}
template <typename T>
__global__ void argmax_kernel(const T* a, int n, int stride, int* out) {
  __shared__ T sv[256];
  __shared__ int si[256];
  T best = T(0);
  int bi = -1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const T v = a[static_cast<long long>(i) * stride];
    if (bi < 0 || v > best) { best = v; bi = i; }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if (threadIdx.x < h) {
      const int oi = si[threadIdx.x + h];
      const T ov = sv[threadIdx.x + h];
      const int mi = si[threadIdx.x];
      if (oi >= 0 && (mi < 0 || ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < mi))) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = si[0] < 0 ? 0 : si[0];
}

static inline int grid_for(int n) { return n <= 0 ? 1 : (n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256; }

template <typename T, typename F>
void run_unary(const T* a, int sa, T* out, int so, int n, F f) {
  if (n <= 0) return;
  strided_unary<T, F><<<grid_for(n), 256>>>(a, sa, out, so, n, f);
}
template <typename T, typename F>
void run_binary(const T* a, int sa, const T* b, int sb, T* c, int sc, int n, F f) {
  if (n <= 0) return;
  strided_binary<T, F><<<grid_for(n), 256>>>(a, sa, b, sb, c, sc, n, f);
}
template <typename T, typename F>
void run_scalar_ptr(const T* a, int sa, const T* s, T* out, int so, int n, F f) {
  if (n <= 0) return;
  strided_scalar_ptr<T, F><<<grid_for(n), 256>>>(a, sa, s, out, so, n, f);
}

template <typename T> constexpr int dt();
template <> constexpr int dt<float>() { return ZB_F32; }
template <> constexpr int dt<double>() { return ZB_F64; }

template <typename T, int OP> struct OpF { __device__ T operator()(T a, T b) const { return OP == 0 ? a + b : OP == 1 ? a - b : OP == 2 ? a * b : a / b; } };
template <typename T, int OP> struct OpSF { T s; __device__ T operator()(T a) const { return OP == 0 ? a + s : OP == 1 ? a - s : OP == 2 ? a * s : a / s; } };

template <typename T, int OP>
void aa(T* a, int sa, T* b, int sb, T* c, int sc, int n) {
  zb_ctx* ctx = compat_ctx();
  if (ctx && sa == 1 && sb == 1 && sc == 1) { zb_binary(ctx, dt<T>(), OP, a, b, c, n); return; }
  run_binary<T>(a, sa, b, sb, c, sc, n, OpF<T, OP>{});
}
template <typename T, int OP>
void as(T* a, int n, int sa, T s, T* out, int so) {
  zb_ctx* ctx = compat_ctx();
  if (ctx && sa == 1 && so == 1) { zb_binary_scalar(ctx, dt<T>(), OP, a, static_cast<double>(s), out, n); return; }
  run_unary<T>(a, sa, out, so, n, OpSF<T, OP>{s});
}
template <typename T, int OP>
void asp(T* a, int n, int sa, T* s, T* out, int so) { run_scalar_ptr<T>(a, sa, s, out, so, n, OpF<T, OP>{}); }

template <typename T> struct ReluS { T alpha; __device__ T operator()(T x) const { return x > T(0) ? x : alpha * x; } };
template <typename T> struct ReluMaskS { T alpha; __device__ T operator()(T x) const { return x > T(0) ? T(1) : alpha * T(-1); } };
template <typename T> struct ClipS { T lo, hi; __device__ T operator()(T x) const { return max(min(x, hi), lo); } };
template <typename T> struct ClipMaskS { T lo, hi; __device__ T operator()(T x) const { return (x >= lo && x <= hi) ? T(1) : T(0); } };
template <typename T> struct PowS { T e; __device__ T operator()(T x) const { return pow(x, e); } };

template <typename T>
void relu_c(T* in, T* out, T alpha, int n, int si, int so) {
  zb_ctx* ctx = compat_ctx();
  if (ctx && si == 1 && so == 1) { zb_relu(ctx, dt<T>(), in, out, static_cast<double>(alpha), n); return; }
  run_unary<T>(in, si, out, so, n, ReluS<T>{alpha});
}
template <typename T>
void relu_mask_c(T* in, T* out, T alpha, int n, int si, int so) {
  zb_ctx* ctx = compat_ctx();
  if (ctx && si == 1 && so == 1) { zb_relu_backward_mask(ctx, dt<T>(), in, out, static_cast<double>(alpha), n); return; }
  run_unary<T>(in, si, out, so, n, ReluMaskS<T>{alpha});
}
template <typename T>
void bias_add_c(const T* in, T* out, int channel_stride, const T* bias, int bias_size, int total) {
  // reference: y[i] = x[i] + b[(i / channel_stride) % bias_size]   (array_array.cu:49-58), NCHW with channel_stride = H*W
  zb_ctx* ctx = compat_ctx();
  if (!ctx || channel_stride <= 0 || bias_size <= 0) return;
  const long long n = static_cast<long long>(total) / (static_cast<long long>(channel_stride) * bias_size);
  zb_conv2d_bias_add(ctx, dt<T>(), ZB_NCHW, in, bias, out, n, bias_size, channel_stride, 1);
}
template <typename T>
void bias_bkwd_c(const T* dout, T* dbias, int N, int C, int H, int W) {
  zb_ctx* ctx = compat_ctx();
  if (ctx) zb_conv2d_bias_bwd(ctx, dt<T>(), ZB_NCHW, dout, dbias, N, C, H, W);
}
template <typename T>
void max_idx_c(T* a, int size, int stride, int* host_out) {
  int* d = nullptr;
  cudaMalloc(&d, sizeof(int));
  argmax_kernel<T><<<1, 256>>>(a, size, stride, d);
  cudaMemcpy(host_out, d, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d);
}

}  // namespace zb

using namespace zb;

#define COMPAT_TYPE(T, SFX)                                                                                                      \
  void relu_##SFX(T* i, T* o, T alpha, int n, int si, int so) { relu_c<T>(i, o, alpha, n, si, so); }                            \
  void relu_backward_mask_##SFX(T* i, T* o, T alpha, int n, int si, int so) { relu_mask_c<T>(i, o, alpha, n, si, so); }         \
  void array_array_add_##SFX(T* a, int sa, T* b, int sb, T* c, int sc, int n) { aa<T, 0>(a, sa, b, sb, c, sc, n); }             \
  void array_array_sub_##SFX(T* a, int sa, T* b, int sb, T* c, int sc, int n) { aa<T, 1>(a, sa, b, sb, c, sc, n); }             \
  void array_array_mul_##SFX(T* a, int sa, T* b, int sb, T* c, int sc, int n) { aa<T, 2>(a, sa, b, sb, c, sc, n); }             \
  void array_array_div_##SFX(T* a, int sa, T* b, int sb, T* c, int sc, int n) { aa<T, 3>(a, sa, b, sb, c, sc, n); }             \
  void array_array_add_assign_##SFX(T* a, int sa, T* b, int sb, int n) { aa<T, 0>(a, sa, b, sb, a, sa, n); }                    \
  void array_array_sub_assign_##SFX(T* a, int sa, T* b, int sb, int n) { aa<T, 1>(a, sa, b, sb, a, sa, n); }                    \
  void array_array_mul_assign_##SFX(T* a, int sa, T* b, int sb, int n) { aa<T, 2>(a, sa, b, sb, a, sa, n); }                    \
  void array_array_div_assign_##SFX(T* a, int sa, T* b, int sb, int n) { aa<T, 3>(a, sa, b, sb, a, sa, n); }                    \
  void conv_bias_add_##SFX(const T* i, T* o, int cs, const T* b, int bs, int tot) { bias_add_c<T>(i, o, cs, b, bs, tot); }      \
  void conv2d_bias_bkwd_##SFX(const T* d, T* db, int N, int C, int H, int W) { bias_bkwd_c<T>(d, db, N, C, H, W); }             \
  void array_scalar_add_##SFX(T* a, int n, int sa, T s, T* o, int so) { as<T, 0>(a, n, sa, s, o, so); }                         \
  void array_scalar_sub_##SFX(T* a, int n, int sa, T s, T* o, int so) { as<T, 1>(a, n, sa, s, o, so); }                         \
  void array_scalar_mul_##SFX(T* a, int n, int sa, T s, T* o, int so) { as<T, 2>(a, n, sa, s, o, so); }                         \
  void array_scalar_div_##SFX(T* a, int n, int sa, T s, T* o, int so) { as<T, 3>(a, n, sa, s, o, so); }                         \
  void array_scalar_add_assign_##SFX(T* a, int n, int st, T s) { as<T, 0>(a, n, st, s, a, st); }                                \
  void array_scalar_sub_assign_##SFX(T* a, int n, int st, T s) { as<T, 1>(a, n, st, s, a, st); }                                \
  void array_scalar_mul_assign_##SFX(T* a, int n, int st, T s) { as<T, 2>(a, n, st, s, a, st); }                                \
  void array_scalar_div_assign_##SFX(T* a, int n, int st, T s) { as<T, 3>(a, n, st, s, a, st); }                                \
  void array_scalar_pointer_add_##SFX(T* a, int n, int sa, T* s, T* o, int so) { asp<T, 0>(a, n, sa, s, o, so); }               \
  void array_scalar_pointer_sub_##SFX(T* a, int n, int sa, T* s, T* o, int so) { asp<T, 1>(a, n, sa, s, o, so); }               \
  void array_scalar_pointer_mul_##SFX(T* a, int n, int sa, T* s, T* o, int so) { asp<T, 2>(a, n, sa, s, o, so); }               \
  void array_scalar_pointer_div_##SFX(T* a, int n, int sa, T* s, T* o, int so) { asp<T, 3>(a, n, sa, s, o, so); }               \
  void array_scalar_pointer_add_assign_##SFX(T* a, int n, int st, T* s) { asp<T, 0>(a, n, st, s, a, st); }                      \
  void array_scalar_pointer_sub_assign_##SFX(T* a, int n, int st, T* s) { asp<T, 1>(a, n, st, s, a, st); }                      \
  void array_scalar_pointer_mul_assign_##SFX(T* a, int n, int st, T* s) { asp<T, 2>(a, n, st, s, a, st); }                      \
  void array_scalar_pointer_div_assign_##SFX(T* a, int n, int st, T* s) { asp<T, 3>(a, n, st, s, a, st); }                      \
  void array_clip_##SFX(T* i, T* o, int n, int si, int so, T lo, T hi) { run_unary<T>(i, si, o, so, n, ClipS<T>{lo, hi}); }     \
  void array_clip_assign_##SFX(T* i, int n, int st, T lo, T hi) { run_unary<T>(i, st, i, st, n, ClipS<T>{lo, hi}); }            \
  void array_clip_backward_##SFX(T* i, T* m, T hi, T lo, int n, int si, int sm) { run_unary<T>(i, si, m, sm, n, ClipMaskS<T>{lo, hi}); } \
  void array_clip_backward_assign_##SFX(T* i, T hi, T lo, int n, int st) { run_unary<T>(i, st, i, st, n, ClipMaskS<T>{lo, hi}); } \
  void array_pow_##SFX(T* a, int n, int sa, T s, T* o, int so) { run_unary<T>(a, sa, o, so, n, PowS<T>{s}); }                   \
  void array_pow_assign_##SFX(T* a, int n, int st, T s) { run_unary<T>(a, st, a, st, n, PowS<T>{s}); }                          \
  void memory_access_##SFX(T* array, int offset, T* result) { cudaMemcpy(result, array + offset, sizeof(T), cudaMemcpyDeviceToHost); } \
  void memory_set_##SFX(T* array, int offset, T value) { cudaMemcpy(array + offset, &value, sizeof(T), cudaMemcpyHostToDevice); } \
  void array_max_idx_##SFX(T* a, int size, int stride, int* out) { max_idx_c<T>(a, size, stride, out); }

#define COMPAT_UNARY(NAME, FN)                                                                                          \
  namespace zb { template <typename T> struct NAME##S { __device__ T operator()(T x) const { return FN(x); } }; }      \
  extern "C" {                                                                                                          \
  void array_##NAME##_float(float* a, int n, int si, float* o, int so) { run_unary<float>(a, si, o, so, n, zb::NAME##S<float>{}); }     \
  void array_##NAME##_double(double* a, int n, int si, double* o, int so) { run_unary<double>(a, si, o, so, n, zb::NAME##S<double>{}); } \
  void array_##NAME##_assign_float(float* a, int n, int st) { run_unary<float>(a, st, a, st, n, zb::NAME##S<float>{}); }                \
  void array_##NAME##_assign_double(double* a, int n, int st) { run_unary<double>(a, st, a, st, n, zb::NAME##S<double>{}); }            \
  }

extern "C" {
COMPAT_TYPE(float, float)
COMPAT_TYPE(double, double)
}
COMPAT_UNARY(sin, sin) COMPAT_UNARY(cos, cos) COMPAT_UNARY(tan, tan)
COMPAT_UNARY(asin, asin) COMPAT_UNARY(acos, acos) COMPAT_UNARY(atan, atan)
COMPAT_UNARY(sinh, sinh) COMPAT_UNARY(cosh, cosh) COMPAT_UNARY(tanh, tanh)
COMPAT_UNARY(abs, fabs) COMPAT_UNARY(sqrt, sqrt) COMPAT_UNARY(exp, exp) COMPAT_UNARY(log, log)
