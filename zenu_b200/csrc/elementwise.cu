// elementwise.cu — HBM-bound elementwise, broadcast, copy/layout and optimizer kernels.
// 128-bit vectorised, grid-stride, 4 independent vectors in flight per thread.
// Replaces the reference's one-element-per-thread strided kernels (zenu-cuda-kernel-sys/kernel/
// activations.cu:3-41, array_array.cu:4-74, array_scalar.cu:66-128), cuBLAS scopy/sscal used as
// memcpy/zero-fill (zenu-cuda/src/cublas/mod.rs:36-77,296-336) and the per-tensor multi-kernel optimizer
// updates (zenu-optimizer/src/sgd.rs:20-30, adam.rs:19-58, adamw.rs:20-69).
#include <algorithm>

#include <vector>

#include "common.cuh"

namespace zb {

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int N = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int N = 2; };

template <typename T>
__device__ __forceinline__ void unpack(const typename Vec<T>::type& v, T* o);
template <> __device__ __forceinline__ void unpack<float>(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
template <> __device__ __forceinline__ void unpack<double>(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
template <typename T>
__device__ __forceinline__ typename Vec<T>::type pack(const T* o);
template <> __device__ __forceinline__ float4 pack<float>(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
template <> __device__ __forceinline__ double2 pack<double>(const double* o) { return make_double2(o[0], o[1]); }

// Generic map kernel: out[i] = f(in0[i], in1[i], i).  NIN = number of input streams actually read.
template <typename T, int NIN, bool VECTOR, typename F>
__global__ void __launch_bounds__(256) map_kernel(F f, T* out, const T* in0, const T* in1, long long n) {  // out may alias in0/in1 (no __restrict__)
  constexpr int VN = Vec<T>::N;
  using V = typename Vec<T>::type;
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  if (VECTOR) {
    const long long nvec = n / VN;
    constexpr int U = 4;
    long long i = tid;
    for (; i + (U - 1) * nthreads < nvec; i += U * nthreads) {
      V a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (NIN >= 1) a[u] = reinterpret_cast<const V*>(in0)[i + u * nthreads];
        if (NIN >= 2) b[u] = reinterpret_cast<const V*>(in1)[i + u * nthreads];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        T av[VN], bv[VN], ov[VN];
        if (NIN >= 1) unpack<T>(a[u], av);
        if (NIN >= 2) unpack<T>(b[u], bv);
#pragma unroll
        for (int e = 0; e < VN; ++e) ov[e] = f(NIN >= 1 ? av[e] : T(0), NIN >= 2 ? bv[e] : T(0), (i + u * nthreads) * VN + e);
        reinterpret_cast<V*>(out)[i + u * nthreads] = pack<T>(ov);
      }
    }
    for (; i < nvec; i += nthreads) {
      T av[VN], bv[VN], ov[VN];
      if (NIN >= 1) unpack<T>(reinterpret_cast<const V*>(in0)[i], av);
      if (NIN >= 2) unpack<T>(reinterpret_cast<const V*>(in1)[i], bv);
#pragma unroll
      for (int e = 0; e < VN; ++e) ov[e] = f(NIN >= 1 ? av[e] : T(0), NIN >= 2 ? bv[e] : T(0), i * VN + e);
      reinterpret_cast<V*>(out)[i] = pack<T>(ov);
    }
    for (long long j = nvec * VN + tid; j < n; j += nthreads) out[j] = f(NIN >= 1 ? in0[j] : T(0), NIN >= 2 ? in1[j] : T(0), j);
  } else {
    for (long long j = tid; j < n; j += nthreads) out[j] = f(NIN >= 1 ? in0[j] : T(0), NIN >= 2 ? in1[j] : T(0), j);
  }
}

static inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T, int NIN, typename F>
int launch_map(zb_ctx* ctx, F f, T* out, const T* in0, const T* in1, long long n) {
  if (n <= 0) return ZB_OK;
  const bool vec = aligned16(out) && aligned16(in0) && aligned16(in1);
  const long long work = vec ? (n + Vec<T>::N - 1) / Vec<T>::N : n;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((work + 1023) / 1024, ctx->sm_count * 8ll)));
  if (plan_dry()) { plan_note("map;"); return ZB_OK; }
  if (vec)
    map_kernel<T, NIN, true, F><<<grid, 256, 0, ctx->stream>>>(f, out, in0, in1, n);
  else
    map_kernel<T, NIN, false, F><<<grid, 256, 0, ctx->stream>>>(f, out, in0, in1, n);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// ------------------------------------------------------------------------------------------------ functors
template <typename T> struct ReluF { T alpha; __device__ T operator()(T x, T, long long) const { return x > T(0) ? x : alpha * x; } };
template <typename T> struct ReluMaskF { T alpha; __device__ T operator()(T x, T, long long) const { return x > T(0) ? T(1) : alpha * T(-1); } };
template <typename T> struct ReluBwdF { T alpha; __device__ T operator()(T x, T dy, long long) const { return x > T(0) ? dy : (alpha * T(-1)) * dy; } };
template <typename T, int OP> struct BinF {
  __device__ T operator()(T a, T b, long long) const { return OP == 0 ? a + b : OP == 1 ? a - b : OP == 2 ? a * b : a / b; }
};
template <typename T, int OP> struct BinScalarF {
  T s;
  __device__ T operator()(T a, T, long long) const { return OP == 0 ? a + s : OP == 1 ? a - s : OP == 2 ? a * s : a / s; }
};
template <typename T, int OP> struct BcastRowsF {  // c[r, j] = a[r, j] op b[j]
  const T* b; long long cols;
  __device__ T operator()(T a, T, long long i) const {
    const T bv = __ldg(b + (i % cols));
    return OP == 0 ? a + bv : OP == 1 ? a - bv : OP == 2 ? a * bv : a / bv;
  }
};
template <typename T> struct BiasNchwF {  // y[n,k,hw] = x + b[k]
  const T* b; long long hw, k;
  __device__ T operator()(T a, T, long long i) const { return a + __ldg(b + ((i / hw) % k)); }
};
template <typename T> struct FillF { T v; __device__ T operator()(T, T, long long) const { return v; } };
template <typename T> struct CopyF { __device__ T operator()(T a, T, long long) const { return a; } };
template <typename T> struct MaskBitsF {  // a[i] where bit i of a 1-bit-per-element array is set, else 0
  const uint32_t* bits;
  __device__ T operator()(T a, T, long long i) const { return ((__ldg(bits + (i >> 5)) >> (i & 31)) & 1u) ? a : T(0); }
};
template <typename T> struct SgdF {  // p -= (g * scale) * lr   (sgd.rs:24-27: grad * lr then sub)
  T lr, scale;
  __device__ T operator()(T p, T g, long long) const { return p - (g * scale) * lr; }
};

template <typename T, template <typename, int> class F, typename... Extra>
int dispatch_binop(zb_ctx* ctx, int op, T* out, const T* a, const T* b, long long n, Extra... extra) {
  switch (op) {
    case ZB_OP_ADD: return launch_map<T, 2>(ctx, F<T, 0>{extra...}, out, a, b, n);
    case ZB_OP_SUB: return launch_map<T, 2>(ctx, F<T, 1>{extra...}, out, a, b, n);
    case ZB_OP_MUL: return launch_map<T, 2>(ctx, F<T, 2>{extra...}, out, a, b, n);
    case ZB_OP_DIV: return launch_map<T, 2>(ctx, F<T, 3>{extra...}, out, a, b, n);
  }
  set_last_error("unknown binary op %d", op);
  return ZB_ERR_INVALID;
}
template <typename T, template <typename, int> class F, typename... Extra>
int dispatch_unop(zb_ctx* ctx, int op, T* out, const T* a, long long n, Extra... extra) {
  switch (op) {
    case ZB_OP_ADD: return launch_map<T, 1>(ctx, F<T, 0>{extra...}, out, a, static_cast<const T*>(nullptr), n);
    case ZB_OP_SUB: return launch_map<T, 1>(ctx, F<T, 1>{extra...}, out, a, static_cast<const T*>(nullptr), n);
    case ZB_OP_MUL: return launch_map<T, 1>(ctx, F<T, 2>{extra...}, out, a, static_cast<const T*>(nullptr), n);
    case ZB_OP_DIV: return launch_map<T, 1>(ctx, F<T, 3>{extra...}, out, a, static_cast<const T*>(nullptr), n);
  }
  set_last_error("unknown binary op %d", op);
  return ZB_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------------ Adam / AdamW
template <typename T>
__global__ void __launch_bounds__(256) adam_kernel(T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m,
                                                   T* __restrict__ v, T lr, T beta1, T beta2, T eps, T wd, int decay,
                                                   T inv_bc1, T inv_bc2, T gscale, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const T gi = g[i] * gscale;
    const T mi = m[i] * beta1 + gi * (T(1) - beta1);
    const T vi = v[i] * beta2 + (gi * gi) * (T(1) - beta2);
    m[i] = mi;
    v[i] = vi;
    const T mh = mi * inv_bc1, vh = vi * inv_bc2;
    const T upd = mh / (sqrt(vh) + eps);
    T pi = p[i];
    if (decay) pi -= (pi * lr) * wd;
    p[i] = pi - upd * lr;
  }
}

// bias corrections from a device table (graph-replayable step, see zb_adam_step_table)
template <typename T>
__global__ void __launch_bounds__(256) adam_table_kernel(T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m, T* __restrict__ v,
                                                         T lr, T beta1, T beta2, T eps, T wd, int decay, const T* __restrict__ table,
                                                         const int* __restrict__ step_index, T gscale, long long n) {
  const int idx = *step_index;
  const T inv_bc1 = table[2 * idx], inv_bc2 = table[2 * idx + 1];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const T gi = g[i] * gscale;
    const T mi = m[i] * beta1 + gi * (T(1) - beta1);
    const T vi = v[i] * beta2 + (gi * gi) * (T(1) - beta2);
    m[i] = mi;
    v[i] = vi;
    const T mh = mi * inv_bc1, vh = vi * inv_bc2;
    const T upd = mh / (sqrt(vh) + eps);
    T pi = p[i];
    if (decay) pi -= (pi * lr) * wd;
    p[i] = pi - upd * lr;
  }
}
__global__ void adam_advance_kernel(int* step_index) { *step_index += 1; }

// ------------------------------------------------------------------------------------------------ layout
// [B][R][C] -> [B][C][R] tiled transpose (NCHW<->NHWC with R = C_ch / HW as appropriate).
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows,
                                                        long long cols) {
  __shared__ T tile[32][33];
  const long long b = blockIdx.z;
  const T* s = src + b * rows * cols;
  T* d = dst + b * rows * cols;
  const long long c0 = static_cast<long long>(blockIdx.x) * 32, r0 = static_cast<long long>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + 8 * i][tx] = s[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long c = c0 + ty + 8 * i, r = r0 + tx;
    if (r < rows && c < cols) d[c * rows + r] = tile[tx][ty + 8 * i];
  }
}

template <typename T>
int transpose_batched(zb_ctx* ctx, const T* src, T* dst, long long batch, long long rows, long long cols) {
  if (batch * rows * cols == 0) return ZB_OK;
  ZB_REQUIRE(batch <= 65535 && (rows + 31) / 32 <= 65535, "transpose: dimension too large for the launch grid");
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32), static_cast<unsigned>(batch));
  plan_note("transpose;");
  ZB_KLAUNCH(ctx, transpose_kernel<T><<<grid, 256, 0, ctx->stream>>>(src, dst, rows, cols));
  return ZB_OK;
}
template int transpose_batched<float>(zb_ctx*, const float*, float*, long long, long long, long long);
template int transpose_batched<double>(zb_ctx*, const double*, double*, long long, long long, long long);

template <typename T>
static int relu_t(zb_ctx* ctx, const void* x, void* y, double alpha, long long n) {
  return launch_map<T, 1>(ctx, ReluF<T>{static_cast<T>(alpha)}, static_cast<T*>(y), static_cast<const T*>(x), static_cast<const T*>(nullptr), n);
}
template <typename T>
static int relu_mask_t(zb_ctx* ctx, const void* x, void* y, double alpha, long long n) {
  return launch_map<T, 1>(ctx, ReluMaskF<T>{static_cast<T>(alpha)}, static_cast<T*>(y), static_cast<const T*>(x), static_cast<const T*>(nullptr), n);
}
template <typename T>
static int relu_bwd_t(zb_ctx* ctx, const void* x, const void* dy, void* dx, double alpha, long long n) {
  return launch_map<T, 2>(ctx, ReluBwdF<T>{static_cast<T>(alpha)}, static_cast<T*>(dx), static_cast<const T*>(x), static_cast<const T*>(dy), n);
}

template <typename T>
static int adam_t(zb_ctx* ctx, void* p, const void* g, void* m, void* v, double lr, double b1, double b2, double eps,
                  double wd, int decay, long long step_t, double gscale, long long n) {
  if (n <= 0) return ZB_OK;
  // bias corrections in T precision like the reference (beta.powf(step), adam.rs:23-25)
  const T b1t = static_cast<T>(pow(static_cast<T>(b1), static_cast<T>(step_t)));
  const T b2t = static_cast<T>(pow(static_cast<T>(b2), static_cast<T>(step_t)));
  const T inv1 = T(1) / (T(1) - b1t), inv2 = T(1) / (T(1) - b2t);
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, ctx->sm_count * 16ll)));
  adam_kernel<T><<<grid, 256, 0, ctx->stream>>>(static_cast<T*>(p), static_cast<const T*>(g), static_cast<T*>(m), static_cast<T*>(v),
                                                static_cast<T>(lr), static_cast<T>(b1), static_cast<T>(b2), static_cast<T>(eps),
                                                static_cast<T>(wd), decay, inv1, inv2, static_cast<T>(gscale), n);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

template <typename T>
static int adam_table_fill_t(zb_ctx* ctx, double b1, double b2, long long first_step, long long count, void* table) {
  std::vector<T> h(static_cast<size_t>(2 * count));
  for (long long i = 0; i < count; ++i) {   // T-precision powers like adam_t (and the reference, adam.rs:23-25)
    const T b1t = static_cast<T>(pow(static_cast<T>(b1), static_cast<T>(first_step + i)));
    const T b2t = static_cast<T>(pow(static_cast<T>(b2), static_cast<T>(first_step + i)));
    h[2 * i] = T(1) / (T(1) - b1t);
    h[2 * i + 1] = T(1) / (T(1) - b2t);
  }
  ZB_CHECK_CUDA(cudaMemcpyAsync(table, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, ctx->stream));
  ZB_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));   // h is pageable and dies here (rare: once per `count` steps)
  return ZB_OK;
}
template <typename T>
static int adam_table_t(zb_ctx* ctx, void* p, const void* g, void* m, void* v, double lr, double b1, double b2, double eps, double wd,
                        int decay, const void* table, const int32_t* step_index, double gscale, long long n) {
  if (n <= 0) return ZB_OK;
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((n + 255) / 256, ctx->sm_count * 16ll)));
  adam_table_kernel<T><<<grid, 256, 0, ctx->stream>>>(static_cast<T*>(p), static_cast<const T*>(g), static_cast<T*>(m), static_cast<T*>(v),
                                                      static_cast<T>(lr), static_cast<T>(b1), static_cast<T>(b2), static_cast<T>(eps),
                                                      static_cast<T>(wd), decay, static_cast<const T*>(table), step_index,
                                                      static_cast<T>(gscale), n);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
// used by api.cu for NCHW bias add
template <typename T>
int bias_add_nchw(zb_ctx* ctx, const T* x, const T* bias, T* y, long long n, long long k, long long hw) {
  return launch_map<T, 1>(ctx, BiasNchwF<T>{bias, hw, k}, y, x, static_cast<const T*>(nullptr), n * k * hw);
}
template int bias_add_nchw<float>(zb_ctx*, const float*, const float*, float*, long long, long long, long long);
template int bias_add_nchw<double>(zb_ctx*, const double*, const double*, double*, long long, long long, long long);

}  // namespace zb

using namespace zb;

#define ZB_DTYPE_SWITCH(dtype, CALL_F32, CALL_F64)            \
  do {                                                        \
    if ((dtype) == ZB_F32) return CALL_F32;                   \
    if ((dtype) == ZB_F64) return CALL_F64;                   \
    zb::set_last_error("unknown dtype %d", int(dtype));       \
    return ZB_ERR_INVALID;                                    \
  } while (0)

// ---- input pipeline: uint8 images -> normalised f32/f64 NCHW batch, int labels -> one-hot targets -----------
// Replaces the reference's host path (per-sample Vec<Variable> + CPU concat + synchronous cudaMemcpy of f32,
// zenu/src/dataset.rs:74-100): the batch crosses PCIe as bytes (4x fewer) and is expanded on the device.
struct NormCoef { float scale[8], shift[8]; };   // out = u8 * scale[c] + shift[c]
template <typename T>
__global__ void __launch_bounds__(256) u8_to_float_kernel(const uint8_t* __restrict__ src, T* __restrict__ dst, long long N, int C,
                                                          long long HW, int src_nhwc, NormCoef nc) {
  const long long total = N * C * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long hw = i % HW;
    const long long ncc = i / HW;
    const int c = static_cast<int>(ncc % C);
    const long long n = ncc / C;
    const uint8_t v = src_nhwc ? src[(n * HW + hw) * C + c] : src[i];
    dst[i] = static_cast<T>(static_cast<float>(v) * nc.scale[c] + nc.shift[c]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) onehot_kernel(const int* __restrict__ labels, T* __restrict__ out, long long N, long long K) {
  const long long total = N * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / K, k = i - n * K;
    out[i] = labels[n] == k ? T(1) : T(0);
  }
}

extern "C" {

int zb_relu(zb_ctx* ctx, int dtype, const void* x, void* y, double alpha, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype, relu_t<float>(ctx, x, y, alpha, n), relu_t<double>(ctx, x, y, alpha, n));
}
int zb_relu_backward_mask(zb_ctx* ctx, int dtype, const void* x, void* mask, double alpha, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype, relu_mask_t<float>(ctx, x, mask, alpha, n), relu_mask_t<double>(ctx, x, mask, alpha, n));
}
int zb_relu_bwd(zb_ctx* ctx, int dtype, const void* x, const void* dy, void* dx, double alpha, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype, relu_bwd_t<float>(ctx, x, dy, dx, alpha, n), relu_bwd_t<double>(ctx, x, dy, dx, alpha, n));
}
int zb_binary(zb_ctx* ctx, int dtype, int op, const void* a, const void* b, void* c, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype,
                  (dispatch_binop<float, BinF>(ctx, op, static_cast<float*>(c), static_cast<const float*>(a), static_cast<const float*>(b), n)),
                  (dispatch_binop<double, BinF>(ctx, op, static_cast<double*>(c), static_cast<const double*>(a), static_cast<const double*>(b), n)));
}
int zb_binary_scalar(zb_ctx* ctx, int dtype, int op, const void* a, double scalar, void* c, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype,
                  (dispatch_unop<float, BinScalarF>(ctx, op, static_cast<float*>(c), static_cast<const float*>(a), n, static_cast<float>(scalar))),
                  (dispatch_unop<double, BinScalarF>(ctx, op, static_cast<double*>(c), static_cast<const double*>(a), n, scalar)));
}
int zb_binary_bcast_rows(zb_ctx* ctx, int dtype, int op, const void* a, const void* b, void* c, int64_t rows, int64_t cols) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype,
                  (dispatch_unop<float, BcastRowsF>(ctx, op, static_cast<float*>(c), static_cast<const float*>(a), rows * cols, static_cast<const float*>(b), static_cast<long long>(cols))),
                  (dispatch_unop<double, BcastRowsF>(ctx, op, static_cast<double*>(c), static_cast<const double*>(a), rows * cols, static_cast<const double*>(b), static_cast<long long>(cols))));
}
int zb_fill(zb_ctx* ctx, int dtype, void* x, double value, int64_t n) {
  ZB_API_RANGE();
  if (value == 0.0 && n > 0) {
    ZB_CHECK_CUDA(cudaMemsetAsync(x, 0, static_cast<size_t>(n) * (dtype == ZB_F64 ? 8 : 4), ctx->stream));
    return ZB_OK;
  }
  ZB_DTYPE_SWITCH(dtype,
                  (launch_map<float, 0>(ctx, FillF<float>{static_cast<float>(value)}, static_cast<float*>(x), static_cast<const float*>(nullptr), static_cast<const float*>(nullptr), n)),
                  (launch_map<double, 0>(ctx, FillF<double>{value}, static_cast<double*>(x), static_cast<const double*>(nullptr), static_cast<const double*>(nullptr), n)));
}
int zb_mask_apply(zb_ctx* ctx, int dtype, const void* x, const void* bits, void* out, int64_t n) {
  ZB_API_RANGE();
  ZB_REQUIRE(bits != nullptr, "zb_mask_apply: bits is NULL");
  ZB_DTYPE_SWITCH(dtype,
                  (launch_map<float, 1>(ctx, MaskBitsF<float>{static_cast<const uint32_t*>(bits)}, static_cast<float*>(out), static_cast<const float*>(x), static_cast<const float*>(nullptr), n)),
                  (launch_map<double, 1>(ctx, MaskBitsF<double>{static_cast<const uint32_t*>(bits)}, static_cast<double*>(out), static_cast<const double*>(x), static_cast<const double*>(nullptr), n)));
}
int zb_copy(zb_ctx* ctx, int dtype, const void* src, void* dst, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype,
                  (launch_map<float, 1>(ctx, CopyF<float>{}, static_cast<float*>(dst), static_cast<const float*>(src), static_cast<const float*>(nullptr), n)),
                  (launch_map<double, 1>(ctx, CopyF<double>{}, static_cast<double*>(dst), static_cast<const double*>(src), static_cast<const double*>(nullptr), n)));
}
int zb_nchw_to_nhwc(zb_ctx* ctx, int dtype, const void* src, void* dst, int64_t n, int64_t c, int64_t h, int64_t w) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype, transpose_batched<float>(ctx, static_cast<const float*>(src), static_cast<float*>(dst), n, c, h * w),
                  transpose_batched<double>(ctx, static_cast<const double*>(src), static_cast<double*>(dst), n, c, h * w));
}
int zb_nhwc_to_nchw(zb_ctx* ctx, int dtype, const void* src, void* dst, int64_t n, int64_t c, int64_t h, int64_t w) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype, transpose_batched<float>(ctx, static_cast<const float*>(src), static_cast<float*>(dst), n, h * w, c),
                  transpose_batched<double>(ctx, static_cast<const double*>(src), static_cast<double*>(dst), n, h * w, c));
}
int zb_sgd_step(zb_ctx* ctx, int dtype, void* param, const void* grad, double lr, double grad_scale, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype,
                  (launch_map<float, 2>(ctx, SgdF<float>{static_cast<float>(lr), static_cast<float>(grad_scale)}, static_cast<float*>(param), static_cast<const float*>(param), static_cast<const float*>(grad), n)),
                  (launch_map<double, 2>(ctx, SgdF<double>{lr, grad_scale}, static_cast<double*>(param), static_cast<const double*>(param), static_cast<const double*>(grad), n)));
}
int zb_adam_step(zb_ctx* ctx, int dtype, void* param, const void* grad, void* m, void* v, double lr, double beta1,
                 double beta2, double eps, double weight_decay, int decay, int64_t step_t, double grad_scale, int64_t n) {
  ZB_API_RANGE();
  ZB_DTYPE_SWITCH(dtype, adam_t<float>(ctx, param, grad, m, v, lr, beta1, beta2, eps, weight_decay, decay, step_t, grad_scale, n),
                  adam_t<double>(ctx, param, grad, m, v, lr, beta1, beta2, eps, weight_decay, decay, step_t, grad_scale, n));
}

int zb_adam_table_fill(zb_ctx* ctx, int dtype, double beta1, double beta2, int64_t first_step, int64_t count, void* table) {
  ZB_API_RANGE();
  ZB_REQUIRE(table != nullptr && first_step >= 1 && count >= 1, "zb_adam_table_fill: bad argument");
  ZB_DTYPE_SWITCH(dtype, adam_table_fill_t<float>(ctx, beta1, beta2, first_step, count, table),
                  adam_table_fill_t<double>(ctx, beta1, beta2, first_step, count, table));
}
int zb_adam_step_table(zb_ctx* ctx, int dtype, void* param, const void* grad, void* m, void* v, double lr, double beta1, double beta2,
                       double eps, double weight_decay, int decay, const void* table, const int32_t* step_index, double grad_scale,
                       int64_t n) {
  ZB_API_RANGE();
  ZB_REQUIRE(table != nullptr && step_index != nullptr, "zb_adam_step_table: NULL table / step index");
  ZB_DTYPE_SWITCH(dtype, adam_table_t<float>(ctx, param, grad, m, v, lr, beta1, beta2, eps, weight_decay, decay, table, step_index, grad_scale, n),
                  adam_table_t<double>(ctx, param, grad, m, v, lr, beta1, beta2, eps, weight_decay, decay, table, step_index, grad_scale, n));
}
int zb_adam_advance(zb_ctx* ctx, int32_t* step_index) {
  ZB_API_RANGE();
  ZB_REQUIRE(step_index != nullptr, "zb_adam_advance: NULL step index");
  adam_advance_kernel<<<1, 1, 0, ctx->stream>>>(step_index);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

int zb_input_u8_to_float(zb_ctx* ctx, int dtype, int src_layout, const void* src_u8, void* dst_nchw, int64_t n, int64_t c, int64_t h,
                         int64_t w, const double* host_mean, const double* host_std) {
  ZB_API_RANGE();
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "input: unknown dtype %d", dtype);
  ZB_REQUIRE(src_layout == ZB_NCHW || src_layout == ZB_NHWC, "input: unknown source layout %d", src_layout);
  ZB_REQUIRE(c >= 1 && c <= 8, "input: 1..8 channels supported (got %lld)", static_cast<long long>(c));
  if (n * c * h * w == 0) return ZB_OK;
  NormCoef nc;
  for (int i = 0; i < 8; ++i) {   // (u8 / 255 - mean) / std
    const double m = host_mean && i < c ? host_mean[i] : 0.0, sd = host_std && i < c ? host_std[i] : 1.0;
    ZB_REQUIRE(sd != 0.0, "input: std[%d] is zero", i);
    nc.scale[i] = static_cast<float>(1.0 / (255.0 * sd));
    nc.shift[i] = static_cast<float>(-m / sd);
  }
  const long long total = n * c * h * w;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  if (dtype == ZB_F32)
    u8_to_float_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(src_u8), static_cast<float*>(dst_nchw), n,
                                                                 static_cast<int>(c), h * w, src_layout == ZB_NHWC, nc);
  else
    u8_to_float_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(src_u8), static_cast<double*>(dst_nchw), n,
                                                                  static_cast<int>(c), h * w, src_layout == ZB_NHWC, nc);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
int zb_onehot(zb_ctx* ctx, int dtype, const void* labels_i32, void* out, int64_t n, int64_t classes) {
  ZB_API_RANGE();
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "onehot: unknown dtype %d", dtype);
  if (n * classes == 0) return ZB_OK;
  const long long total = n * classes;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  if (dtype == ZB_F32) onehot_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const int*>(labels_i32), static_cast<float*>(out), n, classes);
  else onehot_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const int*>(labels_i32), static_cast<double*>(out), n, classes);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
}  // extern "C"
