// reduce_copy.cu — axis reductions and strided copies as first-class ops.
//
// Replaces, for contiguous tensors viewed as [outer][len][inner]:
//   Matrix::sum(axis)        zenu-matrix/src/operation/sum.rs:9-31   (a loop of `len` add_assign launches over the axis)
//   sum_to(source, target)   operation/sum.rs:35-92                  (numpy-style reduction to a broadcastable shape)
//   Matrix::mean(axis)       operation/mean.rs:8-20                  (sum / len)
//   Matrix::variance(axis)   operation/var.rs:18-26                  (biased: mean of squared differences from the mean)
// and the strided element copies behind copy_from / to_default_stride / transpose-materialise
//   CopyBlas::copy_raw       operation/copy_from.rs:9-55             (one cublas{S,D}copy per contiguous run)
// One launch (plus a deterministic fold when the axis is split over slabs) instead of O(len) launches; no atomics.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace zb {

// MODE 0: sum(x) * scale      MODE 1: sum((x - mean[o][i])^2) * scale
// Thread = one output element (o, i); grid.y = slab of the reduced axis.  Consecutive threads walk consecutive `inner` indices, so
// the loads of a warp are contiguous whenever inner >= 32; inner == 1 takes the row kernel below instead.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) axis_reduce_kernel(const T* __restrict__ a, const T* __restrict__ mean, T* __restrict__ out,
                                                          long long outer, long long len, long long inner, long long len_per_slab, T scale) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= outer * inner) return;
  const long long o = idx / inner, i = idx - o * inner;
  const long long l0 = static_cast<long long>(blockIdx.y) * len_per_slab;
  const long long l1 = l0 + len_per_slab < len ? l0 + len_per_slab : len;
  const T* p = a + (o * len) * inner + i;
  const T m = MODE == 1 ? mean[idx] : T(0);
  T acc0 = T(0), acc1 = T(0), acc2 = T(0), acc3 = T(0);
  long long l = l0;
  for (; l + 4 <= l1; l += 4) {   // four independent loads in flight
    const T v0 = p[(l + 0) * inner], v1 = p[(l + 1) * inner], v2 = p[(l + 2) * inner], v3 = p[(l + 3) * inner];
    if (MODE == 0) { acc0 += v0; acc1 += v1; acc2 += v2; acc3 += v3; }
    else { acc0 += (v0 - m) * (v0 - m); acc1 += (v1 - m) * (v1 - m); acc2 += (v2 - m) * (v2 - m); acc3 += (v3 - m) * (v3 - m); }
  }
  for (; l < l1; ++l) {
    const T v = p[l * inner];
    acc0 += MODE == 0 ? v : (v - m) * (v - m);
  }
  const T r = (acc0 + acc1) + (acc2 + acc3);
  out[static_cast<long long>(blockIdx.y) * outer * inner + idx] = gridDim.y == 1 ? r * scale : r;
}

// inner == 1: out[o] = reduce over the contiguous row a[o][0..len): one block per row
template <typename T, int MODE>
__global__ void __launch_bounds__(256) row_reduce_kernel(const T* __restrict__ a, const T* __restrict__ mean, T* __restrict__ out, long long len,
                                                         T scale) {
  __shared__ T red[8];
  const long long o = blockIdx.x;
  const T* p = a + o * len;
  const T m = MODE == 1 ? mean[o] : T(0);
  T acc = T(0);
  for (long long l = threadIdx.x; l < len; l += blockDim.x) {
    const T v = p[l];
    acc += MODE == 0 ? v : (v - m) * (v - m);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
    for (int w = 0; w < 8; ++w) s += red[w];
    out[o] = s * scale;
  }
}

template <typename T>
__global__ void fold_slabs_kernel(const T* __restrict__ partial, T* __restrict__ out, long long n, int slabs, T scale) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    T s = T(0);
    for (int k = 0; k < slabs; ++k) s += partial[static_cast<long long>(k) * n + i];   // fixed order: deterministic
    out[i] = s * scale;
  }
}

template <typename T, int MODE>
static int axis_reduce(zb_ctx* ctx, const T* a, const T* mean, T* out, long long outer, long long len, long long inner, double scale) {
  const long long n_out = outer * inner;
  if (n_out == 0) return ZB_OK;
  if (len == 0) {   // empty axis: the sum is zero (Matrix::zeros + no add_assign, sum.rs:17-25)
    ZB_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(T) * n_out, ctx->stream));
    return ZB_OK;
  }
  if (inner == 1 && len >= 64) {
    ZB_REQUIRE(outer <= 2147483647ll, "reduce: too many rows");
    ZB_KLAUNCH(ctx, row_reduce_kernel<T, MODE><<<static_cast<unsigned>(outer), 256, 0, ctx->stream>>>(a, mean, out, len, static_cast<T>(scale)));
    return ZB_OK;
  }
  const long long blocks_x = (n_out + 255) / 256;
  ZB_REQUIRE(blocks_x <= 2147483647ll, "reduce: output too large");
  // few outputs and a long axis: split the axis over slabs so the SMs fill, fold the slabs afterwards
  int slabs = 1;
  if (blocks_x < 2ll * ctx->sm_count && len >= 256)
    slabs = static_cast<int>(std::min<long long>(std::min<long long>((2ll * ctx->sm_count + blocks_x - 1) / blocks_x, len / 64), 1024));
  slabs = std::max(slabs, 1);
  const long long per = (len + slabs - 1) / slabs;
  slabs = static_cast<int>((len + per - 1) / per);
  T* dst = out;
  if (slabs > 1) {
    void* ws = nullptr;
    const int rc = ctx_workspace(ctx, sizeof(T) * static_cast<size_t>(slabs) * n_out, &ws);
    if (rc != ZB_OK) return rc;
    dst = static_cast<T*>(ws);
  }
  dim3 grid(static_cast<unsigned>(blocks_x), static_cast<unsigned>(slabs));
  ZB_KLAUNCH(ctx, axis_reduce_kernel<T, MODE><<<grid, 256, 0, ctx->stream>>>(a, mean, dst, outer, len, inner, per, static_cast<T>(scale)));
  if (slabs > 1) {
    const int g = static_cast<int>(std::min<long long>(blocks_x, ctx->sm_count * 8ll));
    ZB_KLAUNCH(ctx, fold_slabs_kernel<T><<<g, 256, 0, ctx->stream>>>(dst, out, n_out, slabs, static_cast<T>(scale)));
  }
  return ZB_OK;
}

struct StridedCopyDims {
  int ndim;
  long long shape[8], src_stride[8], dst_stride[8];
};
template <typename T>
__global__ void strided_copy_kernel(const T* __restrict__ src, T* __restrict__ dst, long long n, StridedCopyDims d) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long rem = i, so = 0, dof = 0;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      if (k < d.ndim) {
        const long long c = rem % d.shape[k];
        rem /= d.shape[k];
        so += c * d.src_stride[k];
        dof += c * d.dst_stride[k];
      }
    }
    dst[dof] = src[so];
  }
}

template <typename T>
static int run_sum_to(zb_ctx* ctx, const T* src, const int64_t* s_shape, int s_nd, T* dst, const int64_t* d_shape, int d_nd) {
  // right-aligned like the reference: extra leading axes of the source are summed away, then every axis whose target extent is 1
  std::vector<long long> cur(s_shape, s_shape + s_nd);
  std::vector<long long> tgt(s_nd, 1);
  for (int k = 0; k < d_nd; ++k) tgt[s_nd - d_nd + k] = d_shape[k];
  for (int k = 0; k < s_nd; ++k)
    ZB_REQUIRE(cur[k] == tgt[k] || tgt[k] == 1, "sum_to: target extent %lld does not divide into source extent %lld (axis %d)", tgt[k], cur[k], k);
  long long total = 1;
  for (long long v : cur) total *= v;
  std::vector<int> axes;
  for (int k = 0; k < s_nd; ++k)
    if (cur[k] != tgt[k]) axes.push_back(k);
  if (axes.empty()) {
    if (total > 0 && !plan_dry()) ZB_CHECK_CUDA(cudaMemcpyAsync(dst, src, sizeof(T) * total, cudaMemcpyDeviceToDevice, ctx->stream));
    return ZB_OK;
  }
  // merge runs of adjacent reduced axes into one [outer][len][inner] reduction each; ping-pong through two temporaries
  const T* in = src;
  void* tmp[2] = {nullptr, nullptr};
  int which = 0, rc = ZB_OK;
  size_t a = 0;
  while (a < axes.size() && rc == ZB_OK) {
    size_t b = a;
    while (b + 1 < axes.size() && axes[b + 1] == axes[b] + 1) ++b;
    long long outer = 1, len = 1, inner = 1;
    for (int k = 0; k < axes[a]; ++k) outer *= cur[k];
    for (int k = axes[a]; k <= axes[b]; ++k) len *= cur[k];
    for (int k = axes[b] + 1; k < s_nd; ++k) inner *= cur[k];
    const bool last = b + 1 == axes.size();
    T* out = dst;
    if (!last) {
      if (tmp[which] == nullptr && (rc = (cudaMallocAsync(&tmp[which], sizeof(T) * std::max<long long>(outer * inner, 1), ctx->stream) == cudaSuccess ? ZB_OK : ZB_ERR_CUDA)) != ZB_OK) {
        set_last_error("sum_to: temporary allocation failed");
        break;
      }
      out = static_cast<T*>(tmp[which]);
    }
    rc = axis_reduce<T, 0>(ctx, in, static_cast<const T*>(nullptr), out, outer, len, inner, 1.0);
    for (int k = axes[a]; k <= axes[b]; ++k) cur[k] = 1;
    in = out;
    which ^= 1;
    a = b + 1;
  }
  for (void* t : tmp)
    if (t) cudaFreeAsync(t, ctx->stream);
  return rc;
}

}  // namespace zb

using namespace zb;

#define ZB_RC_DTYPE(dtype, f32_expr, f64_expr)                  \
  do {                                                          \
    if ((dtype) == ZB_F32) return (f32_expr);                   \
    if ((dtype) == ZB_F64) return (f64_expr);                   \
    zb::set_last_error("unknown dtype %d", (dtype));            \
    return ZB_ERR_INVALID;                                      \
  } while (0)

extern "C" {

int zb_sum_axis(zb_ctx* ctx, int dtype, const void* a, void* out, int64_t outer, int64_t len, int64_t inner) {
  ZB_API_RANGE();
  ZB_REQUIRE(outer >= 0 && len >= 0 && inner >= 0, "sum_axis: negative extent");
  ZB_RC_DTYPE(dtype, (axis_reduce<float, 0>(ctx, static_cast<const float*>(a), nullptr, static_cast<float*>(out), outer, len, inner, 1.0)),
              (axis_reduce<double, 0>(ctx, static_cast<const double*>(a), nullptr, static_cast<double*>(out), outer, len, inner, 1.0)));
}

int zb_mean_axis(zb_ctx* ctx, int dtype, const void* a, void* out, int64_t outer, int64_t len, int64_t inner) {
  ZB_API_RANGE();
  ZB_REQUIRE(outer >= 0 && len > 0 && inner >= 0, "mean_axis: empty or negative extent");
  const double s = 1.0 / static_cast<double>(len);
  ZB_RC_DTYPE(dtype, (axis_reduce<float, 0>(ctx, static_cast<const float*>(a), nullptr, static_cast<float*>(out), outer, len, inner, s)),
              (axis_reduce<double, 0>(ctx, static_cast<const double*>(a), nullptr, static_cast<double*>(out), outer, len, inner, s)));
}

int zb_variance_axis(zb_ctx* ctx, int dtype, const void* a, void* out, void* mean_out, int64_t outer, int64_t len, int64_t inner) {
  ZB_API_RANGE();
  ZB_REQUIRE(outer >= 0 && len > 0 && inner >= 0, "variance_axis: empty or negative extent");
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "unknown dtype %d", dtype);
  // mean first (into mean_out when the caller wants it, else into a stream-ordered temporary), then the mean of squared differences
  const size_t esz = dtype == ZB_F64 ? 8 : 4;
  void* mean = mean_out;
  void* tmp = nullptr;
  if (mean == nullptr) {
    if (plan_dry()) tmp = reinterpret_cast<void*>(uintptr_t(11) << 40);
    else ZB_CHECK_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(esz * outer * inner, 16), ctx->stream));
    mean = tmp;
  }
  const double s = 1.0 / static_cast<double>(len);
  int rc = dtype == ZB_F32 ? axis_reduce<float, 0>(ctx, static_cast<const float*>(a), nullptr, static_cast<float*>(mean), outer, len, inner, s)
                           : axis_reduce<double, 0>(ctx, static_cast<const double*>(a), nullptr, static_cast<double*>(mean), outer, len, inner, s);
  if (rc == ZB_OK)
    rc = dtype == ZB_F32
             ? axis_reduce<float, 1>(ctx, static_cast<const float*>(a), static_cast<const float*>(mean), static_cast<float*>(out), outer, len, inner, s)
             : axis_reduce<double, 1>(ctx, static_cast<const double*>(a), static_cast<const double*>(mean), static_cast<double*>(out), outer, len, inner, s);
  if (tmp != nullptr && !plan_dry()) cudaFreeAsync(tmp, ctx->stream);
  return rc;
}

int zb_sum_to(zb_ctx* ctx, int dtype, const void* src, const int64_t* src_shape, int src_ndim, void* dst, const int64_t* dst_shape,
              int dst_ndim) {
  ZB_API_RANGE();
  ZB_REQUIRE(src_shape != nullptr && (dst_shape != nullptr || dst_ndim == 0) && src_ndim >= 0 && src_ndim <= 8 && dst_ndim >= 0,
             "sum_to: bad shape arguments");
  ZB_REQUIRE(src_ndim >= dst_ndim, "sum_to: the source has fewer axes than the target (sum.rs:39-42)");
  for (int k = 0; k < src_ndim; ++k) ZB_REQUIRE(src_shape[k] >= 0, "sum_to: negative extent");
  ZB_RC_DTYPE(dtype, (run_sum_to<float>(ctx, static_cast<const float*>(src), src_shape, src_ndim, static_cast<float*>(dst), dst_shape, dst_ndim)),
              (run_sum_to<double>(ctx, static_cast<const double*>(src), src_shape, src_ndim, static_cast<double*>(dst), dst_shape, dst_ndim)));
}

int zb_copy_strided(zb_ctx* ctx, int dtype, const void* src, void* dst, int ndim, const int64_t* shape, const int64_t* src_strides,
                    const int64_t* dst_strides) {
  ZB_API_RANGE();
  ZB_REQUIRE(ndim >= 0 && ndim <= 8 && (ndim == 0 || (shape && src_strides && dst_strides)), "copy_strided: 0..8 axes with shape and strides");
  ZB_REQUIRE(dtype == ZB_F32 || dtype == ZB_F64, "unknown dtype %d", dtype);
  StridedCopyDims d;
  d.ndim = ndim;
  long long n = 1;
  bool dense = true;
  long long expect = 1;
  for (int k = ndim - 1; k >= 0; --k) {
    ZB_REQUIRE(shape[k] >= 0 && src_strides[k] >= 0 && dst_strides[k] >= 0, "copy_strided: negative extent / stride");
    d.shape[k] = shape[k]; d.src_stride[k] = src_strides[k]; d.dst_stride[k] = dst_strides[k];
    if (shape[k] != 1 && (src_strides[k] != expect || dst_strides[k] != expect)) dense = false;
    expect *= shape[k];
    n *= shape[k];
  }
  for (int k = ndim; k < 8; ++k) { d.shape[k] = 1; d.src_stride[k] = 0; d.dst_stride[k] = 0; }
  if (n == 0) return ZB_OK;
  if (dense) return zb_copy(ctx, dtype, src, dst, n);   // both sides in default stride: the vectorised copy
  const int grid = static_cast<int>(std::min<long long>((n + 255) / 256, ctx->sm_count * 16ll));
  if (dtype == ZB_F32) ZB_KLAUNCH(ctx, strided_copy_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(src), static_cast<float*>(dst), n, d));
  else ZB_KLAUNCH(ctx, strided_copy_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double*>(src), static_cast<double*>(dst), n, d));
  return ZB_OK;
}

}  // extern "C"
