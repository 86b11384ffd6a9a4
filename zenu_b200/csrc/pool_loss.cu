// pool_loss.cu — "next" rows of the scope table: max-pool, global average pool and the
// fused softmax + cross-entropy loss head, so a ResNet training step closes on the device.
//   max-pool: semantics of the reference CPU path (zenu-matrix/src/nn/pool2d.rs:77-150): padding contributes
//             zeros to the max, the first maximum in (kh, kw) order receives the gradient.
//   softmax_xent: zenu-autograd/src/loss/cross_entropy.rs:12-23 (+ softmax.rs / operation/softmax.rs),
//             forward and backward in one pass over the logits.
#include <algorithm>

#include "common.cuh"

namespace zb {

struct PoolGeom {
  long long N, C, H, W, P, Q;
  int kh, kw, sh, sw, ph, pw;
  long long sn, sc, s_h, s_w;      // input strides (elements)
  long long on, oc, o_h, o_w;      // output strides
  int nhwc;
};

__device__ __forceinline__ void pool_decode(const PoolGeom& g, long long i, long long& n, long long& c, long long& p,
                                            long long& q) {
  if (g.nhwc) { c = i % g.C; i /= g.C; q = i % g.Q; i /= g.Q; p = i % g.P; n = i / g.P; }
  else { q = i % g.Q; i /= g.Q; p = i % g.P; i /= g.P; c = i % g.C; n = i / g.C; }
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(256) maxpool_kernel(const PoolGeom g, const T* __restrict__ x, T* __restrict__ y,
                                                      const T* __restrict__ dy, T* __restrict__ dx) {
  const long long total = g.N * g.C * g.P * g.Q;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long n, c, p, q;
    pool_decode(g, i, n, c, p, q);
    const T* xb = x + n * g.sn + c * g.sc;
    T best = T(0);
    long long best_off = -1;
    bool first = true;
    for (int r = 0; r < g.kh; ++r) {
      const long long ih = p * g.sh + r - g.ph;
      for (int s = 0; s < g.kw; ++s) {
        const long long iw = q * g.sw + s - g.pw;
        const bool oob = ih < 0 || ih >= g.H || iw < 0 || iw >= g.W;
        const T v = oob ? T(0) : xb[ih * g.s_h + iw * g.s_w];
        if (first || v > best) {
          best = v;
          best_off = oob ? -1 : (ih * g.s_h + iw * g.s_w);
          first = false;
        }
      }
    }
    const long long o = n * g.on + c * g.oc + p * g.o_h + q * g.o_w;
    if (!BWD) {
      y[o] = best;
    } else if (best_off >= 0) {
      atomicAdd(dx + n * g.sn + c * g.sc + best_off, dy[o]);
    }
  }
}

static PoolGeom pool_geom(int layout, long long n, long long c, long long h, long long w, long long kh, long long kw,
                          long long sh, long long sw, long long ph, long long pw) {
  PoolGeom g;
  g.N = n; g.C = c; g.H = h; g.W = w;
  g.P = (h + 2 * ph - kh) / sh + 1;
  g.Q = (w + 2 * pw - kw) / sw + 1;
  g.kh = static_cast<int>(kh); g.kw = static_cast<int>(kw); g.sh = static_cast<int>(sh); g.sw = static_cast<int>(sw);
  g.ph = static_cast<int>(ph); g.pw = static_cast<int>(pw);
  g.nhwc = (layout == ZB_NHWC);
  if (g.nhwc) {
    g.sn = h * w * c; g.sc = 1; g.s_h = w * c; g.s_w = c;
    g.on = g.P * g.Q * c; g.oc = 1; g.o_h = g.Q * c; g.o_w = c;
  } else {
    g.sn = c * h * w; g.sc = h * w; g.s_h = w; g.s_w = 1;
    g.on = c * g.P * g.Q; g.oc = g.P * g.Q; g.o_h = g.Q; g.o_w = 1;
  }
  return g;
}

template <typename T>
static int maxpool_fwd_t(zb_ctx* ctx, const PoolGeom& g, const T* x, T* y) {
  const long long total = g.N * g.C * g.P * g.Q;
  if (total == 0) return ZB_OK;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  maxpool_kernel<T, false><<<grid, 256, 0, ctx->stream>>>(g, x, y, static_cast<const T*>(nullptr), static_cast<T*>(nullptr));
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
template <typename T>
static int maxpool_bwd_t(zb_ctx* ctx, const PoolGeom& g, const T* x, const T* dy, T* dx) {
  const long long total = g.N * g.C * g.P * g.Q;
  ZB_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(T) * g.N * g.C * g.H * g.W, ctx->stream));
  if (total == 0) return ZB_OK;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  maxpool_kernel<T, true><<<grid, 256, 0, ctx->stream>>>(g, x, static_cast<T*>(nullptr), dy, dx);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// ---- indexed max-pool, NHWC, 4 channels per thread ------------------------------------------------------------
// Forward also records which window tap won (uint8: r*kw+s, 255 = a padding zero won) so that backward is a gather
// over the <= ceil(kh/sh)*ceil(kw/sw) windows covering an input pixel: no memset, no atomics, deterministic.
template <typename T> struct Vec4T;
template <> struct Vec4T<float> { using type = float4; };
template <> struct Vec4T<double> { using type = double4; };

// I = int when every element index fits in 31 bits (64-bit divides dominated the old version: 1.1 TB/s -> HBM-bound now)
// KH x KW > 0: compile-time window (3 x 3 is what the networks use): the nine 128-bit loads of an output are independent of its
// compare chain, and fully unrolled they are all in flight together instead of one per trip of a runtime-count loop
template <typename T, typename I, int KH = 0, int KW = 0>
__global__ void __launch_bounds__(256) maxpool_idx_fwd_kernel(const PoolGeom g, const T* __restrict__ x, T* __restrict__ y,
                                                              uchar4* __restrict__ idx) {
  using V = typename Vec4T<T>::type;
  const I c4n = static_cast<I>(g.C >> 2), Q = static_cast<I>(g.Q), P = static_cast<I>(g.P), H = static_cast<I>(g.H), W = static_cast<I>(g.W);
  const I total = static_cast<I>(g.N) * P * Q * c4n;
  const I s_h = static_cast<I>(g.s_h), s_w = static_cast<I>(g.s_w), sn = static_cast<I>(g.sn);
  for (I i = blockIdx.x * static_cast<I>(blockDim.x) + threadIdx.x; i < total; i += static_cast<I>(gridDim.x) * blockDim.x) {
    const I c4 = i % c4n;
    I t = i / c4n;
    const I q = t % Q; t /= Q;
    const I p = t % P;
    const I n = t / P;
    const T* xb = x + n * sn + c4 * 4;
    T best[4];
    unsigned char bi[4];
    bool first = true;
    const int kh = KH ? KH : g.kh, kw = KW ? KW : g.kw;
#pragma unroll
    for (int r = 0; r < kh; ++r) {
      const I ih = p * g.sh + r - g.ph;
#pragma unroll
      for (int s_ = 0; s_ < kw; ++s_) {
        const I iw = q * g.sw + s_ - g.pw;
        const bool oob = ih < 0 || ih >= H || iw < 0 || iw >= W;
        T v[4] = {T(0), T(0), T(0), T(0)};
        if (!oob) {
          const V vv = *reinterpret_cast<const V*>(xb + ih * s_h + iw * s_w);
          v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
        }
        const unsigned char tap = oob ? 255 : static_cast<unsigned char>(r * kw + s_);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (first || v[e] > best[e]) { best[e] = v[e]; bi[e] = tap; }
        first = false;
      }
    }
    V o;
    o.x = best[0]; o.y = best[1]; o.z = best[2]; o.w = best[3];
    *reinterpret_cast<V*>(y + static_cast<long long>(i) * 4) = o;
    idx[i] = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256) maxpool_idx_bwd_kernel(const PoolGeom g, const T* __restrict__ dy,
                                                              const uchar4* __restrict__ idx, T* __restrict__ dx) {
  using V = typename Vec4T<T>::type;
  const I c4n = static_cast<I>(g.C >> 2), Q = static_cast<I>(g.Q), P = static_cast<I>(g.P), H = static_cast<I>(g.H), W = static_cast<I>(g.W);
  const I total = static_cast<I>(g.N) * H * W * c4n;
  for (I i = blockIdx.x * static_cast<I>(blockDim.x) + threadIdx.x; i < total; i += static_cast<I>(gridDim.x) * blockDim.x) {
    const I c4 = i % c4n;
    I t = i / c4n;
    const I w = t % W; t /= W;
    const I h = t % H;
    const I n = t / H;
    T acc[4] = {T(0), T(0), T(0), T(0)};
    // windows p with p*sh - ph <= h <= p*sh - ph + kh - 1
    I p_lo = (h + g.ph - g.kh + 1 + g.sh - 1);
    p_lo = p_lo <= 0 ? 0 : p_lo / g.sh;
    I p_hi = (h + g.ph) / g.sh;
    if (p_hi > P - 1) p_hi = P - 1;
    I q_lo = (w + g.pw - g.kw + 1 + g.sw - 1);
    q_lo = q_lo <= 0 ? 0 : q_lo / g.sw;
    I q_hi = (w + g.pw) / g.sw;
    if (q_hi > Q - 1) q_hi = Q - 1;
    for (I p = p_lo; p <= p_hi; ++p) {
      const int r = static_cast<int>(h + g.ph - p * g.sh);
      for (I q = q_lo; q <= q_hi; ++q) {
        const int s_ = static_cast<int>(w + g.pw - q * g.sw);
        const unsigned char tap = static_cast<unsigned char>(r * g.kw + s_);
        const I o = ((n * P + p) * Q + q) * c4n + c4;
        const uchar4 wi = idx[o];
        const V gv = *(reinterpret_cast<const V*>(dy) + o);
        if (wi.x == tap) acc[0] += gv.x;
        if (wi.y == tap) acc[1] += gv.y;
        if (wi.z == tap) acc[2] += gv.z;
        if (wi.w == tap) acc[3] += gv.w;
      }
    }
    V o4;
    o4.x = acc[0]; o4.y = acc[1]; o4.z = acc[2]; o4.w = acc[3];
    *reinterpret_cast<V*>(dx + static_cast<long long>(i) * 4) = o4;
  }
}

// Backward for stride (2, 2): one thread owns a 2 x 2 input patch (x 4 channels).  The windows covering the patch are loaded
// once for its four pixels instead of once per pixel (3x3/s2: 4 window reads per patch instead of 9), which is what the
// per-pixel gather above is bound by (L2 read volume, not HBM).
template <typename T>
__global__ void __launch_bounds__(256) maxpool_idx_bwd_s2_kernel(const PoolGeom g, const T* __restrict__ dy,
                                                                 const uchar4* __restrict__ idx, T* __restrict__ dx) {
  using V = typename Vec4T<T>::type;
  const int c4n = static_cast<int>(g.C >> 2), Q = static_cast<int>(g.Q), P = static_cast<int>(g.P), H = static_cast<int>(g.H), W = static_cast<int>(g.W);
  const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
  const int total = static_cast<int>(g.N) * H2 * W2 * c4n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c4 = i % c4n;
    int t = i / c4n;
    const int b = t % W2; t /= W2;
    const int a = t % H2;
    const int n = t / H2;
    const int h0 = 2 * a, w0 = 2 * b;
    T acc[2][2][4];
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[y][x][e] = T(0);
    // windows p with p*2 - ph <= h0 + 1 and p*2 - ph + kh - 1 >= h0
    int p_lo = h0 + g.ph - g.kh + 1;
    p_lo = p_lo <= 0 ? 0 : (p_lo + 1) >> 1;
    int p_hi = (h0 + 1 + g.ph) >> 1;
    if (p_hi > P - 1) p_hi = P - 1;
    int q_lo = w0 + g.pw - g.kw + 1;
    q_lo = q_lo <= 0 ? 0 : (q_lo + 1) >> 1;
    int q_hi = (w0 + 1 + g.pw) >> 1;
    if (q_hi > Q - 1) q_hi = Q - 1;
    for (int p = p_lo; p <= p_hi; ++p) {
      const int r0 = h0 + g.ph - 2 * p;           // filter row of patch row 0 (row 1: r0 + 1)
      for (int q = q_lo; q <= q_hi; ++q) {
        const int s0 = w0 + g.pw - 2 * q;
        const int o = ((n * P + p) * Q + q) * c4n + c4;
        const uchar4 wi = idx[o];
        const V gv = *(reinterpret_cast<const V*>(dy) + o);
        const unsigned char win[4] = {wi.x, wi.y, wi.z, wi.w};
        const T gvv[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int y = 0; y < 2; ++y) {
          const int r = r0 + y;
          if (r < 0 || r >= g.kh) continue;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const int s_ = s0 + x;
            if (s_ < 0 || s_ >= g.kw) continue;
            const unsigned char tap = static_cast<unsigned char>(r * g.kw + s_);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (win[e] == tap) acc[y][x][e] += gvv[e];
          }
        }
      }
    }
#pragma unroll
    for (int y = 0; y < 2; ++y) {
      if (h0 + y >= H) continue;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (w0 + x >= W) continue;
        V o4;
        o4.x = acc[y][x][0]; o4.y = acc[y][x][1]; o4.z = acc[y][x][2]; o4.w = acc[y][x][3];
        *(reinterpret_cast<V*>(dx) + ((static_cast<long long>(n) * H + h0 + y) * W + w0 + x) * c4n + c4) = o4;
      }
    }
  }
}

template <typename T>
static int maxpool_idx_fwd_t(zb_ctx* ctx, const PoolGeom& g, const T* x, T* y, void* idx) {
  const long long total = g.N * g.P * g.Q * (g.C >> 2);
  if (total == 0) return ZB_OK;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 32ll));
  if (g.N * g.H * g.W * g.C < (1ll << 31) - (1ll << 24)) {
    if (g.kh == 3 && g.kw == 3) maxpool_idx_fwd_kernel<T, int, 3, 3><<<grid, 256, 0, ctx->stream>>>(g, x, y, static_cast<uchar4*>(idx));
    else maxpool_idx_fwd_kernel<T, int><<<grid, 256, 0, ctx->stream>>>(g, x, y, static_cast<uchar4*>(idx));
  } else
    maxpool_idx_fwd_kernel<T, long long><<<grid, 256, 0, ctx->stream>>>(g, x, y, static_cast<uchar4*>(idx));
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
template <typename T>
static int maxpool_idx_bwd_t(zb_ctx* ctx, const PoolGeom& g, const T* dy, const void* idx, T* dx) {
  const long long total = g.N * g.H * g.W * (g.C >> 2);
  if (total == 0) return ZB_OK;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 32ll));
  if (g.sh == 2 && g.sw == 2 && g.N * (g.H + 1) * (g.W + 1) * g.C < (1ll << 31) - (1ll << 24)) {
    const long long patches = g.N * ((g.H + 1) / 2) * ((g.W + 1) / 2) * (g.C >> 2);
    const int grid2 = static_cast<int>(std::min<long long>((patches + 255) / 256, ctx->sm_count * 32ll));
    maxpool_idx_bwd_s2_kernel<T><<<grid2, 256, 0, ctx->stream>>>(g, dy, static_cast<const uchar4*>(idx), dx);
  } else if (g.N * g.H * g.W * g.C < (1ll << 31) - (1ll << 24))
    maxpool_idx_bwd_kernel<T, int><<<grid, 256, 0, ctx->stream>>>(g, dy, static_cast<const uchar4*>(idx), dx);
  else
    maxpool_idx_bwd_kernel<T, long long><<<grid, 256, 0, ctx->stream>>>(g, dy, static_cast<const uchar4*>(idx), dx);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// ---- global average pool: x [N][HW][C] (NHWC) or [N][C][HW] (NCHW) -> y [N][C] ------------------------------
template <typename T>
__global__ void __launch_bounds__(256) gap_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long long N, long long C,
                                                      long long HW, int nhwc) {
  const long long total = N * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / C, c = i - n * C;
    T s = T(0);
    if (nhwc) {
      const T* p = x + n * HW * C + c;
      for (long long j = 0; j < HW; ++j) s += p[j * C];
    } else {
      const T* p = x + i * HW;
      for (long long j = 0; j < HW; ++j) s += p[j];
    }
    y[i] = s / static_cast<T>(HW);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) gap_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, long long N, long long C,
                                                      long long HW, int nhwc) {
  const long long total = N * C * HW;
  const T inv = T(1) / static_cast<T>(HW);
  if (nhwc && total < (1ll << 31)) {   // 32-bit index math: this kernel is pure address arithmetic around one store
    const int c_n = static_cast<int>(C), chw = static_cast<int>(C * HW), tot = static_cast<int>(total);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
      const int n = i / chw, c = i % c_n;
      dx[i] = dy[n * c_n + c] * inv;
    }
    return;
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long n, c;
    if (nhwc) { c = i % C; n = i / (C * HW); } else { const long long nc = i / HW; c = nc % C; n = nc / C; }
    dx[i] = dy[n * C + c] * inv;
  }
}

// ---- softmax + cross entropy ------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T block_reduce(T v, T* sh, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (other > v ? other : v) : v + other;
  }
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  T r = sh[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = is_max ? (sh[w] > r ? sh[w] : r) : r + sh[w];
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_xent_rows(const T* __restrict__ z, const T* __restrict__ t, T* __restrict__ row_loss,
                                                         T* __restrict__ dz, long long B, long long K) {
  __shared__ T sh[8];
  const long long b = blockIdx.x;
  const T* zb_ = z + b * K;
  const T* tb = t + b * K;
  T mx = -INFINITY;
  for (long long j = threadIdx.x; j < K; j += blockDim.x) mx = zb_[j] > mx ? zb_[j] : mx;
  mx = block_reduce<T>(mx, sh, true);
  T se = T(0), ts = T(0);
  for (long long j = threadIdx.x; j < K; j += blockDim.x) { se += exp(zb_[j] - mx); ts += tb[j]; }
  se = block_reduce<T>(se, sh, false);
  ts = block_reduce<T>(ts, sh, false);
  T l = T(0);
  for (long long j = threadIdx.x; j < K; j += blockDim.x) {
    const T p = exp(zb_[j] - mx) / se;
    if (tb[j] != T(0)) l += tb[j] * log(p);
    if (dz) dz[b * K + j] = (p * ts - tb[j]) / static_cast<T>(B);
  }
  l = block_reduce<T>(l, sh, false);
  if (threadIdx.x == 0) row_loss[b] = l;
}
template <typename T>
__global__ void loss_finalize(const T* __restrict__ row_loss, T* __restrict__ loss, long long B) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    T s = T(0);
    for (long long b = 0; b < B; ++b) s += row_loss[b];
    *loss = -s / static_cast<T>(B);
  }
}

template <typename T>
static int softmax_xent_t(zb_ctx* ctx, const T* z, const T* t, T* loss, T* dz, long long B, long long K) {
  ZB_REQUIRE(B > 0 && K > 0 && B <= 2147483647ll, "softmax_xent: bad shape");
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(T) * B, &ws);
  if (rc != ZB_OK) return rc;
  softmax_xent_rows<T><<<static_cast<unsigned>(B), 256, 0, ctx->stream>>>(z, t, static_cast<T*>(ws), dz, B, K);
  ZB_LAUNCH_CHECK(ctx);
  loss_finalize<T><<<1, 32, 0, ctx->stream>>>(static_cast<const T*>(ws), loss, B);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

}  // namespace zb

using namespace zb;

// argument validation shared by the pooling entry points (the conv entry points do the same through check_desc, api.cu)
static int check_pool_args(int layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw,
                           int64_t ph, int64_t pw) {
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC, "pool: unknown layout %d", layout);
  ZB_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0 && kh > 0 && kw > 0, "pool: non-positive extent");
  ZB_REQUIRE(sh > 0 && sw > 0 && ph >= 0 && pw >= 0, "pool: bad stride / padding");
  ZB_REQUIRE(h + 2 * ph >= kh && w + 2 * pw >= kw, "pool: window larger than the padded input");
  ZB_REQUIRE(kh <= 255 && kw <= 255 && n * c * h * w < (1ll << 40), "pool: extent too large");
  return ZB_OK;
}

extern "C" {

int zb_maxpool2d_fwd(zb_ctx* ctx, int dtype, int layout, const void* x, void* y, int64_t n, int64_t c, int64_t h, int64_t w,
                     int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph, int64_t pw) {
  ZB_API_RANGE();
  { const int rc = check_pool_args(layout, n, c, h, w, kh, kw, sh, sw, ph, pw); if (rc != ZB_OK) return rc; }
  const PoolGeom g = pool_geom(layout, n, c, h, w, kh, kw, sh, sw, ph, pw);
  if (dtype == ZB_F32) return maxpool_fwd_t<float>(ctx, g, static_cast<const float*>(x), static_cast<float*>(y));
  if (dtype == ZB_F64) return maxpool_fwd_t<double>(ctx, g, static_cast<const double*>(x), static_cast<double*>(y));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}
int zb_maxpool2d_bwd(zb_ctx* ctx, int dtype, int layout, const void* x, const void* dy, void* dx, int64_t n, int64_t c,
                     int64_t h, int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph, int64_t pw) {
  ZB_API_RANGE();
  { const int rc = check_pool_args(layout, n, c, h, w, kh, kw, sh, sw, ph, pw); if (rc != ZB_OK) return rc; }
  const PoolGeom g = pool_geom(layout, n, c, h, w, kh, kw, sh, sw, ph, pw);
  if (dtype == ZB_F32) return maxpool_bwd_t<float>(ctx, g, static_cast<const float*>(x), static_cast<const float*>(dy), static_cast<float*>(dx));
  if (dtype == ZB_F64) return maxpool_bwd_t<double>(ctx, g, static_cast<const double*>(x), static_cast<const double*>(dy), static_cast<double*>(dx));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}
static int check_idx_pool(int layout, int64_t c, int64_t kh, int64_t kw, const void* a, const void* b) {
  ZB_REQUIRE(layout == ZB_NHWC, "indexed max-pool: NHWC only");
  ZB_REQUIRE(c % 4 == 0 && kh * kw < 255, "indexed max-pool: needs C %% 4 == 0 and fewer than 255 taps");
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(a) & 31) == 0 && (reinterpret_cast<uintptr_t>(b) & 31) == 0, "indexed max-pool: tensors must be 32-byte aligned");
  return ZB_OK;
}
int zb_maxpool2d_fwd_idx(zb_ctx* ctx, int dtype, int layout, const void* x, void* y, void* idx, int64_t n, int64_t c, int64_t h,
                         int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph, int64_t pw) {
  ZB_API_RANGE();
  int rc = check_pool_args(layout, n, c, h, w, kh, kw, sh, sw, ph, pw);
  if (rc != ZB_OK) return rc;
  rc = check_idx_pool(layout, c, kh, kw, x, y);
  if (rc != ZB_OK) return rc;
  const PoolGeom g = pool_geom(layout, n, c, h, w, kh, kw, sh, sw, ph, pw);
  if (dtype == ZB_F32) return maxpool_idx_fwd_t<float>(ctx, g, static_cast<const float*>(x), static_cast<float*>(y), idx);
  if (dtype == ZB_F64) return maxpool_idx_fwd_t<double>(ctx, g, static_cast<const double*>(x), static_cast<double*>(y), idx);
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}
int zb_maxpool2d_bwd_idx(zb_ctx* ctx, int dtype, int layout, const void* dy, const void* idx, void* dx, int64_t n, int64_t c,
                         int64_t h, int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph, int64_t pw) {
  ZB_API_RANGE();
  int rc = check_pool_args(layout, n, c, h, w, kh, kw, sh, sw, ph, pw);
  if (rc != ZB_OK) return rc;
  rc = check_idx_pool(layout, c, kh, kw, dy, dx);
  if (rc != ZB_OK) return rc;
  const PoolGeom g = pool_geom(layout, n, c, h, w, kh, kw, sh, sw, ph, pw);
  if (dtype == ZB_F32) return maxpool_idx_bwd_t<float>(ctx, g, static_cast<const float*>(dy), idx, static_cast<float*>(dx));
  if (dtype == ZB_F64) return maxpool_idx_bwd_t<double>(ctx, g, static_cast<const double*>(dy), idx, static_cast<double*>(dx));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}
int zb_gap_fwd(zb_ctx* ctx, int dtype, int layout, const void* x, void* y, int64_t n, int64_t c, int64_t hw) {
  ZB_API_RANGE();
  ZB_REQUIRE((layout == ZB_NCHW || layout == ZB_NHWC) && n >= 0 && c >= 0 && hw > 0, "global average pool: bad layout / extent");
  const long long total = n * c;
  if (total == 0) return ZB_OK;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  if (dtype == ZB_F32) gap_fwd_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(x), static_cast<float*>(y), n, c, hw, layout == ZB_NHWC);
  else if (dtype == ZB_F64) gap_fwd_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double*>(x), static_cast<double*>(y), n, c, hw, layout == ZB_NHWC);
  else { zb::set_last_error("unknown dtype %d", dtype); return ZB_ERR_INVALID; }
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
int zb_gap_bwd(zb_ctx* ctx, int dtype, int layout, const void* dy, void* dx, int64_t n, int64_t c, int64_t hw) {
  ZB_API_RANGE();
  ZB_REQUIRE((layout == ZB_NCHW || layout == ZB_NHWC) && n >= 0 && c >= 0 && hw > 0, "global average pool: bad layout / extent");
  const long long total = n * c * hw;
  if (total == 0) return ZB_OK;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  if (dtype == ZB_F32) gap_bwd_kernel<float><<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(dy), static_cast<float*>(dx), n, c, hw, layout == ZB_NHWC);
  else if (dtype == ZB_F64) gap_bwd_kernel<double><<<grid, 256, 0, ctx->stream>>>(static_cast<const double*>(dy), static_cast<double*>(dx), n, c, hw, layout == ZB_NHWC);
  else { zb::set_last_error("unknown dtype %d", dtype); return ZB_ERR_INVALID; }
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
int zb_softmax_xent(zb_ctx* ctx, int dtype, const void* z, const void* t, void* loss, void* dz, int64_t batch, int64_t classes) {
  ZB_API_RANGE();
  if (dtype == ZB_F32) return softmax_xent_t<float>(ctx, static_cast<const float*>(z), static_cast<const float*>(t), static_cast<float*>(loss), static_cast<float*>(dz), batch, classes);
  if (dtype == ZB_F64) return softmax_xent_t<double>(ctx, static_cast<const double*>(z), static_cast<const double*>(t), static_cast<double*>(loss), static_cast<double*>(dz), batch, classes);
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

}  // extern "C"
