// conv_simt.cu — FFMA (f32) / DFMA (f64) implicit-GEMM kernels: the "fp32 / f64, no tensor core" math mode
// of conv fprop / dgrad / wgrad and GEMM (BASELINE north_star: "f64 runs as FFMA/DFMA kernels with no CPU
// fallback"; 1e-5 tolerance mode for f32).  Also serves shapes the TMA path cannot take (C % 32 != 0,
// unaligned pitches).  Layout-agnostic: every tensor is addressed through logical (n,c,h,w) element strides,
// so NCHW/KCRS (reference contract) and NHWC/KRSC run through the same code.
//
// Replaces the same reference entry points as umma_gemm.cu (conv.cpp:29-187, cublas/mod.rs:84-160).
#include <algorithm>

#include "common.cuh"

namespace zb {

struct Strides4 {
  long long n, c, h, w;
};

struct SimtConvParams {
  long long N, C, H, W, K, R, S, P, Q;
  int pad_h, pad_w, stride_h, stride_w, dil_h, dil_w;
  Strides4 xs, ws /* (k, c, r, s) */, ys;
  long long Mg, Ng, Kg;  // GEMM extents of this pass
  long long k_per_split;
  long long split_stride;  // wgrad split-K: elements between the partial filter gradients of consecutive K slices
};

enum SimtMode { SIMT_FPROP = 0, SIMT_DGRAD = 1, SIMT_WGRAD = 2 };

constexpr int SBM = 64, SBN = 64, SBK = 16;

// A(m, k) element of the implicit GEMM
template <typename T, int MODE>
__device__ __forceinline__ T load_a(const SimtConvParams& p, const T* __restrict__ src, long long m, long long k) {
  if (m >= p.Mg || k >= p.Kg) return T(0);
  if (MODE == SIMT_FPROP) {  // src = x; m = (n,p,q); k = (c,r,s)
    const long long q = m % p.Q, t = m / p.Q, pp = t % p.P, n = t / p.P;
    const long long s = k % p.S, t2 = k / p.S, r = t2 % p.R, c = t2 / p.R;
    const long long ih = pp * p.stride_h - p.pad_h + r * p.dil_h, iw = q * p.stride_w - p.pad_w + s * p.dil_w;
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return T(0);
    return src[n * p.xs.n + c * p.xs.c + ih * p.xs.h + iw * p.xs.w];
  } else if (MODE == SIMT_DGRAD) {  // src = dy; m = (n,ih,iw); k = (ko,r,s)
    const long long iw = m % p.W, t = m / p.W, ih = t % p.H, n = t / p.H;
    const long long s = k % p.S, t2 = k / p.S, r = t2 % p.R, ko = t2 / p.R;
    const long long th = ih + p.pad_h - r * p.dil_h, tw = iw + p.pad_w - s * p.dil_w;
    if (th < 0 || tw < 0 || th % p.stride_h != 0 || tw % p.stride_w != 0) return T(0);
    const long long pp = th / p.stride_h, q = tw / p.stride_w;
    if (pp >= p.P || q >= p.Q) return T(0);
    return src[n * p.ys.n + ko * p.ys.c + pp * p.ys.h + q * p.ys.w];
  } else {  // WGRAD: src = dy; m = ko; k = (n,p,q)
    const long long q = k % p.Q, t = k / p.Q, pp = t % p.P, n = t / p.P;
    return src[n * p.ys.n + m * p.ys.c + pp * p.ys.h + q * p.ys.w];
  }
}

// B(k, j) element
template <typename T, int MODE>
__device__ __forceinline__ T load_b(const SimtConvParams& p, const T* __restrict__ src, long long k, long long j) {
  if (j >= p.Ng || k >= p.Kg) return T(0);
  if (MODE == SIMT_FPROP) {  // src = w; j = ko
    const long long s = k % p.S, t2 = k / p.S, r = t2 % p.R, c = t2 / p.R;
    return src[j * p.ws.n + c * p.ws.c + r * p.ws.h + s * p.ws.w];
  } else if (MODE == SIMT_DGRAD) {  // src = w; j = c
    const long long s = k % p.S, t2 = k / p.S, r = t2 % p.R, ko = t2 / p.R;
    return src[ko * p.ws.n + j * p.ws.c + r * p.ws.h + s * p.ws.w];
  } else {  // WGRAD: src = x; j = (c,r,s); k = (n,p,q)
    const long long q = k % p.Q, t = k / p.Q, pp = t % p.P, n = t / p.P;
    const long long s = j % p.S, t2 = j / p.S, r = t2 % p.R, c = t2 / p.R;
    const long long ih = pp * p.stride_h - p.pad_h + r * p.dil_h, iw = q * p.stride_w - p.pad_w + s * p.dil_w;
    if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return T(0);
    return src[n * p.xs.n + c * p.xs.c + ih * p.xs.h + iw * p.xs.w];
  }
}

template <typename T, int MODE>
__device__ __forceinline__ long long out_index(const SimtConvParams& p, long long m, long long j) {
  if (MODE == SIMT_FPROP) {
    const long long q = m % p.Q, t = m / p.Q, pp = t % p.P, n = t / p.P;
    return n * p.ys.n + j * p.ys.c + pp * p.ys.h + q * p.ys.w;
  } else if (MODE == SIMT_DGRAD) {
    const long long iw = m % p.W, t = m / p.W, ih = t % p.H, n = t / p.H;
    return n * p.xs.n + j * p.xs.c + ih * p.xs.h + iw * p.xs.w;
  } else {
    const long long s = j % p.S, t2 = j / p.S, r = t2 % p.R, c = t2 / p.R;
    return m * p.ws.n + c * p.ws.c + r * p.ws.h + s * p.ws.w;
  }
}

// 64x64 output tile, 16-deep K slices, 256 threads, 4x4 register micro-tile, double-buffered smem.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) simt_conv_kernel(const SimtConvParams p, const T* __restrict__ a_src,
                                                        const T* __restrict__ b_src, const T* __restrict__ bias,
                                                        T* __restrict__ out) {
  __shared__ T sa[2][SBK][SBM + 4];
  __shared__ T sb[2][SBK][SBN + 4];
  const int tid = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * SBM, n0 = static_cast<long long>(blockIdx.y) * SBN;
  const long long kbeg = static_cast<long long>(blockIdx.z) * p.k_per_split;
  const long long kend = (kbeg + p.k_per_split < p.Kg) ? kbeg + p.k_per_split : p.Kg;
  const int tx = tid & 15, ty = tid >> 4;  // micro-tile position: rows ty*4.., cols tx*4..
  // load mapping: A: k = tid & 15, m = (tid >> 4) + 16*i ; B: j = tid & 63, k = (tid >> 6) + 4*i
  const int la_k = tid & 15, la_m = tid >> 4;
  const int lb_j = tid & 63, lb_k = tid >> 6;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  T ra[4], rb[4];
  auto fetch = [&](long long k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long kk = k0 + la_k;
      ra[i] = (kk < kend) ? load_a<T, MODE>(p, a_src, m0 + la_m + 16 * i, kk) : T(0);
      const long long kb = k0 + lb_k + 4 * i;
      rb[i] = (kb < kend) ? load_b<T, MODE>(p, b_src, kb, n0 + lb_j) : T(0);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sa[buf][la_k][la_m + 16 * i] = ra[i];
      sb[buf][lb_k + 4 * i][lb_j] = rb[i];
    }
  };
  if (kbeg < kend) {
    fetch(kbeg);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (long long k0 = kbeg; k0 < kend; k0 += SBK) {
    const bool more = (k0 + SBK) < kend;
    if (more) fetch(k0 + SBK);
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      T av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sa[buf][kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sb[buf][kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.Mg) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= p.Ng) continue;
      const long long o = out_index<T, MODE>(p, m, n);
      if (gridDim.z > 1) {   // split-K: this slice's partial; folded in a fixed order afterwards (no atomics: run-to-run deterministic)
        out[static_cast<long long>(blockIdx.z) * p.split_stride + o] = acc[i][j];
      } else {
        T v = acc[i][j];
        if (MODE == SIMT_FPROP && bias != nullptr) v += bias[n];
        out[o] = v;
      }
    }
  }
}

// dw[i] = sum over the K slices of partial[s][i], s ascending (fixed order)
template <typename T>
__global__ void simt_fold_kernel(const T* __restrict__ partial, T* __restrict__ out, long long n, int splits) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    T acc = T(0);
    for (int s = 0; s < splits; ++s) acc += partial[static_cast<long long>(s) * n + i];
    out[i] = acc;
  }
}

static Strides4 act_strides(int layout, long long C, long long H, long long W) {
  Strides4 s;
  if (layout == ZB_NCHW) { s.n = C * H * W; s.c = H * W; s.h = W; s.w = 1; }
  else { s.n = H * W * C; s.c = 1; s.h = W * C; s.w = C; }
  return s;
}
static Strides4 filt_strides(int layout, long long C, long long R, long long S) {
  Strides4 s;  // (k, c, r, s)
  if (layout == ZB_NCHW) { s.n = C * R * S; s.c = R * S; s.h = S; s.w = 1; }  // KCRS
  else { s.n = R * S * C; s.c = 1; s.h = S * C; s.w = C; }                  // KRSC
  return s;
}

static void fill_params(SimtConvParams& p, int layout, const zb_conv2d_desc* d) {
  p.N = d->n; p.C = d->c; p.H = d->h; p.W = d->w; p.K = d->k; p.R = d->kh; p.S = d->kw;
  p.split_stride = 0;
  p.P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  p.Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  p.pad_h = static_cast<int>(d->pad_h); p.pad_w = static_cast<int>(d->pad_w);
  p.stride_h = static_cast<int>(d->stride_h); p.stride_w = static_cast<int>(d->stride_w);
  p.dil_h = static_cast<int>(d->dil_h); p.dil_w = static_cast<int>(d->dil_w);
  p.xs = act_strides(layout, d->c, d->h, d->w);
  p.ys = act_strides(layout, d->k, p.P, p.Q);
  p.ws = filt_strides(layout, d->c, d->kh, d->kw);
}

template <typename T>
int simt_conv_fprop(zb_ctx* ctx, int layout, const zb_conv2d_desc* d, const T* x, const T* w, const T* bias, T* y) {
  SimtConvParams p;
  fill_params(p, layout, d);
  p.Mg = p.N * p.P * p.Q; p.Ng = p.K; p.Kg = p.C * p.R * p.S; p.k_per_split = p.Kg;
  dim3 grid(ceil_div(p.Mg, SBM), ceil_div(p.Ng, SBN), 1);
  plan_note("simt_conv_fprop<%s> layout=%d;", sizeof(T) == 8 ? "f64" : "f32", layout);
  ZB_KLAUNCH(ctx, simt_conv_kernel<T, SIMT_FPROP><<<grid, 256, 0, ctx->stream>>>(p, x, w, bias, y));
  return ZB_OK;
}

template <typename T>
int simt_conv_dgrad(zb_ctx* ctx, int layout, const zb_conv2d_desc* d, const T* dy, const T* w, T* dx) {
  SimtConvParams p;
  fill_params(p, layout, d);
  p.Mg = p.N * p.H * p.W; p.Ng = p.C; p.Kg = p.K * p.R * p.S; p.k_per_split = p.Kg;
  dim3 grid(ceil_div(p.Mg, SBM), ceil_div(p.Ng, SBN), 1);
  plan_note("simt_conv_dgrad<%s> layout=%d;", sizeof(T) == 8 ? "f64" : "f32", layout);
  ZB_KLAUNCH(ctx, simt_conv_kernel<T, SIMT_DGRAD><<<grid, 256, 0, ctx->stream>>>(p, dy, w, static_cast<const T*>(nullptr), dx));
  return ZB_OK;
}

template <typename T>
int simt_conv_wgrad(zb_ctx* ctx, int layout, const zb_conv2d_desc* d, const T* dy, const T* x, T* dw) {
  SimtConvParams p;
  fill_params(p, layout, d);
  p.Mg = p.K; p.Ng = p.C * p.R * p.S; p.Kg = p.N * p.P * p.Q;
  const long long tiles = static_cast<long long>(ceil_div(p.Mg, SBM)) * ceil_div(p.Ng, SBN);
  long long splits = std::max<long long>(1, std::min<long long>((4ll * ctx->sm_count + tiles - 1) / tiles, p.Kg / 256));
  splits = std::min<long long>(splits, 65535);
  p.k_per_split = ((p.Kg + splits - 1) / splits + SBK - 1) / SBK * SBK;
  splits = (p.Kg + p.k_per_split - 1) / p.k_per_split;
  dim3 grid(ceil_div(p.Mg, SBM), ceil_div(p.Ng, SBN), static_cast<unsigned>(splits));
  plan_note("simt_conv_wgrad<%s> layout=%d splitk=%d ~splits=%lld;", sizeof(T) == 8 ? "f64" : "f32", layout, splits > 1 ? 1 : 0, splits);
  if (splits <= 1) {
    ZB_KLAUNCH(ctx, simt_conv_kernel<T, SIMT_WGRAD><<<grid, 256, 0, ctx->stream>>>(p, dy, x, static_cast<const T*>(nullptr), dw));
    return ZB_OK;
  }
  // split-K: one partial filter gradient per K slice in the scratch arena, then a fixed-order fold (round 1 used atomicAdd here, which
  // made ZB_MATH_FP32 and every f64 wgrad run-to-run nondeterministic)
  const long long total = static_cast<long long>(p.K) * p.C * p.R * p.S;
  void* ws = nullptr;
  const int rc = ctx_workspace(ctx, sizeof(T) * static_cast<size_t>(splits) * total, &ws);
  if (rc != ZB_OK) return rc;
  p.split_stride = total;
  ZB_KLAUNCH(ctx, simt_conv_kernel<T, SIMT_WGRAD><<<grid, 256, 0, ctx->stream>>>(p, dy, x, static_cast<const T*>(nullptr), static_cast<T*>(ws)));
  const int fgrid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 8ll));
  ZB_KLAUNCH(ctx, simt_fold_kernel<T><<<fgrid, 256, 0, ctx->stream>>>(static_cast<const T*>(ws), dw, total, static_cast<int>(splits)));
  return ZB_OK;
}

// ---------------------------------------------------------------------------------------------- GEMM
template <typename T>
__global__ void __launch_bounds__(256) simt_gemm_kernel(int ta, int tb, long long M, long long N, long long K, T alpha,
                                                        const T* __restrict__ a, long long lda,
                                                        const T* __restrict__ b, long long ldb, T beta,
                                                        T* __restrict__ c, long long ldc, const T* __restrict__ bias) {
  __shared__ T sa[SBK][SBM + 4];
  __shared__ T sb[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const long long m0 = static_cast<long long>(blockIdx.x) * SBM, n0 = static_cast<long long>(blockIdx.y) * SBN;
  const int tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (long long k0 = 0; k0 < K; k0 += SBK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      {  // A tile: choose the thread->element map that walks the contiguous dimension
        int mm, kk;
        if (ta) { mm = tid & 63; kk = (tid >> 6) + 4 * i; } else { kk = tid & 15; mm = (tid >> 4) + 16 * i; }
        const long long m = m0 + mm, k = k0 + kk;
        sa[kk][mm] = (m < M && k < K) ? (ta ? a[k * lda + m] : a[m * lda + k]) : T(0);
      }
      {
        int nn, kk;
        if (tb) { kk = tid & 15; nn = (tid >> 4) + 16 * i; } else { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
        const long long n = n0 + nn, k = k0 + kk;
        sb[kk][nn] = (n < N && k < K) ? (tb ? b[n * ldb + k] : b[k * ldb + n]) : T(0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SBK; ++kk) {
      T av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = sa[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = sb[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long n = n0 + tx * 4 + j;
      if (n >= N) continue;
      T v = alpha * acc[i][j];
      if (bias) v += bias[n];
      if (beta != T(0)) v += beta * c[m * ldc + n];
      c[m * ldc + n] = v;
    }
  }
}

template <typename T>
int simt_gemm(zb_ctx* ctx, bool ta, bool tb, long long m, long long n, long long k, T alpha, const T* a, long long lda,
              const T* b, long long ldb, T beta, T* c, long long ldc, const T* bias) {
  dim3 grid(ceil_div(m, SBM), ceil_div(n, SBN), 1);
  plan_note("simt_gemm<%s>;", sizeof(T) == 8 ? "f64" : "f32");
  ZB_KLAUNCH(ctx, simt_gemm_kernel<T><<<grid, 256, 0, ctx->stream>>>(ta ? 1 : 0, tb ? 1 : 0, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, bias));
  return ZB_OK;
}

template int simt_conv_fprop<float>(zb_ctx*, int, const zb_conv2d_desc*, const float*, const float*, const float*, float*);
template int simt_conv_fprop<double>(zb_ctx*, int, const zb_conv2d_desc*, const double*, const double*, const double*, double*);
template int simt_conv_dgrad<float>(zb_ctx*, int, const zb_conv2d_desc*, const float*, const float*, float*);
template int simt_conv_dgrad<double>(zb_ctx*, int, const zb_conv2d_desc*, const double*, const double*, double*);
template int simt_conv_wgrad<float>(zb_ctx*, int, const zb_conv2d_desc*, const float*, const float*, float*);
template int simt_conv_wgrad<double>(zb_ctx*, int, const zb_conv2d_desc*, const double*, const double*, double*);
template int simt_gemm<float>(zb_ctx*, bool, bool, long long, long long, long long, float, const float*, long long, const float*, long long, float, float*, long long, const float*);
template int simt_gemm<double>(zb_ctx*, bool, bool, long long, long long, long long, double, const double*, long long, const double*, long long, double, double*, long long, const double*);

}  // namespace zb
