// batchnorm.cu — BatchNorm2d forward (train / inference) and backward, with optional fused ReLU and
// residual add, for NHWC (native) and NCHW (reference contract) tensors, f32 and f64.
//
// Replaces cudnnBatchNormalizationForwardTraining / Backward / ForwardInference (reference
// zenu-cuda/src/cudnn/batch_norm.rs:51-91,206-249,380-414) with the semantics of the reference CPU path
// (zenu-matrix/src/nn/batch_norm.rs:283-420): eps 1e-10, momentum on the OLD running stat, unbiased running
// variance, saved mean and 1/sqrt(var+eps).  Also hosts the per-channel column reductions used by
// conv2d_bias_bkwd and Matrix::sum(axis 0).
//
// Structure (every kernel is HBM-bound; 128-bit loads, no atomics, deterministic):
//   reduce  : grid (channel groups, row slabs); each thread owns VN channels and walks rows with 4 loads in
//             flight; block-level smem reduction -> partial[slab][stat][C]
//   finalize: one thread per channel folds the slab partials in double precision
//   apply   : same geometry as reduce; per-channel coefficients live in registers
// Statistics use a per-channel shift (the first row) so the single-pass sum / sum-of-squares does not cancel.
// Grids are one wave of the CTAs that are resident for the kernel instantiation at hand (resident_ctas); the order in which a
// kernel walks its rows follows what the previous kernel of the step left in L2 (RowWalk).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace zb {

// epsilon: zb_ctx::bn_eps, 1e-10 by default (zenu-matrix/src/nn/batch_norm.rs:296); the cuDNN-frontend BatchNorm shim passes its own

template <typename T, int VN> struct VecT;
template <> struct VecT<float, 4> { using type = float4; };
template <> struct VecT<float, 1> { using type = float; };
template <> struct VecT<double, 2> { using type = double2; };
template <> struct VecT<double, 1> { using type = double; };

template <typename T, int VN>
__device__ __forceinline__ void ldv(const T* p, T* o) {
  using V = typename VecT<T, VN>::type;
  const V v = *reinterpret_cast<const V*>(p);
  const T* s = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int e = 0; e < VN; ++e) o[e] = s[e];
}
// Load of a line that nobody reads again soon (second pass of a BatchNorm): ld.global.cs marks it evict-first in L2, so the pass's
// own dead lines are displaced before the lines it has yet to reach (left there by the previous kernel) and before its output.
template <typename T, int VN>
__device__ __forceinline__ void ldv_once(const T* p, T* o, bool once) {
  using V = typename VecT<T, VN>::type;
  const V v = once ? __ldcs(reinterpret_cast<const V*>(p)) : *reinterpret_cast<const V*>(p);
  const T* s = reinterpret_cast<const T*>(&v);
#pragma unroll
  for (int e = 0; e < VN; ++e) o[e] = s[e];
}
template <typename T, int VN>
__device__ __forceinline__ void stv(T* p, const T* o) {
  using V = typename VecT<T, VN>::type;
  V v;
  T* s = reinterpret_cast<T*>(&v);
#pragma unroll
  for (int e = 0; e < VN; ++e) s[e] = o[e];
  *reinterpret_cast<V*>(p) = v;
}

// y = ((x - mean) * inv) * gamma + beta with a fixed rounding sequence: the backward of a fused BN+ReLU recomputes
// this value from x to rebuild the ReLU mask, so forward and backward must round identically.
__device__ __forceinline__ float bn_affine(float a, float m, float iv, float g, float b) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(a, m), iv), g, b);
}
__device__ __forceinline__ double bn_affine(double a, double m, double iv, double g, double b) {
  return __fma_rn(__dmul_rn(__dsub_rn(a, m), iv), g, b);
}

// Minimum CTAs per SM the three streaming NHWC kernels (statistics pass, forward apply, backward apply) are compiled for; the grids
// follow the resulting residency through resident_ctas().  3 caps them at 85 registers: the two-stream BN+add+ReLU forward apply
// (92 registers left alone, i.e. 2 CTAs = 64 KB of loads in flight per SM) gets a third CTA, the others may use more registers than
// the default heuristic gives them and stay at 3.  Measured per ResNet-50 step (separate builds, one box): unconstrained 34.48 ms,
// 3: 34.09 ms, 4 (64 registers, spills): 35.63 ms; 8 instead of 4 rows in flight in the single-stream apply: no change.
#ifndef ZB_BN_MIN_CTAS
#define ZB_BN_MIN_CTAS 3
#endif

struct ColGeom {
  int tx, ty, col_groups, slabs;
  long long rows_per_slab, cvecs;
};

// CTAs of `kernel` (256 threads, `smem` dynamic bytes) that are resident on one SM.  The NHWC kernels hold 52 - 92 registers per thread
// (four 128-bit loads per stream in flight), i.e. 2 - 4 CTAs per SM; their grids are sized to exactly ONE wave of that, so that no
// trailing partial wave is left and a ROWS_SWEEP walk really is one front through the tensor (a grid of 8 CTAs per SM, which the
// register file does not hold, ran as 2.67 waves with the last one two thirds empty).
template <typename K>
static int resident_ctas(K kernel, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 256, smem) != cudaSuccess) { cudaGetLastError(); n = 0; }
  return std::min(8, std::max(1, n));
}

static ColGeom col_geom(zb_ctx* ctx, long long rows, long long C, int vn, int resident = 8) {
  ColGeom g;
  g.cvecs = C / vn;
  int tx = 1;
  while (tx < 32 && tx < g.cvecs) tx <<= 1;
  g.tx = tx;
  g.ty = 256 / tx;
  g.col_groups = static_cast<int>((g.cvecs + tx - 1) / tx);
  static int ctas_knob = -1;   // CTAs per SM the grid is sized for, overriding the kernel's own residency (tuning knob: ZENU_B200_BN_CTAS)
  if (ctas_knob < 0) {
    const char* e = getenv("ZENU_B200_BN_CTAS");
    ctas_knob = e ? std::min(8, std::max(1, atoi(e))) : 0;
  }
  const int ctas_per_sm = ctas_knob > 0 ? ctas_knob : resident;
  long long max_slabs = std::max<long long>(1, (ctx->sm_count * static_cast<long long>(ctas_per_sm)) / g.col_groups);
  static int rows_per_thread = -1;   // rows each thread walks per slab (tuning knob: ZENU_B200_BN_ROWS)
  if (rows_per_thread < 0) {
    const char* e = getenv("ZENU_B200_BN_ROWS");
    rows_per_thread = e ? std::max(4, atoi(e)) : 16;
  }
  long long want = (rows + g.ty * static_cast<long long>(rows_per_thread) - 1) / (g.ty * static_cast<long long>(rows_per_thread));
  g.slabs = static_cast<int>(std::max<long long>(1, std::min<long long>(std::min(want, max_slabs), 65535)));
  g.rows_per_slab = (rows + g.slabs - 1) / g.slabs;
  g.slabs = static_cast<int>((rows + g.rows_per_slab - 1) / g.rows_per_slab);
  return g;
}

// Order in which an NHWC kernel of this file walks its rows.  For the apply kernels the result does not depend on it (for the
// reduce kernels only the summation order does); what it decides is which part of the tensor a kernel touches FIRST, i.e.
// whether it finds the lines its producer touched LAST still in the 126 MB L2.
//   ROWS_SLAB_UP     the grid's y index owns one contiguous slab and walks it upwards
//   ROWS_SLAB_DOWN   same slabs, walked downwards: a second pass that starts where a ROWS_SLAB_UP pass over the same slabs ended
//   ROWS_SWEEP_DOWN  small slabs handed out round-robin from the END of the tensor, so the whole grid moves through memory as
//                    one front, last row first: follows a convolution, whose persistent CTAs write their row tiles in
//                    ascending order, and leaves the FIRST rows of its own output in L2 for the next convolution
//   ROWS_SWEEP_UP    the same front, first row first: the second pass after a ROWS_SWEEP_DOWN statistics pass
enum { ROWS_SLAB_UP = 0, ROWS_SLAB_DOWN = 1, ROWS_SWEEP_DOWN = 2, ROWS_SWEEP_UP = 3 };
struct RowWalk {
  long long rows_per_slab, nslabs;
  int sweep, down;
  int once;   // apply kernels: the inputs are read with ldv_once
};
#define ZB_ROW_WALK_BEGIN(walk, rows, ty, ty_n)                                                                        \
  {                                                                                                                    \
    long long zb_s = ((walk).sweep && (walk).down) ? (walk).nslabs - 1 - blockIdx.y : blockIdx.y;                      \
    const long long zb_sstep = !(walk).sweep ? (walk).nslabs                                                           \
                                             : (walk).down ? -static_cast<long long>(gridDim.y) : static_cast<long long>(gridDim.y); \
    const long long step = (walk).down ? -(ty_n) : (ty_n);                                                             \
    for (; zb_s >= 0 && zb_s < (walk).nslabs; zb_s += zb_sstep) {                                                      \
      const long long r0 = zb_s * (walk).rows_per_slab;                                                                \
      const long long r1 = (r0 + (walk).rows_per_slab < (rows)) ? r0 + (walk).rows_per_slab : (rows);                  \
      long long r = (walk).down ? r1 - 1 - (ty) : r0 + (ty);                                                           \
      auto zb_in = [r0, r1](long long q) { return q >= r0 && q < r1; };
#define ZB_ROW_WALK_END \
    }                   \
  }

// Row-order policy (A/B knob ZENU_B200_BN_ORDER): 0 = every kernel walks its slab upwards (round 1); 1 = the forward apply follows
// the conv that produced x, the backward apply walks the statistics pass's slabs downwards; 2 (default) = additionally the backward
// statistics pass sweeps downwards behind the dgrad that produced dy and the backward apply sweeps back up.
static int bn_row_order_mode() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("ZENU_B200_BN_ORDER"); mode = e ? atoi(e) : 2; }
  return mode;
}

static RowWalk make_walk(const ColGeom& g, long long rows, int order) {
  RowWalk w;
  static int once = -1;   // A/B knob: ZENU_B200_BN_LDCS=0 loads everything with the default policy
  if (once < 0) { const char* e = getenv("ZENU_B200_BN_LDCS"); once = e ? atoi(e) : 1; }
  w.once = once;
  w.sweep = order >= ROWS_SWEEP_DOWN;
  w.down = order == ROWS_SLAB_DOWN || order == ROWS_SWEEP_DOWN;
  if (w.sweep) {
    static int per = -1;   // rows every thread row takes of a small slab (tuning knob: ZENU_B200_BN_SWEEP_ROWS)
    if (per < 0) { const char* e = getenv("ZENU_B200_BN_SWEEP_ROWS"); per = e ? std::max(4, atoi(e)) : 8; }
    w.rows_per_slab = g.ty * static_cast<long long>(per);
    w.nslabs = (rows + w.rows_per_slab - 1) / w.rows_per_slab;
  } else {
    w.rows_per_slab = g.rows_per_slab;
    w.nslabs = g.slabs;
  }
  return w;
}

// ---- functors for the NHWC column reduce: NS statistics from up to three input streams -------------------
template <typename T, int VN>
struct StatsF {  // sum(x - shift), sum((x - shift)^2); shift = first row
  static constexpr int NS = 2, NIN = 1;
  const T* x0;
  long long sstride;  // element stride between the first elements of consecutive channels (1 NHWC, H*W NCHW)
  T shift[VN];
  __device__ void init(long long c0) {
#pragma unroll
    for (int e = 0; e < VN; ++e) shift[e] = x0[(c0 + e) * sstride];
  }
  __device__ void operator()(const T* a, const T*, const T*, T (*acc)[VN], long long) const {
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      const T d = a[e] - shift[e];
      acc[0][e] += d;
      acc[1][e] += d * d;
    }
  }
};
template <typename T, int VN>
struct SumF {  // plain column sum
  static constexpr int NS = 1, NIN = 1;
  __device__ void init(long long) {}
  __device__ void operator()(const T* a, const T*, const T*, T (*acc)[VN], long long) const {
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[0][e] += a[e];
  }
};
// MASK 5 = as 4 without writing the masked gradient: the residual branch takes (dy, bits) as it is (a "lazy" masked gradient that
// its consumer -- a dgrad epilogue or the shortcut's BatchNorm backward -- masks on the fly): 5 + 2/32 tensor passes instead of 6 + 1/32
// MASK: 0 = plain BN backward, 1 = ReLU mask from the saved output y (third stream), 2 = ReLU mask recomputed from x
// (fused BN+ReLU without residual: y > 0 <=> bn_affine(x) > 0, so y is never read), 3 = as 1, and the masked gradient
// dy' is also written to gm_out (it IS the residual-branch gradient; the apply pass then reads it instead of dy and y:
// 7 tensor passes instead of 8 for the fused BN+add+ReLU backward)
template <typename T, int VN, int MASK>
struct BnBwdF {  // sum(dy'), sum(dy' * xhat); inputs: x, dy, y(mask)
  static constexpr int NS = 2, NIN = (MASK == 1 || MASK == 3) ? 3 : 2;
  T* gm_out;
  const uint32_t* bits;   // MASK == 4: as 3, with the ReLU mask read from the 1-bit-per-element array the forward wrote
  const T* mean;
  const T* inv;
  const T* gamma;
  const T* beta;
  T m[VN], iv[VN], gm[VN], bt[VN];
  __device__ void init(long long c0) {
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      m[e] = mean[c0 + e]; iv[e] = inv[c0 + e];
      if (MASK == 2) { gm[e] = gamma[c0 + e]; bt[e] = beta[c0 + e]; }
    }
  }
  __device__ void operator()(const T* x, const T* dy, const T* y, T (*acc)[VN], long long off) const {
    T gv[VN];
    unsigned nib = 0u;
    if (MASK == 4 || MASK == 5) nib = __ldg(bits + (off >> 5)) >> (off & 31);   // VN consecutive bits (VN divides 32, off % VN == 0)
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      bool keep = true;
      if (MASK == 1 || MASK == 3) keep = y[e] > T(0);
      if (MASK == 2) keep = bn_affine(x[e], m[e], iv[e], gm[e], bt[e]) > T(0);
      if (MASK == 4 || MASK == 5) keep = ((nib >> e) & 1u) != 0u;
      const T g = keep ? dy[e] : T(0);
      gv[e] = g;
      acc[0][e] += g;
      acc[1][e] += g * ((x[e] - m[e]) * iv[e]);
    }
    if (MASK == 3 || MASK == 4) stv<T, VN>(gm_out + off, gv);
  }
};

// partial layout: [slab][stat][C]
template <typename T, int VN, typename F>
__global__ void __launch_bounds__(256, ZB_BN_MIN_CTAS) col_reduce_nhwc(F f, const T* __restrict__ in0, const T* __restrict__ in1,
                                                       const T* __restrict__ in2, T* __restrict__ partial, long long rows,
                                                       long long C, RowWalk walk, int tx_n, int ty_n) {
  constexpr int NS = F::NS;
  extern __shared__ unsigned char red_raw[];
  T* red = reinterpret_cast<T*>(red_raw);  // [ty][NS][tx*VN]
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const long long cv = static_cast<long long>(blockIdx.x) * tx_n + tx;
  const bool active = cv * VN < C;
  const long long c0 = cv * VN;
  T acc[NS][VN];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int e = 0; e < VN; ++e) acc[s][e] = T(0);
  if (active) {
    f.init(c0);
    constexpr int U = 4;
    ZB_ROW_WALK_BEGIN(walk, rows, ty, ty_n)
    for (; zb_in(r + (U - 1) * step); r += U * step) {
      T a[U][VN], b[U][VN], c[U][VN];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long off = (r + u * step) * C + c0;
        ldv<T, VN>(in0 + off, a[u]);
        if (F::NIN >= 2) ldv<T, VN>(in1 + off, b[u]);
        if (F::NIN >= 3) ldv<T, VN>(in2 + off, c[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) f(a[u], b[u], c[u], acc, (r + u * step) * C + c0);
    }
    for (; zb_in(r); r += step) {
      T a[VN], b[VN], c[VN];
      const long long off = r * C + c0;
      ldv<T, VN>(in0 + off, a);
      if (F::NIN >= 2) ldv<T, VN>(in1 + off, b);
      if (F::NIN >= 3) ldv<T, VN>(in2 + off, c);
      f(a, b, c, acc, off);
    }
    ZB_ROW_WALK_END
  }
  const int row_elems = NS * tx_n * VN;
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int e = 0; e < VN; ++e) red[ty * row_elems + s * tx_n * VN + tx * VN + e] = acc[s][e];
  __syncthreads();
  // tree over ty
  for (int half = ty_n >> 1; half > 0; half >>= 1) {
    if (ty < half) {
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int e = 0; e < VN; ++e) {
          const int idx = s * tx_n * VN + tx * VN + e;
          red[ty * row_elems + idx] += red[(ty + half) * row_elems + idx];
        }
    }
    __syncthreads();
  }
  if (ty == 0 && active) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int e = 0; e < VN; ++e)
        partial[(static_cast<long long>(blockIdx.y) * NS + s) * C + c0 + e] = red[s * tx_n * VN + tx * VN + e];
  }
}

// NCHW: block = (channel, slab of images); threads stride over the H*W plane of each image.
template <typename T, typename F>
__global__ void __launch_bounds__(256) col_reduce_nchw(F f, const T* __restrict__ in0, const T* __restrict__ in1,
                                                       const T* __restrict__ in2, T* __restrict__ partial, long long N,
                                                       long long C, long long HW, long long imgs_per_slab) {
  constexpr int NS = F::NS;
  __shared__ T red[NS][8];
  const long long c = blockIdx.x;
  const long long n0 = static_cast<long long>(blockIdx.y) * imgs_per_slab;
  const long long n1 = (n0 + imgs_per_slab < N) ? n0 + imgs_per_slab : N;
  T acc[NS][1];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s][0] = T(0);
  f.init(c);
  for (long long n = n0; n < n1; ++n) {
    const long long base = (n * C + c) * HW;
    for (long long i = threadIdx.x; i < HW; i += blockDim.x) {
      T a[1], b[1], d[1];
      a[0] = in0[base + i];
      if (F::NIN >= 2) b[0] = in1[base + i];
      if (F::NIN >= 3) d[0] = in2[base + i];
      f(a, b, d, acc, base + i);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const T v = warp_sum(acc[s][0]);
    if (lane == 0) red[s][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      T v = T(0);
      for (int w = 0; w < 8; ++w) v += red[s][w];
      partial[(static_cast<long long>(blockIdx.y) * NS + s) * C + c] = v;
    }
  }
}

// ---- finalize kernels ---------------------------------------------------------------------------------
// Block = 32 channels x 8 slab lanes (256 threads): the slab partials of a channel are folded by 8 threads with
// coalesced loads, then combined through shared memory in double precision.  (A single thread per channel walking
// ~1000 slabs serially cost more than the reduce pass itself on 64-channel layers.)
constexpr int kFinC = 8, kFinS = 128;
template <typename T, int NS>
__device__ __forceinline__ void fold_partials(const T* __restrict__ partial, int slabs, long long C, long long c, bool active,
                                              double (&out)[NS]) {
  // thread = (channel lc, slab lane sl); a warp holds 4 slab lanes x 8 channels: two xor-shuffles fold them, then the 32 warp
  // sums of a channel are added from shared memory.  (8 channels x 128 slab lanes per block: ~10 dependent load rounds of the
  // previous 32 x 32 arrangement become ~3, and a 64-channel layer gets 8 blocks instead of 2: these ~100 tiny kernels per step
  // sit between every BatchNorm reduce and apply pass.)
  __shared__ double sh[NS][kFinS / 4][kFinC];
  const int lc = threadIdx.x & (kFinC - 1), sl = threadIdx.x / kFinC;
  double acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = 0.0;
  if (active) {
#pragma unroll 4
    for (int i = sl; i < slabs; i += kFinS)
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[s] += static_cast<double>(partial[(static_cast<long long>(i) * NS + s) * C + c]);
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], 8);
    acc[s] += __shfl_xor_sync(0xffffffffu, acc[s], 16);
    if ((threadIdx.x & 31) < kFinC) sh[s][threadIdx.x >> 5][lc] = acc[s];
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < kFinS / 4; ++j) v += sh[s][j][lc];
    out[s] = v;
  }
}
#define ZB_FIN_GRID(C) ceil_div((C), kFinC), kFinC * kFinS

// coef layout in workspace: [0]=mean [1]=inv_std (fwd) ; bwd: [0]=gamma*inv [1]=c1 [2]=c2
template <typename T>
__global__ void __launch_bounds__(kFinC * kFinS)
bn_fwd_finalize(const T* __restrict__ partial, int slabs, long long C, double count, double momentum, double eps,
                const T* x_first_row, long long shift_stride, T* run_mean, T* __restrict__ run_var,
                T* __restrict__ saved_mean, T* __restrict__ saved_inv, T* __restrict__ coef) {
  const long long c = blockIdx.x * static_cast<long long>(kFinC) + (threadIdx.x & (kFinC - 1));
  double st[2];
  fold_partials<T, 2>(partial, slabs, C, c, c < C, st);
  if (c >= C || threadIdx.x >= kFinC) return;
  const double s = st[0], ss = st[1];
  const double shift = static_cast<double>(x_first_row[c * shift_stride]);
  const double dm = s / count;
  const double mean = shift + dm;
  double var = ss / count - dm * dm;
  if (var < 0.0) var = 0.0;
  const double inv = 1.0 / sqrt(var + eps);
  if (run_mean) run_mean[c] = static_cast<T>(mean * (1.0 - momentum) + static_cast<double>(run_mean[c]) * momentum);
  if (run_var) {
    const double unbiased = var * (count / (count - 1.0));
    run_var[c] = static_cast<T>(unbiased * (1.0 - momentum) + static_cast<double>(run_var[c]) * momentum);
  }
  if (saved_mean) saved_mean[c] = static_cast<T>(mean);
  if (saved_inv) saved_inv[c] = static_cast<T>(inv);
  coef[c] = static_cast<T>(mean);
  coef[C + c] = static_cast<T>(inv);
}

template <typename T>
__global__ void bn_infer_coef(const T* __restrict__ mean, const T* __restrict__ var, long long C, double eps, T* __restrict__ coef) {
  const long long c = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (c >= C) return;
  coef[c] = mean[c];
  coef[C + c] = static_cast<T>(1.0 / sqrt(static_cast<double>(var[c]) + eps));
}

template <typename T>
__global__ void __launch_bounds__(kFinC * kFinS)
bn_bwd_finalize(const T* __restrict__ partial, int slabs, long long C, double count, const T* __restrict__ scale,
                const T* __restrict__ inv, T* __restrict__ dscale, T* __restrict__ dbias, T* __restrict__ coef) {
  const long long c = blockIdx.x * static_cast<long long>(kFinC) + (threadIdx.x & (kFinC - 1));
  double st[2];
  fold_partials<T, 2>(partial, slabs, C, c, c < C, st);
  if (c >= C || threadIdx.x >= kFinC) return;
  const double s = st[0], sx = st[1];
  dbias[c] = static_cast<T>(s);
  dscale[c] = static_cast<T>(sx);
  coef[c] = static_cast<T>(static_cast<double>(scale[c]) * static_cast<double>(inv[c]));
  coef[C + c] = static_cast<T>(s / count);
  coef[2 * C + c] = static_cast<T>(sx / count);
}

template <typename T>
__global__ void __launch_bounds__(kFinC * kFinS)
sum_finalize(const T* __restrict__ partial, int slabs, long long C, T* __restrict__ out) {
  const long long c = blockIdx.x * static_cast<long long>(kFinC) + (threadIdx.x & (kFinC - 1));
  double st[1];
  fold_partials<T, 1>(partial, slabs, C, c, c < C, st);
  if (c >= C || threadIdx.x >= kFinC) return;
  out[c] = static_cast<T>(st[0]);
}

// ---- apply kernels ------------------------------------------------------------------------------------
// forward: y = ((x - mean) * inv) * gamma + beta  [+ res] [relu]
// ReLU mask, 1 bit per element (bit index = NHWC element index): each thread owns 4 consecutive bits, the 8 lanes that share
// a 32-bit word (consecutive tx, same row) OR their nibbles together and one of them stores the word.  Needs C % 32 == 0.
__device__ __forceinline__ void store_mask_nibble(uint32_t* __restrict__ mask, long long off, unsigned nib) {
  const int l7 = threadIdx.x & 7;
  unsigned w = nib << (l7 * 4);
  const unsigned grp = 0xffu << ((threadIdx.x & 31) & ~7);
  w |= __shfl_xor_sync(grp, w, 1);
  w |= __shfl_xor_sync(grp, w, 2);
  w |= __shfl_xor_sync(grp, w, 4);
  if (l7 == 0) mask[off >> 5] = w;
}

template <typename T, int VN, bool RELU, bool RES>
__global__ void __launch_bounds__(256, ZB_BN_MIN_CTAS) bn_apply_nhwc(const T* __restrict__ x, const T* __restrict__ res,
                                                     T* __restrict__ y, const T* __restrict__ coef,
                                                     const T* __restrict__ gamma, const T* __restrict__ beta,
                                                     long long rows, long long C, RowWalk walk, int tx_n,
                                                     int ty_n, uint32_t* __restrict__ mask = nullptr) {
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const long long c0 = (static_cast<long long>(blockIdx.x) * tx_n + tx) * VN;
  if (c0 >= C) return;
  T m[VN], iv[VN], g[VN], b[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) { m[e] = coef[c0 + e]; iv[e] = coef[C + c0 + e]; g[e] = gamma[c0 + e]; b[e] = beta[c0 + e]; }
  ZB_ROW_WALK_BEGIN(walk, rows, ty, ty_n)
  constexpr int U = 4;
  for (; zb_in(r + (U - 1) * step); r += U * step) {
    T a[U][VN], rr[U][VN];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long off = (r + u * step) * C + c0;
      ldv_once<T, VN>(x + off, a[u], walk.once);
      if (RES) ldv_once<T, VN>(res + off, rr[u], walk.once);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T o[VN];
      unsigned nib = 0u;
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        T v = bn_affine(a[u][e], m[e], iv[e], g[e], b[e]);
        if (RES) v += rr[u][e];
        if (RELU) { nib |= (v > T(0) ? 1u : 0u) << e; v = v > T(0) ? v : T(0); }
        o[e] = v;
      }
      stv<T, VN>(y + (r + u * step) * C + c0, o);
      if (RELU && VN == 4 && mask != nullptr) store_mask_nibble(mask, (r + u * step) * C + c0, nib);
    }
  }
  for (; zb_in(r); r += step) {
    T a[VN], rr[VN], o[VN];
    const long long off = r * C + c0;
    ldv_once<T, VN>(x + off, a, walk.once);
    if (RES) ldv_once<T, VN>(res + off, rr, walk.once);
    unsigned nib = 0u;
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      T v = bn_affine(a[e], m[e], iv[e], g[e], b[e]);
      if (RES) v += rr[e];
      if (RELU) { nib |= (v > T(0) ? 1u : 0u) << e; v = v > T(0) ? v : T(0); }
      o[e] = v;
    }
    stv<T, VN>(y + off, o);
    if (RELU && VN == 4 && mask != nullptr) store_mask_nibble(mask, off, nib);
  }
  ZB_ROW_WALK_END
}

// backward: dx = coef * (dy' - c1 - xhat * c2);  dres = dy'
template <typename T, int VN, int MASK, bool DRES>
__global__ void __launch_bounds__(256, ZB_BN_MIN_CTAS) bn_bwd_apply_nhwc(const T* __restrict__ x, const T* __restrict__ dy,
                                                         const T* __restrict__ y, T* __restrict__ dx,
                                                         T* __restrict__ dres, const T* __restrict__ mean,
                                                         const T* __restrict__ inv, const T* __restrict__ coef,
                                                         const T* __restrict__ gamma, const T* __restrict__ beta,
                                                         long long rows, long long C, RowWalk walk, int tx_n,
                                                         int ty_n, const uint32_t* __restrict__ bits = nullptr) {
  // MASK 4: the gradient is dy masked by a 1-bit-per-element array (bit index = NHWC element index)
  const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
  const long long c0 = (static_cast<long long>(blockIdx.x) * tx_n + tx) * VN;
  if (c0 >= C) return;
  T m[VN], iv[VN], k0[VN], k1[VN], k2[VN], ga[VN], be[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) {
    m[e] = mean[c0 + e]; iv[e] = inv[c0 + e];
    k0[e] = coef[c0 + e]; k1[e] = coef[C + c0 + e]; k2[e] = coef[2 * C + c0 + e];
    if (MASK == 2) { ga[e] = gamma[c0 + e]; be[e] = beta[c0 + e]; }
  }
  ZB_ROW_WALK_BEGIN(walk, rows, ty, ty_n)
  constexpr int U = MASK == 1 ? 2 : 4;
  for (; zb_in(r + (U - 1) * step); r += U * step) {
    T a[U][VN], g[U][VN], yy[U][VN];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long off = (r + u * step) * C + c0;
      ldv_once<T, VN>(x + off, a[u], walk.once);
      ldv_once<T, VN>(dy + off, g[u], walk.once);
      if (MASK == 1) ldv_once<T, VN>(y + off, yy[u], walk.once);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      T o[VN], gm[VN];
      unsigned nib = 0u;
      if (MASK == 4) { const long long off = (r + u * step) * C + c0; nib = __ldg(bits + (off >> 5)) >> (off & 31); }
#pragma unroll
      for (int e = 0; e < VN; ++e) {
        bool keep = true;
        if (MASK == 1) keep = yy[u][e] > T(0);
        if (MASK == 2) keep = bn_affine(a[u][e], m[e], iv[e], ga[e], be[e]) > T(0);
        if (MASK == 4) keep = ((nib >> e) & 1u) != 0u;
        gm[e] = keep ? g[u][e] : T(0);
        o[e] = k0[e] * (gm[e] - k1[e] - ((a[u][e] - m[e]) * iv[e]) * k2[e]);
      }
      const long long off = (r + u * step) * C + c0;
      stv<T, VN>(dx + off, o);
      if (DRES) stv<T, VN>(dres + off, gm);
    }
  }
  for (; zb_in(r); r += step) {
    T a[VN], g[VN], yy[VN], o[VN], gm[VN];
    const long long off = r * C + c0;
    ldv_once<T, VN>(x + off, a, walk.once);
    ldv_once<T, VN>(dy + off, g, walk.once);
    if (MASK == 1) ldv_once<T, VN>(y + off, yy, walk.once);
    unsigned nib = 0u;
    if (MASK == 4) nib = __ldg(bits + (off >> 5)) >> (off & 31);
#pragma unroll
    for (int e = 0; e < VN; ++e) {
      bool keep = true;
      if (MASK == 1) keep = yy[e] > T(0);
      if (MASK == 2) keep = bn_affine(a[e], m[e], iv[e], ga[e], be[e]) > T(0);
      if (MASK == 4) keep = ((nib >> e) & 1u) != 0u;
      gm[e] = keep ? g[e] : T(0);
      o[e] = k0[e] * (gm[e] - k1[e] - ((a[e] - m[e]) * iv[e]) * k2[e]);
    }
    stv<T, VN>(dx + off, o);
    if (DRES) stv<T, VN>(dres + off, gm);
  }
  ZB_ROW_WALK_END
}

// NCHW apply kernels: one block per (n, c) plane.
template <typename T, bool RELU, bool RES>
__global__ void __launch_bounds__(256) bn_apply_nchw(const T* __restrict__ x, const T* __restrict__ res, T* __restrict__ y,
                                                     const T* __restrict__ coef, const T* __restrict__ gamma,
                                                     const T* __restrict__ beta, long long C, long long HW) {
  const long long plane = blockIdx.x;
  const long long c = plane % C;
  const T m = coef[c], iv = coef[C + c], g = gamma[c], b = beta[c];
  const long long base = plane * HW;
  for (long long i = threadIdx.x; i < HW; i += blockDim.x) {
    T v = bn_affine(x[base + i], m, iv, g, b);
    if (RES) v += res[base + i];
    if (RELU) v = v > T(0) ? v : T(0);
    y[base + i] = v;
  }
}
template <typename T, int MASK, bool DRES>
__global__ void __launch_bounds__(256) bn_bwd_apply_nchw(const T* __restrict__ x, const T* __restrict__ dy,
                                                         const T* __restrict__ y, T* __restrict__ dx, T* __restrict__ dres,
                                                         const T* __restrict__ mean, const T* __restrict__ inv,
                                                         const T* __restrict__ coef, const T* __restrict__ gamma,
                                                         const T* __restrict__ beta, long long C, long long HW) {
  const long long plane = blockIdx.x;
  const long long c = plane % C;
  const T m = mean[c], iv = inv[c], k0 = coef[c], k1 = coef[C + c], k2 = coef[2 * C + c];
  const T ga = MASK == 2 ? gamma[c] : T(0), be = MASK == 2 ? beta[c] : T(0);
  const long long base = plane * HW;
  for (long long i = threadIdx.x; i < HW; i += blockDim.x) {
    bool keep = true;
    if (MASK == 1) keep = y[base + i] > T(0);
    if (MASK == 2) keep = bn_affine(x[base + i], m, iv, ga, be) > T(0);
    const T gm = keep ? dy[base + i] : T(0);
    dx[base + i] = k0 * (gm - k1 - ((x[base + i] - m) * iv) * k2);
    if (DRES) dres[base + i] = gm;
  }
}

// ------------------------------------------------------------------------------------------------ host
static inline bool al16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T> constexpr int vec_n() { return 16 / sizeof(T); }

// Runs a column reduce (NHWC or NCHW) into `partial` ([slabs][NS][C]); returns the slab count.
template <typename T, template <typename, int> class FT, typename Init>
static int run_col_reduce(zb_ctx* ctx, int layout, long long N, long long C, long long HW, const T* in0, const T* in1,
                          const T* in2, T* partial, long long partial_cap_elems, Init init, int* slabs_out,
                          int order = ROWS_SLAB_UP) {
  const long long rows = N * HW;
  if (layout == ZB_NHWC) {
    constexpr int VN = vec_n<T>();
    const bool vec = (C % VN == 0) && al16(in0) && al16(in1) && al16(in2);
    if (vec) {
      using F = FT<T, VN>;
      F f; init(f);
      static const int occ = resident_ctas(col_reduce_nhwc<T, VN, F>, sizeof(T) * 256 * F::NS * VN);
      ColGeom g = col_geom(ctx, rows, C, VN, occ);
      ZB_REQUIRE(static_cast<long long>(g.slabs) * F::NS * C <= partial_cap_elems, "bn: partial buffer too small");
      dim3 grid(g.col_groups, g.slabs);
      const size_t smem = sizeof(T) * g.ty * F::NS * g.tx * VN;
      col_reduce_nhwc<T, VN, F><<<grid, 256, smem, ctx->stream>>>(f, in0, in1, in2, partial, rows, C, make_walk(g, rows, order), g.tx, g.ty);
      ZB_LAUNCH_CHECK(ctx);
      *slabs_out = g.slabs;
    } else {
      using F = FT<T, 1>;
      F f; init(f);
      static const int occ = resident_ctas(col_reduce_nhwc<T, 1, F>, sizeof(T) * 256 * F::NS);
      ColGeom g = col_geom(ctx, rows, C, 1, occ);
      ZB_REQUIRE(static_cast<long long>(g.slabs) * F::NS * C <= partial_cap_elems, "bn: partial buffer too small");
      dim3 grid(g.col_groups, g.slabs);
      const size_t smem = sizeof(T) * g.ty * F::NS * g.tx;
      col_reduce_nhwc<T, 1, F><<<grid, 256, smem, ctx->stream>>>(f, in0, in1, in2, partial, rows, C, make_walk(g, rows, order), g.tx, g.ty);
      ZB_LAUNCH_CHECK(ctx);
      *slabs_out = g.slabs;
    }
  } else {
    using F = FT<T, 1>;
    F f; init(f);
    ZB_REQUIRE(C <= 2147483647ll, "bn: too many channels");
    long long slabs = std::max<long long>(1, std::min<long long>(N, (ctx->sm_count * 8ll + C - 1) / C));
    const long long per = (N + slabs - 1) / slabs;
    slabs = (N + per - 1) / per;
    ZB_REQUIRE(slabs * F::NS * C <= partial_cap_elems, "bn: partial buffer too small");
    dim3 grid(static_cast<unsigned>(C), static_cast<unsigned>(slabs));
    col_reduce_nchw<T, F><<<grid, 256, 0, ctx->stream>>>(f, in0, in1, in2, partial, N, C, HW, per);
    ZB_LAUNCH_CHECK(ctx);
    *slabs_out = static_cast<int>(slabs);
  }
  return ZB_OK;
}

template <typename T, int VN> using StatsFT = StatsF<T, VN>;
template <typename T, int VN> using SumFT = SumF<T, VN>;
template <typename T, int VN> using BnBwdMaskFT = BnBwdF<T, VN, 1>;
template <typename T, int VN> using BnBwdNoMaskFT = BnBwdF<T, VN, 0>;
template <typename T, int VN> using BnBwdRecomputeFT = BnBwdF<T, VN, 2>;
template <typename T, int VN> using BnBwdMaskStoreFT = BnBwdF<T, VN, 3>;
template <typename T, int VN> using BnBwdBitsStoreFT = BnBwdF<T, VN, 4>;
template <typename T, int VN> using BnBwdBitsFT = BnBwdF<T, VN, 5>;

// Upper bound on the slab count run_col_reduce may pick (sizes the partial buffer).
static long long max_slabs(zb_ctx* ctx, int layout, long long N, long long C) {
  if (layout == ZB_NHWC) return ctx->sm_count * 8ll;
  return std::max<long long>(1, std::min<long long>(N, (ctx->sm_count * 8ll + C - 1) / C));
}

template <typename T, bool RELU, bool RES>
static int launch_apply(zb_ctx* ctx, int layout, long long N, long long C, long long HW, const T* x, const T* res, T* y,
                        const T* coef, const T* gamma, const T* beta, uint32_t* mask = nullptr, int order = ROWS_SLAB_UP) {
  const long long rows = N * HW;
  if (layout == ZB_NHWC) {
    constexpr int VN = vec_n<T>();
    const bool vec = (C % VN == 0) && al16(x) && al16(res) && al16(y);
    if (vec) {
      static const int occ = resident_ctas(bn_apply_nhwc<T, VN, RELU, RES>, 0);
      ColGeom g = col_geom(ctx, rows, C, VN, occ);
      dim3 grid(g.col_groups, g.slabs);
      ZB_REQUIRE(mask == nullptr || (sizeof(T) == 4 && RELU && C % 32 == 0), "bn: ReLU bit mask needs f32, relu and C %% 32 == 0");
      bn_apply_nhwc<T, VN, RELU, RES><<<grid, 256, 0, ctx->stream>>>(x, res, y, coef, gamma, beta, rows, C, make_walk(g, rows, order), g.tx, g.ty, mask);
    } else {
      ZB_REQUIRE(mask == nullptr, "bn: ReLU bit mask needs 16-byte aligned NHWC tensors");
      static const int occ = resident_ctas(bn_apply_nhwc<T, 1, RELU, RES>, 0);
      ColGeom g = col_geom(ctx, rows, C, 1, occ);
      dim3 grid(g.col_groups, g.slabs);
      bn_apply_nhwc<T, 1, RELU, RES><<<grid, 256, 0, ctx->stream>>>(x, res, y, coef, gamma, beta, rows, C, make_walk(g, rows, order), g.tx, g.ty);
    }
  } else {
    ZB_REQUIRE(mask == nullptr, "bn: ReLU bit mask is NHWC only");
    bn_apply_nchw<T, RELU, RES><<<static_cast<unsigned>(N * C), 256, 0, ctx->stream>>>(x, res, y, coef, gamma, beta, C, HW);
  }
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

template <typename T>
static int dispatch_apply(zb_ctx* ctx, int layout, long long N, long long C, long long HW, const T* x, const T* res, T* y,
                          const T* coef, const T* gamma, const T* beta, int relu, uint32_t* mask = nullptr,
                          int order = ROWS_SLAB_UP) {
  ZB_REQUIRE(mask == nullptr || relu, "bn: a ReLU bit mask without relu");
  if (relu && res) return launch_apply<T, true, true>(ctx, layout, N, C, HW, x, res, y, coef, gamma, beta, mask, order);
  if (relu) return launch_apply<T, true, false>(ctx, layout, N, C, HW, x, res, y, coef, gamma, beta, mask, order);
  if (res) return launch_apply<T, false, true>(ctx, layout, N, C, HW, x, res, y, coef, gamma, beta, nullptr, order);
  return launch_apply<T, false, false>(ctx, layout, N, C, HW, x, res, y, coef, gamma, beta, nullptr, order);
}

template <typename T>
static int bn_fwd_train_t(zb_ctx* ctx, int layout, long long N, long long C, long long H, long long W, double momentum,
                          const T* x, const T* scale, const T* bias, T* run_mean, T* run_var, T* saved_mean,
                          T* saved_inv, T* y, const T* res, int relu, const T* pre_partial = nullptr, int pre_rows = 0,
                          const T* pre_shift = nullptr, uint32_t* relu_mask = nullptr) {
  // pre_partial != NULL: the statistics pass already happened inside the producing conv's epilogue
  // (zb_conv2d_fprop_bnstats): [pre_rows][2][C] partial sums of (x - pre_shift) and its square
  ZB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "bn: empty tensor");
  ZB_REQUIRE(N * H * W < 2147483647ll * 64, "bn: tensor too large");
  const long long HW = H * W;
  const long long ms = max_slabs(ctx, layout, N, C);
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(T) * (ms * 2 * C + 2 * C), &ws);
  if (rc != ZB_OK) return rc;
  T* partial = static_cast<T*>(ws);
  T* coef = partial + ms * 2 * C;
  int slabs = 0;
  prof_begin(ctx, PROF_BN);
  if (pre_partial != nullptr) {
    bn_fwd_finalize<T><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(pre_partial, pre_rows, C, static_cast<double>(N * HW), momentum, ctx->bn_eps, pre_shift,
                                                                 1, run_mean, run_var, saved_mean, saved_inv, coef);
    ZB_LAUNCH_CHECK(ctx);
  } else {
    rc = run_col_reduce<T, StatsFT>(ctx, layout, N, C, HW, x, static_cast<const T*>(nullptr), static_cast<const T*>(nullptr),
                                    partial, ms * 2 * C, [&](auto& f) { f.x0 = x; f.sstride = (layout == ZB_NHWC ? 1 : HW); }, &slabs);
    if (rc != ZB_OK) return rc;
    // shift used by the reduce = first row (NHWC: x[c]) or first element of channel c in image 0 (NCHW: x[c*HW])
    bn_fwd_finalize<T><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(partial, slabs, C, static_cast<double>(N * HW), momentum, ctx->bn_eps, x,
                                                                 layout == ZB_NHWC ? 1 : HW, run_mean, run_var, saved_mean,
                                                                 saved_inv, coef);
    ZB_LAUNCH_CHECK(ctx);
  }
  // statistics from the conv epilogue: x was written by a convolution, last row tiles last; own statistics pass: same slabs
  rc = dispatch_apply<T>(ctx, layout, N, C, HW, x, res, y, coef, scale, bias, relu, relu_mask,
                         bn_row_order_mode() == 0 ? ROWS_SLAB_UP : pre_partial != nullptr ? ROWS_SWEEP_DOWN : ROWS_SLAB_DOWN);
  // algorithmic bytes: x read twice (once when the statistics came with the conv) + y written (+ residual read)
  prof_end(ctx, PROF_BN, static_cast<double>(N * C * HW) * sizeof(T) * ((res ? 4.0 : 3.0) - (pre_partial ? 1.0 : 0.0)));
  return rc;
}

template <typename T>
static int bn_fwd_infer_t(zb_ctx* ctx, int layout, long long N, long long C, long long H, long long W, const T* x,
                          const T* scale, const T* bias, const T* mean, const T* var, T* y) {
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(T) * 2 * C, &ws);
  if (rc != ZB_OK) return rc;
  T* coef = static_cast<T*>(ws);
  bn_infer_coef<T><<<ceil_div(C, 128), 128, 0, ctx->stream>>>(mean, var, C, ctx->bn_eps, coef);
  ZB_LAUNCH_CHECK(ctx);
  return dispatch_apply<T>(ctx, layout, N, C, H * W, x, static_cast<const T*>(nullptr), y, coef, scale, bias, 0);
}

template <typename T, int MASK, bool DRES>
static int launch_bwd_apply(zb_ctx* ctx, int layout, long long N, long long C, long long HW, const T* x, const T* dy,
                            const T* y, T* dx, T* dres, const T* mean, const T* inv, const T* coef, const T* gamma,
                            const T* beta, const uint32_t* bits = nullptr) {
  const long long rows = N * HW;
  const int order = bn_row_order_mode() == 0 ? ROWS_SLAB_UP : bn_row_order_mode() == 1 ? ROWS_SLAB_DOWN : ROWS_SWEEP_UP;
  if (layout == ZB_NHWC) {
    constexpr int VN = vec_n<T>();
    const bool vec = (C % VN == 0) && al16(x) && al16(dy) && al16(y) && al16(dx) && al16(dres);
    ZB_REQUIRE(MASK != 4 || (vec && sizeof(T) == 4 && bits != nullptr), "bn bwd: the bit-mask apply needs aligned f32 NHWC tensors");
    if (vec) {
      static const int occ = resident_ctas(bn_bwd_apply_nhwc<T, VN, MASK, DRES>, 0);
      ColGeom g = col_geom(ctx, rows, C, VN, occ);
      dim3 grid(g.col_groups, g.slabs);
      bn_bwd_apply_nhwc<T, VN, MASK, DRES><<<grid, 256, 0, ctx->stream>>>(x, dy, y, dx, dres, mean, inv, coef, gamma, beta, rows, C, make_walk(g, rows, order), g.tx, g.ty, bits);
    } else {
      static const int occ = resident_ctas(bn_bwd_apply_nhwc<T, 1, MASK, DRES>, 0);
      ColGeom g = col_geom(ctx, rows, C, 1, occ);
      dim3 grid(g.col_groups, g.slabs);
      bn_bwd_apply_nhwc<T, 1, MASK, DRES><<<grid, 256, 0, ctx->stream>>>(x, dy, y, dx, dres, mean, inv, coef, gamma, beta, rows, C, make_walk(g, rows, order), g.tx, g.ty);
    }
  } else {
    bn_bwd_apply_nchw<T, MASK, DRES><<<static_cast<unsigned>(N * C), 256, 0, ctx->stream>>>(x, dy, y, dx, dres, mean, inv, coef, gamma, beta, C, HW);
  }
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

template <typename T>
static int bn_bwd_t(zb_ctx* ctx, int layout, long long N, long long C, long long H, long long W, const T* x, const T* dy,
                    const T* scale, const T* saved_mean, const T* saved_inv, T* dx, T* dscale, T* dbias, const T* y,
                    T* dres, const T* relu_bias = nullptr, const uint32_t* relu_mask = nullptr) {
  // relu_mask != NULL: the ReLU mask comes from the 1-bit array the fused forward wrote (y is not read); needs dres
  // relu_bias != NULL: backward of relu(bn(x)) with the ReLU mask recomputed from x, scale and this bias (y unused)
  ZB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "bn: empty tensor");
  ZB_REQUIRE(!(relu_bias && (y || dres)), "bn bwd: mask recomputation excludes the y / residual-gradient arguments");
  const long long HW = H * W;
  const long long ms = max_slabs(ctx, layout, N, C);
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(T) * (ms * 2 * C + 5 * C), &ws);
  if (rc != ZB_OK) return rc;
  T* partial = static_cast<T*>(ws);
  T* coef = partial + ms * 2 * C;  // 3*C
  T* stats = coef + 3 * C;         // 2*C: recomputed mean / inv when not supplied
  int slabs = 0;
  const T* mean = saved_mean;
  const T* inv = saved_inv;
  if (mean == nullptr || inv == nullptr) {
    // zenu-matrix/src/nn/batch_norm.rs:355-368: recompute the batch statistics from x
    rc = run_col_reduce<T, StatsFT>(ctx, layout, N, C, HW, x, static_cast<const T*>(nullptr), static_cast<const T*>(nullptr),
                                    partial, ms * 2 * C, [&](auto& f) { f.x0 = x; f.sstride = (layout == ZB_NHWC ? 1 : HW); }, &slabs);
    if (rc != ZB_OK) return rc;
    bn_fwd_finalize<T><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(partial, slabs, C, static_cast<double>(N * HW), 0.0, ctx->bn_eps, x,
                                                                 layout == ZB_NHWC ? 1 : HW, static_cast<T*>(nullptr),
                                                                 static_cast<T*>(nullptr), static_cast<T*>(nullptr),
                                                                 static_cast<T*>(nullptr), stats);
    ZB_LAUNCH_CHECK(ctx);
    if (mean == nullptr) mean = stats;
    if (inv == nullptr) inv = stats + C;
  }
  prof_begin(ctx, PROF_BN);
  const int rorder = bn_row_order_mode() >= 2 ? ROWS_SWEEP_DOWN : ROWS_SLAB_UP;   // dy was written by a dgrad, last row tiles last
  if (relu_bias != nullptr)
    rc = run_col_reduce<T, BnBwdRecomputeFT>(ctx, layout, N, C, HW, x, dy, static_cast<const T*>(nullptr), partial, ms * 2 * C,
                                             [&](auto& f) { f.mean = mean; f.inv = inv; f.gamma = scale; f.beta = relu_bias; }, &slabs, rorder);
  else if (relu_mask != nullptr && dres == nullptr)   // the masked gradient is not materialised (see BnBwdF, MASK 5)
    rc = run_col_reduce<T, BnBwdBitsFT>(ctx, layout, N, C, HW, x, dy, static_cast<const T*>(nullptr), partial, ms * 2 * C,
                                        [&](auto& f) { f.mean = mean; f.inv = inv; f.gm_out = nullptr; f.bits = relu_mask; }, &slabs, rorder);
  else if (relu_mask != nullptr)
    rc = run_col_reduce<T, BnBwdBitsStoreFT>(ctx, layout, N, C, HW, x, dy, static_cast<const T*>(nullptr), partial, ms * 2 * C,
                                             [&](auto& f) { f.mean = mean; f.inv = inv; f.gm_out = dres; f.bits = relu_mask; }, &slabs, rorder);
  else if (y != nullptr && dres != nullptr)
    rc = run_col_reduce<T, BnBwdMaskStoreFT>(ctx, layout, N, C, HW, x, dy, y, partial, ms * 2 * C,
                                             [&](auto& f) { f.mean = mean; f.inv = inv; f.gm_out = dres; }, &slabs, rorder);
  else if (y != nullptr)
    rc = run_col_reduce<T, BnBwdMaskFT>(ctx, layout, N, C, HW, x, dy, y, partial, ms * 2 * C,
                                        [&](auto& f) { f.mean = mean; f.inv = inv; }, &slabs, rorder);
  else
    rc = run_col_reduce<T, BnBwdNoMaskFT>(ctx, layout, N, C, HW, x, dy, static_cast<const T*>(nullptr), partial, ms * 2 * C,
                                          [&](auto& f) { f.mean = mean; f.inv = inv; }, &slabs, rorder);
  if (rc != ZB_OK) return rc;
  bn_bwd_finalize<T><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(partial, slabs, C, static_cast<double>(N * HW), scale, inv,
                                                               dscale, dbias, coef);
  ZB_LAUNCH_CHECK(ctx);
  if (relu_bias != nullptr) rc = launch_bwd_apply<T, 2, false>(ctx, layout, N, C, HW, x, dy, y, dx, dres, mean, inv, coef, scale, relu_bias);
  else if (relu_mask != nullptr && dres == nullptr)
    rc = launch_bwd_apply<T, 4, false>(ctx, layout, N, C, HW, x, dy, static_cast<const T*>(nullptr), dx, static_cast<T*>(nullptr), mean, inv, coef, scale, relu_bias, relu_mask);
  else if ((y != nullptr || relu_mask != nullptr) && dres != nullptr)   // dres already holds the masked gradient (written by the reduce pass)
    rc = launch_bwd_apply<T, 0, false>(ctx, layout, N, C, HW, x, dres, static_cast<const T*>(nullptr), dx, static_cast<T*>(nullptr), mean, inv, coef, scale, relu_bias);
  else if (y != nullptr) rc = launch_bwd_apply<T, 1, false>(ctx, layout, N, C, HW, x, dy, y, dx, dres, mean, inv, coef, scale, relu_bias);
  else if (dres != nullptr) rc = launch_bwd_apply<T, 0, true>(ctx, layout, N, C, HW, x, dy, y, dx, dres, mean, inv, coef, scale, relu_bias);
  else rc = launch_bwd_apply<T, 0, false>(ctx, layout, N, C, HW, x, dy, y, dx, dres, mean, inv, coef, scale, relu_bias);
  // algorithmic bytes: x, dy read twice + dx written (+ y read twice for the ReLU mask, + dres written); the fused
  // BN+add+ReLU backward reads x twice, dy and y once, writes and re-reads the masked gradient, writes dx: 7 passes
  prof_end(ctx, PROF_BN, static_cast<double>(N * C * HW) * sizeof(T) *
                             (relu_mask ? (dres ? 6.0 + 1.0 / 32.0 : 5.0 + 2.0 / 32.0) : (y && dres) ? 7.0 : 5.0 + (y ? 2.0 : 0.0) + (dres ? 1.0 : 0.0)));
  return rc;
}


// ---- network stem: BatchNorm + ReLU + 3x3 / stride 2 / pad 1 max-pool as ONE pass each way (NHWC f32) ----------------------------
// The stem's BatchNorm output is the largest activation of a ResNet (256 x 112 x 112 x 64 f32 = 822 MB at batch 256) and the only
// thing that reads it is the max-pool.  Forward: every pooled output normalises its 9 taps from x on the fly (the values, the
// first-max tie rule and the "a padding zero won" code 255 are exactly those of zb_bn2d_fwd_train(relu) followed by
// zb_maxpool2d_fwd_idx), so the BN output is never written or re-read.  Backward: the gradient of the BN output is the max-pool's
// gather over the <= 4 windows that cover a pixel; both BN-backward passes (statistics, then dx) rebuild it per 2 x 2 input patch
// from the pooled gradient and the winner codes (4 window reads per 4 pixels, L2-resident: they are 1/4 + 1/16 of the tensor)
// instead of reading an 822 MB dy that a separate max-pool backward would first have to write.  ReLU mask: recomputed from x
// (bn_affine(x) > 0), as in zb_bn2d_relu_bwd.
// Forward thread = (pooled output, 4 channels), nine independent 128-bit loads in flight.  (A sliding-window form that carries the
// column two neighbouring windows share - 6 taps normalised per output instead of 9 - measured SLOWER: 0.63 vs 0.37 ms on the
// 256 x 112 x 112 x 64 stem, 99 registers and a serial walk per thread against 48 registers and 5 resident blocks per SM here; so
// did giving a block a 4 x 4 patch of outputs instead of a 1 x 16 strip: 0.44 ms, the L1 reuse it adds costs more in coalescing.)
__global__ void __launch_bounds__(256) bn_relu_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ coef,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               float* __restrict__ y, uchar4* __restrict__ idx, int N, int H, int W,
                                                               int C, int P, int Q) {
  const int c4n = C >> 2;
  const int total = N * P * Q * c4n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c4 = i % c4n;
    int t = i / c4n;
    const int q = t % Q; t /= Q;
    const int p = t % P;
    const int n = t / P;
    const float4 m = __ldg(reinterpret_cast<const float4*>(coef) + c4), iv = __ldg(reinterpret_cast<const float4*>(coef + C) + c4);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    const float* xb = x + static_cast<long long>(n) * H * W * C + c4 * 4;
    float best[4];
    unsigned char bi[4];
    bool first = true;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = p * 2 + r - 1;
#pragma unroll
      for (int s_ = 0; s_ < 3; ++s_) {
        const int iw = q * 2 + s_ - 1;
        const bool oob = ih < 0 || ih >= H || iw < 0 || iw >= W;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (!oob) {
          const float4 vv = *reinterpret_cast<const float4*>(xb + (static_cast<long long>(ih) * W + iw) * C);
          v[0] = bn_affine(vv.x, m.x, iv.x, g.x, b.x); v[1] = bn_affine(vv.y, m.y, iv.y, g.y, b.y);
          v[2] = bn_affine(vv.z, m.z, iv.z, g.z, b.z); v[3] = bn_affine(vv.w, m.w, iv.w, g.w, b.w);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : 0.f;
        }
        const unsigned char tap = oob ? 255 : static_cast<unsigned char>(r * 3 + s_);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (first || v[e] > best[e]) { best[e] = v[e]; bi[e] = tap; }
        first = false;
      }
    }
    *reinterpret_cast<float4*>(y + static_cast<long long>(i) * 4) = make_float4(best[0], best[1], best[2], best[3]);
    idx[i] = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
  }
}

// APPLY = false: per-block partial sums [block][2][C] of g and g * xhat (g = ReLU-masked gathered gradient); APPLY = true: dx.
// Thread = (2 x 2 input patch (2a.., 2b..), 4 channels); the channel group of a thread is fixed (256 % (C / 4) == 0), so its sums stay
// in registers.  3x3 / 2 / 1 windows: the patch is covered by windows (a, b), (a, b+1), (a+1, b), (a+1, b+1) and its pixels sit at
// fixed taps of each: (0,0) <- code 4 of w00; (0,1) <- 5 of w00, 3 of w01; (1,0) <- 7 of w00, 1 of w10; (1,1) <- 8 of w00, 6 of w01,
// 2 of w10, 0 of w11 (added in zb_maxpool2d_bwd_idx's order).  All twelve loads of a patch are issued before any is used.
template <bool APPLY>
__global__ void __launch_bounds__(256, 3) bn_relu_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dyp,
                                                                  const uchar4* __restrict__ idx, const float* __restrict__ mean,
                                                                  const float* __restrict__ inv, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ coef,
                                                                  float* __restrict__ dx, float* __restrict__ partial, int N, int H,
                                                                  int W, int C, int P, int Q) {
  __shared__ float red[APPLY ? 1 : 2][APPLY ? 1 : 256][4];
  const int c4n = C >> 2, G = 256 / c4n;
  const int c4 = threadIdx.x % c4n, grp = threadIdx.x / c4n;
  const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
  const int patches = N * H2 * W2;
  const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean) + c4), iv4 = __ldg(reinterpret_cast<const float4*>(inv) + c4);
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + c4), b4 = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  const float m[4] = {m4.x, m4.y, m4.z, m4.w}, iv[4] = {iv4.x, iv4.y, iv4.z, iv4.w};
  const float ga[4] = {g4.x, g4.y, g4.z, g4.w}, be[4] = {b4.x, b4.y, b4.z, b4.w};
  float k0[4] = {0.f, 0.f, 0.f, 0.f}, k1[4] = {0.f, 0.f, 0.f, 0.f}, k2[4] = {0.f, 0.f, 0.f, 0.f};
  if (APPLY) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(coef) + c4), b = __ldg(reinterpret_cast<const float4*>(coef + C) + c4);
    const float4 c = __ldg(reinterpret_cast<const float4*>(coef + 2 * C) + c4);
    k0[0] = a.x; k0[1] = a.y; k0[2] = a.z; k0[3] = a.w;
    k1[0] = b.x; k1[1] = b.y; k1[2] = b.z; k1[3] = b.w;
    k2[0] = c.x; k2[1] = c.y; k2[2] = c.z; k2[3] = c.w;
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int pi = blockIdx.x * G + grp; pi < patches; pi += gridDim.x * G) {
    const int bq = pi % W2;
    int t = pi / W2;
    const int a = t % H2;
    const int n = t / H2;
    const int h0 = 2 * a, w0 = 2 * bq;
    const bool row1 = h0 + 1 < H, col1 = w0 + 1 < W;          // the patch's second row / column exists
    const bool win_p0 = a < P, win_q0 = bq < Q;               // (odd H: the last patch row can lie below the last window)
    const bool win_p1 = a + 1 < P, win_q1 = bq + 1 < Q;
    const float* xp = x + ((static_cast<long long>(n) * H + h0) * W + w0) * C + c4 * 4;
    const long long rs = static_cast<long long>(W) * C;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 x00 = *reinterpret_cast<const float4*>(xp);
    const float4 x01 = col1 ? *reinterpret_cast<const float4*>(xp + C) : zero4;
    const float4 x10 = row1 ? *reinterpret_cast<const float4*>(xp + rs) : zero4;
    const float4 x11 = (row1 && col1) ? *reinterpret_cast<const float4*>(xp + rs + C) : zero4;
    const int pa = win_p0 ? a : P - 1, pb = win_p1 ? a + 1 : P - 1, qa = win_q0 ? bq : Q - 1, qb = win_q1 ? bq + 1 : Q - 1;
    const int o00 = ((n * P + pa) * Q + qa) * c4n + c4, o01 = ((n * P + pa) * Q + qb) * c4n + c4;
    const int o10 = ((n * P + pb) * Q + qa) * c4n + c4, o11 = ((n * P + pb) * Q + qb) * c4n + c4;
    const uchar4 i00 = __ldg(idx + o00), i01 = __ldg(idx + o01), i10 = __ldg(idx + o10), i11 = __ldg(idx + o11);
    const float4 d00 = __ldg(reinterpret_cast<const float4*>(dyp) + o00), d01 = __ldg(reinterpret_cast<const float4*>(dyp) + o01);
    const float4 d10 = __ldg(reinterpret_cast<const float4*>(dyp) + o10), d11 = __ldg(reinterpret_cast<const float4*>(dyp) + o11);
    const bool v00 = win_p0 && win_q0, v01 = win_p0 && win_q1, v10 = win_p1 && win_q0, v11 = win_p1 && win_q1;
    const unsigned char w00[4] = {i00.x, i00.y, i00.z, i00.w}, w01[4] = {i01.x, i01.y, i01.z, i01.w};
    const unsigned char w10[4] = {i10.x, i10.y, i10.z, i10.w}, w11[4] = {i11.x, i11.y, i11.z, i11.w};
    const float g00[4] = {d00.x, d00.y, d00.z, d00.w}, g01[4] = {d01.x, d01.y, d01.z, d01.w};
    const float g10[4] = {d10.x, d10.y, d10.z, d10.w}, g11[4] = {d11.x, d11.y, d11.z, d11.w};
    float acc[2][2][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[0][0][e] = (v00 && w00[e] == 4) ? g00[e] : 0.f;
      float t01 = (v00 && w00[e] == 5) ? g00[e] : 0.f;
      if (v01 && w01[e] == 3) t01 += g01[e];
      acc[0][1][e] = t01;
      float t10 = (v00 && w00[e] == 7) ? g00[e] : 0.f;
      if (v10 && w10[e] == 1) t10 += g10[e];
      acc[1][0][e] = t10;
      float t11 = (v00 && w00[e] == 8) ? g00[e] : 0.f;
      if (v01 && w01[e] == 6) t11 += g01[e];
      if (v10 && w10[e] == 2) t11 += g10[e];
      if (v11 && w11[e] == 0) t11 += g11[e];
      acc[1][1][e] = t11;
    }
    const float4 xv[2][2] = {{x00, x01}, {x10, x11}};
#pragma unroll
    for (int yy = 0; yy < 2; ++yy) {
#pragma unroll
      for (int xx = 0; xx < 2; ++xx) {
        const bool exists = (yy == 0 || row1) && (xx == 0 || col1);
        const float xs[4] = {xv[yy][xx].x, xv[yy][xx].y, xv[yy][xx].z, xv[yy][xx].w};
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool keep = exists && bn_affine(xs[e], m[e], iv[e], ga[e], be[e]) > 0.f;
          const float gg = keep ? acc[yy][xx][e] : 0.f;
          const float xh = (xs[e] - m[e]) * iv[e];
          if (APPLY) o[e] = k0[e] * (gg - k1[e] - xh * k2[e]);
          else { s1[e] += gg; s2[e] += gg * xh; }
        }
        if (APPLY && exists) *reinterpret_cast<float4*>(dx + (xp - x) + yy * rs + xx * C) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { red[0][threadIdx.x][e] = s1[e]; red[1][threadIdx.x][e] = s2[e]; }
    __syncthreads();
    for (int half = G >> 1; half > 0; half >>= 1) {
      if (grp < half) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          red[0][threadIdx.x][e] += red[0][threadIdx.x + half * c4n][e];
          red[1][threadIdx.x][e] += red[1][threadIdx.x + half * c4n][e];
        }
      }
      __syncthreads();
    }
    if (grp == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        partial[(static_cast<long long>(blockIdx.x) * 2 + 0) * C + c4 * 4 + e] = red[0][threadIdx.x][e];
        partial[(static_cast<long long>(blockIdx.x) * 2 + 1) * C + c4 * 4 + e] = red[1][threadIdx.x][e];
      }
    }
  }
}

static bool bn_pool_supported(long long N, long long C, long long H, long long W) {
  const long long c4n = C / 4;
  return C % 4 == 0 && c4n >= 1 && c4n <= 256 && 256 % c4n == 0 && H >= 2 && W >= 2 && N * (H + 1) * (W + 1) * C < (1ll << 31) - (1ll << 24);
}

static int bn_relu_pool_fwd_f32(zb_ctx* ctx, long long N, long long C, long long H, long long W, double momentum, const float* x,
                                const float* scale, const float* bias, float* run_mean, float* run_var, float* saved_mean,
                                float* saved_inv, float* y_pool, void* idx, const float* pre_partial, int pre_rows, const float* pre_shift) {
  const long long P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1, HW = H * W;
  const long long ms = max_slabs(ctx, ZB_NHWC, N, C);
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(float) * (ms * 2 * C + 2 * C), &ws);
  if (rc != ZB_OK) return rc;
  float* partial = static_cast<float*>(ws);
  float* coef = partial + ms * 2 * C;
  prof_begin(ctx, PROF_BN);
  if (pre_partial != nullptr) {
    bn_fwd_finalize<float><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(pre_partial, pre_rows, C, static_cast<double>(N * HW), momentum, ctx->bn_eps,
                                                                   pre_shift, 1, run_mean, run_var, saved_mean, saved_inv, coef);
    ZB_LAUNCH_CHECK(ctx);
  } else {
    int slabs = 0;
    rc = run_col_reduce<float, StatsFT>(ctx, ZB_NHWC, N, C, HW, x, static_cast<const float*>(nullptr), static_cast<const float*>(nullptr),
                                        partial, ms * 2 * C, [&](auto& f) { f.x0 = x; f.sstride = 1; }, &slabs);
    if (rc != ZB_OK) return rc;
    bn_fwd_finalize<float><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(partial, slabs, C, static_cast<double>(N * HW), momentum, ctx->bn_eps, x, 1,
                                                                   run_mean, run_var, saved_mean, saved_inv, coef);
    ZB_LAUNCH_CHECK(ctx);
  }
  const long long total = N * P * Q * (C / 4);
  // whole waves of the 5 CTAs per SM that are resident (48 registers x 256 threads)
  static const bool whole_waves = []() { const char* e = getenv("ZENU_B200_POOL_GRID"); return e == nullptr || atoi(e) != 0; }();
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * (whole_waves ? 30ll : 32ll))));
  bn_relu_pool_fwd_kernel<<<grid, 256, 0, ctx->stream>>>(x, coef, scale, bias, y_pool, static_cast<uchar4*>(idx), static_cast<int>(N),
                                                         static_cast<int>(H), static_cast<int>(W), static_cast<int>(C), static_cast<int>(P),
                                                         static_cast<int>(Q));
  ZB_LAUNCH_CHECK(ctx);
  // algorithmic bytes: x read (twice when the statistics did not come with the conv), pooled output + winner codes written
  prof_end(ctx, PROF_BN, static_cast<double>(N * C * HW) * 4.0 * (pre_partial ? 1.0 : 2.0) + static_cast<double>(N * C * P * Q) * 5.0);
  return ZB_OK;
}

static int bn_relu_pool_bwd_f32(zb_ctx* ctx, long long N, long long C, long long H, long long W, const float* x, const float* dyp,
                                const void* idx, const float* scale, const float* bias, const float* mean, const float* inv, float* dx,
                                float* dscale, float* dbias) {
  const long long P = (H + 2 - 3) / 2 + 1, Q = (W + 2 - 3) / 2 + 1;
  const long long ms = ctx->sm_count * 8ll;
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(float) * (ms * 2 * C + 3 * C), &ws);
  if (rc != ZB_OK) return rc;
  float* partial = static_cast<float*>(ws);
  float* coef = partial + ms * 2 * C;
  const int G = static_cast<int>(256 / (C / 4));
  const long long patches = N * ((H + 1) / 2) * ((W + 1) / 2);
  // one full wave: the kernels hold 80 registers x 256 threads, i.e. 3 CTAs per SM are resident, and every CTA strides over the patches
  // (a grid of sm_count * 8 was 2.67 waves, the last one two thirds empty).  ZENU_B200_POOL_GRID=0: the old grid (A/B knob)
  static const bool one_wave = []() { const char* e = getenv("ZENU_B200_POOL_GRID"); return e == nullptr || atoi(e) != 0; }();
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((patches + G - 1) / G, one_wave ? ctx->sm_count * 3ll : ms)));
  prof_begin(ctx, PROF_BN);
  bn_relu_pool_bwd_kernel<false><<<grid, 256, 0, ctx->stream>>>(x, dyp, static_cast<const uchar4*>(idx), mean, inv, scale, bias, nullptr, nullptr,
                                                                partial, static_cast<int>(N), static_cast<int>(H), static_cast<int>(W),
                                                                static_cast<int>(C), static_cast<int>(P), static_cast<int>(Q));
  ZB_LAUNCH_CHECK(ctx);
  bn_bwd_finalize<float><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(partial, grid, C, static_cast<double>(N * H * W), scale, inv, dscale, dbias, coef);
  ZB_LAUNCH_CHECK(ctx);
  bn_relu_pool_bwd_kernel<true><<<grid, 256, 0, ctx->stream>>>(x, dyp, static_cast<const uchar4*>(idx), mean, inv, scale, bias, coef, dx, nullptr,
                                                               static_cast<int>(N), static_cast<int>(H), static_cast<int>(W),
                                                               static_cast<int>(C), static_cast<int>(P), static_cast<int>(Q));
  ZB_LAUNCH_CHECK(ctx);
  // algorithmic bytes: x read twice, dx written, pooled gradient + winner codes read twice
  prof_end(ctx, PROF_BN, static_cast<double>(N * C * H * W) * 4.0 * 3.0 + static_cast<double>(N * C * P * Q) * 2.0 * 5.0);
  return ZB_OK;
}

// out[c] = sum over (n, hw) of a — conv bias gradient / Matrix::sum(axis 0) on a [rows][cols] matrix (NHWC, HW = 1)
template <typename T>
int channel_sum(zb_ctx* ctx, int layout, long long N, long long C, long long HW, const T* a, T* out) {
  if (N * C * HW == 0) return ZB_OK;
  const long long ms = max_slabs(ctx, layout, N, C);
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(T) * ms * C, &ws);
  if (rc != ZB_OK) return rc;
  T* partial = static_cast<T*>(ws);
  int slabs = 0;
  rc = run_col_reduce<T, SumFT>(ctx, layout, N, C, HW, a, static_cast<const T*>(nullptr), static_cast<const T*>(nullptr),
                                partial, ms * C, [&](auto&) {}, &slabs);
  if (rc != ZB_OK) return rc;
  sum_finalize<T><<<ZB_FIN_GRID(C), 0, ctx->stream>>>(partial, slabs, C, out);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}
template int channel_sum<float>(zb_ctx*, int, long long, long long, long long, const float*, float*);
template int channel_sum<double>(zb_ctx*, int, long long, long long, long long, const double*, double*);

}  // namespace zb

using namespace zb;

extern "C" {

int zb_bn2d_fwd_train(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, double momentum,
                      const void* x, const void* scale, const void* bias, void* running_mean, void* running_var,
                      void* saved_mean, void* saved_inv_std, void* y, const void* residual, int relu) {
  ZB_API_RANGE();
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC, "bn: unknown layout %d", layout);
  if (dtype == ZB_F32)
    return bn_fwd_train_t<float>(ctx, layout, n, c, h, w, momentum, static_cast<const float*>(x), static_cast<const float*>(scale),
                                 static_cast<const float*>(bias), static_cast<float*>(running_mean), static_cast<float*>(running_var),
                                 static_cast<float*>(saved_mean), static_cast<float*>(saved_inv_std), static_cast<float*>(y),
                                 static_cast<const float*>(residual), relu);
  if (dtype == ZB_F64)
    return bn_fwd_train_t<double>(ctx, layout, n, c, h, w, momentum, static_cast<const double*>(x), static_cast<const double*>(scale),
                                  static_cast<const double*>(bias), static_cast<double*>(running_mean), static_cast<double*>(running_var),
                                  static_cast<double*>(saved_mean), static_cast<double*>(saved_inv_std), static_cast<double*>(y),
                                  static_cast<const double*>(residual), relu);
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

int zb_bn2d_fwd_train_prestats(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, double momentum,
                               const void* x, const void* scale, const void* bias, void* running_mean, void* running_var,
                               void* saved_mean, void* saved_inv_std, void* y, const void* residual, int relu,
                               const void* stat_partial, int64_t stat_rows, const void* shift) {
  ZB_API_RANGE();
  ZB_REQUIRE(stat_partial != nullptr && shift != nullptr && stat_rows > 0, "bn prestats: missing statistics");
  return zb_bn2d_fwd_train_fused(ctx, dtype, layout, n, c, h, w, momentum, x, scale, bias, running_mean, running_var, saved_mean,
                                 saved_inv_std, y, residual, relu, stat_partial, stat_rows, shift, nullptr);
}

int64_t zb_bn2d_relu_mask_words(int64_t n, int64_t c, int64_t h, int64_t w) { return (n * c * h * w + 31) / 32; }

int zb_bn2d_fwd_train_fused(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, double momentum,
                            const void* x, const void* scale, const void* bias, void* running_mean, void* running_var,
                            void* saved_mean, void* saved_inv_std, void* y, const void* residual, int relu,
                            const void* stat_partial, int64_t stat_rows, const void* shift, void* relu_mask) {
  ZB_API_RANGE();
  ZB_REQUIRE(layout == ZB_NHWC && dtype == ZB_F32, "bn fused forward: NHWC f32 only");
  ZB_REQUIRE(stat_partial == nullptr || (shift != nullptr && stat_rows > 0 && stat_rows < (1 << 30)), "bn fused forward: bad statistics");
  ZB_REQUIRE(relu_mask == nullptr || (relu && c % 32 == 0), "bn fused forward: the ReLU bit mask needs relu and C %% 32 == 0");
  return bn_fwd_train_t<float>(ctx, layout, n, c, h, w, momentum, static_cast<const float*>(x), static_cast<const float*>(scale),
                               static_cast<const float*>(bias), static_cast<float*>(running_mean), static_cast<float*>(running_var),
                               static_cast<float*>(saved_mean), static_cast<float*>(saved_inv_std), static_cast<float*>(y),
                               static_cast<const float*>(residual), relu, static_cast<const float*>(stat_partial),
                               static_cast<int>(stat_rows), static_cast<const float*>(shift), static_cast<uint32_t*>(relu_mask));
}

int zb_bn2d_bwd_mask(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x, const void* dy,
                     const void* scale, const void* saved_mean, const void* saved_inv_std, void* dx, void* dscale, void* dbias,
                     const void* relu_mask, void* dres) {
  ZB_API_RANGE();
  ZB_REQUIRE(layout == ZB_NHWC && dtype == ZB_F32 && c % 32 == 0, "bn bwd (bit mask): NHWC f32 with C %% 32 == 0 only");
  ZB_REQUIRE(relu_mask != nullptr && saved_mean != nullptr && saved_inv_std != nullptr,
             "bn bwd (bit mask): the mask and the saved statistics are required");
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dres) & 15) == 0,
             "bn bwd (bit mask): tensors must be 16-byte aligned");
  return bn_bwd_t<float>(ctx, layout, n, c, h, w, static_cast<const float*>(x), static_cast<const float*>(dy),
                         static_cast<const float*>(scale), static_cast<const float*>(saved_mean), static_cast<const float*>(saved_inv_std),
                         static_cast<float*>(dx), static_cast<float*>(dscale), static_cast<float*>(dbias), nullptr, static_cast<float*>(dres),
                         nullptr, static_cast<const uint32_t*>(relu_mask));
}

int zb_bn2d_fwd_infer(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                      const void* scale, const void* bias, const void* mean, const void* var, void* y) {
  ZB_API_RANGE();
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC, "bn: unknown layout %d", layout);
  if (dtype == ZB_F32)
    return bn_fwd_infer_t<float>(ctx, layout, n, c, h, w, static_cast<const float*>(x), static_cast<const float*>(scale),
                                 static_cast<const float*>(bias), static_cast<const float*>(mean), static_cast<const float*>(var),
                                 static_cast<float*>(y));
  if (dtype == ZB_F64)
    return bn_fwd_infer_t<double>(ctx, layout, n, c, h, w, static_cast<const double*>(x), static_cast<const double*>(scale),
                                  static_cast<const double*>(bias), static_cast<const double*>(mean), static_cast<const double*>(var),
                                  static_cast<double*>(y));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

int zb_bn2d_bwd(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                const void* dy, const void* scale, const void* saved_mean, const void* saved_inv_std, void* dx,
                void* dscale, void* dbias, const void* y, void* dres) {
  ZB_API_RANGE();
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC, "bn: unknown layout %d", layout);
  if (dtype == ZB_F32)
    return bn_bwd_t<float>(ctx, layout, n, c, h, w, static_cast<const float*>(x), static_cast<const float*>(dy),
                           static_cast<const float*>(scale), static_cast<const float*>(saved_mean),
                           static_cast<const float*>(saved_inv_std), static_cast<float*>(dx), static_cast<float*>(dscale),
                           static_cast<float*>(dbias), static_cast<const float*>(y), static_cast<float*>(dres));
  if (dtype == ZB_F64)
    return bn_bwd_t<double>(ctx, layout, n, c, h, w, static_cast<const double*>(x), static_cast<const double*>(dy),
                            static_cast<const double*>(scale), static_cast<const double*>(saved_mean),
                            static_cast<const double*>(saved_inv_std), static_cast<double*>(dx), static_cast<double*>(dscale),
                            static_cast<double*>(dbias), static_cast<const double*>(y), static_cast<double*>(dres));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

int zb_bn2d_relu_bwd(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                     const void* dy, const void* scale, const void* bias, const void* saved_mean, const void* saved_inv_std,
                     void* dx, void* dscale, void* dbias) {
  ZB_API_RANGE();
  ZB_REQUIRE(layout == ZB_NCHW || layout == ZB_NHWC, "bn: unknown layout %d", layout);
  ZB_REQUIRE(bias != nullptr && saved_mean != nullptr && saved_inv_std != nullptr, "bn relu bwd: bias and saved statistics are required");
  if (dtype == ZB_F32)
    return bn_bwd_t<float>(ctx, layout, n, c, h, w, static_cast<const float*>(x), static_cast<const float*>(dy),
                           static_cast<const float*>(scale), static_cast<const float*>(saved_mean),
                           static_cast<const float*>(saved_inv_std), static_cast<float*>(dx), static_cast<float*>(dscale),
                           static_cast<float*>(dbias), nullptr, nullptr, static_cast<const float*>(bias));
  if (dtype == ZB_F64)
    return bn_bwd_t<double>(ctx, layout, n, c, h, w, static_cast<const double*>(x), static_cast<const double*>(dy),
                            static_cast<const double*>(scale), static_cast<const double*>(saved_mean),
                            static_cast<const double*>(saved_inv_std), static_cast<double*>(dx), static_cast<double*>(dscale),
                            static_cast<double*>(dbias), nullptr, nullptr, static_cast<const double*>(bias));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

static int check_bn_pool(int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k, int64_t stride, int64_t pad) {
  ZB_REQUIRE(n > 0 && c > 0 && h > 0 && w > 0, "bn + max-pool: empty tensor");
  if (dtype != ZB_F32 || layout != ZB_NHWC || k != 3 || stride != 2 || pad != 1 || !bn_pool_supported(n, c, h, w)) {
    zb::set_last_error("bn + relu + max-pool: served for f32 NHWC, 3x3 / stride 2 / pad 1 windows and C / 4 a power of two <= 256");
    return ZB_ERR_UNSUPPORTED;
  }
  return ZB_OK;
}

int zb_bn2d_relu_maxpool_fwd_train(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k,
                                   int64_t stride, int64_t pad, double momentum, const void* x, const void* scale, const void* bias,
                                   void* running_mean, void* running_var, void* saved_mean, void* saved_inv_std, void* y_pool,
                                   void* pool_idx, const void* stat_partial, int64_t stat_rows, const void* shift) {
  ZB_API_RANGE();
  const int rc = check_bn_pool(dtype, layout, n, c, h, w, k, stride, pad);
  if (rc != ZB_OK) return rc;
  ZB_REQUIRE(x && scale && bias && y_pool && pool_idx, "bn + relu + max-pool: null argument");
  ZB_REQUIRE(al16(x) && al16(y_pool) && al16(scale) && al16(bias), "bn + relu + max-pool: tensors must be 16-byte aligned");
  ZB_REQUIRE(stat_partial == nullptr || (shift != nullptr && stat_rows > 0 && stat_rows < (1 << 30)), "bn + relu + max-pool: bad statistics");
  return bn_relu_pool_fwd_f32(ctx, n, c, h, w, momentum, static_cast<const float*>(x), static_cast<const float*>(scale),
                              static_cast<const float*>(bias), static_cast<float*>(running_mean), static_cast<float*>(running_var),
                              static_cast<float*>(saved_mean), static_cast<float*>(saved_inv_std), static_cast<float*>(y_pool), pool_idx,
                              static_cast<const float*>(stat_partial), static_cast<int>(stat_rows), static_cast<const float*>(shift));
}

int zb_bn2d_relu_maxpool_bwd(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k, int64_t stride,
                             int64_t pad, const void* x, const void* dy_pool, const void* pool_idx, const void* scale, const void* bias,
                             const void* saved_mean, const void* saved_inv_std, void* dx, void* dscale, void* dbias) {
  ZB_API_RANGE();
  const int rc = check_bn_pool(dtype, layout, n, c, h, w, k, stride, pad);
  if (rc != ZB_OK) return rc;
  ZB_REQUIRE(x && dy_pool && pool_idx && scale && bias && saved_mean && saved_inv_std && dx && dscale && dbias,
             "bn + relu + max-pool backward: null argument");
  ZB_REQUIRE(al16(x) && al16(dy_pool) && al16(dx) && al16(scale) && al16(bias) && al16(saved_mean) && al16(saved_inv_std),
             "bn + relu + max-pool backward: tensors must be 16-byte aligned");
  return bn_relu_pool_bwd_f32(ctx, n, c, h, w, static_cast<const float*>(x), static_cast<const float*>(dy_pool), pool_idx,
                              static_cast<const float*>(scale), static_cast<const float*>(bias), static_cast<const float*>(saved_mean),
                              static_cast<const float*>(saved_inv_std), static_cast<float*>(dx), static_cast<float*>(dscale),
                              static_cast<float*>(dbias));
}

int zb_conv2d_bias_bwd(zb_ctx* ctx, int dtype, int layout, const void* dy, void* dbias, int64_t n, int64_t k, int64_t h,
                       int64_t w) {
  ZB_API_RANGE();
  if (dtype == ZB_F32) return channel_sum<float>(ctx, layout, n, k, h * w, static_cast<const float*>(dy), static_cast<float*>(dbias));
  if (dtype == ZB_F64) return channel_sum<double>(ctx, layout, n, k, h * w, static_cast<const double*>(dy), static_cast<double*>(dbias));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

int zb_sum_rows(zb_ctx* ctx, int dtype, const void* a, void* out, int64_t rows, int64_t cols) {
  ZB_API_RANGE();
  if (dtype == ZB_F32) return channel_sum<float>(ctx, ZB_NHWC, rows, cols, 1, static_cast<const float*>(a), static_cast<float*>(out));
  if (dtype == ZB_F64) return channel_sum<double>(ctx, ZB_NHWC, rows, cols, 1, static_cast<const double*>(a), static_cast<double*>(out));
  zb::set_last_error("unknown dtype %d", dtype);
  return ZB_ERR_INVALID;
}

}  // extern "C"
