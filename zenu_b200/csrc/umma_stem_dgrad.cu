// umma_stem_dgrad.cu — backward-data of small-C convolutions (C <= 4: the 7x7/s2 ResNet stem, the first CIFAR layer) on tcgen05.
//
//   dX[n][h][w][c] = sum over (r, s, k) of dY[n][p][q][k] * W[k][r][s][c],   p*sh = h + ph - r*dh,  q*sw = w + pw - s
//
// GEMM view per input row h: D_h[q][s*4 + c] = sum over the filter rows r that reach h and over k of
// dY[n, p(h, r), q, k] * Wd[s*4 + c][r*K + k]  (M = Q <= 128 output columns, N = 32 = 8 filter columns x 4 padded channels);
// the epilogue overlap-adds the filter columns (w = q*sw - pw + s) into the dense row dX[n][h].
// The first version handled one input row per tile and re-fetched every dY row for each of the R/sh input rows it reaches
// (and the 4 KB filter tile with it): 8.3 GB of L2->SM traffic for the ResNet stem, 2.3 ms.  Here a tile is TH = 8 consecutive
// input rows with one TMEM accumulator each: a dY row lands in smem once per tile and feeds every accumulator it reaches, and
// the whole transformed filter (R x K/32 tiles of 4 KB) stays resident in smem for the life of the persistent CTA.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace zb {

using namespace ptx;

constexpr int kSdTH = 8;        // input rows (accumulators) per tile; 2 buffers x 8 x 32 columns = all 512 TMEM columns
constexpr int kSdMaxStages = 8;

struct StemDgradParams {
  int N, H, W, C, P, Q, K, R, S;
  int sh, sw, ph, pw, dh;
  int h_blocks, total_tiles, chunks, stages;
  int b_tiles;          // R * chunks resident filter tiles
  float* dx;
  float beta;
  int* err_flag;
};

__device__ __forceinline__ int floor_div_i(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

__global__ void __launch_bounds__(192, 1)
stem_dgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ StemDgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                 // b_tiles x 4 KB, tile (r, chunk) at (r * chunks + chunk) * 4096
  uint8_t* sA = sB + p.b_tiles * 4096;                // stages x 16 KB dY rows [128 q][32 k]
  uint8_t* sE = sA + p.stages * 16384;                // 2 x 16 KB epilogue staging [128 q][32]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sE + 2 * 16384);
  uint64_t* empty_bar = full_bar + kSdMaxStages;
  uint64_t* b_bar = empty_bar + kSdMaxStages;
  uint64_t* tfull_bar = b_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile int* err = p.err_flag;
  if (warp == 0) {
    if (elect_one()) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < kSdMaxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(b_bar, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(b_bar, static_cast<uint32_t>(p.b_tiles) * 4096u);
      for (int r = 0; r < p.R; ++r)
        for (int c = 0; c < p.chunks; ++c) tma_load_2d(sB + (r * p.chunks + c) * 4096, &tmB, b_bar, r * p.K + c * 32, 0);
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
        const int img = tile / p.h_blocks, hb = tile - img * p.h_blocks;
        const int h0 = hb * kSdTH, h1 = min(h0 + kSdTH, p.H) - 1;
        const int p_lo = -floor_div_i(-(h0 + p.ph - (p.R - 1) * p.dh), p.sh), p_hi = floor_div_i(h1 + p.ph, p.sh);
        for (int pr = p_lo; pr <= p_hi && ok; ++pr)
          for (int c = 0; c < p.chunks; ++c) {
            if (!mbar_wait(&empty_bar[stage], phase ^ 1, err)) { ok = false; break; }
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.Q) * 128u);
            tma_load_4d(sA + stage * 16384, &tmA, &full_bar[stage], c * 32, 0, pr, img);   // rows p < 0 or >= P: zero fill
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(kUmmaBM, 32, 0, 0);
      const uint32_t a0 = smem_u32(sA);
      const uint64_t db_first = make_smem_desc(smem_u32(sB), 16, 1024, kSmemLayoutSw128);
      const uint64_t da_base = make_smem_desc(a0, 16, 1024, kSmemLayoutSw128);   // + stage * 1024 (16 KB stages, address field in 16-byte units)
      const uint32_t db_row_step = static_cast<uint32_t>(p.chunks * 256);         // next filter row of the transformed filter
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      bool ok = mbar_wait(b_bar, 0, err);
      for (int tile = blockIdx.x; tile < p.total_tiles && ok; tile += gridDim.x) {
        const int img = tile / p.h_blocks, hb = tile - img * p.h_blocks;
        const int h0 = hb * kSdTH, h1 = min(h0 + kSdTH, p.H) - 1;
        const int p_lo = -floor_div_i(-(h0 + p.ph - (p.R - 1) * p.dh), p.sh), p_hi = floor_div_i(h1 + p.ph, p.sh);
        if (!mbar_wait(&tempty_bar[acc], acc_phase ^ 1, err)) break;
        tc_fence_after();
        uint32_t started = 0;
        for (int pr = p_lo; pr <= p_hi && ok; ++pr)
          for (int c = 0; c < p.chunks; ++c) {
            if (!mbar_wait(&full_bar[stage], phase, err)) { ok = false; break; }
            tc_fence_after();
            // (the issuing thread is the critical path here: N = 32 MMAs retire faster than scalar code can describe them, so
            // descriptors are advanced by adding to their 14-bit address field and the dil_h == 1 case avoids the divisions)
            const uint64_t da0 = da_base + static_cast<uint32_t>(stage * (16384 >> 4));
            const int t0 = h0 + p.ph - pr * p.sh;   // = r * dil_h for the filter row that links input row h0 to dY row pr
            if (p.dh == 1) {
              // rows of the tile this dY row reaches: 0 <= t0 + hl < R; one MMA group each, descriptors / TMEM address by adds
              const int hl_lo = max(0, -t0), hl_hi = min(h1 - h0, p.R - 1 - t0);
              uint64_t db0 = db_first + static_cast<uint32_t>(((t0 + hl_lo) * p.chunks + c) * 256);   // 4096-byte tiles, >> 4
              uint32_t d_tmem = tmem_base + static_cast<uint32_t>((acc * kSdTH + hl_lo) * 32);
              for (int hl = hl_lo; hl <= hl_hi; ++hl, db0 += db_row_step, d_tmem += 32) {
                umma_tf32(d_tmem, da0, db0, idesc, (started >> hl) & 1u);
                umma_tf32(d_tmem, da0 + 2, db0 + 2, idesc, 1u);
                umma_tf32(d_tmem, da0 + 4, db0 + 4, idesc, 1u);
                umma_tf32(d_tmem, da0 + 6, db0 + 6, idesc, 1u);
              }
              if (hl_hi >= hl_lo) started |= ((2u << hl_hi) - 1u) & ~((1u << hl_lo) - 1u);
            } else {
              int t = t0;
              for (int hl = 0; hl <= h1 - h0; ++hl, ++t) {
                if (t < 0 || t % p.dh != 0) continue;
                const int r = t / p.dh;
                if (r >= p.R) continue;
                const uint64_t db0 = db_first + static_cast<uint64_t>((r * p.chunks + c) * 256);
                const uint32_t d_tmem = tmem_base + (acc * kSdTH + hl) * 32;
                umma_tf32(d_tmem, da0, db0, idesc, (started >> hl) & 1u);
                umma_tf32(d_tmem, da0 + 2, db0 + 2, idesc, 1u);
                umma_tf32(d_tmem, da0 + 4, db0 + 4, idesc, 1u);
                umma_tf32(d_tmem, da0 + 6, db0 + 6, idesc, 1u);
                started |= 1u << hl;
              }
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        if (!ok) break;
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // epilogue: per input row, accumulator D[q][s*4+c] -> shared [128 q][32] (XOR-swizzled 16-byte groups) -> one thread per
    // output pixel w overlap-adds the <= ceil(S / sw) filter columns that reach it (q steps down as s steps up by sw)
    const int ew = warp & 3;
    int acc = 0, rowbuf = 0;
    uint32_t acc_phase = 0;
    // the filter columns reaching pixel w do not depend on the row: first column / its q for this thread's (up to) two pixels
    int w_sx0[2], w_q0[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int w = ew * 32 + lane + i * 128;
      w_sx0[i] = (w + p.pw) % p.sw;
      w_q0[i] = (w + p.pw - w_sx0[i]) / p.sw;
    }
    const bool fast_taps = p.sw == 2 && p.S <= 8;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int img = tile / p.h_blocks, hb = tile - img * p.h_blocks;
      const int h0 = hb * kSdTH, h1 = min(h0 + kSdTH, p.H) - 1;
      if (!mbar_wait(&tfull_bar[acc], acc_phase, err)) break;
      tc_fence_after();
      for (int hl = 0; hl <= h1 - h0; ++hl) {
        float4* stage4 = reinterpret_cast<float4*>(sE + rowbuf * 16384);
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + (acc * kSdTH + hl) * 32, v);
        tmem_ld_wait();
        const int qrow = ew * 32 + lane;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          stage4[qrow * 8 + (g ^ (qrow & 7))] = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                                                            __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
        if (hl == h1 - h0) {   // last accumulator of the tile has been read: the MMA warp may reuse this TMEM buffer
          tc_fence_before();
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (hl == h1 - h0 && lane == 0) mbar_arrive(&tempty_bar[acc]);
        float* out_row = p.dx + (static_cast<long long>(img) * p.H + h0 + hl) * (static_cast<long long>(p.W) * p.C);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int w = ew * 32 + lane + i * 128;
          if (w >= p.W) break;
          float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
          int q = w_q0[i];
          if (fast_taps) {
            // S <= 8, stride 2 (the ResNet stem: 7 / 2): at most four filter columns reach a pixel.  Their four 128-bit smem reads are
            // issued together (the runtime-bound loop below serialises one read + add per trip: ncu had the epilogue warps waiting on
            // exactly that chain) and summed in the loop's order; a column that does not exist contributes +0.
            float4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int sx = w_sx0[i] + 2 * j, qj = q - j;
              const bool ok = sx < p.S && qj >= 0 && qj < p.Q;
              t[j] = ok ? stage4[qj * 8 + (sx ^ (qj & 7))] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { sum.x += t[j].x; sum.y += t[j].y; sum.z += t[j].z; sum.w += t[j].w; }
          } else {
            for (int sx = w_sx0[i]; sx < p.S; sx += p.sw, --q) {
              if (q >= 0 && q < p.Q) {
                const float4 t = stage4[q * 8 + (sx ^ (q & 7))];
                sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
              }
            }
          }
          float* o = out_row + w * p.C;
          const float vals4[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < p.C) o[c] = p.beta != 0.f ? vals4[c] + p.beta * o[c] : vals4[c];
        }
        rowbuf ^= 1;   // the next row goes to the other staging buffer: one barrier per row is enough
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// wd[j = s*4+c][r*K + k] = w[k][r][s][c] (KRSC), zero elsewhere
__global__ void stem_dgrad_filter_kernel(const float* __restrict__ w, float* __restrict__ wd, int K, int R, int S, int C) {
  const int total = 32 * R * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % K;
    const int r = (i / K) % R;
    const int j = i / (K * R);
    const int sidx = j >> 2, c = j & 3;
    wd[i] = (sidx < S && c < C) ? w[((static_cast<long long>(k) * R + r) * S + sidx) * C + c] : 0.f;
  }
}

static CUtensorMapDataType sd_dtype() {
  const char* e = getenv("ZENU_B200_TMA_F32");
  return (e && e[0] == '1') ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
}

// Returns ZB_ERR_UNSUPPORTED (nothing launched) when the geometry is not served by this kernel.
int umma_conv_stem_dgrad(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* w, float* dx, float beta) {
  if (ZB_ENV_FLAG("ZENU_B200_NO_STEM_DGRAD")) return ZB_ERR_UNSUPPORTED;
  if (!(d->c <= 4 && d->kw <= 8 && d->dil_w == 1 && d->k % 32 == 0 && d->kh <= 16 && d->stride_h <= 8 && d->stride_w <= 8))
    return ZB_ERR_UNSUPPORTED;
  const long long P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  const long long Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  if (Q > kUmmaBM || d->w > 256 || P <= 0 || Q <= 0 || (reinterpret_cast<uintptr_t>(dy) & 15) != 0) return ZB_ERR_UNSUPPORTED;
  // every input row must be reached by at least one filter row (otherwise its accumulator would never be written)
  for (int a = 0; a < d->stride_h; ++a) {
    int cnt = 0;
    for (int r = 0; r < d->kh; ++r) cnt += ((a - r * d->dil_h) % d->stride_h + d->stride_h) % d->stride_h == 0;
    if (cnt == 0) return ZB_ERR_UNSUPPORTED;
  }
  StemDgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = static_cast<int>(d->n); p.H = static_cast<int>(d->h); p.W = static_cast<int>(d->w); p.C = static_cast<int>(d->c);
  p.P = static_cast<int>(P); p.Q = static_cast<int>(Q); p.K = static_cast<int>(d->k); p.R = static_cast<int>(d->kh); p.S = static_cast<int>(d->kw);
  p.sh = static_cast<int>(d->stride_h); p.sw = static_cast<int>(d->stride_w); p.ph = static_cast<int>(d->pad_h); p.pw = static_cast<int>(d->pad_w);
  p.dh = static_cast<int>(d->dil_h);
  p.chunks = p.K / 32;
  p.b_tiles = p.R * p.chunks;
  const int budget = 227 * 1024 - 1024 - 512 - 2 * 16384;
  if (p.b_tiles * 4096 + 3 * 16384 > budget) return ZB_ERR_UNSUPPORTED;
  p.stages = std::min(kSdMaxStages, (budget - p.b_tiles * 4096) / 16384);
  p.h_blocks = ceil_div(d->h, kSdTH);
  if (d->n * p.h_blocks > 0x3fffffffll) return ZB_ERR_UNSUPPORTED;
  p.total_tiles = static_cast<int>(d->n) * p.h_blocks;
  p.dx = dx; p.beta = beta; p.err_flag = ctx->err_flag;
  const size_t wd_bytes = sizeof(float) * 32 * p.R * p.K;
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, wd_bytes, &ws);
  if (rc != ZB_OK) return rc;
  float* wd = static_cast<float*>(ws);
  {
    const int total = 32 * p.R * p.K;
    plan_note("stem_dgrad_filter;");
    ZB_KLAUNCH(ctx, stem_dgrad_filter_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(w, wd, p.K, p.R, p.S, p.C));
  }
  CUtensorMap ma, mb;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->k), static_cast<cuuint64_t>(Q), static_cast<cuuint64_t>(P), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->k) * 4, static_cast<cuuint64_t>(Q) * d->k * 4, static_cast<cuuint64_t>(P) * Q * d->k * 4};
    cuuint32_t box[4] = {32, static_cast<cuuint32_t>(Q), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ctx->encode_tiled(&ma, sd_dtype(), 4, const_cast<float*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (stem dgrad dY rows) failed (%d)", int(r)); return ZB_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(p.R) * p.K, 32};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(p.R) * p.K * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&mb, sd_dtype(), 2, wd, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (stem dgrad filter) failed (%d)", int(r)); return ZB_ERR_CUDA; }
  }
  const size_t smem = static_cast<size_t>(p.b_tiles) * 4096 + static_cast<size_t>(p.stages) * 16384 + 2 * 16384 + 512 + 1024;
  static SmemOptIn opt_in;
  { const int rc2 = smem_opt_in(ctx, opt_in, stem_dgrad_kernel, smem); if (rc2 != ZB_OK) return rc2; }
  const int grid = std::min(p.total_tiles, ctx->sm_count);
  plan_note("stem_dgrad stages=%d b_tiles=%d beta=%d ~tiles=%d ~grid=%d;", p.stages, p.b_tiles, beta != 0.f ? 1 : 0, p.total_tiles, grid);
  if (plan_dry()) return ZB_OK;
  prof_begin(ctx, PROF_TENSOR);
  stem_dgrad_kernel<<<grid, 192, smem, ctx->stream>>>(ma, mb, p);
  prof_end(ctx, PROF_TENSOR, 2.0 * d->n * P * Q * d->k * d->c * d->kh * d->kw);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

}  // namespace zb
