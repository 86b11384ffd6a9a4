// umma_wgrad.cu — halo-reuse backward-filter kernel for stride-1 R x S convolutions (tcgen05 kind::tf32).
//
//   dW[k][r][s][c] = sum over (n, p, q) of dY[n][p][q][k] * X[n][p + r - pad_h][q + s - pad_w][c]
//
// The implicit-GEMM wgrad (umma_gemm.cu, one tile group per filter tap) re-reads dY and X once per tap: 18 tensor passes
// through L2 for a 3x3 filter, which is what bounds the 64- and 128-channel layers (67 TFLOP/s on 64 -> 64 @ 56x56).
// Here the reduction dimension (pixels) is walked in raster order: one TMA box lands (tp + R - 1) x Wr input pixels of a
// 32-channel chunk (Wr = W + 2*pad_w, halo columns / rows arrive as TMA zero fill) and one box per 32 output channels lands
// the matching tp x Wr raster of dY (columns q >= Q zero-filled).  Both are MN-major operands (row = pixel = GEMM-K, 128 bytes
// = 32 channels = GEMM-M/N, SWIZZLE_128B_BASE32B).  Output pixel j pairs with input pixel j + r*Wr + s, so
//   * filter row r    = the same X raster read through a descriptor that starts r*Wr rows (128 B each) later, and
//   * filter column s = the next 32-column group of N with a leading-dimension stride of ONE row (LBO = 128 B): the S
//     N-groups of one MMA are the same smem rows shifted by 0..S-1 pixels.
// (The swizzle is a function of the absolute smem address, so row-granular descriptor starts and overlapping N groups are
// legal: measured in round 1 with a descriptor-shift probe build, git history: tools/probe_mn_shift.py.)  One MMA = [128 k] x [S*32 (s, c)] x [8 pixels]; R accumulators of S*32
// columns live in TMEM for the whole kernel: each CTA owns (32-channel chunk, 128-k tile, pixel range) and drains TMEM once.
// dY and X are each fetched once per CTA role instead of once per tap.  Partials [split][K][R*S*C] -> deterministic reduce.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace zb {

using namespace ptx;

int umma_chain_limit();

struct WgHaloParams {
  int Wr, tp, p_tiles, R, S;
  int kt_pad;           // pixels (GEMM-K rows) consumed per tile, multiple of 8
  int a_box_bytes;      // smem bytes reserved per 32-k box of dY (1024-aligned)
  int a_tx_bytes;       // bytes one dY box delivers (tp * Wr * 128)
  int x_slot_bytes;     // smem bytes reserved for the X raster
  int x_tx_bytes;       // (tp + R - 1) * Wr * 128
  int a_boxes;          // dY boxes reserved per stage (min(4, ceil(K / 32)))
  int stage_bytes, stages, ring_bytes;
  int c_chunks, k_tiles, splits;
  int total_tiles, tiles_per_split;   // pixel tiles = N * p_tiles
  // small feature maps (7x7): `stack` images per tile, their rasters ([pad rows, H rows] x [pad columns, W columns], bottom / right padding
  // = the next row's / image's top / left padding) back to back, img_bytes each, one TMA box per image and operand (see UmmaParams::halo_stack)
  int stack, img_bytes;
  int K, C;
  int lower_w, lower_h;
  float* partial;       // [splits][K][R*S*C]
  long long split_stride, ld;
  int* err_flag;
};

__global__ void __launch_bounds__(192, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ WgHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.ring_bytes);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* done_bar = empty_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile int* err = p.err_flag;

  // role of this CTA: adjacent CTAs share the pixel range and the k tile (same dY boxes in flight -> L2 hits)
  int id = blockIdx.x;
  const int cc = id % p.c_chunks; id /= p.c_chunks;
  const int kt = id % p.k_tiles;
  const int split = id / p.k_tiles;
  const int c0 = cc * 32, m0 = kt * 128;
  const int k_boxes = min(4, (p.K - m0 + 31) / 32);
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.total_tiles, t_begin + p.tiles_per_split);
  const int acc_cols = p.S * 32;
  const uint32_t tmem_cols = (p.R * acc_cols <= 32) ? 32u : (p.R * acc_cols <= 64) ? 64u : (p.R * acc_cols <= 128) ? 128u
                             : (p.R * acc_cols <= 256) ? 256u : 512u;

  // Rows that TMA never writes (K padding of the dY boxes, the tail of the X slot, boxes of a ragged k tile) are read by the
  // MMAs and must hold finite values (they meet zeros on the other side): clear the ring once.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.ring_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 0) {
    if (elect_one()) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < 4; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = static_cast<uint32_t>(k_boxes) * p.a_tx_bytes + p.x_tx_bytes;
      for (int tile = t_begin; tile < t_end; ++tile) {
        if (!mbar_wait(&empty_bar[stage], phase ^ 1, err)) break;
        mbar_arrive_expect_tx(&full_bar[stage], bytes);
        uint8_t* sA = smem + stage * p.stage_bytes;
        uint8_t* sX = sA + p.a_boxes * p.a_box_bytes;
        if (p.stack > 1) {   // an image past the end of the batch arrives as zero fill on both sides
          for (int g = 0; g < p.stack; ++g) {
            const int img = tile * p.stack + g;
            for (int j = 0; j < k_boxes; ++j)
              tma_load_4d(sA + j * p.a_box_bytes + g * p.img_bytes, &tmA, &full_bar[stage], m0 + 32 * j, 0, 0, img);
            tma_load_4d(sX + g * p.img_bytes, &tmB, &full_bar[stage], c0, p.lower_w, p.lower_h, img);
          }
        } else {
          const int img = tile / p.p_tiles, pt = tile - img * p.p_tiles;
          const int p0 = pt * p.tp;
          for (int j = 0; j < k_boxes; ++j) tma_load_4d(sA + j * p.a_box_bytes, &tmA, &full_bar[stage], m0 + 32 * j, 0, p0, img);
          tma_load_4d(sX, &tmB, &full_bar[stage], c0, p.lower_w, p0 + p.lower_h, img);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(kUmmaBM, acc_cols, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      const int ksteps = p.kt_pad >> 3;
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t a_desc0 = make_smem_desc(0, p.a_box_bytes, 512, kSmemLayoutSw128Base32);
      const uint64_t b_desc0 = make_smem_desc(0, 128, 512, kSmemLayoutSw128Base32);
      const uint32_t row_step = static_cast<uint32_t>(p.Wr * 128) >> 4;   // one filter row = Wr raster rows of 128 bytes
      bool ok = true;
      for (int tile = t_begin; tile < t_end; ++tile) {
        if (!mbar_wait(&full_bar[stage], phase, err)) { ok = false; break; }
        tc_fence_after();
        // descriptors advanced by integer adds on their address field (the issuing thread is on the critical path)
        const uint32_t a_off = (smem0 + stage * p.stage_bytes) >> 4;
        const uint64_t da0 = a_desc0 + a_off;
        const uint64_t db0 = b_desc0 + a_off + static_cast<uint32_t>((p.a_boxes * p.a_box_bytes) >> 4);
        // One thread issues every MMA and its instruction stream is what bounds these N = S*32 <= 96 wide MMAs (~45 cycles per
        // tcgen05.mma with ready-made descriptors, ~100 through a runtime-count inner loop: tools/probes/mma_rate.cu), so the
        // 3-row filter gets a fully unrolled body: three MMAs per K step on precomputed descriptor offsets.
        if (p.R == 3) {
          const uint32_t t1 = tmem_base + acc_cols, t2 = tmem_base + 2 * acc_cols;
          const uint64_t db1 = db0 + row_step, db2 = db0 + 2 * row_step;
          const uint32_t first = tile > t_begin ? 1u : 0u;
          umma_tf32(tmem_base, da0, db0, idesc, first);
          umma_tf32(t1, da0, db1, idesc, first);
          umma_tf32(t2, da0, db2, idesc, first);
#pragma unroll 4
          for (int i = 1; i < ksteps; ++i) {
            const uint32_t o = static_cast<uint32_t>(i * 64);
            umma_tf32(tmem_base, da0 + o, db0 + o, idesc, 1u);
            umma_tf32(t1, da0 + o, db1 + o, idesc, 1u);
            umma_tf32(t2, da0 + o, db2 + o, idesc, 1u);
          }
        } else {
          for (int i = 0; i < ksteps; ++i) {
            const uint64_t da = da0 + static_cast<uint32_t>(i * 64);
            uint64_t db = db0 + static_cast<uint32_t>(i * 64);
            const uint32_t accum = (tile > t_begin || i > 0) ? 1u : 0u;
            for (int r = 0; r < p.R; ++r, db += row_step) umma_tf32(tmem_base + r * acc_cols, da, db, idesc, accum);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (ok) umma_commit(done_bar);
    }
  } else {
    // epilogue: lane = k row of the tile, 32 consecutive channels of tap (r, s) per tcgen05.ld
    const int ew = warp & 3;
    const int k = m0 + ew * 32 + lane;
    if (mbar_wait(done_bar, 0, err)) {
      tc_fence_after();
      float* dst_row = p.partial + static_cast<long long>(split) * p.split_stride + static_cast<long long>(k) * p.ld + c0;
      for (int r = 0; r < p.R; ++r)
        for (int s = 0; s < p.S; ++s) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + r * acc_cols + s * 32, v);
          tmem_ld_wait();
          if (k < p.K) {
            float4* dst = reinterpret_cast<float4*>(dst_row + static_cast<long long>(r * p.S + s) * p.C);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                                   __uint_as_float(v[q * 4 + 3]));
          }
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// out[i] = sum_s partial[s][i] + beta * out[i], 4 elements per thread
__global__ void wgrad_reduce_kernel(const float4* __restrict__ partial, float4* __restrict__ out, long long n4, long long split_stride4,
                                    int splits, float beta) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
#pragma unroll 2
    for (; s + 2 <= splits; s += 2) {   // independent accumulators: several 128-bit loads in flight per thread
      const float4 v = partial[s * split_stride4 + i], u = partial[(s + 1) * split_stride4 + i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      b.x += u.x; b.y += u.y; b.z += u.z; b.w += u.w;
    }
    if (s < splits) {
      const float4 v = partial[s * split_stride4 + i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    if (beta != 0.f) {
      const float4 o = out[i];
      a.x += beta * o.x; a.y += beta * o.y; a.z += beta * o.z; a.w += beta * o.w;
    }
    out[i] = a;
  }
}

static CUtensorMapDataType wg_dtype() {
  static int raw = -1;
  if (raw < 0) {
    const char* e = getenv("ZENU_B200_TMA_F32");
    raw = (e && e[0] == '1') ? 1 : 0;
  }
  return raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
}

// [N][H][W][C] NHWC tensor, box = 32 channels x box_w x box_h pixels of one image, MN-major swizzle
static int make_raster_map(zb_ctx* ctx, CUtensorMap* map, const float* base, long long N, long long H, long long W, long long C,
                           int box_w, int box_h) {
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4, static_cast<cuuint64_t>(H) * W * C * 4};
  cuuint32_t box[4] = {32, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(map, wg_dtype(), 4, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (wgrad raster) failed (%d): NHWC=%lldx%lldx%lldx%lld box=%dx%d", int(r), N, H, W, C, box_w, box_h);
    return ZB_ERR_CUDA;
  }
  return ZB_OK;
}

// Returns ZB_ERR_UNSUPPORTED (nothing launched) when the geometry is not served; the caller then uses the per-tap kernel.
int umma_conv_wgrad_halo(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* x, float* dw, float beta) {
  const int R = static_cast<int>(d->kh), S = static_cast<int>(d->kw);
  const int ph = static_cast<int>(d->pad_h), pw = static_cast<int>(d->pad_w);
  if (ZB_ENV_FLAG("ZENU_B200_NO_WGRAD_HALO") || umma_chain_limit() > 0) return ZB_ERR_UNSUPPORTED;
  if (d->stride_h != 1 || d->stride_w != 1 || d->dil_h != 1 || d->dil_w != 1 || R * S < 2 || R > 8 || R * S * 32 > 512 ||
      d->c % 32 != 0 || d->k % 4 != 0 || (reinterpret_cast<uintptr_t>(dw) & 15) != 0)
    return ZB_ERR_UNSUPPORTED;
  const long long P = d->h + 2 * ph - R + 1, Q = d->w + 2 * pw - S + 1;
  const long long Wr = d->w + 2 * pw;
  if (P <= 0 || Q <= 0 || Wr > 256 || d->n * P > 0x3fffffffll) return ZB_ERR_UNSUPPORTED;
  WgHaloParams p;
  memset(&p, 0, sizeof(p));
  p.R = R; p.S = S; p.Wr = static_cast<int>(Wr);
  p.K = static_cast<int>(d->k); p.C = static_cast<int>(d->c);
  p.a_boxes = static_cast<int>(std::min<long long>(4, (d->k + 31) / 32));
  p.stack = 1;
  const int budget = 227 * 1024 - 1024 - 256;   // alignment slack, barriers
  bool found = false;
  // tiny images (7x7): one image is <= 64 raster pixels, 42 KB of loads per 24 MMAs (L2 bound: the per-tap kernel wins), so several
  // images share a tile: their rasters are stacked with the padding rows / columns shared between neighbours
  const bool tiny = P * Wr <= 64 && !ZB_ENV_FLAG("ZENU_B200_WGRAD_HALO_ALL");
  if (tiny) {
    const int Ws = static_cast<int>(d->w) + pw, Hs = static_cast<int>(d->h) + ph;
    if (ZB_ENV_FLAG("ZENU_B200_NO_HALO_STACK") || S - 1 - pw > pw || R - 1 - ph > ph || (Hs * Ws) % 8 != 0) return ZB_ERR_UNSUPPORTED;
    for (int pass = 0; pass < 2 && !found; ++pass)   // first choice: three stages in flight, else two
    for (int G = static_cast<int>(std::min<long long>(8, d->n)); G >= 2 && !found; --G) {
      const int kt = G * Hs * Ws;   // raster pixels (GEMM-K rows) per tile, a multiple of 8
      const int a_box = (kt * 128 + 1023) & ~1023;
      const int x_rows = (R - 1) * Ws + (S - 1) + kt;
      const int x_slot = (x_rows * 128 + 1023) & ~1023;
      const int stage = p.a_boxes * a_box + x_slot;
      const int tail = std::max(0, 4 * a_box - stage);
      if ((pass == 0 ? 3 : 2) * stage + tail > budget) continue;
      p.stack = G; p.img_bytes = Hs * Ws * 128;
      p.Wr = Ws; p.tp = Hs; p.p_tiles = 1; p.kt_pad = kt; p.a_box_bytes = a_box; p.a_tx_bytes = kt * 128;
      p.x_slot_bytes = x_slot; p.x_tx_bytes = kt * 128;
      p.stage_bytes = stage;
      p.stages = std::min(4, (budget - tail) / stage);
      p.ring_bytes = p.stages * stage + tail;
      found = true;
    }
    if (!found) return ZB_ERR_UNSUPPORTED;
  }
  for (int tp0 = static_cast<int>(std::min<long long>(P, 256 - R + 1)); tp0 >= 1 && !found; --tp0) {
    const int p_tiles = ceil_div(P, tp0);
    const int tp = ceil_div(P, p_tiles);
    const int kt = tp * p.Wr, kt_pad = (kt + 7) & ~7;
    const int a_box = (kt_pad * 128 + 1023) & ~1023;
    const int x_rows = std::max((tp + R - 1) * p.Wr, (R - 1) * p.Wr + (S - 1) + kt_pad);
    const int x_slot = (x_rows * 128 + 1023) & ~1023;
    const int stage = p.a_boxes * a_box + x_slot;
    const int tail = std::max(0, 4 * a_box - stage);   // an M = 128 descriptor reads 4 boxes even when the k tile has fewer
    if (2 * stage + tail > budget) continue;
    p.tp = tp; p.p_tiles = p_tiles; p.kt_pad = kt_pad; p.a_box_bytes = a_box; p.a_tx_bytes = kt * 128;
    p.x_slot_bytes = x_slot; p.x_tx_bytes = (tp + R - 1) * p.Wr * 128;
    p.stage_bytes = stage;
    p.stages = std::min(4, (budget - tail) / stage);
    p.ring_bytes = p.stages * stage + tail;
    found = true;
  }
  if (!found) return ZB_ERR_UNSUPPORTED;
  p.c_chunks = static_cast<int>(d->c / 32);
  p.k_tiles = ceil_div(d->k, kUmmaBM);
  p.total_tiles = p.stack > 1 ? ceil_div(d->n, p.stack) : static_cast<int>(d->n) * p.p_tiles;
  const int roles = p.c_chunks * p.k_tiles;
  int splits = std::max(1, std::min(p.total_tiles, (ctx->sm_count + roles - 1) / roles));
  if (roles * splits > ctx->sm_count && splits > 1 && roles * (splits - 1) >= (ctx->sm_count * 3) / 4) --splits;   // one wave, >= 75 % full
  p.tiles_per_split = ceil_div(p.total_tiles, splits);
  p.splits = ceil_div(p.total_tiles, p.tiles_per_split);
  p.lower_w = -pw; p.lower_h = -ph;
  p.ld = static_cast<long long>(R) * S * d->c;
  p.split_stride = d->k * p.ld;
  p.err_flag = ctx->err_flag;
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, sizeof(float) * static_cast<size_t>(p.splits) * p.split_stride, &ws);
  if (rc != ZB_OK) return rc;
  p.partial = static_cast<float*>(ws);
  CUtensorMap ma, mb;
  // (stacked: both boxes are one image's Wr x tp raster; dY rows >= P / columns >= Q and X's padding arrive as zero fill)
  if ((rc = make_raster_map(ctx, &ma, dy, d->n, P, Q, d->k, p.Wr, p.tp)) != ZB_OK) return rc;
  if ((rc = make_raster_map(ctx, &mb, x, d->n, d->h, d->w, d->c, p.Wr, p.stack > 1 ? p.tp : p.tp + R - 1)) != ZB_OK) return rc;
  const size_t smem = static_cast<size_t>(p.ring_bytes) + 256 + 1024;
  static SmemOptIn opt_in;
  { const int rc2 = smem_opt_in(ctx, opt_in, wgrad_halo_kernel, smem); if (rc2 != ZB_OK) return rc2; }
  plan_note("wgrad_halo R=%d S=%d a_boxes=%d stages=%d roles=%d tp=%d stack=%d beta=%d ~splits=%d ~grid=%d;wgrad_reduce;", R, S, p.a_boxes, p.stages, roles,
            p.tp, p.stack, beta != 0.f ? 1 : 0, p.splits, roles * p.splits);
  if (plan_dry()) return ZB_OK;
  prof_begin(ctx, PROF_TENSOR);
  wgrad_halo_kernel<<<roles * p.splits, 192, smem, ctx->stream>>>(ma, mb, p);
  prof_end(ctx, PROF_TENSOR, 2.0 * d->n * P * Q * d->k * d->c * R * S);
  ZB_LAUNCH_CHECK(ctx);
  const long long n4 = p.split_stride / 4;
  const int grid = static_cast<int>(std::min<long long>((n4 + 255) / 256, ctx->sm_count * 8ll));
  wgrad_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(p.partial), reinterpret_cast<float4*>(dw), n4,
                                                     n4, p.splits, beta);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

}  // namespace zb
