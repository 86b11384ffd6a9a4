// umma_stem_wgrad.cu — backward-filter of small-C convolutions (C <= 4: the 7x7/s2 ResNet stem, the first CIFAR layer) on tcgen05.
//
//   dW[k][r][s][c] = sum over (n, p, q) of dY[n][p][q][k] * X[n][p*sh - ph + r*1][q*sw - pw + s][c]
//
// GEMM view (umma_gemm.cu "small-C conv"): M = k, N = R x 32 = for every filter row r the 32-float sliding window "8 filter
// columns x 4 padded channels", K = output pixels walked in 32-pixel runs of one output row.  Both operands are MN-major boxes of
// [32 pixels][128 bytes]: A = dY (one box per 32 output channels), B = one window box per filter row, taken from input row
// ih = p*sh - ph + r.  The one-run-per-K-block form re-fetches all R window boxes for every output row although consecutive output
// rows share R - sh of them (7x7/s2: 5 of 7): 36 KB of loads per 32 pixels, 4.1 GB of L2->SM traffic for the ResNet stem, 0.78 ms.
// Here a CTA walks a column strip (image, 32-pixel run) down the output rows: window boxes live in a ring indexed by input row,
// each step loads only the sh new ones (+ the dY boxes), and the R boxes of a step are R consecutive ring slots, so one MMA with
// N = R*32 and LBO = one box covers them (two MMAs when the ring wraps).  The accumulator [k][R*32] stays in TMEM for the whole
// kernel; every CTA writes one partial, reduced by the same deterministic unpack kernel as before.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace zb {

using namespace ptx;

constexpr int kSwStages = 8;     // steps in flight (dY boxes + the step's new window boxes share one barrier pair)

struct StemWgradParams {
  int N, H, P, Q, K, R;
  int sh, ph;
  int q_runs, strips;            // 32-pixel runs per output row; strips = N * q_runs
  int k_boxes;                   // dY boxes per step (ceil(K / 32) <= 4)
  int ring;                      // window-box ring slots (4 KB each), >= R + sh * (kSwStages + 1)
  int a_stage_bytes;             // k_boxes * 4096
  float* partial;                // [gridDim.x][K][R*32]
  int* err_flag;
};

__global__ void __launch_bounds__(192, 1)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ StemWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                   // ring x 4 KB window boxes, slot = (p*sh + r) % ring
  uint8_t* sA = sB + p.ring * 4096;                     // kSwStages x k_boxes x 4 KB dY boxes
  // an M = 128 descriptor reads 4 dY boxes even when the k tile has fewer: the boxes past the last stage fall into this pad
  const int ring_bytes = p.ring * 4096 + kSwStages * p.a_stage_bytes + (4 - p.k_boxes) * 4096;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + ring_bytes);
  uint64_t* empty_bar = full_bar + kSwStages;
  uint64_t* done_bar = empty_bar + kSwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile int* err = p.err_flag;
  // boxes of a ragged k tile are never written by TMA but are read by the M = 128 MMAs: keep them finite
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < ring_bytes / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 0) {
    if (elect_one()) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < kSwStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = static_cast<int>(blockIdx.x) < p.strips;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int strip = blockIdx.x; strip < p.strips && ok; strip += gridDim.x) {
        const int img = strip / p.q_runs, q0 = (strip - img * p.q_runs) * 32;
        for (int pr = 0; pr < p.P && ok; ++pr) {
          // a new strip overwrites ring slots the previous strip's last steps may still be reading: drain first
          if (pr == 0 && strip != static_cast<int>(blockIdx.x)) {   // = the commit of the most recent step (commits retire in order)
            const int prev = stage == 0 ? kSwStages - 1 : stage - 1;
            const uint32_t prev_phase = stage == 0 ? phase ^ 1 : phase;
            if (!mbar_wait(&empty_bar[prev], prev_phase, err)) { ok = false; break; }
          }
          if (!mbar_wait(&empty_bar[stage], phase ^ 1, err)) { ok = false; break; }
          const int r_lo = pr == 0 ? 0 : p.R - p.sh;     // window boxes this step brings in
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.k_boxes + p.R - r_lo) * 4096u);
          const int pix = (img * p.P + pr) * p.Q + q0;
          for (int j = 0; j < p.k_boxes; ++j) tma_load_2d(sA + stage * p.a_stage_bytes + j * 4096, &tmA, &full_bar[stage], 32 * j, pix);
          for (int r = r_lo; r < p.R; ++r)   // rows outside the image: zero fill
            tma_load_4d(sB + ((pr * p.sh + r) % p.ring) * 4096, &tmB, &full_bar[stage], 0, q0, pr * p.sh - p.ph + r, img);
          if (++stage == kSwStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      const uint64_t a_desc0 = make_smem_desc(0, 4096, 512, kSmemLayoutSw128Base32);
      const uint64_t b_desc0 = make_smem_desc(0, 4096, 512, kSmemLayoutSw128Base32);
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true, first = true;
      for (int strip = blockIdx.x; strip < p.strips && ok; strip += gridDim.x) {
        for (int pr = 0; pr < p.P; ++pr) {
          if (!mbar_wait(&full_bar[stage], phase, err)) { ok = false; break; }
          tc_fence_after();
          const int start = (pr * p.sh) % p.ring;
          const int n1 = min(p.R, p.ring - start), n2 = p.R - n1;   // boxes before / after the ring wraps
          const uint32_t id1 = make_idesc_tf32(kUmmaBM, n1 * 32, 1, 1);
          const uint32_t id2 = n2 > 0 ? make_idesc_tf32(kUmmaBM, n2 * 32, 1, 1) : 0u;
          const uint64_t da = a_desc0 + ((a0 + stage * p.a_stage_bytes) >> 4);
          const uint64_t db1 = b_desc0 + ((b0 + start * 4096) >> 4), db2 = b_desc0 + (b0 >> 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {   // 4 x 8 pixels
            const uint32_t accum = (first && i == 0) ? 0u : 1u;
            umma_tf32(tmem_base, da + i * 64, db1 + i * 64, id1, accum);
            if (n2 > 0) umma_tf32(tmem_base + n1 * 32, da + i * 64, db2 + i * 64, id2, accum);
          }
          first = false;
          umma_commit(&empty_bar[stage]);
          if (++stage == kSwStages) { stage = 0; phase ^= 1; }
        }
      }
      if (ok && has_work) umma_commit(done_bar);
    }
  } else {
    // epilogue: lane = k row, 32 window elements of filter row r per tcgen05.ld; a CTA without work writes zeros
    const int ew = warp & 3;
    const int k = ew * 32 + lane;
    float* dst_row = p.partial + (static_cast<long long>(blockIdx.x) * p.K + k) * (p.R * 32);
    bool ready = true;
    if (has_work) {
      ready = mbar_wait(done_bar, 0, err);
      tc_fence_after();
    }
    if (ready) {
      for (int r = 0; r < p.R; ++r) {
        uint32_t v[32];
        if (has_work) {
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + r * 32, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0u;
        }
        if (k < p.K) {
          float4* dst = reinterpret_cast<float4*>(dst_row + r * 32);
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
    tmem_dealloc(tmem_base, 256);
  }
}

static CUtensorMapDataType sw_dtype() {
  const char* e = getenv("ZENU_B200_TMA_F32");
  return (e && e[0] == '1') ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
}

// xp: the packed NHWC4 input ([N][H][Wp][4], see smallc_pack_input_kernel); part: [*splits_out][K][R*32] partials.
// Returns ZB_ERR_UNSUPPORTED (nothing launched) when the geometry is not served.
int umma_conv_stem_wgrad(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* xp, long long Wp, long long P, long long Q,
                         float* part, int max_splits, int* splits_out) {
  if (ZB_ENV_FLAG("ZENU_B200_NO_STEM_WGRAD")) return ZB_ERR_UNSUPPORTED;
  const int R = static_cast<int>(d->kh), sh = static_cast<int>(d->stride_h);
  if (R > 8 || d->dil_h != 1 || sh > R || d->k > 128 || d->k % 4 != 0 || (reinterpret_cast<uintptr_t>(dy) & 15) != 0) return ZB_ERR_UNSUPPORTED;
  StemWgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = static_cast<int>(d->n); p.H = static_cast<int>(d->h); p.P = static_cast<int>(P); p.Q = static_cast<int>(Q);
  p.K = static_cast<int>(d->k); p.R = R; p.sh = sh; p.ph = static_cast<int>(d->pad_h);
  p.q_runs = ceil_div(Q, 32);
  if (d->n * p.q_runs > 0x3fffffffll || d->n * P * Q > 0x7fffffffll) return ZB_ERR_UNSUPPORTED;
  p.strips = static_cast<int>(d->n) * p.q_runs;
  p.k_boxes = ceil_div(d->k, 32);
  p.ring = R + sh * (kSwStages + 2);
  p.a_stage_bytes = p.k_boxes * 4096;
  const int grid = std::min(std::min(p.strips, ctx->sm_count), max_splits);
  if (grid < 1) return ZB_ERR_UNSUPPORTED;
  p.partial = part;
  p.err_flag = ctx->err_flag;
  const size_t smem = static_cast<size_t>(p.ring) * 4096 + static_cast<size_t>(kSwStages) * p.a_stage_bytes + (4 - p.k_boxes) * 4096 + 256 + 1024;
  if (smem > 200 * 1024) return ZB_ERR_UNSUPPORTED;
  CUtensorMap ma, mb;
  {   // dY as [N*P*Q][K] row-major, box = 32 output channels x 32 pixels, MN-major swizzle
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d->k), static_cast<cuuint64_t>(d->n * P * Q)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d->k) * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode_tiled(&ma, sw_dtype(), 2, const_cast<float*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (stem wgrad dY) failed (%d)", int(r)); return ZB_ERR_CUDA; }
  }
  {   // sliding windows: element (j, q, h, n) = xp[n][h][q*sw*4 + j]; box = 32 floats x 32 windows of one input row
    cuuint64_t dims[4] = {32, static_cast<cuuint64_t>(Q), static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->stride_w) * 16, static_cast<cuuint64_t>(Wp) * 16, static_cast<cuuint64_t>(d->h) * Wp * 16};
    cuuint32_t box[4] = {32, 32, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ctx->encode_tiled(&mb, sw_dtype(), 4, const_cast<float*>(xp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (stem wgrad windows) failed (%d)", int(r)); return ZB_ERR_CUDA; }
  }
  static SmemOptIn opt_in;
  { const int rc2 = smem_opt_in(ctx, opt_in, stem_wgrad_kernel, smem); if (rc2 != ZB_OK) return rc2; }
  plan_note("stem_wgrad ring=%d k_boxes=%d ~strips=%d ~grid=%d;", p.ring, p.k_boxes, p.strips, grid);
  *splits_out = grid;
  if (plan_dry()) return ZB_OK;
  prof_begin(ctx, PROF_TENSOR);
  stem_wgrad_kernel<<<grid, 192, smem, ctx->stream>>>(ma, mb, p);
  prof_end(ctx, PROF_TENSOR, 2.0 * d->n * P * Q * d->k * d->c * d->kh * d->kw);
  ZB_LAUNCH_CHECK(ctx);
  *splits_out = grid;
  return ZB_OK;
}

}  // namespace zb
