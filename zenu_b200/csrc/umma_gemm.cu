// umma_gemm.cu — the tensor-core hot path: one persistent, warp-specialised tcgen05 kernel that serves
//   * Linear / generic row-major GEMM (all four transpose combinations),
//   * Conv2d fprop / dgrad as implicit GEMM (TMA im2col loads of NHWC activations),
//   * Conv2d wgrad as implicit GEMM with the pixel dimension as the reduction (MN-major operands, split-K),
// plus the host-side planners that turn a conv/GEMM request into tensor maps + UmmaParams.
//
// Replaces: cuDNN-frontend conv graphs (reference conv.cpp:29-187 via graph_conv.rs:40-268) and
// cublas{S}gemm_v2_64 (zenu-cuda/src/cublas/mod.rs:84-160).  f32 operands are consumed as TF32
// (kind::tf32, fp32 accumulation in TMEM).
//
// Kernel anatomy (192 threads, 1 CTA / SM, grid = min(tiles, SMs), static round-robin tile schedule):
//   warp 0   : TMA producer   — one elected lane fills a STAGES-deep smem ring (A 128x32, B BNx32 fp32, SWIZZLE_128B)
//   warp 1   : MMA issuer     — one elected lane issues 4 x tcgen05.mma (K = 8 each) per stage into a
//                               double-buffered TMEM accumulator (2 x BN columns); also owns TMEM alloc/dealloc
//   warps 2-5: epilogue       — tcgen05.ld 32x32b.x32 -> registers -> alpha/bias/beta -> 128-bit global stores
// Pipelines: full/empty mbarriers (TMA <-> MMA), tmem_full/tmem_empty (MMA <-> epilogue).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace zb {

using namespace ptx;

constexpr int kOldDepth = 3;   // 32-column chunks of the old output tile in flight per epilogue warp (beta != 0 launches)
template <int BN, int STAGES, bool OLD = false, int CL = 1>
struct UmmaSmem {
  static constexpr int A_BYTES = kUmmaBM * kUmmaBK * 4;  // 16 KB
  static constexpr int B_BYTES = (BN / CL) * kUmmaBK * 4;   // CTA pair (CL = 2): each CTA holds half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;   // 4 warps x [32 rows][32 cols] fp32 staging (coalesced stores)
  static constexpr int EPI_BYTES = 4 * 4096;
  static constexpr int STAT_OFFSET = EPI_OFFSET + EPI_BYTES;   // 4 warps x [sum, sum of squares, shift][BN] fp32 (fused BatchNorm statistics)
  static constexpr int STAT_BYTES = 4 * 3 * BN * 4;
  static constexpr int OLD_OFFSET = STAT_OFFSET + STAT_BYTES;   // beta launches: 4 warps x kOldDepth x 4 KB cp.async landing buffers
  static constexpr int OLD_BYTES = OLD ? 4 * kOldDepth * 4096 : 0;
  static constexpr int BAR_OFFSET = OLD_OFFSET + OLD_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;  // + alignment slack
  // a pair always takes the whole TMEM so that both CTAs get the same base (the leader's MMA addresses both with one column offset)
  static constexpr int TMEM_COLS = CL > 1 ? 512 : (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

struct TileCoord {
  int m_blk, n_blk, tap, split;
};
template <bool V> struct FullTag { static constexpr bool value = V; };
template <int V> struct IntTag { static constexpr int value = V; };
struct TileDiv {   // the divisors of a tile id, as multiply-high constants (built once per role)
  FastDiv n, t, m;
  int m_groups;
};
__device__ __forceinline__ TileDiv make_tile_div(const UmmaParams& p, int cl) {
  TileDiv d;
  d.m_groups = (p.m_tiles + cl - 1) / cl;
  d.n = fastdiv_make(p.n_tiles);
  d.t = fastdiv_make(p.tap_tiles);
  d.m = fastdiv_make(d.m_groups);
  return d;
}
__device__ __forceinline__ TileCoord decode_tile(const UmmaParams& p, int tile, const TileDiv& d) {
  TileCoord t;
  int q = fastdiv(tile, d.n);
  t.n_blk = tile - q * p.n_tiles;
  tile = q;
  q = fastdiv(tile, d.t);
  t.tap = tile - q * p.tap_tiles;
  tile = q;
  q = fastdiv(tile, d.m);
  t.m_blk = tile - q * p.m_tiles;
  t.split = q;
  return t;
}

// Cluster walk (CL CTAs on CL consecutive row blocks of the same column block / tap / split): work item q of the cluster ->
// this CTA's tile.  m_tiles need not be a multiple of CL for the halo kernel (a row block past the end is a dummy).
__device__ __forceinline__ TileCoord decode_tile_cluster(const UmmaParams& p, int q, int cl, int rank, const TileDiv& d) {
  TileCoord t;
  int u = fastdiv(q, d.n);
  t.n_blk = q - u * p.n_tiles;
  q = u;
  u = fastdiv(q, d.t);
  t.tap = q - u * p.tap_tiles;
  q = u;
  u = fastdiv(q, d.m);
  t.m_blk = (q - u * d.m_groups) * cl + rank;
  t.split = u;
  return t;
}

// K-block range of a tile: split-K slices, or (small-C dgrad) the filter rows that reach input row h
__device__ __forceinline__ void tile_kb_range(const UmmaParams& p, const TileCoord& tc, int& kb_begin, int& kb_end) {
  if (p.a_mode == A_ROWS_K) {
    const int h = tc.m_blk % p.dg_H;
    kb_begin = 0;
    kb_end = p.dg_cnt[(h + p.dg_ph) % p.dg_sh] * p.c_chunks;
  } else if (p.batch_mode == 1) {   // the split index is an image: every tile reduces over the whole K range
    kb_begin = 0;
    kb_end = p.kb_total;
  } else {
    kb_begin = tc.split * p.kb_per_split;
    kb_end = min(kb_begin + p.kb_per_split, p.kb_total);
  }
}

// ---------------------------------------------------------------------------------------------- epilogue (warps 2-5)
// Shared by the implicit-GEMM kernel and the halo-reuse conv kernel: drains finished TMEM accumulators tile by tile.
template <int BN, bool MASKABLE = false, bool DEEP = false>
__device__ __forceinline__ void epilogue_role(const UmmaParams& p, uint8_t* smem, int epi_offset, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t tmem_base, int warp, int lane, volatile int* err,
                                              bool halo = false, uint8_t* old_smem = nullptr, int n_acc = 2, int group = 1,
                                              int cluster = 1, bool pair_mma = false) {
  // pair_mma: the accumulators are written by the cta_group::2 MMAs of the cluster's leader CTA (rank 0), which waits for the
  // epilogues of BOTH CTAs before it overwrites a TMEM buffer: every warp releases a buffer on the leader's tempty barrier
  // cluster > 1 (halo kernel with multicast filter tiles): the CTAs of a cluster walk tile groups in lockstep: group q ->
  // column block q % n_tiles, row block (q / n_tiles) * cluster + rank; a row block past the end is a dummy (nothing stored)
  // n_acc TMEM accumulators of BN columns are drained in rotation; a CTA takes tiles in groups of `group` consecutive ids
  // (group > 1: the multi-row stem kernel finishes `group` accumulators per scheduling step)
  const int total_tiles = p.m_tiles * p.n_tiles * p.tap_tiles * p.splits;
  struct LL { int EPI_OFFSET; } Lv{epi_offset};
    // ================================ epilogue ================================
    // Each warp owns 32 accumulator rows (its TMEM lane quarter).  Per 32-column chunk: tcgen05.ld (lane = row) ->
    // XOR-swizzled smem staging -> read back so that 8 lanes cover one row's 128 bytes -> coalesced 128-bit global
    // stores (4 full cache lines per instruction) with alpha / bias / beta applied on the way out.
    const int ew = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    float4* stage4 = reinterpret_cast<float4*>(smem + Lv.EPI_OFFSET + ew * 4096);
    const uint32_t stage_u32 = smem_u32(stage4);
    const bool vec_ok = ((p.ldd & 3) == 0) && ((p.tap_col_stride & 3) == 0) && ((p.split_stride & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) &&
                        (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    const bool partial = p.splits > 1 && p.batch_mode != 1;
    const int sub_row = lane >> 3, piece = lane & 7;
    float* stat_w = reinterpret_cast<float*>(smem + Lv.EPI_OFFSET + 4 * 4096) + ew * 3 * BN;   // this warp's [3][BN]: sum, sum sq, shift
    const bool stats = p.stat_partial != nullptr;
    if (stats) {   // the grid is a multiple of n_tiles: this CTA only ever sees column block blockIdx.x % n_tiles
      const int nb0 = ((blockIdx.x / cluster) % p.n_tiles) * BN;
      for (int j = lane; j < BN; j += 32) {
        stat_w[j] = 0.f;
        stat_w[BN + j] = 0.f;
        stat_w[2 * BN + j] = nb0 + j < p.N ? __ldg(p.stat_shift + nb0 + j) : 0.f;
      }
      __syncwarp();
    }
    const int groups = total_tiles / group;
    const TileDiv tdiv = make_tile_div(p, cluster);
    const FastDiv fd_group = fastdiv_make(group);
    const FastDiv fd_win_img = fastdiv_make(p.win_p_tiles * p.win_q_tiles), fd_win_qt = fastdiv_make(p.win_q_tiles),
                  fd_win_bq = fastdiv_make(p.win_box_q);
    const FastDiv fd_pq = fastdiv_make(p.conv_P * p.conv_Q), fd_q = fastdiv_make(p.conv_Q);
    const FastDiv fd_stack = fastdiv_make(p.halo_stack_rows > 0 ? p.halo_stack_rows : 1);
    const int cl_id = blockIdx.x / cluster, n_cl = gridDim.x / cluster, cl_rank = blockIdx.x % cluster;
    const int cl_groups = ((p.m_tiles + cluster - 1) / cluster) * p.n_tiles * p.tap_tiles * p.splits;
    for (int step = 0;; ++step) {
      TileCoord tc;
      bool tile_exists = true;
      if (cluster == 1) {
        const int sg = fastdiv(step, fd_group);
        const int grp = blockIdx.x + sg * gridDim.x;
        if (grp >= groups) break;
        tc = decode_tile(p, grp * group + (step - sg * group), tdiv);
      } else {
        const int q = cl_id + step * n_cl;
        if (q >= cl_groups) break;
        tc = decode_tile_cluster(p, q, cluster, cl_rank, tdiv);
        tile_exists = tc.m_blk < p.m_tiles;
      }
      const int m0 = tc.m_blk * kUmmaBM, n0 = tc.n_blk * BN;
      // element offset of this lane's row inside D (or -1 when the row does not exist)
      long long my_off = -1;
      {
        const int l = ew * 32 + lane;
        if (p.out_mode == OUT_WINDOW) {
          const int per_img = p.win_p_tiles * p.win_q_tiles;
          int img = fastdiv(tc.m_blk, fd_win_img);
          const int rem = tc.m_blk - img * per_img;
          const int pt = fastdiv(rem, fd_win_qt), qt = rem - pt * p.win_q_tiles;
          int pl = fastdiv(l, fd_win_bq);
          const int ql = l - pl * p.win_box_q;
          bool img_ok = true;
          if (p.halo_stack > 1) {   // stacked tile: window row pl = image pl / stack_rows of the group, output row pl % stack_rows
            const int g = fastdiv(pl, fd_stack);
            pl -= g * p.halo_stack_rows;
            img = tc.m_blk * p.halo_stack + g;
            img_ok = g < p.halo_stack && img < p.batch_n;
          }
          const int pp = pt * p.win_box_p + pl, qq = qt * p.win_box_q + ql;
          if (tile_exists && img_ok && pl < p.win_box_p && pp < p.conv_P && qq < p.conv_Q) {
            if (p.scat_sy != 0)   // a parity class of a strided dgrad: window (pp, qq) is input pixel (pp * sy + oy, qq * sx + ox)
              my_off = ((static_cast<long long>(img) * p.scat_OH + pp * p.scat_sy + p.scat_oy) * p.scat_OW + qq * p.scat_sx + p.scat_ox) * p.ldd;
            else
              my_off = ((static_cast<long long>(img) * p.conv_P + pp) * p.conv_Q + qq) * p.ldd;
          }
        } else {
          const int row = m0 + l;
          if (row < p.M) {
            long long orow = row;
            if (p.out_mode == OUT_SCATTER) {
              const int pq = p.conv_P * p.conv_Q;
              const int img = fastdiv(row, fd_pq), rem = row - img * pq;
              const int pp = fastdiv(rem, fd_q), qq = rem - pp * p.conv_Q;
              orow = (static_cast<long long>(img) * p.scat_OH + pp * p.scat_sy + p.scat_oy) * p.scat_OW + qq * p.scat_sx + p.scat_ox;
            }
            my_off = orow * p.ldd;
          }
        }
        if (my_off >= 0) my_off += static_cast<long long>(tc.split) * p.split_stride + tc.tap * p.tap_col_stride;
      }
      // accumulator flushes of this tile (see UmmaParams::chain_kb): all but the first accumulate into D
      int n_sub = 1;
      if (halo) {
        if (p.halo_chain > 0) n_sub = (p.c_chunks + p.halo_chain - 1) / p.halo_chain;
      } else if (p.chain_kb > 0) {
        int kb0, kb1;
        tile_kb_range(p, tc, kb0, kb1);
        n_sub = max(1, (kb1 - kb0 + p.chain_kb - 1) / p.chain_kb);
      }
      if (p.out_mode == OUT_WDGRAD) {
        if (!mbar_wait(&tfull_bar[acc], acc_phase, err)) break;
        tc_fence_after();
        // accumulator D[q][s*4+c] (N = 32) -> shared [128 q][32] tile (all four warps) -> every thread overlap-adds the
        // filter columns s that reach its output element and stores the dense row dX[n][h][0..W)[0..C) coalesced
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN, r);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 8; ++v)
          stage4[lane * 8 + (v ^ (lane & 7))] = make_float4(__uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]),
                                                            __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // TMEM buffer is free: the MMA warp may start the next tile
        // one thread per output pixel w: the <= ceil(S / stride_w) filter columns that reach it are 128-bit smem reads (4 channels
        // of one tap), q steps down by one as the column steps up by stride_w
        const float4* tile4 = reinterpret_cast<const float4*>(smem + Lv.EPI_OFFSET);
        int kb0, kb1;
        tile_kb_range(p, tc, kb0, kb1);
        float* out_row = p.D + static_cast<long long>(tc.m_blk) * (p.dg_W * p.dg_C);
        for (int w = ew * 32 + lane; w < p.dg_W; w += 128) {
          float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kb1 > kb0) {
            const int sx0 = (w + p.dg_pw) % p.dg_sw;
            int q = (w + p.dg_pw - sx0) / p.dg_sw;
            for (int sx = sx0; sx < p.dg_S; sx += p.dg_sw, --q) {
              if (q >= 0 && q < p.conv_Q) {
                const float4 v = tile4[q * 8 + (sx ^ (q & 7))];
                sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
              }
            }
          }
          float* o = out_row + w * p.dg_C;
          const float vals4[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c < p.dg_C) o[c] = p.beta != 0.f ? vals4[c] + p.beta * o[c] : vals4[c];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the tile is rewritten by the next accumulator
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      // rows this lane stores after the transpose (8 per chunk, the same 8 for every chunk of the tile), as 32-bit element
      // offsets from the warp's first existing row (-1: the row does not exist: ragged last tile, halo positions).  The
      // epilogue warps are instruction-bound on the output-dominated layers (ncu: ~0.28 IPC per scheduler, no single hot
      // stall), so the per-chunk code is kept lean: one 64-bit base, IMAD.WIDE addressing, and a branch-free variant for
      // tiles whose 32 rows all exist.
      const unsigned valid_mask = __ballot_sync(0xffffffffu, my_off >= 0);
      long long base_off = 0;
      int rel = -1;
      bool tile_vec = vec_ok;
      if (valid_mask != 0u) {
        base_off = __shfl_sync(0xffffffffu, my_off, __ffs(valid_mask) - 1);
        const long long dlt = my_off - base_off;
        if (__any_sync(0xffffffffu, my_off >= 0 && (dlt < 0 || dlt > 0x3fffffffll))) tile_vec = false;
        rel = my_off >= 0 ? static_cast<int>(dlt) : -1;
      }
      int off32[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) off32[it] = __shfl_sync(0xffffffffu, rel, it * 4 + sub_row);
      const bool full = valid_mask == 0xffffffffu;
      float* const wb = p.D + base_off + n0 + piece * 4;   // element (row it, chunk c) = wb[off32[it] + c * 32 ..+3]
      const uint32_t stat_u32 = smem_u32(stat_w) + piece * 16;
      bool dead = false;
#pragma unroll 1
      for (int sub = 0; sub < n_sub; ++sub) {
      const float e_alpha = partial ? 1.f : p.alpha;
      const float e_beta = sub == 0 ? (partial ? 0.f : p.beta) : 1.f;
      const float* e_bias = (sub == 0 && !partial) ? p.bias : nullptr;
      const bool use_beta = e_beta != 0.f;
      const bool plain = e_alpha == 1.f && e_bias == nullptr && !use_beta;   // store the accumulator as is
      // beta != 0 (gradient fan-in, 3xTF32 passes): the old tile is fetched ahead of use, the first chunk(s) before the
      // accumulator is even complete, so the read latency hides behind the MMAs / the previous chunk's stores.
      // deep variant (old_smem != nullptr): kOldDepth chunks in flight per warp through cp.async into warp-private smem slots
      // (slot = [chunk % depth][row group it][lane], 16 bytes each); one register chunk ahead is not enough memory-level
      // parallelism for the K = 64..256 gradient fan-in GEMMs, whose time is the read-modify-write of the output
      float4 olds[8];
      const bool prefetch = use_beta && tile_vec;
      // DEEP (the OLD = true kernel variants): old tiles always come through the cp.async ring; the one-chunk-ahead register prefetch
      // below is compiled out there (its 32 registers made those variants spill once the mask nibbles were added)
      const bool deep = DEEP && prefetch && old_smem != nullptr;
      // old values through a bit mask (UmmaParams::old_bits): first flush of the tile only, like the caller's beta itself
      // (MASKABLE: only the kernel variants that serve beta launches carry the masking code and its registers)
      const bool use_mask = MASKABLE && use_beta && sub == 0 && !partial && p.old_bits != nullptr;
      const uint32_t old_u32 = deep ? smem_u32(old_smem) + ew * (kOldDepth * 4096) + lane * 16 : 0u;
      auto issue_old = [&](int c) {
        if (n0 + c * 32 + 32 <= p.N && c < BN / 32) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (off32[it] >= 0)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(old_u32 + ((c % kOldDepth) * 8 + it) * 512),
                           "l"(wb + (off32[it] + c * 32)) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      // mask words of the whole tile (32 rows x BN / 32 words per warp) land in this warp's statistics slot (never used by a beta
      // launch) through 4-byte cp.async: lane (sub_row, piece) fetches word (row it * 4 + sub_row, chunk piece).  They join the first
      // group of the old-tile ring, so the first chunk's wait covers them and no chunk waits on a global load of its own.
      constexpr int kMaskCh = BN / 32;
      const bool mask_tile = use_mask && tile_vec;
      const uint32_t mask_u32 = smem_u32(stat_w);
      if (mask_tile) {
        if (piece < kMaskCh && n0 + piece * 32 + 32 <= p.N) {
#pragma unroll
          for (int it = 0; it < 8; ++it)
            if (off32[it] >= 0) {
              const long long idx = base_off + n0 + off32[it] + piece * 32;   // a multiple of 32 (ldd % 32 == 0: checked at launch)
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(mask_u32 + ((it * 4 + sub_row) * kMaskCh + piece) * 4),
                           "l"(p.old_bits + (idx >> 5)) : "memory");
            }
        }
        if (!deep) asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (deep) {
#pragma unroll
        for (int c = 0; c < kOldDepth; ++c) issue_old(c);
      } else if (!DEEP && prefetch && n0 + 32 <= p.N) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
          olds[it] = off32[it] >= 0 ? *reinterpret_cast<const float4*>(wb + off32[it]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (!mbar_wait(&tfull_bar[acc], acc_phase, err)) { dead = true; break; }
      tc_fence_after();
      if (mask_tile && !deep) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
      // FULL: every row of this warp exists (no per-row predicates)
      auto chunk_loop = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n0 + c * 32;
          if (col0 >= p.N) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int v = 0; v < 8; ++v)
            sts128(stage_u32 + lane * 128 + ((v ^ (lane & 7)) << 4), __uint_as_float(r[v * 4]), __uint_as_float(r[v * 4 + 1]),
                   __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
          __syncwarp();
          const int col = col0 + piece * 4;
          if (tile_vec && col0 + 32 <= p.N) {   // warp-uniform fast path: whole 32-column chunk, 128-bit stores
            float4 vals[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + sub_row;
              vals[it] = lds128(stage_u32 + rr * 128 + ((piece ^ (rr & 7)) << 4));
            }
            const int coff = c * 32;
            // chunk base kept as ONE opaque 64-bit register: each store address is a single IMAD.WIDE (off32 * 4 + cb) instead of a
            // 64-bit offset sum rebuilt per store (5 instructions each in the SASS of the previous form)
            float* cb = wb + coff;
            asm volatile("" : "+l"(cb));
            if (plain) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                if (FULL || off32[it] >= 0) stg128(cb + off32[it], vals[it]);
            } else {
              float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (e_bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(e_bias + col));
              float4 cur[8];
              if (deep) {
                asm volatile("cp.async.wait_group %0;" ::"n"(kOldDepth - 1) : "memory");
                if (mask_tile && c == 0) __syncwarp();   // the tile's mask words (other lanes' copies, first group) are visible
#pragma unroll
                for (int it = 0; it < 8; ++it)
                  cur[it] = (FULL || off32[it] >= 0) ? lds128(old_u32 + ((c % kOldDepth) * 8 + it) * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
                issue_old(c + kOldDepth);   // refills the slot just read
              } else if (!DEEP) {
#pragma unroll
                for (int it = 0; it < 8; ++it) cur[it] = olds[it];
                if (use_beta && c + 1 < BN / 32 && col0 + 64 <= p.N) {   // next chunk's old values: in flight during this chunk's stores
#pragma unroll
                  for (int it = 0; it < 8; ++it)
                    olds[it] = (FULL || off32[it] >= 0) ? *reinterpret_cast<const float4*>(wb + (off32[it] + coff + 32))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
              if (use_mask) {
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const uint32_t nb = lds32(mask_u32 + ((it * 4 + sub_row) * kMaskCh + c) * 4) >> (piece * 4);
                  cur[it].x = (nb & 1u) ? cur[it].x : 0.f; cur[it].y = (nb & 2u) ? cur[it].y : 0.f;
                  cur[it].z = (nb & 4u) ? cur[it].z : 0.f; cur[it].w = (nb & 8u) ? cur[it].w : 0.f;
                }
              }
              const float2 al2 = make_float2(e_alpha, e_alpha), be2 = make_float2(e_beta, e_beta);
              const float2 bva = make_float2(bv.x, bv.y), bvb = make_float2(bv.z, bv.w);
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                if (!FULL && off32[it] < 0) continue;
                // alpha * acc + bias (+ beta * old) as packed fused multiply-adds: the roundings of the scalar fmaf forms
                float2 oa = fma_f32x2(al2, make_float2(vals[it].x, vals[it].y), bva);
                float2 ob = fma_f32x2(al2, make_float2(vals[it].z, vals[it].w), bvb);
                if (use_beta) {
                  oa = fma_f32x2(be2, make_float2(cur[it].x, cur[it].y), oa);
                  ob = fma_f32x2(be2, make_float2(cur[it].z, cur[it].w), ob);
                }
                const float4 o = make_float4(oa.x, oa.y, ob.x, ob.y);
                vals[it] = o;
                stg128(cb + off32[it], o);
              }
            }
            if (stats) {   // BatchNorm statistics of the values just stored (rows that exist only)
              const float4 sh = lds128(stat_u32 + (2 * BN + coff) * 4);
              // packed fp32 pairs (FADD2 / FFMA2): d = v - shift, s1 += d, s2 = fma(d, d, s2) with the roundings of the scalar forms
              const float2 sha = make_float2(sh.x, sh.y), shb = make_float2(sh.z, sh.w);
              float2 s1a = make_float2(0.f, 0.f), s1b = s1a, s2a = s1a, s2b = s1a;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                if (!FULL && off32[it] < 0) continue;
                const float2 da = sub_f32x2(make_float2(vals[it].x, vals[it].y), sha), db = sub_f32x2(make_float2(vals[it].z, vals[it].w), shb);
                s1a = add_f32x2(s1a, da); s1b = add_f32x2(s1b, db);
                s2a = fma_f32x2(da, da, s2a); s2b = fma_f32x2(db, db, s2b);
              }
#pragma unroll
              for (int o = 8; o <= 16; o <<= 1) {
                s1a = add_f32x2(s1a, make_float2(__shfl_xor_sync(0xffffffffu, s1a.x, o), __shfl_xor_sync(0xffffffffu, s1a.y, o)));
                s1b = add_f32x2(s1b, make_float2(__shfl_xor_sync(0xffffffffu, s1b.x, o), __shfl_xor_sync(0xffffffffu, s1b.y, o)));
                s2a = add_f32x2(s2a, make_float2(__shfl_xor_sync(0xffffffffu, s2a.x, o), __shfl_xor_sync(0xffffffffu, s2a.y, o)));
                s2b = add_f32x2(s2b, make_float2(__shfl_xor_sync(0xffffffffu, s2b.x, o), __shfl_xor_sync(0xffffffffu, s2b.y, o)));
              }
              if (sub_row == 0) {
                const float4 t1 = lds128(stat_u32 + coff * 4), t2 = lds128(stat_u32 + (BN + coff) * 4);
                const float2 u1a = add_f32x2(make_float2(t1.x, t1.y), s1a), u1b = add_f32x2(make_float2(t1.z, t1.w), s1b);
                const float2 u2a = add_f32x2(make_float2(t2.x, t2.y), s2a), u2b = add_f32x2(make_float2(t2.z, t2.w), s2b);
                sts128(stat_u32 + coff * 4, u1a.x, u1a.y, u1b.x, u1b.y);
                sts128(stat_u32 + (BN + coff) * 4, u2a.x, u2a.y, u2b.x, u2b.y);
              }
            }
          } else {   // ragged chunk or unaligned output: element-wise (kept out of registers: rare path)
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + sub_row;
              const long long off = __shfl_sync(0xffffffffu, my_off, rr);
              const float4 v4 = lds128(stage_u32 + rr * 128 + ((piece ^ (rr & 7)) << 4));
              if (off < 0) continue;
              float* dst = p.D + off + col;
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (col + e < p.N) {
                  float val = e == 0 ? v4.x : (e == 1 ? v4.y : (e == 2 ? v4.z : v4.w));
                  val = e_alpha * val + (e_bias != nullptr ? __ldg(e_bias + col + e) : 0.f);
                  if (use_beta) {
                    const long long idx = off + col + e;
                    const bool keep = !use_mask || ((__ldg(p.old_bits + (idx >> 5)) >> (idx & 31)) & 1u) != 0u;
                    if (keep) val += e_beta * dst[e];
                  }
                  dst[e] = val;
                }
            }
          }
          __syncwarp();  // staging is rewritten by the next chunk
        }
      };
      if (full) chunk_loop(FullTag<true>{}); else chunk_loop(FullTag<false>{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair_mma) mbar_arrive_remote(mapa_shared(smem_u32(&tempty_bar[acc]), 0)); else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
      if (dead) break;
    }
    if (stats) {   // one row of partials per epilogue warp: [row][2][N]
      __syncwarp();
      const int n_blk = (blockIdx.x / cluster) % p.n_tiles;
      const long long row = (cluster == 1 ? static_cast<long long>(blockIdx.x / p.n_tiles)
                                          : static_cast<long long>((blockIdx.x / cluster) / p.n_tiles) * cluster + blockIdx.x % cluster) * 4 + ew;
      for (int j = lane; j < BN; j += 32) {
        const int col = n_blk * BN + j;
        if (col < p.N) {
          p.stat_partial[(row * 2 + 0) * p.N + col] = stat_w[j];
          p.stat_partial[(row * 2 + 1) * p.N + col] = stat_w[BN + j];
        }
      }
    }
}

// CL = 2: CTA pairs (clusters of two CTAs, tcgen05 cta_group::2).  The pair owns a 256 x BN tile: each CTA loads its own 128 rows of
// A and HALF of the B tile into its own shared memory, the leader (rank 0) issues M = 256 MMAs that read both shared memories and
// write each CTA's 128 accumulator rows into that CTA's TMEM, and each CTA's epilogue drains its own TMEM.  A K block then costs a
// CTA A + B/2 = 32 KB of TMA traffic into its shared memory instead of 48 KB (BN = 256): measured, an SM ingests at most ~70-80 B/clk
// through TMA (tools/probes/tma_bw), and a private 128 x 256 fp32 tile needs 96 B/clk at the full tensor rate.  (Sharing B by TMA
// multicast between two independent CTAs was measured first: it lowers the L2 reads but not the bytes an SM ingests, and bought
// nothing.)  Barriers: both producers count their bytes on the LEADER's full barrier (cta_group::2 TMA loads), the leader's
// tcgen05.commit is multicast onto both CTAs' empty / tmem-full barriers, both epilogues arrive on the leader's tmem-empty barrier.
template <int BN, int STAGES, bool OLD = false, int CL = 1>
__global__ void __launch_bounds__(192, 1)
umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ UmmaParams p) {
  using L = UmmaSmem<BN, STAGES, OLD, CL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile int* err = p.err_flag;

  if (warp == 0) {
    if (elect_one()) {
      prefetch_tensormap(&tmA);
      prefetch_tensormap(&tmB);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], 4 * CL);
      }
      fence_barrier_init();
    }
    __syncwarp();
    if (CL == 1) {
      tmem_alloc(tmem_slot, L::TMEM_COLS);
      tmem_relinquish();
    } else {
      tmem_alloc_pair(tmem_slot, L::TMEM_COLS);
      tmem_relinquish_pair();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work items: tiles (CL = 1), or pairs of row blocks walked by the cluster (m_tiles % CL == 0, checked by the host)
  const int total_tiles = (p.m_tiles / CL) * p.n_tiles * p.tap_tiles * p.splits;
  const int cl_rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int w_first = blockIdx.x / CL, w_step = gridDim.x / CL;
  const TileDiv tdiv = make_tile_div(p, CL);
  const bool a_mn = (p.a_mode == A_TILED_MN);
  const bool b_mn = (p.b_mode != B_TILED_K);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int pq = p.conv_P * p.conv_Q;
      for (int tile = w_first; tile < total_tiles; tile += w_step) {
        const TileCoord tc = CL > 1 ? decode_tile_cluster(p, tile, CL, cl_rank, tdiv) : decode_tile(p, tile, tdiv);
        const int m0 = tc.m_blk * kUmmaBM, n0 = tc.n_blk * BN;
        int kb_begin, kb_end;
        tile_kb_range(p, tc, kb_begin, kb_end);
        int a_w = 0, a_h = 0, a_n = 0;
        if (p.a_mode == A_IM2COL_K) {
          const int img = m0 / pq, rem = m0 - img * pq;
          const int pp = rem / p.conv_Q, qq = rem - pp * p.conv_Q;
          a_w = p.lower_w + qq * p.stride_w;
          a_h = p.lower_h + pp * p.stride_h;
          a_n = img;
        } else if (p.a_mode == A_WINDOW_K) {
          const int per_img = p.win_p_tiles * p.win_q_tiles;
          const int img = tc.m_blk / per_img, rem = tc.m_blk - img * per_img;
          const int pt = rem / p.win_q_tiles, qt = rem - pt * p.win_q_tiles;
          a_w = qt * p.win_box_q;                                   // first output column of the tile
          a_h = pt * p.win_box_p * p.stride_h + p.lower_h;          // input row of filter row 0
          a_n = img;
        } else if (p.a_mode == A_ROWS_K) {
          a_n = tc.m_blk / p.dg_H;
          a_h = tc.m_blk - a_n * p.dg_H;                            // input row h of dX
        }
        // MN-major operands come in 32-wide boxes; a pair CTA loads the boxes of its own half of the B tile (columns nB0 ...)
        const int nB0 = n0 + cl_rank * (BN / CL);
        const int a_boxes = a_mn ? min(4, (p.M - m0 + 31) / 32) : 0;
        const int b_boxes = b_mn ? max(0, min(BN / CL / 32, (p.N - nB0 + 31) / 32)) : 0;
        const uint32_t a_bytes = a_mn ? a_boxes * 4096u
                                      : (p.a_mode == A_WINDOW_K ? uint32_t(p.win_box_q * p.win_box_p) * 128u
                                         : (p.a_mode == A_ROWS_K ? uint32_t(p.win_box_q) * 128u : uint32_t(L::A_BYTES)));
        uint32_t bytes = a_bytes + (b_mn ? b_boxes * 4096u : uint32_t(L::B_BYTES));
        if (CL > 1) {   // the leader's full barrier also counts what the peer (row block m0 + 128, second half of B) loads
          const int pa = a_mn ? min(4, (p.M - (m0 + kUmmaBM) + 31) / 32) : 0;
          const int pb = b_mn ? max(0, min(BN / CL / 32, (p.N - (n0 + BN / CL) + 31) / 32)) : 0;
          bytes += (a_mn ? pa * 4096u : uint32_t(L::A_BYTES)) + (b_mn ? pb * 4096u : uint32_t(L::B_BYTES));
        }
        // cta_group::2 loads signal the barrier at this offset in the leader CTA
        const uint32_t full0 = CL > 1 ? mapa_shared(smem_u32(full_bar), 0) : 0u;
        bool ok = true;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1, err)) { ok = false; break; }
          if (CL == 1 || cl_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], bytes);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          const uint32_t fullp = full0 + stage * 8;
          // ---- A operand
          // batched modes (see UmmaParams::batch_mode): image of this K block / tile and the K block inside the image
          const int bimg = p.batch_mode == 2 ? kb / p.kb_per_batch : (p.batch_mode == 1 ? tc.split : 0);
          const int bkb = p.batch_mode == 2 ? kb - bimg * p.kb_per_batch : kb;
          if (p.a_mode == A_TILED_K) {
            if (CL == 1) tma_load_2d(sA, &tmA, &full_bar[stage], bkb * kUmmaBK, m0 + bimg * p.a_batch_rows);
            else tma_load_2d_pair(sA, &tmA, fullp, kb * kUmmaBK, m0);
          } else if (p.a_mode == A_IM2COL_K) {
            const int tap = kb / p.c_chunks, c0 = (kb - tap * p.c_chunks) * kUmmaBK;
            if (CL == 1) tma_load_im2col_4d(sA, &tmA, &full_bar[stage], c0, a_w, a_h, a_n, p.tap_w[tap], p.tap_h[tap]);
            else tma_load_im2col_4d_pair(sA, &tmA, fullp, c0, a_w, a_h, a_n, p.tap_w[tap], p.tap_h[tap]);
          } else if (p.a_mode == A_ROWS_K) {
            const int i = kb / p.c_chunks;
            const int r = p.dg_r[(a_h + p.dg_ph) % p.dg_sh][i];
            const int prow = (a_h + p.dg_ph - r * p.dg_dh) / p.dg_sh;   // exact by construction of dg_r; may be < 0 or >= P (zero fill)
            tma_load_4d(sA, &tmA, &full_bar[stage], (kb - i * p.c_chunks) * kUmmaBK, 0, prow, a_n);
          } else if (p.a_mode == A_WINDOW_K) {
            tma_load_4d(sA, &tmA, &full_bar[stage], 0, a_w, a_h + p.tap_h[kb], a_n);   // kb = filter row
          } else {
            int pix = kb * kUmmaBK;
            if (p.b_mode == B_WINDOW_MN) {  // K blocks are 32-pixel runs inside one output row
              const int row = kb / p.win_qblocks, qb = kb - row * p.win_qblocks;
              pix = row * p.conv_Q + qb * 32;
            }
            for (int j = 0; j < a_boxes; ++j) {
              if (CL == 1) tma_load_2d(sA + j * 4096, &tmA, &full_bar[stage], m0 + 32 * j, pix);
              else tma_load_2d_pair(sA + j * 4096, &tmA, fullp, m0 + 32 * j, pix);
            }
          }
          // ---- B operand
          if (p.b_mode == B_TILED_K) {
            int k0 = kb * kUmmaBK;
            if (p.a_mode == A_IM2COL_K) {
              const int tap = kb / p.c_chunks;
              k0 = tap * p.b_tap_stride + (kb - tap * p.c_chunks) * kUmmaBK;
            } else if (p.a_mode == A_ROWS_K) {
              const int i = kb / p.c_chunks;
              k0 = p.dg_r[(a_h + p.dg_ph) % p.dg_sh][i] * p.b_tap_stride + (kb - i * p.c_chunks) * kUmmaBK;
            }
            if (CL == 1)
              tma_load_2d(sB, &tmB, &full_bar[stage], p.batch_mode == 2 ? bkb * kUmmaBK : k0, n0 + (p.batch_mode == 2 ? bimg * p.b_batch_rows : 0));
            else   // this CTA's half of the tile rows (the map's box is BN / 2 rows)
              tma_load_2d_pair(sB, &tmB, fullp, k0, nB0);
          } else if (p.b_mode == B_TILED_MN) {
            if (CL == 1) {
              for (int j = 0; j < b_boxes; ++j)
                tma_load_2d(sB + j * 4096, &tmB, &full_bar[stage], n0 + 32 * j, kb * kUmmaBK + (p.batch_mode == 1 ? bimg * p.b_batch_rows : 0));
            } else {
              for (int j = 0; j < b_boxes; ++j) tma_load_2d_pair(sB + j * 4096, &tmB, fullp, nB0 + 32 * j, kb * kUmmaBK);
            }
          } else if (p.b_mode == B_WINDOW_MN) {  // K index = pixel (32-pixel run of one output row), N index = window element
            const int row = kb / p.win_qblocks, qb = kb - row * p.win_qblocks;
            const int img = row / p.conv_P, pp = row - img * p.conv_P;
            for (int j = 0; j < b_boxes; ++j)   // one box per filter row folded into N (ntaps rows per tile group)
              tma_load_4d(sB + j * 4096, &tmB, &full_bar[stage], 0, qb * 32,
                          pp * p.stride_h + p.lower_h + p.tap_h[tc.tap * p.ntaps + j], img);
          } else {  // B_IM2COL_MN: K index = base pixel, N index = channel
            const int pix = kb * kUmmaBK;
            const int img = pix / pq, rem = pix - img * pq;
            const int pp = rem / p.conv_Q, qq = rem - pp * p.conv_Q;
            const int bw = p.lower_w + qq * p.stride_w, bh = p.lower_h + pp * p.stride_h;
            if (CL == 1) {
              for (int j = 0; j < b_boxes; ++j)
                tma_load_im2col_4d(sB + j * 4096, &tmB, &full_bar[stage], n0 + 32 * j, bw, bh, img, p.tap_w[tc.tap], p.tap_h[tc.tap]);
            } else {
              for (int j = 0; j < b_boxes; ++j)
                tma_load_im2col_4d_pair(sB + j * 4096, &tmB, fullp, nB0 + 32 * j, bw, bh, img, p.tap_w[tc.tap], p.tap_h[tc.tap]);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if ((CL == 1 || cl_rank == 0) && elect_one()) {   // pair: the leader issues for both CTAs
      const uint32_t idesc = make_idesc_tf32(kUmmaBM * CL, BN, a_mn ? 1 : 0, b_mn ? 1 : 0);
      // The issuing thread is on the critical path of small-N tiles (an N = 64 MMA is ~32 cycles of tensor work): descriptors are
      // formed once per launch and advanced by adding to their 14-bit address field (smem addresses < 256 KB: no carry out; in a
      // cluster the shared-window address of rank 1 carries the CTA rank in bit 24, which must not leak into the LBO field).
      const uint64_t a_desc0 = (a_mn ? make_smem_desc(0, 4096, 512, kSmemLayoutSw128Base32)
                                     : make_smem_desc(0, 16, 1024, kSmemLayoutSw128)) +
                               ((smem_u32(smem) & 0x3FFFFu) >> 4);
      const uint64_t b_desc0 = (b_mn ? make_smem_desc(0, 4096, 512, kSmemLayoutSw128Base32)
                                     : make_smem_desc(0, 16, 1024, kSmemLayoutSw128)) +
                               (((smem_u32(smem) & 0x3FFFFu) + L::A_BYTES) >> 4);
      const uint32_t a_kstep = a_mn ? (1024 >> 4) : (32 >> 4), b_kstep = b_mn ? (1024 >> 4) : (32 >> 4);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = w_first; tile < total_tiles; tile += w_step) {
        const TileCoord tc = CL > 1 ? decode_tile_cluster(p, tile, CL, cl_rank, tdiv) : decode_tile(p, tile, tdiv);
        int kb_begin, kb_end;
        tile_kb_range(p, tc, kb_begin, kb_end);
        bool ok = true;
        int kb = kb_begin;
        do {   // one pass per accumulator flush (a single one unless chain_kb limits the chain length)
        const int sub_begin = kb;
        const int sub_end = p.chain_kb > 0 ? min(kb_end, kb + p.chain_kb) : kb_end;
        if (!(CL == 1 ? mbar_wait(&tempty_bar[acc], acc_phase ^ 1, err) : mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1, err))) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (; kb < sub_end; ++kb) {
          if (!mbar_wait(&full_bar[stage], phase, err)) { ok = false; break; }
          tc_fence_after();
          const uint32_t st_off = static_cast<uint32_t>(stage * L::STAGE_BYTES) >> 4;
          const uint64_t da = a_desc0 + st_off, db = b_desc0 + st_off;
          if (CL == 1) {
            umma_tf32(d_tmem, da, db, idesc, kb > sub_begin ? 1u : 0u);
            umma_tf32(d_tmem, da + a_kstep, db + b_kstep, idesc, 1u);
            umma_tf32(d_tmem, da + 2 * a_kstep, db + 2 * b_kstep, idesc, 1u);
            umma_tf32(d_tmem, da + 3 * a_kstep, db + 3 * b_kstep, idesc, 1u);
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          } else {
            umma_tf32_pair(d_tmem, da, db, idesc, kb > sub_begin ? 1u : 0u);
            umma_tf32_pair(d_tmem, da + a_kstep, db + b_kstep, idesc, 1u);
            umma_tf32_pair(d_tmem, da + 2 * a_kstep, db + 2 * b_kstep, idesc, 1u);
            umma_tf32_pair(d_tmem, da + 3 * a_kstep, db + 3 * b_kstep, idesc, 1u);
            umma_commit_pair(&empty_bar[stage]);  // ... in both CTAs
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        if (CL == 1) umma_commit(&tfull_bar[acc]); else umma_commit_pair(&tfull_bar[acc]);  // accumulator complete -> epilogue(s)
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        } while (kb < kb_end);
        if (!ok) break;
      }
    }
  } else {
    epilogue_role<BN, OLD, OLD>(p, smem, L::EPI_OFFSET, tfull_bar, tempty_bar, tmem_base, warp, lane, err, false, OLD ? smem + L::OLD_OFFSET : nullptr,
                           2, 1, CL, CL > 1);
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // nobody leaves while the leader's MMAs may still read this CTA's smem or a peer signals its barriers
  if (warp == 1) {
    tc_fence_after();
    if (CL == 1) tmem_dealloc(tmem_base, L::TMEM_COLS); else tmem_dealloc_pair(tmem_base, L::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------- halo-reuse conv kernel
// Stride-1 R x S convolutions (fprop, and dgrad written as a conv over dY with the flipped filter).  The implicit-GEMM kernel
// above fetches the activation tile once per filter tap (9x for 3x3) and is bound by L2->SM traffic on the 64/128-channel
// layers.  Here one TMA box lands a raster of (tp + R - 1) x Wr input pixels (Wr = W + 2*pad: halo columns come in as TMA
// zero fill) for a 32-channel chunk, each pixel one 128-byte row of the K-major SWIZZLE_128B layout; the M tile is the raster
// positions of tp output rows, and tap (r, s) is the SAME smem data read through a descriptor that starts r*Wr + s rows later
// (the hardware swizzle is a function of the absolute smem address, so a start that is 128- but not 1024-byte aligned is
// legal: measured in round 1 with a descriptor-shift probe build, git history: tools/probe_desc_shift.py).  Filter tiles stream through their own ring, or stay resident for the
// whole persistent CTA when the filter fits (64 -> 64 channels).  Output positions in halo columns are computed and dropped.
// CL > 1: thread-block cluster of CL CTAs working on CL consecutive row blocks of the same column block in lockstep.  The
// streamed filter tiles are what bounds the >= 128-channel 3x3 layers (a 128-pixel tile re-reads the whole [BN][taps*C] filter
// slice from L2: 0.6-4.7 MB per tile, 8-14 TB/s of L2->SM traffic at the measured speeds), so each CTA fetches 1/CL of every
// filter tile and TMA-multicasts it into all CL shared memories; a ring slot is refilled once the MMA warps of ALL CTAs have
// released it (tcgen05.commit multicast onto every CTA's b_empty barrier).  Rasters stay private.
// PAIR (with CL = 2): the two CTAs form a tcgen05 CTA pair instead (cta_group::2, see umma_kernel): each CTA lands its own raster and
// HALF of every filter tile (BN / 2 rows) in its own shared memory, the leader issues M = 256 MMAs for both row blocks, commits are
// multicast, both epilogues release TMEM on the leader's barrier.  One instruction then covers two row blocks, which is what the
// N <= 128 layers need: a cta_group::1 MMA has an issue floor of 77 cycles whatever N is (41 % of the tensor rate at N = 64), a
// cta_group::2 MMA one of 46 cycles for twice the rows (70 % at N = 64, 100 % at N = 128; tools/probes/mma_rate_2cta.cu).
constexpr int kHaloMaxB = 24;
template <int BN, int CL = 1, bool PAIR = false>
__global__ void __launch_bounds__(192, 1)
halo_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ UmmaParams p) {
  static_assert(!PAIR || CL == 2, "a CTA pair is a cluster of two");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * 128;   // filter tile bytes in THIS CTA's shared memory
  constexpr int TMEM_COLS = PAIR ? 512 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.halo_slots * p.halo_slot_bytes;
  const int epi_off = p.halo_slots * p.halo_slot_bytes + p.halo_b_stages * B_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + epi_off + 4 * 4096 + 4 * 3 * BN * 4);   // after the epilogue staging + statistics
  uint64_t* a_empty = a_full + 4;
  uint64_t* b_full = a_empty + 4;
  uint64_t* b_empty = b_full + kHaloMaxB;
  uint64_t* tfull_bar = b_empty + kHaloMaxB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile int* err = p.err_flag;
  if (warp == 0) {
    if (elect_one()) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < kHaloMaxB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], PAIR ? 1 : CL); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], PAIR ? 8 : 4); }
      fence_barrier_init();
    }
    __syncwarp();
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  if (p.halo_stack > 1) {   // the rows below the last image of every slot are its bottom padding: zero, and no TMA load ever writes them
    for (int sl = 0; sl < p.halo_slots; ++sl) {
      uint4* z = reinterpret_cast<uint4*>(sA + sl * p.halo_slot_bytes + p.halo_raster_bytes);
      const int n16 = (p.halo_slot_bytes - p.halo_raster_bytes) / 16;
      for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's (async-proxy) operand reads
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers' barriers are initialised before anyone multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int taps = p.ntaps, chunks = p.c_chunks;
  const bool resident = p.halo_b_resident != 0;
  // tile walk: group q -> column block q % n_tiles, row block (q / n_tiles) * CL + rank (past the end: a dummy, loads are zero fill)
  const int cl_rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cl_id = blockIdx.x / CL, n_cl = gridDim.x / CL;
  const int total_q = ((p.m_tiles + CL - 1) / CL) * p.n_tiles;
  constexpr uint16_t kClMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (warp == 0) {
    if (elect_one()) {
      int ai = 0, bi = 0;
      uint32_t aph = 0, bph = 0;
      const bool lead = !PAIR || cl_rank == 0;   // pair: both CTAs' loads are counted on the leader's barriers
      const uint32_t afull0 = PAIR ? mapa_shared(smem_u32(a_full), 0) : 0u, bfull0 = PAIR ? mapa_shared(smem_u32(b_full), 0) : 0u;
      constexpr uint32_t kShare = PAIR ? 2u : 1u;
      if (resident) {  // the whole [BN][taps*C] filter slice, once per CTA (n_tiles == 1); pair: this CTA's half of the rows
        if (lead) mbar_arrive_expect_tx(&b_full[0], kShare * static_cast<uint32_t>(taps * chunks) * B_BYTES);
        for (int c = 0; c < chunks; ++c)
          for (int t = 0; t < taps; ++t) {
            if (PAIR) tma_load_2d_pair(sB + (c * taps + t) * B_BYTES, &tmB, bfull0, t * p.b_tap_stride + c * kUmmaBK, cl_rank * (BN / 2));
            else tma_load_2d(sB + (c * taps + t) * B_BYTES, &tmB, &b_full[0], t * p.b_tap_stride + c * kUmmaBK, 0);
          }
      }
      bool ok = true;
      for (int q = cl_id; q < total_q && ok; q += n_cl) {
        const int n_blk = q % p.n_tiles, m_blk = (q / p.n_tiles) * CL + cl_rank;
        const bool stacked = p.halo_stack > 1;
        const int img = stacked ? m_blk * p.halo_stack : m_blk / p.win_p_tiles, pt = stacked ? 0 : m_blk - img * p.win_p_tiles;
        const int h0 = pt * p.win_box_p * p.halo_stride;
        for (int c = 0; c < chunks && ok; ++c) {
          if (!mbar_wait(&a_empty[ai], aph ^ 1, err)) { ok = false; break; }
          if (lead) mbar_arrive_expect_tx(&a_full[ai], kShare * static_cast<uint32_t>(p.halo_raster_bytes));
          // one raster (stride 1), one per input parity class (stride 2), or one per image of a stacked tile (an image past the end
          // of the batch arrives as zero fill)
          for (int pl = 0; pl < p.halo_planes; ++pl) {
            uint8_t* dst = sA + ai * p.halo_slot_bytes + pl * p.halo_plane_bytes;
            const int im = stacked ? img + pl : img;
            if (PAIR) tma_load_4d_pair(dst, &tmA, afull0 + ai * 8, c * kUmmaBK, p.halo_dw[pl], h0 + p.halo_dh[pl], im);
            else tma_load_4d(dst, &tmA, &a_full[ai], c * kUmmaBK, p.halo_dw[pl], h0 + p.halo_dh[pl], im);
          }
          if (++ai == p.halo_slots) { ai = 0; aph ^= 1; }
          if (!resident) {
            for (int t = 0; t < taps; ++t) {
              if (!mbar_wait(&b_empty[bi], bph ^ 1, err)) { ok = false; break; }
              if (lead) mbar_arrive_expect_tx(&b_full[bi], kShare * B_BYTES);
              if (PAIR)   // this CTA's half of the tile rows, into its own shared memory only
                tma_load_2d_pair(sB + bi * B_BYTES, &tmB, bfull0 + bi * 8, t * p.b_tap_stride + c * kUmmaBK, n_blk * BN + cl_rank * (BN / 2));
              else if (CL == 1)
                tma_load_2d(sB + bi * B_BYTES, &tmB, &b_full[bi], t * p.b_tap_stride + c * kUmmaBK, n_blk * BN);
              else   // this CTA's 1/CL of the tile rows (the map's box is BN / CL rows), delivered to every CTA of the cluster
                tma_load_2d_multicast(sB + bi * B_BYTES + cl_rank * (B_BYTES / CL), &tmB, &b_full[bi], t * p.b_tap_stride + c * kUmmaBK,
                                      n_blk * BN + cl_rank * (BN / CL), kClMask);
              if (++bi == p.halo_b_stages) { bi = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if ((!PAIR || cl_rank == 0) && elect_one()) {   // pair: the leader issues for both CTAs
      const uint32_t idesc = make_idesc_tf32(PAIR ? 2 * kUmmaBM : kUmmaBM, BN, 0, 0);
      int ai = 0, bi = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_phase = 0;
      bool ok = true;
      if (resident) ok = mbar_wait(&b_full[0], 0, err);
      const uint32_t a_addr0 = smem_u32(sA) & 0x3FFFFu;   // (rank bits of a cluster CTA's shared-window address dropped)
      // descriptors advanced by integer adds on their address field (see umma_kernel): ~6 instructions per MMA instead of ~30
      const uint64_t a_desc0 = make_smem_desc(0, 16, 1024, kSmemLayoutSw128);
      const uint64_t b_desc0 = make_smem_desc(smem_u32(sB), 16, 1024, kSmemLayoutSw128);
      uint32_t tapoff[9];   // descriptor offset of tap t (whole 128-byte raster rows), in registers for the unrolled 3x3 path
#pragma unroll
      for (int t = 0; t < 9; ++t) tapoff[t] = static_cast<uint32_t>(p.tap_w[t]) * 8u;
      for (int q = cl_id; q < total_q && ok; q += n_cl) {
        for (int c = 0; c < chunks && ok;) {   // one pass per accumulator flush (halo_chain channel chunks each)
        const int c_begin = c;
        const int c_end = p.halo_chain > 0 ? min(chunks, c + p.halo_chain) : chunks;
        if (!(PAIR ? mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1, err) : mbar_wait(&tempty_bar[acc], acc_phase ^ 1, err))) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (; c < c_end && ok; ++c) {
          if (!mbar_wait(&a_full[ai], aph, err)) { ok = false; break; }
          tc_fence_after();
          const uint64_t a_slot = a_desc0 + ((a_addr0 + ai * p.halo_slot_bytes) >> 4);
          // The issuing thread is the bottleneck of the small-N layers, not the tensor pipe: one thread needs ~45 cycles per
          // tcgen05.mma when the descriptors are ready-made registers, but ~100 when every tap goes through a constant-memory table
          // lookup, R2UR moves and loop control (tools/probes/mma_rate.cu: 98 -> 48 clk/MMA at N = 64, 98 -> 64 = the full tensor rate
          // at N = 128).  3x3 filters therefore take a fully unrolled path: tap offsets held in registers, 36 MMAs back to back per
          // channel chunk when the filter is resident.
          auto mma4 = [&](uint64_t da, uint64_t db, uint32_t first_acc) {
            if (PAIR) {
              umma_tf32_pair(d_tmem, da, db, idesc, first_acc);
              umma_tf32_pair(d_tmem, da + 2, db + 2, idesc, 1u);
              umma_tf32_pair(d_tmem, da + 4, db + 4, idesc, 1u);
              umma_tf32_pair(d_tmem, da + 6, db + 6, idesc, 1u);
            } else {
              umma_tf32(d_tmem, da, db, idesc, first_acc);
              umma_tf32(d_tmem, da + 2, db + 2, idesc, 1u);
              umma_tf32(d_tmem, da + 4, db + 4, idesc, 1u);
              umma_tf32(d_tmem, da + 6, db + 6, idesc, 1u);
            }
          };
          auto b_release = [&]() {
            if (PAIR) umma_commit_pair(&b_empty[bi]);
            else if (CL == 1) umma_commit(&b_empty[bi]);
            else umma_commit_multicast(&b_empty[bi], kClMask);
            if (++bi == p.halo_b_stages) { bi = 0; bph ^= 1; }
          };
          // NT taps fully unrolled (offsets in registers): 9 = 3x3 filters, 4 / 2 = the multi-tap parity classes of a stride-2 dgrad
          auto taps_unrolled = [&](auto tag) -> bool {
            constexpr int NT = decltype(tag)::value;
            if (resident) {
              uint64_t db = b_desc0 + static_cast<uint32_t>(c * NT * (B_BYTES >> 4));
              mma4(a_slot + tapoff[0], db, c > c_begin ? 1u : 0u);
#pragma unroll
              for (int t = 1; t < NT; ++t) {
                db += B_BYTES >> 4;
                mma4(a_slot + tapoff[t], db, 1u);
              }
            } else {
#pragma unroll
              for (int t = 0; t < NT; ++t) {
                if (!mbar_wait(&b_full[bi], bph, err)) return false;
                tc_fence_after();
                mma4(a_slot + tapoff[t], b_desc0 + static_cast<uint32_t>(bi * (B_BYTES >> 4)), (c > c_begin || t > 0) ? 1u : 0u);
                b_release();
              }
            }
            return true;
          };
          if (taps == 9) {
            if (!taps_unrolled(IntTag<9>{})) { ok = false; break; }
          } else if (taps == 4) {
            if (!taps_unrolled(IntTag<4>{})) { ok = false; break; }
          } else if (taps == 2) {
            if (!taps_unrolled(IntTag<2>{})) { ok = false; break; }
          } else {
#pragma unroll 1
            for (int t = 0; t < taps; ++t) {
              uint64_t db;
              if (resident) {
                db = b_desc0 + static_cast<uint32_t>((c * taps + t) * (B_BYTES >> 4));
              } else {
                if (!mbar_wait(&b_full[bi], bph, err)) { ok = false; break; }
                tc_fence_after();
                db = b_desc0 + static_cast<uint32_t>(bi * (B_BYTES >> 4));
              }
              mma4(a_slot + static_cast<uint32_t>(p.tap_w[t]) * 8u, db, (c > c_begin || t > 0) ? 1u : 0u);   // tap = whole 128-byte rows
              if (!resident) b_release();
            }
          }
          if (PAIR) umma_commit_pair(&a_empty[ai]); else umma_commit(&a_empty[ai]);
          if (++ai == p.halo_slots) { ai = 0; aph ^= 1; }
        }
        if (!ok) break;
        if (PAIR) umma_commit_pair(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        }
        if (!ok) break;
      }
    }
  } else {
    epilogue_role<BN, true>(p, smem, epi_off, tfull_bar, tempty_bar, tmem_base, warp, lane, err, true, nullptr, 2, 1, CL, PAIR);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast into this CTA or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------- multi-row stem fprop
// Small-C (C <= 4) forward conv, several output rows per tile.  The sliding-window box of an input row (Q windows x 32 floats,
// see "small-C conv" below) depends only on the input row, so with TP output rows per tile one box feeds every (output row,
// filter row) pair that reads it: (TP-1)*sh + R boxes per TP rows instead of R per row (7x7/s2, TP = 4: 3.25 instead of 7), and
// the packed filter ([K][R*32], R tiles of BN x 32) stays resident in smem instead of being re-fetched per tile.  The
// one-row-per-tile form was bound by exactly that L2->SM traffic (154 KB per 28 KB of output).  TP accumulators per tile, two
// tile sets in TMEM (2*TP*BN <= 512 columns); the shared epilogue drains them as consecutive OUT_WINDOW tiles.
constexpr int kStemMaxStages = 8;
template <int BN>
__global__ void __launch_bounds__(192, 1)
stem_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ UmmaParams p) {
  constexpr int TP = 512 / (2 * BN);
  constexpr int B_TILE = BN * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int R = p.ntaps;                       // filter rows
  uint8_t* sB = smem;                          // R resident filter tiles
  uint8_t* sA = sB + R * B_TILE;               // halo_slots stages x 16 KB window boxes
  const int epi_off = R * B_TILE + p.halo_slots * 16384;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + epi_off + 4 * 4096 + 4 * 3 * BN * 4);
  uint64_t* empty_bar = full_bar + kStemMaxStages;
  uint64_t* b_bar = empty_bar + kStemMaxStages;
  uint64_t* tfull_bar = b_bar + 1;             // 2 * TP
  uint64_t* tempty_bar = tfull_bar + 2 * TP;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2 * TP);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  volatile int* err = p.err_flag;
  if (warp == 0) {
    if (elect_one()) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < kStemMaxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      mbar_init(b_bar, 1);
      for (int i = 0; i < 2 * TP; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
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
  // tiles of the epilogue = (image, output row) with rows padded to a multiple of TP per image; a scheduling step = TP rows
  const int p_groups = p.win_p_tiles / TP;                 // row groups per image
  const int groups = (p.m_tiles / TP);                     // = N * p_groups
  const int sh = p.stride_h, dh = p.dg_dh;                 // filter-row dilation

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(b_bar, static_cast<uint32_t>(R) * B_TILE);
      for (int r = 0; r < R; ++r) tma_load_2d(sB + r * B_TILE, &tmB, b_bar, r * 32, 0);
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int g = blockIdx.x; g < groups && ok; g += gridDim.x) {
        const int img = g / p_groups, p0 = (g - img * p_groups) * TP;
        const int ih_lo = p0 * sh + p.lower_h, ih_hi = (p0 + TP - 1) * sh + p.lower_h + (R - 1) * dh;
        for (int ih = ih_lo; ih <= ih_hi; ++ih) {
          if (!mbar_wait(&empty_bar[stage], phase ^ 1, err)) { ok = false; break; }
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.win_box_q) * 128u);
          tma_load_4d(sA + stage * 16384, &tmA, &full_bar[stage], 0, 0, ih, img);   // rows outside the image: zero fill
          if (++stage == p.halo_slots) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(kUmmaBM, BN, 0, 0);
      const uint32_t a0 = smem_u32(sA);
      const uint64_t db_first = make_smem_desc(smem_u32(sB), 16, 1024, kSmemLayoutSw128);
      const uint64_t da_base = make_smem_desc(a0, 16, 1024, kSmemLayoutSw128);   // + stage * 1024 (16 KB slots, 16-byte units)
      int stage = 0, set = 0;
      uint32_t phase = 0, set_phase = 0;
      bool ok = mbar_wait(b_bar, 0, err);
      for (int g = blockIdx.x; g < groups && ok; g += gridDim.x) {
        const int img = g / p_groups, p0 = (g - img * p_groups) * TP;
        (void)img;
        const int ih_lo = p0 * sh + p.lower_h, ih_hi = (p0 + TP - 1) * sh + p.lower_h + (R - 1) * dh;
        for (int pl = 0; pl < TP && ok; ++pl)
          if (!mbar_wait(&tempty_bar[set * TP + pl], set_phase ^ 1, err)) ok = false;
        if (!ok) break;
        tc_fence_after();
        uint32_t started = 0;
        for (int ih = ih_lo; ih <= ih_hi; ++ih) {
          if (!mbar_wait(&full_bar[stage], phase, err)) { ok = false; break; }
          tc_fence_after();
          const uint64_t da0 = da_base + static_cast<uint32_t>(stage * (16384 >> 4));
          const int t0 = ih - ih_lo;                       // = r * dil_h for output row p0 (pl = 0); decreases by sh per row
          if (dh == 1) {
            // (the issuing thread bounds these N = BN <= 128 MMAs: unrolled over the TP rows, the accumulate flag is "not filter
            // row 0" because every output row meets its filter rows in ascending order)
#pragma unroll
            for (int pl = 0; pl < TP; ++pl) {
              const int r = t0 - pl * sh;
              if (r >= 0 && r < R) {
                const uint64_t db0 = db_first + static_cast<uint32_t>(r * (B_TILE >> 4));
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((set * TP + pl) * BN);
                umma_tf32(d_tmem, da0, db0, idesc, r > 0 ? 1u : 0u);
                umma_tf32(d_tmem, da0 + 2, db0 + 2, idesc, 1u);
                umma_tf32(d_tmem, da0 + 4, db0 + 4, idesc, 1u);
                umma_tf32(d_tmem, da0 + 6, db0 + 6, idesc, 1u);
              }
            }
          } else {
            int t = t0;
            for (int pl = 0; pl < TP; ++pl, t -= sh) {
              if (t < 0 || t % dh != 0) continue;
              const int r = t / dh;
              if (r >= R) continue;
              const uint64_t db0 = db_first + static_cast<uint64_t>(r * (B_TILE >> 4));
              const uint32_t d_tmem = tmem_base + (set * TP + pl) * BN;
              umma_tf32(d_tmem, da0, db0, idesc, (started >> pl) & 1u);
              umma_tf32(d_tmem, da0 + 2, db0 + 2, idesc, 1u);
              umma_tf32(d_tmem, da0 + 4, db0 + 4, idesc, 1u);
              umma_tf32(d_tmem, da0 + 6, db0 + 6, idesc, 1u);
              started |= 1u << pl;
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.halo_slots) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        for (int pl = 0; pl < TP; ++pl) umma_commit(&tfull_bar[set * TP + pl]);
        set ^= 1;
        if (set == 0) set_phase ^= 1;
      }
    }
  } else {
    epilogue_role<BN>(p, smem, epi_off, tfull_bar, tempty_bar, tmem_base, warp, lane, err, false, nullptr, 2 * TP, TP);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[i] = alpha * sum_s partial[s][i] + bias[col] + beta * out[i]   (deterministic split-K reduction)
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, long long rows,
                                     long long cols, long long ldo, long long split_stride, int splits, float alpha,
                                     float beta, const float* __restrict__ bias) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    // four independent partial sums, 8 loads in flight (the loop used to be one dependent L2 round trip per split)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int s = 0;
#pragma unroll 2
    for (; s + 4 <= splits; s += 4) {
      const float v0 = partial[(s + 0) * split_stride + i], v1 = partial[(s + 1) * split_stride + i];
      const float v2 = partial[(s + 2) * split_stride + i], v3 = partial[(s + 3) * split_stride + i];
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    for (; s < splits; ++s) a0 += partial[s * split_stride + i];
    const float acc = (a0 + a1) + (a2 + a3);
    const long long r = i / cols, c = i - r * cols;
    float v = alpha * acc;
    if (bias) v += bias[c];
    if (beta != 0.f) v += beta * out[r * ldo + c];
    out[r * ldo + c] = v;
  }
}

// Wt[c][t][k] = W[k][taps[t]][c]  (KRSC source); used by dgrad so that the filter is K-major for the GEMM.
struct TapList {
  int rs[kUmmaMaxTaps];  // r*S+s of each tap, in GEMM-K order
};
__global__ void dgrad_filter_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int RS, int C, int ntaps,
                                    const TapList taps) {
  const long long total = static_cast<long long>(C) * ntaps * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const long long r = i / K;
    const int t = static_cast<int>(r % ntaps), c = static_cast<int>(r / ntaps);
    wt[i] = w[(static_cast<long long>(k) * RS + taps.rs[t]) * C + c];
  }
}

// All parity classes of a strided dgrad in one launch: their tap sets partition the R*S taps, so the class filters ([C][ntaps][K]
// each, back to back in the workspace) are one permutation of the filter.
struct ClassTable {
  int n;
  long long start[17];   // element offset of class i's filter in wt (start[n] = total)
  int ntaps[16], tap_base[16];
  int rs[kUmmaMaxTaps];  // r*S+s of tap (tap_base[i] + t) of class i
};
__global__ void dgrad_filter_classes_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int RS, int C, const ClassTable tb) {
  const long long total = tb.start[tb.n];
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    int ci = 0;
    while (ci + 1 < tb.n && i >= tb.start[ci + 1]) ++ci;
    const long long j = i - tb.start[ci];
    const int k = static_cast<int>(j % K);
    const long long r = j / K;
    const int t = static_cast<int>(r % tb.ntaps[ci]), c = static_cast<int>(r / tb.ntaps[ci]);
    wt[i] = w[(static_cast<long long>(k) * RS + tb.rs[tb.tap_base[ci] + t]) * C + c];
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static CUtensorMapDataType operand_dtype() {
  // TFLOAT32 makes the TMA unit round fp32 -> tf32 while loading (instead of the MMA truncating the low 13
  // mantissa bits); ZENU_B200_TMA_F32=1 selects raw FLOAT32 loads for A/B comparison of the two behaviours.
  static int raw = -1;
  if (raw < 0) {
    const char* e = getenv("ZENU_B200_TMA_F32");
    raw = (e && e[0] == '1') ? 1 : 0;
  }
  return raw ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
}

// 2-D row-major matrix [outer][inner], box (box_inner <= 32 fp32 = 128 B swizzle span, box_outer <= 256).
static int make_map_2d(zb_ctx* ctx, CUtensorMap* map, const float* base, long long inner, long long outer,
                       long long pitch_elems, int box_inner, int box_outer, bool mn_major = false) {
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand must be 16-byte aligned");
  ZB_REQUIRE((pitch_elems * 4) % 16 == 0, "TMA operand row pitch must be a multiple of 16 bytes (got %lld elems)", pitch_elems);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch_elems) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode_tiled(map, operand_dtype(), 2, const_cast<float*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): inner=%lld outer=%lld pitch=%lld box=%dx%d", int(r), inner, outer,
                   pitch_elems, box_inner, box_outer);
    return ZB_ERR_CUDA;
  }
  return ZB_OK;
}

// NHWC activation tensor seen as (C, W, H, N) for im2col loads: `pixels` base pixels x 32 channels per load.
static int make_map_im2col(zb_ctx* ctx, CUtensorMap* map, const float* base, long long N, long long H, long long W,
                           long long C, int lower_w, int lower_h, int upper_w, int upper_h, int stride_w, int stride_h,
                           int pixels, bool mn_major = false) {
  ZB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand must be 16-byte aligned");
  ZB_REQUIRE(C % 4 == 0, "im2col TMA needs C %% 4 == 0 (got %lld)", C);
  ZB_REQUIRE(lower_w >= -128 && lower_w <= 127 && lower_h >= -128 && lower_h <= 127 && upper_w >= -128 &&
                 upper_w <= 127 && upper_h >= -128 && upper_h <= 127,
             "im2col corner out of the 8-bit range");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 4, static_cast<cuuint64_t>(W) * C * 4,
                           static_cast<cuuint64_t>(H) * W * C * 4};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride_w), static_cast<cuuint32_t>(stride_h), 1};
  CUresult r = ctx->encode_im2col(map, operand_dtype(), 4, const_cast<float*>(base), dims, strides, lower, upper,
                                  /*channelsPerPixel=*/32, /*pixelsPerColumn=*/static_cast<cuuint32_t>(pixels), estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeIm2col failed (%d): NHWC=%lldx%lldx%lldx%lld lower=(%d,%d) upper=(%d,%d) stride=(%d,%d)",
                   int(r), N, H, W, C, lower_w, lower_h, upper_w, upper_h, stride_w, stride_h);
    return ZB_ERR_CUDA;
  }
  // Same small-tensor driver workaround CUTLASS applies (copy_traits_sm90_im2col.hpp): for drivers <= 13.1 and
  // tensors under 128 KiB, bit 21 of the second descriptor word must be cleared.
  if (ctx->driver_version <= 13010 && N * H * W * C * 4 < 131072) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
  return ZB_OK;
}

// Launches that accumulate BatchNorm statistics use a grid that is a multiple of n_tiles (see UmmaParams::stat_partial).
static int stat_grid(zb_ctx* ctx, int tiles, int n_tiles) {
  const int g = std::min(tiles, ctx->sm_count);
  return std::max(n_tiles, g / n_tiles * n_tiles);
}
struct StatRequest {   // optional fused-BN-statistics request of a conv fprop
  const float* shift = nullptr;
  float* partial = nullptr;
  int* rows = nullptr;   // out: rows of [2][K] partials written (0 = the planner could not fuse them)
};
// CTA pairs (umma_kernel CL = 2, see the kernel): for the wide tiles, whose operand traffic into shared memory is what bounds them.
// Measured (tools/yardstick_gemm.py, tools/bench_conv.py): 10-16 % faster from K = 256 up at BN = 256 (50176 x 1024 x 256 0.073 -> 0.061 ms,
// 50176 x 512 x 1024 0.106 -> 0.093), still slower on the shortest-K, purely output-bound problems (the leader waits for the epilogues
// of BOTH CTAs before it reuses a TMEM buffer: 200704 x 512 x 128 0.095 -> 0.103, 802816 x 256 x 64 0.192 -> 0.198) and neutral to
// slightly slower at BN = 128, so only the former run on pairs.
static int pair_cl(const UmmaParams& p, int bn) {
  static int mode = -1;   // ZENU_B200_PAIR: 0 = never, 1 = BN = 256 with >= 8 K blocks per tile (default), 2 = whenever legal (BN >= 128)
  if (mode < 0) {
    const char* e = getenv("ZENU_B200_PAIR");
    mode = e ? atoi(e) : 1;
  }
  if (mode == 0 || bn < 128 || p.batch_mode != 0) return 1;
  if (mode == 1 && (bn < 256 || p.kb_per_split < 8)) return 1;
  if (p.m_tiles < 2 || (p.m_tiles & 1)) return 1;
  if (p.a_mode != A_TILED_K && p.a_mode != A_IM2COL_K && p.a_mode != A_TILED_MN) return 1;
  if (p.b_mode != B_TILED_K && p.b_mode != B_TILED_MN && p.b_mode != B_IM2COL_MN) return 1;
  if (p.b_mode == B_TILED_K && p.hb_base == nullptr) return 1;
  if (p.out_mode != OUT_ROWS && p.out_mode != OUT_SCATTER) return 1;
  return 2;
}
// persistent grid of a launch: one CTA per SM (or per tile), a multiple of n_tiles (of cl * n_tiles) with fused statistics
static int launch_grid(zb_ctx* ctx, const UmmaParams& p, int cl) {
  const int tiles = p.m_tiles * p.n_tiles * p.tap_tiles * p.splits;
  if (cl == 1) return p.stat_partial ? stat_grid(ctx, tiles, p.n_tiles) : std::min(tiles, ctx->sm_count);
  int clusters = std::min(tiles / cl, ctx->sm_count / cl);
  if (p.stat_partial) clusters = std::max(p.n_tiles, clusters / p.n_tiles * p.n_tiles);
  return clusters * cl;
}
static void stat_attach(zb_ctx* ctx, UmmaParams& p, const StatRequest* st, int tiles, long long kout, const float* y, const float* bias,
                        int pair_bn = 0) {
  if (st == nullptr || st->partial == nullptr) return;
  *st->rows = 0;
  if (kout % 32 != 0 || p.chain_kb > 0 || p.splits > 1 || (reinterpret_cast<uintptr_t>(y) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(bias) & 15) != 0 || (reinterpret_cast<uintptr_t>(st->shift) & 15) != 0 || p.n_tiles > ctx->sm_count)
    return;
  p.stat_partial = st->partial;
  p.stat_shift = st->shift;
  *st->rows = (pair_bn > 0 ? launch_grid(ctx, p, pair_cl(p, pair_bn)) : stat_grid(ctx, tiles, p.n_tiles)) / p.n_tiles * 4;
}

// plan-trace segment of one umma_kernel launch ('~' = value depends on the batch size, not on the kernel variant)
static void note_umma(const UmmaParams& p, int bn, int stages, bool old, int cl, int grid) {
  if (tl_plan == nullptr) return;
  plan_note("umma<bn=%d,stages=%d,old=%d,cl=%d> a_mode=%d b_mode=%d out_mode=%d n_tiles=%d tap_tiles=%d ntaps=%d splitk=%d beta=%d bias=%d stats=%d chain=%d batch_mode=%d "
            "~m_tiles=%d ~splits=%d ~kb_per_split=%d ~grid=%d;",
            bn, stages, old ? 1 : 0, cl, p.a_mode, p.b_mode, p.out_mode, p.n_tiles, p.tap_tiles, p.ntaps, (p.splits > 1 && p.batch_mode != 1) ? 1 : 0,
            p.beta != 0.f ? 1 : 0, p.bias != nullptr ? 1 : 0, p.stat_partial != nullptr ? 1 : 0, p.chain_kb, p.batch_mode, p.m_tiles, p.splits,
            p.kb_per_split, grid);
}

template <int BN, int STAGES, bool OLD = false>
static int launch_cfg(zb_ctx* ctx, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p) {
  using L = UmmaSmem<BN, STAGES, OLD>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static SmemOptIn opt_in;
  { const int rc = smem_opt_in(ctx, opt_in, umma_kernel<BN, STAGES, OLD>, L::TOTAL); if (rc != ZB_OK) return rc; }
  const int grid = launch_grid(ctx, p, 1);
  note_umma(p, BN, STAGES, OLD, 1, grid);
  if (plan_dry()) return ZB_OK;
  // algorithmic FLOPs of this launch: 2 * M * N * K over all taps (K counted in 32-wide blocks as issued)
  prof_begin(ctx, PROF_TENSOR);
  umma_kernel<BN, STAGES, OLD><<<grid, 192, L::TOTAL, ctx->stream>>>(a, b, p);
  prof_end(ctx, PROF_TENSOR, p.prof_flops);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// the same launch on CTA pairs; a K-major B map is re-encoded with a box of BN / 2 rows (each CTA fetches its half)
template <int BN, int STAGES, bool OLD = false>
static int launch_cfg_pair(zb_ctx* ctx, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p) {
  using L = UmmaSmem<BN, STAGES, OLD, 2>;
  static_assert(L::TOTAL <= 227 * 1024, "shared memory budget");
  static SmemOptIn opt_in;
  { const int rc = smem_opt_in(ctx, opt_in, umma_kernel<BN, STAGES, OLD, 2>, L::TOTAL); if (rc != ZB_OK) return rc; }
  CUtensorMap b2 = b;
  if (p.b_mode == B_TILED_K) {
    const int rc = make_map_2d(ctx, &b2, p.hb_base, p.hb_inner, p.hb_outer, p.hb_pitch, 32, BN / 2);
    if (rc != ZB_OK) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(launch_grid(ctx, p, 2));
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  note_umma(p, BN, STAGES, OLD, 2, static_cast<int>(cfg.gridDim.x));
  if (plan_dry()) return ZB_OK;
  prof_begin(ctx, PROF_TENSOR);
  ZB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, umma_kernel<BN, STAGES, OLD, 2>, a, b2, p));
  prof_end(ctx, PROF_TENSOR, p.prof_flops);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// K-major B operand [outer = N][inner = K]: records the matrix in p for launch_cfg_pair
static int make_map_bk(zb_ctx* ctx, CUtensorMap* map, UmmaParams& p, const float* base, long long inner, long long outer,
                       long long pitch_elems, int bn) {
  p.hb_base = base; p.hb_inner = inner; p.hb_outer = outer; p.hb_pitch = pitch_elems;
  return make_map_2d(ctx, map, base, inner, outer, pitch_elems, 32, bn);
}

static int pick_bn(long long n) { return n <= 32 ? 32 : (n <= 64 ? 64 : (n <= 128 ? 128 : 256)); }

static int umma_launch(zb_ctx* ctx, int bn, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p) {
  // launches whose epilogue reads the old output tile (beta != 0, chained 3xTF32 flushes): a shallower operand ring makes
  // room for the deep old-tile prefetch buffers (these are short-K, output-bound problems)
  const bool deep_old = (p.beta != 0.f || p.chain_kb > 0) && p.out_mode != OUT_WDGRAD && !ZB_ENV_FLAG("ZENU_B200_NO_DEEP_BETA");
  if (p.old_bits != nullptr && (p.stat_partial != nullptr || (p.ldd & 31) != 0)) {
    set_last_error("umma: masked accumulate needs ldd % 32 == 0 and no fused statistics");
    return ZB_ERR_UNSUPPORTED;
  }
  if (p.old_bits != nullptr && !deep_old) { set_last_error("umma: masked accumulate needs the deep-beta kernel variants"); return ZB_ERR_UNSUPPORTED; }
  if (pair_cl(p, bn) == 2) {
    if (deep_old) return bn == 128 ? launch_cfg_pair<128, 6, true>(ctx, a, b, p) : launch_cfg_pair<256, 4, true>(ctx, a, b, p);
    return bn == 128 ? launch_cfg_pair<128, 8>(ctx, a, b, p) : launch_cfg_pair<256, 6>(ctx, a, b, p);
  }
  if (deep_old) {
    switch (bn) {
      case 32: return launch_cfg<32, 7, true>(ctx, a, b, p);
      case 64: return launch_cfg<64, 6, true>(ctx, a, b, p);
      case 128: return launch_cfg<128, 4, true>(ctx, a, b, p);
      default: return launch_cfg<256, 3, true>(ctx, a, b, p);
    }
  }
  switch (bn) {
    case 32: return launch_cfg<32, 10>(ctx, a, b, p);
    case 64: return launch_cfg<64, 8>(ctx, a, b, p);
    case 128: return launch_cfg<128, 6>(ctx, a, b, p);
    default: return launch_cfg<256, 4>(ctx, a, b, p);
  }
}

// Chooses a split-K factor so that a problem with few output tiles still fills the SMs.
static int pick_splits(zb_ctx* ctx, long long tiles, int kb_total, int min_kb_per_split) {
  if (tiles >= ctx->sm_count || kb_total < 2 * min_kb_per_split) return 1;
  // largest factor that keeps tiles * splits within ONE full wave of the persistent grid: two waves halve the K range per CTA
  // but double the partial-buffer traffic (write + deterministic reduce), which measured slower (wgrad 6.97 -> 6.74 ms / step)
  static int waves = -1;   // tuning knob: ZENU_B200_SPLIT_WAVES
  if (waves < 0) {
    const char* e = getenv("ZENU_B200_SPLIT_WAVES");
    waves = e ? std::max(1, atoi(e)) : 1;
  }
  long long want = (static_cast<long long>(waves) * ctx->sm_count) / tiles;
  long long cap = kb_total / min_kb_per_split;
  long long s = std::max(1ll, std::min(want, cap));
  return static_cast<int>(std::min<long long>(s, 2ll * ctx->sm_count));
}

static void finish_split_fields(UmmaParams& p, int splits) {
  p.splits = splits;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
}

static int run_with_splits(zb_ctx* ctx, int bn, const CUtensorMap& a, const CUtensorMap& b, UmmaParams p, long long rows,
                           long long cols, float* out, long long ldo, float alpha, float beta, const float* bias) {
  if (p.splits <= 1) {
    p.split_stride = 0;
    p.alpha = alpha;
    p.beta = beta;
    p.bias = bias;
    return umma_launch(ctx, bn, a, b, p);
  }
  // partial buffers: [splits][rows][cols] dense
  void* ws = nullptr;
  const size_t need = sizeof(float) * static_cast<size_t>(p.splits) * rows * cols;
  int rc = ctx_workspace(ctx, need, &ws);
  if (rc != ZB_OK) return rc;
  UmmaParams q = p;
  q.D = static_cast<float*>(ws);
  q.ldd = cols;
  q.split_stride = rows * cols;
  q.alpha = 1.f;
  q.beta = 0.f;
  q.bias = nullptr;
  rc = umma_launch(ctx, bn, a, b, q);
  if (rc != ZB_OK) return rc;
  const long long total = rows * cols;
  const int block = 256;
  const int grid = static_cast<int>(std::min<long long>((total + block - 1) / block, ctx->sm_count * 8ll));
  plan_note("splitk_reduce;");
  ZB_KLAUNCH(ctx, splitk_reduce_kernel<<<grid, block, 0, ctx->stream>>>(static_cast<const float*>(ws), out, rows, cols, ldo, rows * cols,
                                                                     q.splits, alpha, beta, bias));
  return ZB_OK;
}

// Accumulation-chain limit applied to every launch planned by this thread (set around the three 3xTF32 passes by api.cu).
static thread_local int tl_chain_kb = 0;
void umma_set_chain_limit(int kb) { tl_chain_kb = kb; }
int umma_chain_limit() { return tl_chain_kb; }
int umma_conv_wgrad_halo(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);  // umma_wgrad.cu
int umma_conv_stem_dgrad(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, float*, float);  // umma_stem_dgrad.cu
int umma_conv_stem_wgrad(zb_ctx*, const zb_conv2d_desc*, const float*, const float*, long long, long long, long long, float*, int, int*);  // umma_stem_wgrad.cu

static void init_params(UmmaParams& p, zb_ctx* ctx) {
  memset(&p, 0, sizeof(p));
  p.chain_kb = tl_chain_kb;
  p.tap_tiles = 1;
  p.splits = 1;
  p.ntaps = 1;
  p.c_chunks = 1;
  p.conv_P = p.conv_Q = 1;
  p.stride_w = p.stride_h = 1;
  p.alpha = 1.f;
  p.err_flag = ctx->err_flag;
}

// ---------------------------------------------------------------------------------------------- GEMM
// Row-major C[m,n] = alpha * op(A) * op(B) + beta * C (+ bias[n]).  Returns ZB_ERR_UNSUPPORTED when the operands
// do not meet TMA alignment rules (caller then uses the SIMT kernel).
int umma_gemm(zb_ctx* ctx, bool trans_a, bool trans_b, long long m, long long n, long long k, float alpha,
              const float* a, long long lda, const float* b, long long ldb, float beta, float* c, long long ldc,
              const float* bias, const uint32_t* old_bits) {
  if ((lda & 3) || (ldb & 3) || (reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) ||
      m <= 0 || n <= 0 || k <= 0 || m > 0x7fffffffll || n > 0x7fffffffll || k > 0x7fffffffll) {
    set_last_error("umma_gemm: operands not TMA-compatible");
    return ZB_ERR_UNSUPPORTED;
  }
  const int bn = pick_bn(n);
  CUtensorMap ma, mb;
  int rc;
  UmmaParams p;
  init_params(p, ctx);
  if (!trans_a) {  // A stored [m][k]
    rc = make_map_2d(ctx, &ma, a, k, m, lda, 32, kUmmaBM);
    p.a_mode = A_TILED_K;
  } else {  // A stored [k][m]
    rc = make_map_2d(ctx, &ma, a, m, k, lda, 32, kUmmaBK, true);
    p.a_mode = A_TILED_MN;
  }
  if (rc != ZB_OK) return rc;
  if (trans_b) {  // B stored [n][k]
    rc = make_map_bk(ctx, &mb, p, b, k, n, ldb, bn);
    p.b_mode = B_TILED_K;
  } else {  // B stored [k][n]
    rc = make_map_2d(ctx, &mb, b, n, k, ldb, 32, kUmmaBK, true);
    p.b_mode = B_TILED_MN;
  }
  if (rc != ZB_OK) return rc;
  p.M = static_cast<int>(m);
  p.N = static_cast<int>(n);
  p.m_tiles = ceil_div(m, kUmmaBM);
  p.n_tiles = ceil_div(n, bn);
  p.kb_total = ceil_div(k, kUmmaBK);
  p.out_mode = OUT_ROWS;
  p.D = c;
  p.ldd = ldc;
  p.prof_flops = 2.0 * m * n * k;
  finish_split_fields(p, pick_splits(ctx, static_cast<long long>(p.m_tiles) * p.n_tiles, p.kb_total, 8));
  if (old_bits != nullptr) {
    if (p.splits > 1 || beta == 0.f) { set_last_error("umma_gemm: masked accumulate needs an unsplit beta launch"); return ZB_ERR_UNSUPPORTED; }
    p.old_bits = old_bits;
  }
  return run_with_splits(ctx, bn, ma, mb, p, m, n, c, ldc, alpha, beta, bias);
}

// ---------------------------------------------------------------------------------------------- NCHW pointwise convolutions
// The reference contract (NCHW activations, KCRS filters; zenu-matrix/src/nn/conv/interface.rs:270-281) served WITHOUT layout staging
// for 1x1 / stride-1 / unpadded convolutions: in NCHW the pixel dimension is the contiguous one, so per image
//   fprop  Y_n[K][HW]  = W[K][C]    * X_n[C][HW]      A = W, K-major;        B = X_n, N-contiguous (MN-major)
//   dgrad  dX_n[C][HW] = W^T[C][K]  * dY_n[K][HW]     A = W read M-contiguous; B = dY_n, MN-major
//   wgrad  dW[K][C]    = sum_n dY_n[K][HW] * X_n[C][HW]^T   both operands K-major, the reduction runs over (image, pixel)
// and [batch][rows][HW] tensors are plain 2-D matrices [batch * rows][HW] whose tensor-map row coordinate carries the image index
// (UmmaParams::batch_mode).  One launch for the whole batch; outputs land in NCHW as they are.  Needs HW % 4 == 0 (16-byte TMA row
// pitch) and a reduction length that is a whole number of 32-wide K blocks where K blocks would otherwise run into the next image.
bool umma_conv1x1_nchw_supported(const zb_conv2d_desc* d, int pass) {
  if (d->kh != 1 || d->kw != 1 || d->stride_h != 1 || d->stride_w != 1 || d->pad_h != 0 || d->pad_w != 0) return false;
  const long long HW = d->h * d->w;
  if (HW % 4 != 0 || HW < 32 || d->n > 65535 || ZB_ENV_FLAG("ZENU_B200_NO_NCHW_DIRECT")) return false;
  if (pass == 0) return d->c % 32 == 0;            // fprop: K blocks walk the input channels of ONE image
  if (pass == 1) return d->k % 32 == 0;            // dgrad: K blocks walk the output channels
  return true;                                     // wgrad: K blocks walk pixels; columns past HW are TMA zero fill
}

int umma_conv1x1_nchw_fprop(zb_ctx* ctx, const zb_conv2d_desc* d, const float* x, const float* w, float* y, float beta) {
  const long long HW = d->h * d->w, C = d->c, K = d->k, N = d->n;
  const int bn = pick_bn(HW);
  CUtensorMap ma, mb;
  UmmaParams p;
  init_params(p, ctx);
  int rc = make_map_2d(ctx, &ma, w, C, K, C, 32, kUmmaBM);
  if (rc != ZB_OK) return rc;
  if ((rc = make_map_2d(ctx, &mb, x, HW, N * C, HW, 32, kUmmaBK, true)) != ZB_OK) return rc;
  p.a_mode = A_TILED_K; p.b_mode = B_TILED_MN; p.out_mode = OUT_ROWS;
  p.M = static_cast<int>(K); p.N = static_cast<int>(HW);
  p.m_tiles = ceil_div(K, kUmmaBM); p.n_tiles = ceil_div(HW, bn);
  p.kb_total = static_cast<int>(C / kUmmaBK); p.kb_per_split = p.kb_total;
  p.batch_mode = 1; p.splits = static_cast<int>(N); p.b_batch_rows = static_cast<int>(C);
  p.split_stride = K * HW;
  p.D = y; p.ldd = HW; p.alpha = 1.f; p.beta = beta;
  p.prof_flops = 2.0 * N * HW * K * C;
  return umma_launch(ctx, bn, ma, mb, p);
}

int umma_conv1x1_nchw_dgrad(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* w, float* dx, float beta) {
  const long long HW = d->h * d->w, C = d->c, K = d->k, N = d->n;
  const int bn = pick_bn(HW);
  CUtensorMap ma, mb;
  UmmaParams p;
  init_params(p, ctx);
  int rc = make_map_2d(ctx, &ma, w, C, K, C, 32, kUmmaBK, true);   // A[m = c][k] = W[k][c]: stored with M contiguous
  if (rc != ZB_OK) return rc;
  if ((rc = make_map_2d(ctx, &mb, dy, HW, N * K, HW, 32, kUmmaBK, true)) != ZB_OK) return rc;
  p.a_mode = A_TILED_MN; p.b_mode = B_TILED_MN; p.out_mode = OUT_ROWS;
  p.M = static_cast<int>(C); p.N = static_cast<int>(HW);
  p.m_tiles = ceil_div(C, kUmmaBM); p.n_tiles = ceil_div(HW, bn);
  p.kb_total = static_cast<int>(K / kUmmaBK); p.kb_per_split = p.kb_total;
  p.batch_mode = 1; p.splits = static_cast<int>(N); p.b_batch_rows = static_cast<int>(K);
  p.split_stride = C * HW;
  p.D = dx; p.ldd = HW; p.alpha = 1.f; p.beta = beta;
  p.prof_flops = 2.0 * N * HW * K * C;
  return umma_launch(ctx, bn, ma, mb, p);
}

int umma_conv1x1_nchw_wgrad(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* x, float* dw, float beta) {
  const long long HW = d->h * d->w, C = d->c, K = d->k, N = d->n;
  const int bn = pick_bn(C);
  CUtensorMap ma, mb;
  UmmaParams p;
  init_params(p, ctx);
  int rc = make_map_2d(ctx, &ma, dy, HW, N * K, HW, 32, kUmmaBM);
  if (rc != ZB_OK) return rc;
  if ((rc = make_map_2d(ctx, &mb, x, HW, N * C, HW, 32, bn)) != ZB_OK) return rc;
  p.a_mode = A_TILED_K; p.b_mode = B_TILED_K; p.out_mode = OUT_ROWS;
  p.M = static_cast<int>(K); p.N = static_cast<int>(C);
  p.m_tiles = ceil_div(K, kUmmaBM); p.n_tiles = ceil_div(C, bn);
  p.batch_mode = 2; p.kb_per_batch = ceil_div(HW, kUmmaBK); p.a_batch_rows = static_cast<int>(K); p.b_batch_rows = static_cast<int>(C);
  if (N * p.kb_per_batch > 0x3fffffffll) { set_last_error("umma NCHW wgrad: reduction too long"); return ZB_ERR_UNSUPPORTED; }
  p.kb_total = static_cast<int>(N * p.kb_per_batch);
  p.D = dw; p.ldd = C;
  p.prof_flops = 2.0 * N * HW * K * C;
  finish_split_fields(p, pick_splits(ctx, static_cast<long long>(p.m_tiles) * p.n_tiles, p.kb_total, 16));
  return run_with_splits(ctx, bn, ma, mb, p, K, C, dw, C, 1.f, beta, nullptr);
}

// ---------------------------------------------------------------------------------------------- halo conv planner
// A custom tap set for the halo kernel (one parity class of a strided dgrad): `ntaps` taps at raster offsets (off_h, off_w) >= 0 from
// the tile's first input position (lower_h + first output row, lower_w), out_h x out_w outputs per image written to pixel
// (p * sy + oy, q * sx + ox) of an OH x OW image.
struct HaloTapSet {
  int ntaps, lower_h, lower_w, sy, oy, sx, ox;
  long long out_h, out_w, OH, OW;
  int off_h[kUmmaMaxTaps], off_w[kUmmaMaxTaps];
};
struct HaloPlan {
  int Wr, tp, p_tiles, slots, slot_bytes, raster_bytes, b_stages, resident, bn;
  int stride, planes, plane_bytes, rows_pl;                 // stride-2 form: four parity planes per slot (see UmmaParams::halo_planes)
  int stack, stack_rows;                                    // small maps: `stack` images per tile (see UmmaParams::halo_stack)
  int dh[4], dw[4], row_par[8], row_off[8], col_par[8], col_off[8];
  int pair;   // run on CTA pairs (halo_conv_kernel<BN, 2, true>): each CTA holds half of every filter tile
  size_t smem;
};
// in: [N][H][W][Cin] NHWC; filt: [Kout][R*S*Cin] (tap-major, channel-minor); out: [N][P][Q][Kout] with P = H + 2*ph - R + 1.
static bool halo_plan(zb_ctx* ctx, long long N, long long H, long long W, long long Cin, long long Kout, int R, int S, int ph, int pw,
                      HaloPlan* hp, int stride = 1, const HaloTapSet* ts = nullptr) {
  (void)ctx;
  if (ZB_ENV_FLAG("ZENU_B200_NO_HALO")) return false;
  if (stride != 1 && (stride != 2 || ZB_ENV_FLAG("ZENU_B200_NO_HALO_S2"))) return false;
  if (ts != nullptr && stride != 1) return false;
  const long long P = ts ? ts->out_h : (H + 2 * ph - R) / stride + 1, Q = ts ? ts->out_w : (W + 2 * pw - S) / stride + 1;
  const int ntaps = ts ? ts->ntaps : R * S;
  if (ntaps < 2 || ntaps > kUmmaMaxTaps || Cin % 32 != 0 || Kout % 4 != 0 || P <= 0 || Q <= 0 || (!ts && (ph < 0 || pw < 0))) return false;
  hp->stride = stride;
  hp->planes = 1;
  hp->dh[0] = ts ? ts->lower_h : -ph; hp->dw[0] = ts ? ts->lower_w : -pw;
  int max_row_off = R - 1, max_col_off = S - 1;
  if (ts) {
    max_row_off = 0; max_col_off = 0;
    for (int t = 0; t < ts->ntaps; ++t) { max_row_off = std::max(max_row_off, ts->off_h[t]); max_col_off = std::max(max_col_off, ts->off_w[t]); }
  }
  if (stride == 2) {
    // tap (r, s) reads input (2p + r - ph, 2q + s - pw): parity class ((r - ph) mod 2, (s - pw) mod 2), and inside the class' dense
    // plane (start = the smallest r - ph of that parity) the position (p + row_off, q + col_off)
    if (R < 2 || S < 2 || R > 8 || S > 8) return false;   // both parities must occur in each dimension (4 planes)
    int base_h[2] = {1 << 20, 1 << 20}, base_w[2] = {1 << 20, 1 << 20};
    for (int r = 0; r < R; ++r) { const int e = r - ph, par = ((e % 2) + 2) % 2; base_h[par] = std::min(base_h[par], e); }
    for (int s_ = 0; s_ < S; ++s_) { const int e = s_ - pw, par = ((e % 2) + 2) % 2; base_w[par] = std::min(base_w[par], e); }
    max_row_off = 0; max_col_off = 0;
    for (int r = 0; r < R; ++r) { const int e = r - ph, par = ((e % 2) + 2) % 2; hp->row_par[r] = par; hp->row_off[r] = (e - base_h[par]) / 2; max_row_off = std::max(max_row_off, hp->row_off[r]); }
    for (int s_ = 0; s_ < S; ++s_) { const int e = s_ - pw, par = ((e % 2) + 2) % 2; hp->col_par[s_] = par; hp->col_off[s_] = (e - base_w[par]) / 2; max_col_off = std::max(max_col_off, hp->col_off[s_]); }
    hp->planes = 4;
    for (int pl = 0; pl < 4; ++pl) { hp->dh[pl] = base_h[pl >> 1]; hp->dw[pl] = base_w[pl & 1]; }
  }
  const long long Wr = (stride == 1 && !ts) ? W + 2 * pw : Q + max_col_off;  // raster width (stride 1: = Q + S - 1)
  if (Wr > kUmmaBM || stride * (Wr - 1) + 1 > 256) return false;
  int tp = static_cast<int>(std::min<long long>(P, kUmmaBM / Wr));
  // balance the row blocks of an image (14 rows, tp 8 -> 7 + 7 instead of 8 + 6)
  const int p_tiles = ceil_div(P, tp);
  tp = ceil_div(P, p_tiles);
  int rows_pl = tp + max_row_off;   // raster rows per plane (stride 1: tp + R - 1)
  if (stride * (rows_pl - 1) + 1 > 256) return false;
  hp->stack = 1; hp->stack_rows = 0;
  int Wr_i = static_cast<int>(Wr);
  if (static_cast<double>(P * Q) / (static_cast<double>(p_tiles) * kUmmaBM) < 0.6) {   // too many dead MMA rows for one image per tile
    // small maps (7x7): several images per tile, padding rows / columns shared between neighbours (UmmaParams::halo_stack)
    if (stride != 1 || ts != nullptr || ZB_ENV_FLAG("ZENU_B200_NO_HALO_STACK")) return false;
    if (S - 1 - pw > pw || R - 1 - ph > ph || pw < 0 || ph < 0) return false;   // the shared zeros must cover both sides
    const int Ws = static_cast<int>(W) + pw, Hs = static_cast<int>(H) + ph;
    if ((Hs * Ws) % 8 != 0) return false;                                       // every image's raster starts on a swizzle-pattern boundary
    int G = 1;
    while (G < 4 && (static_cast<long long>(G) * Hs + P) * Ws <= kUmmaBM) ++G;   // rows ((G - 1) * Hs + P) * Ws of the tile are live
    if (G < 2 || N < G || static_cast<double>(G * P * Q) / kUmmaBM < 0.6) return false;
    hp->stack = G; hp->stack_rows = Hs;
    Wr_i = Ws; tp = static_cast<int>(P); rows_pl = Hs;
    hp->Wr = Wr_i; hp->tp = tp; hp->p_tiles = 1; hp->rows_pl = rows_pl;
  } else {
    hp->Wr = Wr_i; hp->tp = tp; hp->p_tiles = p_tiles; hp->rows_pl = rows_pl;
  }
  hp->bn = pick_bn(Kout);
  // CTA pairs: ZENU_B200_HALO_PAIR = 0 never, 1 only the N <= 128 layers, 2 (default) every layer.  Measured (tools/bench_conv.py --only
  // 3x3, fprop / dgrad): 256 -> 256 @14x14 0.113 -> 0.101 / 0.116 -> 0.105 ms, 128 -> 128 @28x28 0.145 -> 0.137 / 0.149 -> 0.138,
  // 64 -> 64 @56x56 0.211 -> 0.205 / 0.212 -> 0.205 (that layer is not bound by the MMA issue rate after all).
  static const int pair_mode = []() { const char* e = getenv("ZENU_B200_HALO_PAIR"); return e ? atoi(e) : 2; }();
  hp->pair = (pair_mode > 0 && hp->bn >= 64 && (pair_mode > 1 || hp->bn <= 128) && (N / hp->stack) * hp->p_tiles >= 2 && !ZB_ENV_FLAG("ZENU_B200_NO_CLUSTER")) ? 1 : 0;
  const int plane_data = rows_pl * hp->Wr * 128;
  if (hp->stack > 1) {
    hp->planes = hp->stack;            // one TMA box per image, back to back
    hp->plane_bytes = plane_data;      // (a multiple of 1024: checked above)
    for (int g = 1; g < hp->stack; ++g) { hp->dh[g] = hp->dh[0]; hp->dw[g] = hp->dw[0]; }
  } else {
    hp->plane_bytes = hp->planes == 1 ? 0 : ((plane_data + 1023) & ~1023);   // every plane starts on a swizzle-pattern boundary
  }
  hp->raster_bytes = hp->planes * plane_data;   // bytes the TMA loads of one slot deliver
  // rows a tap descriptor may touch beyond its raster (tap offset + 127 is the last row an MMA reads) stay inside the slot
  const int last_row = max_row_off * hp->Wr + max_col_off + kUmmaBM;
  hp->slot_bytes = ((hp->planes - 1) * hp->plane_bytes + std::max(plane_data, last_row * 128) + 1023) & ~1023;
  if (hp->stack > 1)   // + the zero rows below the last image (its bottom padding and the wrap of its last row), never written by TMA
    hp->slot_bytes = (std::max(hp->stack * plane_data + ((R - 1 - ph) * hp->Wr + S) * 128, last_row * 128) + 1023) & ~1023;
  const int b_bytes = hp->bn * 128 / (hp->pair ? 2 : 1);
  const int chunks = static_cast<int>(Cin / 32);
  const int budget = 227 * 1024 - 1024 - 16384 - 48 * hp->bn - 1024;  // alignment slack, epilogue staging, BN statistics, barriers
  const int n_tiles = ceil_div(Kout, hp->bn);
  hp->resident = 0;
  if (n_tiles == 1 && ntaps * chunks <= kHaloMaxB && ntaps * chunks * b_bytes + 2 * hp->slot_bytes <= budget) {
    hp->resident = 1;
    hp->b_stages = ntaps * chunks;
    hp->slots = std::min(4, (budget - hp->b_stages * b_bytes) / hp->slot_bytes);
  } else {
    hp->slots = std::min(3, std::max(2, chunks >= 2 ? 3 : 2));
    int left = budget - hp->slots * hp->slot_bytes;
    if (left < 4 * b_bytes) { hp->slots = 2; left = budget - 2 * hp->slot_bytes; }
    hp->b_stages = std::min(kHaloMaxB, left / b_bytes);
    if (hp->b_stages < 3) return false;
  }
  hp->smem = static_cast<size_t>(hp->slots) * hp->slot_bytes + static_cast<size_t>(hp->b_stages) * b_bytes + 16384 + 48 * hp->bn + 1024 + 1024;
  return true;
}

template <int BN, int CL, bool PAIR = false>
static int halo_launch_bn(zb_ctx* ctx, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p, size_t smem, int grid) {
  static SmemOptIn opt_in;
  { const int rc = smem_opt_in(ctx, opt_in, halo_conv_kernel<BN, CL, PAIR>, smem); if (rc != ZB_OK) return rc; }
  plan_note("halo_conv<bn=%d,cl=%d,pair=%d> resident=%d slots=%d b_stages=%d n_tiles=%d ntaps=%d c_chunks=%d beta=%d bias=%d stats=%d chain=%d tp=%d "
            "planes=%d stack=%d ~m_tiles=%d ~grid=%d;", BN, CL, PAIR ? 1 : 0, p.halo_b_resident, p.halo_slots, p.halo_b_stages, p.n_tiles, p.ntaps, p.c_chunks,
            p.beta != 0.f ? 1 : 0, p.bias != nullptr ? 1 : 0, p.stat_partial != nullptr ? 1 : 0, p.halo_chain, p.win_box_p, p.halo_planes, p.halo_stack,
            p.m_tiles, grid);
  if (plan_dry()) return ZB_OK;
  prof_begin(ctx, PROF_TENSOR);
  if (CL == 1) {
    halo_conv_kernel<BN, CL, PAIR><<<grid, 192, smem, ctx->stream>>>(a, b, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ZB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, halo_conv_kernel<BN, CL, PAIR>, a, b, p));
  }
  prof_end(ctx, PROF_TENSOR, p.prof_flops);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// tap_r/tap_s: raster offset of tap t (filter row / column in the orientation of `filt`)
static int umma_conv_halo(zb_ctx* ctx, const HaloPlan& hp, long long N, long long H, long long W, long long Cin, long long Kout, int R,
                          int S, int ph, int pw, const float* in, const float* filt, const float* bias, float* out, float beta,
                          double flops, const StatRequest* st = nullptr, const uint32_t* old_bits = nullptr, const HaloTapSet* ts = nullptr) {
  const int st_ = hp.stride;
  const long long P = ts ? ts->out_h : (H + 2 * ph - R) / st_ + 1, Q = ts ? ts->out_w : (W + 2 * pw - S) / st_ + 1;
  CUtensorMap ma, mb;
  {
    ZB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, "TMA operand must be 16-byte aligned");
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(Cin), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(Cin) * 4, static_cast<cuuint64_t>(W) * Cin * 4, static_cast<cuuint64_t>(H) * W * Cin * 4};
    // stride 2: the box spans stride * (n - 1) + 1 positions and the element strides pick every second one, i.e. Wr x rows_pl pixels of
    // ONE parity plane land as a dense raster; which plane is decided by the parity of the start coordinate the producer passes
    cuuint32_t box[4] = {32, static_cast<cuuint32_t>(st_ * (hp.Wr - 1) + 1), static_cast<cuuint32_t>(st_ * (hp.rows_pl - 1) + 1), 1};
    cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(st_), static_cast<cuuint32_t>(st_), 1};
    CUresult r = ctx->encode_tiled(&ma, operand_dtype(), 4, const_cast<float*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (halo raster) failed (%d)", int(r)); return ZB_ERR_CUDA; }
  }
  const int taps = ts ? ts->ntaps : R * S;
  // streamed filter + at least two row blocks: pairs of CTAs share every filter tile through TMA multicast (halo_conv_kernel CL = 2)
  const int cl = (hp.pair || (!hp.resident && hp.bn >= 64 && static_cast<long long>(N) * hp.p_tiles >= 2 && !ZB_ENV_FLAG("ZENU_B200_NO_CLUSTER"))) ? 2 : 1;
  int rc = make_map_2d(ctx, &mb, filt, static_cast<long long>(taps) * Cin, Kout, static_cast<long long>(taps) * Cin, 32, hp.bn / cl);
  if (rc != ZB_OK) return rc;
  UmmaParams p;
  init_params(p, ctx);
  p.a_mode = A_TILED_K;   // (unused by the halo kernel; the epilogue only looks at out_mode)
  p.b_mode = B_TILED_K;
  p.out_mode = OUT_WINDOW;
  p.M = static_cast<int>(N * P * Q);
  p.N = static_cast<int>(Kout);
  p.m_tiles = hp.stack > 1 ? ceil_div(N, hp.stack) : static_cast<int>(N) * hp.p_tiles;
  p.n_tiles = ceil_div(Kout, hp.bn);
  p.win_box_q = hp.Wr; p.win_box_p = hp.tp; p.win_q_tiles = 1; p.win_p_tiles = hp.p_tiles;
  p.halo_stack = hp.stack; p.halo_stack_rows = hp.stack_rows;
  p.batch_n = static_cast<int>(N);
  p.conv_P = static_cast<int>(P); p.conv_Q = static_cast<int>(Q);
  p.lower_w = -pw; p.lower_h = -ph;
  p.ntaps = taps;
  p.c_chunks = static_cast<int>(Cin / 32);
  p.b_tap_stride = static_cast<int>(Cin);
  if (ts) {
    for (int t = 0; t < taps; ++t) p.tap_w[t] = static_cast<uint16_t>(ts->off_h[t] * hp.Wr + ts->off_w[t]);
    p.scat_OH = static_cast<int>(ts->OH); p.scat_OW = static_cast<int>(ts->OW);
    p.scat_sy = ts->sy; p.scat_oy = ts->oy; p.scat_sx = ts->sx; p.scat_ox = ts->ox;
  }
  for (int r = 0; r < R && !ts; ++r)
    for (int sx = 0; sx < S; ++sx)
      p.tap_w[r * S + sx] = st_ == 1 ? static_cast<uint16_t>(r * hp.Wr + sx)
                                     : static_cast<uint16_t>((hp.row_par[r] * 2 + hp.col_par[sx]) * (hp.plane_bytes / 128) + hp.row_off[r] * hp.Wr +
                                                             hp.col_off[sx]);
  p.halo_planes = hp.planes; p.halo_plane_bytes = hp.plane_bytes; p.halo_stride = st_;
  for (int pl = 0; pl < 4; ++pl) { p.halo_dh[pl] = hp.dh[pl % hp.planes]; p.halo_dw[pl] = hp.dw[pl % hp.planes]; }
  if (hp.stack > 1) p.halo_stride = 1;
  p.halo_slots = hp.slots; p.halo_slot_bytes = hp.slot_bytes; p.halo_raster_bytes = hp.raster_bytes;
  p.halo_b_stages = hp.b_stages; p.halo_b_resident = hp.resident;
  p.halo_chain = p.chain_kb > 0 ? std::max(1, p.chain_kb / taps) : 0;
  p.kb_total = taps * p.c_chunks;
  p.prof_flops = flops;
  p.D = out;
  p.ldd = Kout;
  p.alpha = 1.f; p.beta = beta; p.bias = bias;
  p.old_bits = beta != 0.f ? old_bits : nullptr;
  if (p.old_bits != nullptr && (st != nullptr || Kout % 32 != 0)) {   // the mask words use the epilogue's statistics slot
    set_last_error("umma halo: masked accumulate needs Kout % 32 == 0 and no fused statistics");
    return ZB_ERR_UNSUPPORTED;
  }
  finish_split_fields(p, 1);
  p.split_stride = 0;
  stat_attach(ctx, p, st, p.m_tiles * p.n_tiles, Kout, out, bias);
  if (cl == 1) {
    const int grid = p.stat_partial ? stat_grid(ctx, p.m_tiles * p.n_tiles, p.n_tiles) : std::min(p.m_tiles * p.n_tiles, ctx->sm_count);
    switch (hp.bn) {
      case 32: return halo_launch_bn<32, 1>(ctx, ma, mb, p, hp.smem, grid);
      case 64: return halo_launch_bn<64, 1>(ctx, ma, mb, p, hp.smem, grid);
      case 128: return halo_launch_bn<128, 1>(ctx, ma, mb, p, hp.smem, grid);
      default: return halo_launch_bn<256, 1>(ctx, ma, mb, p, hp.smem, grid);
    }
  }
  // clusters walk groups of `cl` row blocks; with fused statistics the cluster count is a multiple of n_tiles (one column block per CTA)
  const int groups = ceil_div(p.m_tiles, cl) * p.n_tiles;
  int clusters = std::min(groups, ctx->sm_count / cl);
  if (p.stat_partial) {
    clusters = std::max(p.n_tiles, clusters / p.n_tiles * p.n_tiles);
    *st->rows = clusters / p.n_tiles * cl * 4;
  }
  const int grid = clusters * cl;
  if (hp.pair) {
    switch (hp.bn) {
      case 64: return halo_launch_bn<64, 2, true>(ctx, ma, mb, p, hp.smem, grid);
      case 128: return halo_launch_bn<128, 2, true>(ctx, ma, mb, p, hp.smem, grid);
      default: return halo_launch_bn<256, 2, true>(ctx, ma, mb, p, hp.smem, grid);
    }
  }
  switch (hp.bn) {
    case 64: return halo_launch_bn<64, 2>(ctx, ma, mb, p, hp.smem, grid);
    case 128: return halo_launch_bn<128, 2>(ctx, ma, mb, p, hp.smem, grid);
    default: return halo_launch_bn<256, 2>(ctx, ma, mb, p, hp.smem, grid);
  }
}

// ---------------------------------------------------------------------------------------------- conv
bool umma_conv_supported(const zb_conv2d_desc* d) {
  if (d->c % 32 != 0 || d->k % 4 != 0) return false;
  if (d->kh * d->kw > kUmmaMaxTaps) return false;
  if (d->dil_h * (d->kh - 1) > 255 || d->dil_w * (d->kw - 1) > 255) return false;
  if (d->pad_h > 127 || d->pad_w > 127) return false;
  return true;
}

// y[N,P,Q,K] = conv(x[N,H,W,C], w[K,R,S,C]) (+bias)
int umma_conv_fprop_nhwc(zb_ctx* ctx, const zb_conv2d_desc* d, const float* x, const float* w, const float* bias,
                         float* y, float beta, const float* stat_shift, float* stat_partial, int* stat_rows) {
  StatRequest st;
  st.shift = stat_shift; st.partial = stat_partial; st.rows = stat_rows;
  if (stat_rows) *stat_rows = 0;
  if (!umma_conv_supported(d)) { set_last_error("umma fprop: shape unsupported"); return ZB_ERR_UNSUPPORTED; }
  const long long P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  const long long Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  const long long M = d->n * P * Q;
  const int bn = pick_bn(d->k);
  const int taps = static_cast<int>(d->kh * d->kw);
  if (d->stride_h == d->stride_w && (d->stride_h == 1 || d->stride_h == 2) && d->dil_h == 1 && d->dil_w == 1 && taps > 1) {
    HaloPlan hp;
    if (halo_plan(ctx, d->n, d->h, d->w, d->c, d->k, static_cast<int>(d->kh), static_cast<int>(d->kw), static_cast<int>(d->pad_h),
                  static_cast<int>(d->pad_w), &hp, static_cast<int>(d->stride_h)))
      return umma_conv_halo(ctx, hp, d->n, d->h, d->w, d->c, d->k, static_cast<int>(d->kh), static_cast<int>(d->kw),
                            static_cast<int>(d->pad_h), static_cast<int>(d->pad_w), x, w, bias, y, beta, 2.0 * M * d->k * d->c * taps, &st);
  }
  CUtensorMap ma, mb;
  UmmaParams p;
  init_params(p, ctx);
  int rc;
  const bool pointwise = (taps == 1 && d->stride_h == 1 && d->stride_w == 1 && d->pad_h == 0 && d->pad_w == 0);
  if (pointwise) {
    rc = make_map_2d(ctx, &ma, x, d->c, M, d->c, 32, kUmmaBM);
    p.a_mode = A_TILED_K;
  } else {
    rc = make_map_im2col(ctx, &ma, x, d->n, d->h, d->w, d->c, -static_cast<int>(d->pad_w), -static_cast<int>(d->pad_h),
                         static_cast<int>(d->pad_w - d->dil_w * (d->kw - 1)), static_cast<int>(d->pad_h - d->dil_h * (d->kh - 1)),
                         static_cast<int>(d->stride_w), static_cast<int>(d->stride_h), kUmmaBM);
    p.a_mode = A_IM2COL_K;
  }
  if (rc != ZB_OK) return rc;
  rc = make_map_bk(ctx, &mb, p, w, static_cast<long long>(taps) * d->c, d->k, static_cast<long long>(taps) * d->c, bn);
  if (rc != ZB_OK) return rc;
  p.b_mode = B_TILED_K;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(d->k);
  p.m_tiles = ceil_div(M, kUmmaBM);
  p.n_tiles = ceil_div(d->k, bn);
  p.conv_P = static_cast<int>(P);
  p.conv_Q = static_cast<int>(Q);
  p.lower_w = -static_cast<int>(d->pad_w);
  p.lower_h = -static_cast<int>(d->pad_h);
  p.stride_w = static_cast<int>(d->stride_w);
  p.stride_h = static_cast<int>(d->stride_h);
  p.ntaps = taps;
  p.c_chunks = static_cast<int>(d->c / 32);
  p.b_tap_stride = static_cast<int>(d->c);
  for (int r = 0; r < d->kh; ++r)
    for (int s = 0; s < d->kw; ++s) {
      p.tap_w[r * d->kw + s] = static_cast<uint16_t>(s * d->dil_w);
      p.tap_h[r * d->kw + s] = static_cast<uint16_t>(r * d->dil_h);
    }
  p.kb_total = taps * p.c_chunks;
  p.prof_flops = 2.0 * M * d->k * d->c * taps;  // = 2*N*P*Q*K*C*R*S
  p.out_mode = OUT_ROWS;
  p.D = y;
  p.ldd = d->k;
  finish_split_fields(p, 1);
  if (beta == 0.f) stat_attach(ctx, p, &st, p.m_tiles * p.n_tiles, d->k, y, bias, bn);
  return run_with_splits(ctx, bn, ma, mb, p, M, d->k, y, d->k, 1.f, beta, bias);
}

// dx[N,H,W,C] = dgrad(dy[N,P,Q,K], w[K,R,S,C]).  Stride 1: one implicit GEMM over dy with the flipped filter.
// Stride s > 1: one implicit GEMM per output parity class (h % s, w % s), scattered into dx.
int umma_conv_dgrad_nhwc(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* w, float* dx, float beta,
                         const uint32_t* old_bits) {
  // old_bits (beta launches): the old dx counts through a 1-bit-per-element mask (UmmaParams::old_bits)
  if (old_bits != nullptr && (beta == 0.f || d->c % 32 != 0)) { set_last_error("umma dgrad: masked accumulate unsupported here"); return ZB_ERR_UNSUPPORTED; }
  if (d->k % 32 != 0 || d->kh * d->kw > kUmmaMaxTaps) {
    set_last_error("umma dgrad: shape unsupported");
    return ZB_ERR_UNSUPPORTED;
  }
  const long long P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  const long long Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  const int sh = static_cast<int>(d->stride_h), sw = static_cast<int>(d->stride_w);
  const int R = static_cast<int>(d->kh), S = static_cast<int>(d->kw);
  if (R == 1 && S == 1 && sh == 1 && sw == 1 && d->pad_h == 0 && d->pad_w == 0 && d->c % 4 == 0 && d->k % 4 == 0) {
    // pointwise: dx[pixels][C] = dy[pixels][K] * w[K][C]: a plain GEMM on the filter as stored (no transform, tiled loads)
    return umma_gemm(ctx, false, false, d->n * P * Q, d->c, d->k, 1.f, dy, d->k, w, d->c, beta, dx, d->c, nullptr, old_bits);
  }
  if (sh == 1 && sw == 1 && d->dil_h == 1 && d->dil_w == 1 && R * S > 1 && d->pad_h <= R - 1 && d->pad_w <= S - 1) {
    // stride 1: dx = conv(dy, flipped filter) with padding R-1-pad; halo-reuse kernel over the dY raster
    HaloPlan hp;
    const int ph2 = R - 1 - static_cast<int>(d->pad_h), pw2 = S - 1 - static_cast<int>(d->pad_w);
    if (halo_plan(ctx, d->n, P, Q, d->k, d->c, R, S, ph2, pw2, &hp)) {
      void* ws = nullptr;
      const size_t elems = static_cast<size_t>(d->c) * R * S * d->k;
      int rc2 = ctx_workspace(ctx, elems * sizeof(float), &ws);
      if (rc2 != ZB_OK) return rc2;
      float* wt = static_cast<float*>(ws);
      TapList tl;
      for (int t = 0; t < R * S; ++t) tl.rs[t] = R * S - 1 - t;   // tap (r', s') of the flipped filter = (R-1-r', S-1-s')
      const long long total = static_cast<long long>(elems);
      const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 8ll));
      plan_note("dgrad_filter;");
      ZB_KLAUNCH(ctx, dgrad_filter_kernel<<<grid, 256, 0, ctx->stream>>>(w, wt, static_cast<int>(d->k), R * S, static_cast<int>(d->c), R * S, tl));
      return umma_conv_halo(ctx, hp, d->n, P, Q, d->k, d->c, R, S, ph2, pw2, dy, wt, nullptr, dx, beta,
                            2.0 * d->n * P * Q * d->k * d->c * R * S, nullptr, old_bits);
    }
  }
  const int bn = pick_bn(d->c);

  // Enumerate parity classes and their taps first so that unsupported geometry fails before any launch.
  struct ClassPlan {
    int a, b, Ha, Wb, lower_h, lower_w, upper_h, upper_w, ntaps;
    int tap_rs[kUmmaMaxTaps];
    int off_h[kUmmaMaxTaps], off_w[kUmmaMaxTaps];
  };
  std::vector<ClassPlan> plans;
  bool need_zero = false;
  for (int a = 0; a < sh; ++a)
    for (int b = 0; b < sw; ++b) {
      ClassPlan cp;
      cp.a = a; cp.b = b; cp.ntaps = 0;
      cp.Ha = static_cast<int>((d->h - a + sh - 1) / sh);
      cp.Wb = static_cast<int>((d->w - b + sw - 1) / sw);
      if (cp.Ha <= 0 || cp.Wb <= 0) continue;
      int min_h = 1 << 30, min_w = 1 << 30;
      for (int r = 0; r < R; ++r) {
        const long long th = a + d->pad_h - r * d->dil_h;
        if (((th % sh) + sh) % sh != 0) continue;
        for (int s = 0; s < S; ++s) {
          const long long tw = b + d->pad_w - s * d->dil_w;
          if (((tw % sw) + sw) % sw != 0) continue;
          const int oh = static_cast<int>(th >= 0 ? th / sh : -((-th) / sh));
          const int ow = static_cast<int>(tw >= 0 ? tw / sw : -((-tw) / sw));
          cp.tap_rs[cp.ntaps] = r * S + s;
          cp.off_h[cp.ntaps] = oh;
          cp.off_w[cp.ntaps] = ow;
          min_h = std::min(min_h, oh);
          min_w = std::min(min_w, ow);
          cp.ntaps++;
        }
      }
      if (cp.ntaps == 0) { need_zero = true; continue; }
      cp.lower_h = min_h; cp.lower_w = min_w;
      cp.upper_h = cp.Ha - static_cast<int>(P) + min_h;
      cp.upper_w = cp.Wb - static_cast<int>(Q) + min_w;
      for (int t = 0; t < cp.ntaps; ++t) {
        cp.off_h[t] -= min_h;
        cp.off_w[t] -= min_w;
        if (cp.off_h[t] > 255 || cp.off_w[t] > 255) { set_last_error("umma dgrad: tap offset too large"); return ZB_ERR_UNSUPPORTED; }
      }
      if (cp.lower_h < -128 || cp.lower_h > 127 || cp.lower_w < -128 || cp.lower_w > 127 || cp.upper_h < -128 ||
          cp.upper_h > 127 || cp.upper_w < -128 || cp.upper_w > 127) {
        set_last_error("umma dgrad: corner out of range");
        return ZB_ERR_UNSUPPORTED;
      }
      plans.push_back(cp);
    }
  if (need_zero && old_bits != nullptr) {   // pixels no launch reaches would keep their old value unmasked
    set_last_error("umma dgrad: masked accumulate needs every dx pixel covered by a parity class");
    return ZB_ERR_UNSUPPORTED;
  }
  if (need_zero && beta == 0.f) {
    plan_note("memset_dx;");
    if (!plan_dry()) ZB_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * d->n * d->h * d->w * d->c, ctx->stream));
  }

  // workspace: transformed filters for all classes + tap index lists
  size_t wt_elems = 0;
  for (auto& cp : plans) wt_elems += static_cast<size_t>(d->c) * cp.ntaps * d->k;
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, wt_elems * sizeof(float), &ws);
  if (rc != ZB_OK) return rc;
  float* wt_base = static_cast<float*>(ws);

  if (!plans.empty()) {   // the transformed filters of every class, one launch
    if (plans.size() > 16) { set_last_error("umma dgrad: too many parity classes"); return ZB_ERR_UNSUPPORTED; }
    ClassTable tb;
    memset(&tb, 0, sizeof(tb));
    tb.n = static_cast<int>(plans.size());
    int tap_base = 0;
    for (size_t ci = 0; ci < plans.size(); ++ci) {
      tb.ntaps[ci] = plans[ci].ntaps;
      tb.tap_base[ci] = tap_base;
      for (int t = 0; t < plans[ci].ntaps; ++t) tb.rs[tap_base + t] = plans[ci].tap_rs[t];
      tap_base += plans[ci].ntaps;
      tb.start[ci + 1] = tb.start[ci] + static_cast<long long>(d->c) * plans[ci].ntaps * d->k;
    }
    const long long total = tb.start[tb.n];
    const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * 8ll)));
    plan_note("dgrad_filter(classes=%d);", tb.n);
    ZB_KLAUNCH(ctx, dgrad_filter_classes_kernel<<<grid, 256, 0, ctx->stream>>>(w, wt_base, static_cast<int>(d->k), R * S, static_cast<int>(d->c), tb));
  }
  size_t wt_off = 0;
  for (size_t ci = 0; ci < plans.size(); ++ci) {
    const ClassPlan& cp = plans[ci];
    float* wt = wt_base + wt_off;
    wt_off += static_cast<size_t>(d->c) * cp.ntaps * d->k;
    if (cp.ntaps >= 2 && !ZB_ENV_FLAG("ZENU_B200_NO_HALO_DGRAD_CLASS")) {
      // a class with several taps is a stride-1 conv over dY with that tap subset: on the halo kernel its dY raster is fetched once per
      // channel chunk instead of once per tap (outputs scattered to the class' pixels by the epilogue)
      HaloTapSet ts;
      ts.ntaps = cp.ntaps; ts.lower_h = cp.lower_h; ts.lower_w = cp.lower_w;
      ts.sy = sh; ts.oy = cp.a; ts.sx = sw; ts.ox = cp.b;
      ts.out_h = cp.Ha; ts.out_w = cp.Wb; ts.OH = d->h; ts.OW = d->w;
      for (int t = 0; t < cp.ntaps; ++t) { ts.off_h[t] = cp.off_h[t]; ts.off_w[t] = cp.off_w[t]; }
      HaloPlan hp;
      if (halo_plan(ctx, d->n, P, Q, d->k, d->c, 1, 1, 0, 0, &hp, 1, &ts)) {
        rc = umma_conv_halo(ctx, hp, d->n, P, Q, d->k, d->c, 1, 1, 0, 0, dy, wt, nullptr, dx, beta,
                            2.0 * d->n * cp.Ha * cp.Wb * static_cast<double>(d->c) * d->k * cp.ntaps, nullptr, beta != 0.f ? old_bits : nullptr, &ts);
        if (rc != ZB_OK) return rc;
        continue;
      }
    }
    CUtensorMap ma, mb;
    rc = make_map_im2col(ctx, &ma, dy, d->n, P, Q, d->k, cp.lower_w, cp.lower_h, cp.upper_w, cp.upper_h, 1, 1, kUmmaBM);
    if (rc != ZB_OK) return rc;
    UmmaParams p;
    init_params(p, ctx);
    rc = make_map_bk(ctx, &mb, p, wt, static_cast<long long>(cp.ntaps) * d->k, d->c, static_cast<long long>(cp.ntaps) * d->k, bn);
    if (rc != ZB_OK) return rc;
    const long long M = d->n * static_cast<long long>(cp.Ha) * cp.Wb;
    p.a_mode = A_IM2COL_K;
    p.b_mode = B_TILED_K;
    p.M = static_cast<int>(M);
    p.N = static_cast<int>(d->c);
    p.m_tiles = ceil_div(M, kUmmaBM);
    p.n_tiles = ceil_div(d->c, bn);
    p.conv_P = cp.Ha;
    p.conv_Q = cp.Wb;
    p.lower_w = cp.lower_w;
    p.lower_h = cp.lower_h;
    p.ntaps = cp.ntaps;
    p.c_chunks = static_cast<int>(d->k / 32);
    p.b_tap_stride = static_cast<int>(d->k);
    for (int t = 0; t < cp.ntaps; ++t) {
      p.tap_w[t] = static_cast<uint16_t>(cp.off_w[t]);
      p.tap_h[t] = static_cast<uint16_t>(cp.off_h[t]);
    }
    p.kb_total = cp.ntaps * p.c_chunks;
    p.prof_flops = 2.0 * M * d->c * d->k * cp.ntaps;  // summed over parity classes = 2*N*P*Q*K*C*R*S up to borders
    p.D = dx;
    p.ldd = d->c;
    if (sh == 1 && sw == 1) {
      p.out_mode = OUT_ROWS;
    } else {
      p.out_mode = OUT_SCATTER;
      p.scat_OH = static_cast<int>(d->h);
      p.scat_OW = static_cast<int>(d->w);
      p.scat_sy = sh; p.scat_oy = cp.a;
      p.scat_sx = sw; p.scat_ox = cp.b;
    }
    finish_split_fields(p, 1);
    p.split_stride = 0;
    p.beta = beta;  // dx = dgrad + beta * dx: gradient fan-in (residual shortcuts) without a separate add pass
    p.old_bits = beta != 0.f ? old_bits : nullptr;   // every pixel of dx belongs to exactly one parity class: each old value is read once
    rc = umma_launch(ctx, bn, ma, mb, p);
    if (rc != ZB_OK) return rc;
  }
  return ZB_OK;
}

// dw[K,R,S,C] = wgrad(dy[N,P,Q,K], x[N,H,W,C]); reduction over N*P*Q pixels, split-K + deterministic reduce.
int umma_conv_wgrad_nhwc(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* x, float* dw, float beta) {
  if (d->c % 32 != 0 || d->k % 4 != 0 || d->kh * d->kw > kUmmaMaxTaps || d->pad_h > 127 || d->pad_w > 127 ||
      d->dil_h * (d->kh - 1) > 255 || d->dil_w * (d->kw - 1) > 255) {
    set_last_error("umma wgrad: shape unsupported");
    return ZB_ERR_UNSUPPORTED;
  }
  const long long P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  const long long Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  const long long NPQ = d->n * P * Q;
  const int taps = static_cast<int>(d->kh * d->kw);
  const int bn = pick_bn(d->c);
  int rc = umma_conv_wgrad_halo(ctx, d, dy, x, dw, beta);   // stride-1 R x S filters: every tap from one smem raster
  if (rc != ZB_ERR_UNSUPPORTED) return rc;
  CUtensorMap ma, mb;
  UmmaParams p;
  init_params(p, ctx);
  rc = make_map_2d(ctx, &ma, dy, d->k, NPQ, d->k, 32, kUmmaBK, true);
  if (rc != ZB_OK) return rc;
  p.a_mode = A_TILED_MN;
  const bool pointwise = (taps == 1 && d->stride_h == 1 && d->stride_w == 1 && d->pad_h == 0 && d->pad_w == 0);
  if (pointwise) {
    rc = make_map_2d(ctx, &mb, x, d->c, NPQ, d->c, 32, kUmmaBK, true);
    p.b_mode = B_TILED_MN;
  } else {
    rc = make_map_im2col(ctx, &mb, x, d->n, d->h, d->w, d->c, -static_cast<int>(d->pad_w), -static_cast<int>(d->pad_h),
                         static_cast<int>(d->pad_w - d->dil_w * (d->kw - 1)), static_cast<int>(d->pad_h - d->dil_h * (d->kh - 1)),
                         static_cast<int>(d->stride_w), static_cast<int>(d->stride_h), kUmmaBK, true);
    p.b_mode = B_IM2COL_MN;
  }
  if (rc != ZB_OK) return rc;
  p.M = static_cast<int>(d->k);
  p.N = static_cast<int>(d->c);
  p.m_tiles = ceil_div(d->k, kUmmaBM);
  p.n_tiles = ceil_div(d->c, bn);
  p.tap_tiles = taps;
  p.conv_P = static_cast<int>(P);
  p.conv_Q = static_cast<int>(Q);
  p.lower_w = -static_cast<int>(d->pad_w);
  p.lower_h = -static_cast<int>(d->pad_h);
  p.stride_w = static_cast<int>(d->stride_w);
  p.stride_h = static_cast<int>(d->stride_h);
  p.ntaps = taps;
  for (int r = 0; r < d->kh; ++r)
    for (int s = 0; s < d->kw; ++s) {
      p.tap_w[r * d->kw + s] = static_cast<uint16_t>(s * d->dil_w);
      p.tap_h[r * d->kw + s] = static_cast<uint16_t>(r * d->dil_h);
    }
  p.kb_total = ceil_div(NPQ, kUmmaBK);
  p.prof_flops = 2.0 * NPQ * d->k * d->c * taps;
  p.out_mode = OUT_ROWS;
  p.D = dw;
  p.ldd = static_cast<long long>(taps) * d->c;
  p.tap_col_stride = d->c;
  const long long tiles = static_cast<long long>(p.m_tiles) * p.n_tiles * taps;
  finish_split_fields(p, pick_splits(ctx, tiles, p.kb_total, 16));
  if (p.splits <= 1) {
    p.split_stride = 0;
    p.beta = beta;
    return umma_launch(ctx, bn, ma, mb, p);
  }
  // partial buffers keep the [K][taps*C] shape of dw
  const long long rows = d->k, cols = static_cast<long long>(taps) * d->c;
  void* ws = nullptr;
  rc = ctx_workspace(ctx, sizeof(float) * static_cast<size_t>(p.splits) * rows * cols, &ws);
  if (rc != ZB_OK) return rc;
  UmmaParams q = p;
  q.D = static_cast<float*>(ws);
  q.split_stride = rows * cols;
  rc = umma_launch(ctx, bn, ma, mb, q);
  if (rc != ZB_OK) return rc;
  const long long total = rows * cols;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 8ll));
  plan_note("splitk_reduce;");
  ZB_KLAUNCH(ctx, splitk_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(static_cast<const float*>(ws), dw, rows, cols, cols, rows * cols,
                                                                   q.splits, 1.f, beta, nullptr));
  return ZB_OK;
}


// ---------------------------------------------------------------------------------------------- small-C conv
// Convs with C <= 4 input channels (the 7x7/s2 ResNet stem, the first layer of the CIFAR CNN): a 32-channel K block
// does not exist, so the input is repacked once into zero-padded NHWC4 rows [N][H][Wp][4] and one K block is the
// 32-float sliding window "8 taps x 4 channels of one filter row" starting at padded column q*stride_w.  Consecutive
// output pixels' windows overlap in memory; a tiled tensor map whose q-dimension stride (16*stride_w bytes) is smaller
// than the window (128 bytes) expresses exactly that, so one TMA box still lands a [pixels][32] K-major tile.
// Filter rows are the GEMM-K blocks (K = R x 32, taps s >= S and channel 3 carry zero weights).
bool umma_conv_smallc_supported(const zb_conv2d_desc* d) {
  return d->c <= 4 && d->kw <= 8 && d->dil_w == 1 && d->kh <= kUmmaMaxTaps && d->k % 4 == 0 && d->pad_h <= 127 &&
         d->dil_h * (d->kh - 1) <= 255;
}

// xp[n][h][wp][0..3] = x[n][c][h][wp - pad_w] (zero outside); src is NHWC (c_stride 1) or NCHW
__global__ void smallc_pack_input_kernel(const float* __restrict__ x, float4* __restrict__ xp, long long N, int C, int H, int W,
                                         int Wp, int pad_w, int nchw) {
  const long long total = N * H * Wp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wp = static_cast<int>(i % Wp);
    const long long nh = i / Wp;
    const int w = wp - pad_w;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (w >= 0 && w < W) {
      if (nchw) {
        const long long n = nh / H, h = nh - n * H;
        for (int c = 0; c < C; ++c) v[c] = __ldg(x + ((n * C + c) * H + h) * W + w);
      } else {
        const float* src = x + (nh * W + w) * C;
        for (int c = 0; c < C; ++c) v[c] = __ldg(src + c);
      }
    }
    xp[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// wp[k][r][s*4+c] = w[k][r][s][c] (KRSC), zero elsewhere
__global__ void smallc_pack_filter_kernel(const float* __restrict__ w, float* __restrict__ wp, int K, int R, int S, int C) {
  const int total = K * R * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i & 31, kr = i >> 5;
    const int sidx = j >> 2, c = j & 3;
    wp[i] = (sidx < S && c < C) ? w[(static_cast<long long>(kr) * S + sidx) * C + c] : 0.f;
  }
}
// dw[k][r][s][c] = alpha-free sum over split partials of dwp[split][k][r][s*4+c]
__global__ void smallc_unpack_dw_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int K, int R, int S, int C,
                                        int splits, long long split_stride, float beta) {
  const int total = K * R * S * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    const int sidx = (i / C) % S;
    const int kr = i / (C * S);
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp) acc += dwp[sp * split_stride + static_cast<long long>(kr) * 32 + sidx * 4 + c];
    dw[i] = beta != 0.f ? acc + beta * dw[i] : acc;
  }
}

static int make_map_window(zb_ctx* ctx, CUtensorMap* map, const float* base, long long N, long long H, long long Wp,
                           long long Q, int stride_w, int box_q, int box_p, bool mn_major) {
  cuuint64_t dims[4] = {32, static_cast<cuuint64_t>(Q), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(stride_w) * 16, static_cast<cuuint64_t>(Wp) * 16,
                           static_cast<cuuint64_t>(H) * Wp * 16};
  cuuint32_t box[4] = {32, static_cast<cuuint32_t>(box_q), static_cast<cuuint32_t>(box_p), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(map, operand_dtype(), 4, const_cast<float*>(base), dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (sliding window) failed (%d): N=%lld H=%lld Wp=%lld Q=%lld stride=%d box=%dx%d", int(r), N, H,
                   Wp, Q, stride_w, box_q, box_p);
    return ZB_ERR_CUDA;
  }
  return ZB_OK;
}

struct SmallcGeom {
  long long P, Q, Wp;
  size_t xp_bytes, wp_bytes;
};
static SmallcGeom smallc_geom(const zb_conv2d_desc* d) {
  SmallcGeom g;
  g.P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  g.Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  // padded row: pad_w zero pixels, the W data pixels, then zeros so that the last window (8 pixels from padded column
  // (Q-1)*stride_w) stays inside the row; window rows with q >= Q are out of the map's bounds and never touch memory
  g.Wp = std::max<long long>(d->pad_w + d->w, (g.Q - 1) * d->stride_w + 8);
  g.xp_bytes = (static_cast<size_t>(d->n) * d->h * g.Wp * 16 + 1023) & ~size_t(1023);
  g.wp_bytes = (static_cast<size_t>(d->k) * d->kh * 32 * 4 + 1023) & ~size_t(1023);
  return g;
}
static int smallc_pack_input(zb_ctx* ctx, const zb_conv2d_desc* d, const SmallcGeom& g, const float* x, int x_nchw, float* xp) {
  const long long total = d->n * d->h * g.Wp;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, ctx->sm_count * 16ll));
  plan_note("smallc_pack_input(nchw=%d);", x_nchw);
  ZB_KLAUNCH(ctx, smallc_pack_input_kernel<<<grid, 256, 0, ctx->stream>>>(x, reinterpret_cast<float4*>(xp), d->n, static_cast<int>(d->c),
                                                                       static_cast<int>(d->h), static_cast<int>(d->w), static_cast<int>(g.Wp),
                                                                       static_cast<int>(d->pad_w), x_nchw));
  return ZB_OK;
}

template <int BN>
static int stem_fprop_launch(zb_ctx* ctx, const CUtensorMap& a, const CUtensorMap& b, const UmmaParams& p, int grid, size_t smem) {
  static SmemOptIn opt_in;   // per instantiation: each kernel needs its own opt-in to > 48 KB of dynamic shared memory
  { const int rc = smem_opt_in(ctx, opt_in, stem_fprop_kernel<BN>, smem); if (rc != ZB_OK) return rc; }
  plan_note("stem_fprop<bn=%d> stages=%d taps=%d beta=%d bias=%d stats=%d ~m_tiles=%d ~grid=%d;", BN, p.halo_slots, p.ntaps, p.beta != 0.f ? 1 : 0,
            p.bias != nullptr ? 1 : 0, p.stat_partial != nullptr ? 1 : 0, p.m_tiles, grid);
  if (plan_dry()) return ZB_OK;
  prof_begin(ctx, PROF_TENSOR);
  stem_fprop_kernel<BN><<<grid, 192, smem, ctx->stream>>>(a, b, p);
  prof_end(ctx, PROF_TENSOR, p.prof_flops);
  ZB_LAUNCH_CHECK(ctx);
  return ZB_OK;
}

// y[N,P,Q,K] (NHWC) = conv(x, w[K,R,S,C]) (+bias); x is NHWC (x_nchw = 0) or NCHW (x_nchw = 1)
int umma_conv_smallc_fprop(zb_ctx* ctx, const zb_conv2d_desc* d, const float* x, int x_nchw, const float* w, const float* bias,
                           float* y, float beta, const float* stat_shift, float* stat_partial, int* stat_rows) {
  StatRequest st;
  st.shift = stat_shift; st.partial = stat_partial; st.rows = stat_rows;
  if (stat_rows) *stat_rows = 0;
  if (!umma_conv_smallc_supported(d)) { set_last_error("umma small-C fprop: shape unsupported"); return ZB_ERR_UNSUPPORTED; }
  const SmallcGeom g = smallc_geom(d);
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, g.xp_bytes + g.wp_bytes, &ws);
  if (rc != ZB_OK) return rc;
  float* xp = static_cast<float*>(ws);
  float* wp = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + g.xp_bytes);
  if ((rc = smallc_pack_input(ctx, d, g, x, x_nchw, xp)) != ZB_OK) return rc;
  {
    const int total = static_cast<int>(d->k * d->kh * 32);
    plan_note("smallc_pack_filter;");
    ZB_KLAUNCH(ctx, smallc_pack_filter_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(w, wp, static_cast<int>(d->k), static_cast<int>(d->kh),
                                                                                         static_cast<int>(d->kw), static_cast<int>(d->c)));
  }
  const int bn = pick_bn(d->k);
  if (g.Q <= kUmmaBM && bn <= 128 && d->k <= bn && !ZB_ENV_FLAG("ZENU_B200_NO_STEM_FPROP")) {
    // several output rows per tile, resident filter (stem_fprop_kernel)
    const int TP = 512 / (2 * bn), R = static_cast<int>(d->kh);
    const int budget = 227 * 1024 - 1024 - 512 - 16384 - 48 * bn;
    const int stages = std::min(kStemMaxStages, (budget - R * bn * 128) / 16384);
    const long long P_pad = (g.P + TP - 1) / TP * TP;
    if (stages >= 3 && d->n * P_pad < 0x3fffffffll) {
      UmmaParams p;
      init_params(p, ctx);
      p.chain_kb = 0;   // K = R blocks: a short accumulation chain in any math mode
      CUtensorMap ma, mb;
      if ((rc = make_map_window(ctx, &ma, xp, d->n, d->h, g.Wp, g.Q, static_cast<int>(d->stride_w), static_cast<int>(g.Q), 1, false)) != ZB_OK) return rc;
      if ((rc = make_map_2d(ctx, &mb, wp, d->kh * 32, d->k, d->kh * 32, 32, bn)) != ZB_OK) return rc;
      p.a_mode = A_WINDOW_K; p.b_mode = B_TILED_K; p.out_mode = OUT_WINDOW;
      p.win_box_q = static_cast<int>(g.Q); p.win_box_p = 1; p.win_q_tiles = 1; p.win_p_tiles = static_cast<int>(P_pad);
      p.M = static_cast<int>(d->n * g.P * g.Q); p.N = static_cast<int>(d->k);
      p.m_tiles = static_cast<int>(d->n * P_pad); p.n_tiles = 1;
      p.conv_P = static_cast<int>(g.P); p.conv_Q = static_cast<int>(g.Q);
      p.lower_h = -static_cast<int>(d->pad_h); p.stride_h = static_cast<int>(d->stride_h); p.stride_w = static_cast<int>(d->stride_w);
      p.dg_dh = static_cast<int>(d->dil_h);
      p.ntaps = R; p.halo_slots = stages; p.kb_total = R;
      p.prof_flops = 2.0 * p.M * d->k * d->c * d->kh * d->kw;
      p.D = y; p.ldd = d->k; p.alpha = 1.f; p.beta = beta; p.bias = bias;
      finish_split_fields(p, 1);
      p.split_stride = 0;
      const int groups = p.m_tiles / TP;
      const int grid = std::min(groups, ctx->sm_count);
      if (beta == 0.f) {
        stat_attach(ctx, p, &st, p.m_tiles, d->k, y, bias);
        if (p.stat_partial != nullptr) *st.rows = grid * 4;   // one partial row per epilogue warp of each launched CTA
      }
      const size_t smem = static_cast<size_t>(R) * bn * 128 + static_cast<size_t>(stages) * 16384 + 16384 + 48 * bn + 512 + 1024;
      switch (bn) {
        case 32: return stem_fprop_launch<32>(ctx, ma, mb, p, grid, smem);
        case 64: return stem_fprop_launch<64>(ctx, ma, mb, p, grid, smem);
        default: return stem_fprop_launch<128>(ctx, ma, mb, p, grid, smem);
      }
    }
  }
  UmmaParams p;
  init_params(p, ctx);
  p.win_box_q = static_cast<int>(std::min<long long>(g.Q, kUmmaBM));
  p.win_box_p = (d->stride_h == 1) ? static_cast<int>(std::max<long long>(1, std::min<long long>(kUmmaBM / p.win_box_q, g.P))) : 1;
  p.win_q_tiles = ceil_div(g.Q, p.win_box_q);
  p.win_p_tiles = ceil_div(g.P, p.win_box_p);
  CUtensorMap ma, mb;
  if ((rc = make_map_window(ctx, &ma, xp, d->n, d->h, g.Wp, g.Q, static_cast<int>(d->stride_w), p.win_box_q, p.win_box_p, false)) != ZB_OK) return rc;
  if ((rc = make_map_2d(ctx, &mb, wp, d->kh * 32, d->k, d->kh * 32, 32, bn)) != ZB_OK) return rc;
  p.a_mode = A_WINDOW_K;
  p.b_mode = B_TILED_K;
  p.out_mode = OUT_WINDOW;
  p.M = static_cast<int>(d->n * g.P * g.Q);
  p.N = static_cast<int>(d->k);
  p.m_tiles = static_cast<int>(d->n) * p.win_p_tiles * p.win_q_tiles;
  p.n_tiles = ceil_div(d->k, bn);
  p.conv_P = static_cast<int>(g.P);
  p.conv_Q = static_cast<int>(g.Q);
  p.lower_h = -static_cast<int>(d->pad_h);
  p.stride_h = static_cast<int>(d->stride_h);
  p.stride_w = static_cast<int>(d->stride_w);
  for (int r = 0; r < d->kh; ++r) p.tap_h[r] = static_cast<uint16_t>(r * d->dil_h);
  p.kb_total = static_cast<int>(d->kh);
  p.prof_flops = 2.0 * p.M * d->k * d->c * d->kh * d->kw;
  p.D = y;
  p.ldd = d->k;
  finish_split_fields(p, 1);
  if (beta == 0.f) stat_attach(ctx, p, &st, p.m_tiles * p.n_tiles, d->k, y, bias);
  return run_with_splits(ctx, bn, ma, mb, p, p.M, d->k, y, d->k, 1.f, beta, bias);
}

// dw[K,R,S,C] = wgrad(dy[N,P,Q,K] (NHWC), x); reduction over pixels in 32-pixel runs of one output row.
// Up to 8 filter rows are folded into GEMM-N (N = R x 32 window elements share every dY tile), more rows fall back to one
// tile group per filter row.
int umma_conv_smallc_wgrad(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* x, int x_nchw, float* dw,
                           float beta) {
  if (!umma_conv_smallc_supported(d)) { set_last_error("umma small-C wgrad: shape unsupported"); return ZB_ERR_UNSUPPORTED; }
  const SmallcGeom g = smallc_geom(d);
  const long long NPQ = d->n * g.P * g.Q;
  const int R = static_cast<int>(d->kh);
  const int fold = R <= 8 ? R : 1;           // filter rows per tile group
  const int bn = pick_bn(fold * 32);
  UmmaParams p;
  init_params(p, ctx);
  p.win_qblocks = ceil_div(g.Q, 32);
  p.m_tiles = ceil_div(d->k, kUmmaBM);
  p.n_tiles = 1;
  p.ntaps = fold;
  p.tap_tiles = R / fold;
  p.kb_total = static_cast<int>(d->n * g.P) * p.win_qblocks;
  const long long tiles = static_cast<long long>(p.m_tiles) * p.tap_tiles;
  finish_split_fields(p, pick_splits(ctx, tiles, p.kb_total, 16));
  const long long rows = d->k, cols = static_cast<long long>(R) * 32;
  const int max_parts = std::max(p.splits, ctx->sm_count);   // the strip-walking kernel writes one partial per CTA
  const size_t part_bytes = (sizeof(float) * static_cast<size_t>(max_parts) * rows * cols + 1023) & ~size_t(1023);
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, g.xp_bytes + part_bytes, &ws);
  if (rc != ZB_OK) return rc;
  float* xp = static_cast<float*>(ws);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + g.xp_bytes);
  if ((rc = smallc_pack_input(ctx, d, g, x, x_nchw, xp)) != ZB_OK) return rc;
  if (fold == R && p.chain_kb == 0) {   // column strips walked down the output rows, window boxes reused (umma_stem_wgrad.cu)
    int parts = 0;
    rc = umma_conv_stem_wgrad(ctx, d, dy, xp, g.Wp, g.P, g.Q, part, max_parts, &parts);
    if (rc == ZB_OK) {
      const int total = static_cast<int>(d->k * d->kh * d->kw * d->c);
      plan_note("smallc_unpack_dw;");
      ZB_KLAUNCH(ctx, smallc_unpack_dw_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(part, dw, static_cast<int>(d->k), static_cast<int>(d->kh),
                                                                                         static_cast<int>(d->kw), static_cast<int>(d->c), parts,
                                                                                         rows * cols, beta));
      return ZB_OK;
    }
    if (rc != ZB_ERR_UNSUPPORTED) return rc;
  }
  CUtensorMap ma, mb;
  if ((rc = make_map_2d(ctx, &ma, dy, d->k, NPQ, d->k, 32, kUmmaBK, true)) != ZB_OK) return rc;
  if ((rc = make_map_window(ctx, &mb, xp, d->n, d->h, g.Wp, g.Q, static_cast<int>(d->stride_w), 32, 1, true)) != ZB_OK) return rc;
  p.a_mode = A_TILED_MN;
  p.b_mode = B_WINDOW_MN;
  p.out_mode = OUT_ROWS;
  p.M = static_cast<int>(d->k);
  p.N = fold * 32;
  p.conv_P = static_cast<int>(g.P);
  p.conv_Q = static_cast<int>(g.Q);
  p.lower_h = -static_cast<int>(d->pad_h);
  p.stride_h = static_cast<int>(d->stride_h);
  p.stride_w = static_cast<int>(d->stride_w);
  for (int r = 0; r < R; ++r) p.tap_h[r] = static_cast<uint16_t>(r * d->dil_h);
  p.prof_flops = 2.0 * NPQ * d->k * d->c * d->kh * d->kw;
  p.D = part;
  p.ldd = cols;
  p.tap_col_stride = fold * 32;
  p.split_stride = p.splits > 1 ? rows * cols : 0;
  p.alpha = 1.f;
  if ((rc = umma_launch(ctx, bn, ma, mb, p)) != ZB_OK) return rc;
  const int total = static_cast<int>(d->k * d->kh * d->kw * d->c);
  plan_note("smallc_unpack_dw;");
  ZB_KLAUNCH(ctx, smallc_unpack_dw_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(part, dw, static_cast<int>(d->k), static_cast<int>(d->kh),
                                                                                     static_cast<int>(d->kw), static_cast<int>(d->c), p.splits,
                                                                                     rows * cols, beta));
  return ZB_OK;
}

// wd[j = s*4+c][r*K + k] = w[k][r][s][c] (KRSC), zero elsewhere: the K-major B operand of the small-C dgrad
__global__ void smallc_pack_dgrad_filter_kernel(const float* __restrict__ w, float* __restrict__ wd, int K, int R, int S, int C) {
  const int total = 32 * R * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % K;
    const int r = (i / K) % R;
    const int j = i / (K * R);
    const int sidx = j >> 2, c = j & 3;
    wd[i] = (sidx < S && c < C) ? w[((static_cast<long long>(k) * R + r) * S + sidx) * C + c] : 0.f;
  }
}

bool umma_conv_smallc_dgrad_supported(const zb_conv2d_desc* d) {
  if (!(d->c <= 4 && d->kw <= 8 && d->dil_w == 1 && d->k % 32 == 0 && d->stride_h <= 4 && d->kh <= 16)) return false;
  const long long Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  if (Q > kUmmaBM) return false;
  // every input-row parity class must be reached by at least one filter row
  for (int a = 0; a < d->stride_h; ++a) {
    int cnt = 0;
    for (int r = 0; r < d->kh; ++r) cnt += ((a - r * d->dil_h) % d->stride_h + d->stride_h) % d->stride_h == 0;
    if (cnt == 0) return false;
  }
  return true;
}

// dx[N,H,W,C] (NHWC, C <= 4) = dgrad(dy[N,P,Q,K], w[K,R,S,C]).  Tile = one input row (n, h): the accumulator
// D[q][s*4+c] = sum over the filter rows r reaching h and over k of dY[n, p(h,r), q, k] * w[k,r,s,c]; the epilogue overlap-adds
// the filter columns.  dY is fetched ~R/stride_h times instead of R*S times (general parity-class path).
int umma_conv_smallc_dgrad(zb_ctx* ctx, const zb_conv2d_desc* d, const float* dy, const float* w, float* dx, float beta) {
  if (!umma_conv_smallc_dgrad_supported(d)) { set_last_error("umma small-C dgrad: shape unsupported"); return ZB_ERR_UNSUPPORTED; }
  {   // 8 input rows per tile, filter resident in smem (umma_stem_dgrad.cu); falls through to the row-per-tile kernel otherwise
    const int rc0 = umma_conv_stem_dgrad(ctx, d, dy, w, dx, beta);
    if (rc0 != ZB_ERR_UNSUPPORTED) return rc0;
  }
  const long long P = zb_conv_out_size(d->h, d->kh, d->pad_h, d->stride_h, d->dil_h);
  const long long Q = zb_conv_out_size(d->w, d->kw, d->pad_w, d->stride_w, d->dil_w);
  const int R = static_cast<int>(d->kh);
  const size_t wd_bytes = sizeof(float) * 32 * R * d->k;
  void* ws = nullptr;
  int rc = ctx_workspace(ctx, wd_bytes, &ws);
  if (rc != ZB_OK) return rc;
  float* wd = static_cast<float*>(ws);
  {
    const int total = static_cast<int>(32 * R * d->k);
    plan_note("smallc_pack_dgrad_filter;");
    ZB_KLAUNCH(ctx, smallc_pack_dgrad_filter_kernel<<<(total + 255) / 256, 256, 0, ctx->stream>>>(w, wd, static_cast<int>(d->k), R, static_cast<int>(d->kw),
                                                                                               static_cast<int>(d->c)));
  }
  UmmaParams p;
  init_params(p, ctx);
  p.win_box_q = static_cast<int>(Q);
  CUtensorMap ma, mb;
  {
    ZB_REQUIRE((reinterpret_cast<uintptr_t>(dy) & 15) == 0, "TMA operand must be 16-byte aligned");
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->k), static_cast<cuuint64_t>(Q), static_cast<cuuint64_t>(P), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(d->k) * 4, static_cast<cuuint64_t>(Q) * d->k * 4, static_cast<cuuint64_t>(P) * Q * d->k * 4};
    cuuint32_t box[4] = {32, static_cast<cuuint32_t>(Q), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = ctx->encode_tiled(&ma, operand_dtype(), 4, const_cast<float*>(dy), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled (dY rows) failed (%d)", int(r)); return ZB_ERR_CUDA; }
  }
  if ((rc = make_map_2d(ctx, &mb, wd, static_cast<long long>(R) * d->k, 32, static_cast<long long>(R) * d->k, 32, 32)) != ZB_OK) return rc;
  p.a_mode = A_ROWS_K;
  p.b_mode = B_TILED_K;
  p.out_mode = OUT_WDGRAD;
  p.M = static_cast<int>(d->n * d->h) * kUmmaBM;
  p.N = 32;
  p.m_tiles = static_cast<int>(d->n * d->h);
  p.n_tiles = 1;
  p.conv_P = static_cast<int>(P);
  p.conv_Q = static_cast<int>(Q);
  p.c_chunks = static_cast<int>(d->k / 32);
  p.b_tap_stride = static_cast<int>(d->k);
  p.dg_H = static_cast<int>(d->h); p.dg_W = static_cast<int>(d->w); p.dg_C = static_cast<int>(d->c); p.dg_S = static_cast<int>(d->kw);
  p.dg_sw = static_cast<int>(d->stride_w); p.dg_pw = static_cast<int>(d->pad_w);
  p.dg_sh = static_cast<int>(d->stride_h); p.dg_ph = static_cast<int>(d->pad_h); p.dg_dh = static_cast<int>(d->dil_h);
  for (int a = 0; a < p.dg_sh; ++a) {   // class a = (h + pad_h) % stride_h: rows r with (a - r*dil_h) % stride_h == 0
    int cnt = 0;
    for (int r = 0; r < R; ++r)
      if (((a - r * p.dg_dh) % p.dg_sh + p.dg_sh) % p.dg_sh == 0) p.dg_r[a][cnt++] = static_cast<uint8_t>(r);
    p.dg_cnt[a] = static_cast<uint8_t>(cnt);
  }
  p.kb_total = R * p.c_chunks;  // upper bound; the per-tile range comes from dg_cnt
  p.prof_flops = 2.0 * d->n * P * Q * d->k * d->c * d->kh * d->kw;
  p.D = dx;
  p.ldd = d->c;
  finish_split_fields(p, 1);
  p.split_stride = 0;
  p.beta = beta;
  p.chain_kb = 0;  // the overlap-add epilogue drains one accumulator per tile (K <= R * k/32 blocks: a short chain anyway)
  return umma_launch(ctx, 32, ma, mb, p);
}

}  // namespace zb
