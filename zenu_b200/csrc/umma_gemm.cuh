// umma_gemm.cuh — parameter block shared by the tcgen05 implicit-GEMM kernel and its host planners.
#pragma once
#include <cstdint>
#include <cuda.h>

namespace zb {

constexpr int kUmmaBM = 128;      // accumulator rows per CTA tile (TMEM lanes)
constexpr int kUmmaBK = 32;       // fp32 elements of K per pipeline stage (= one 128-byte swizzle row)
constexpr int kUmmaMaxTaps = 64;  // filter taps addressable by one launch (7x7 = 49)

enum UmmaAMode : int {
  A_TILED_K = 0,   // A[M][K] row-major, K contiguous            (Linear, 1x1 conv, dgrad 1x1)
  A_IM2COL_K = 1,  // NHWC activations through TMA im2col, K = (tap, channel chunk)   (fprop / dgrad 3x3)
  A_TILED_MN = 2,  // A^T stored: [K][M] row-major, M contiguous (wgrad: dY[pixels][Kout])
  A_ROWS_K = 4,    // small-C dgrad: dY[N][P][Q][K] rows through a 4-D tiled map, one (image, output row p) per K block
  A_WINDOW_K = 3,  // small-C conv (C <= 4, e.g. the 7x7 stem): packed NHWC4 input, one 32-float sliding window
                   // (8 taps x 4 channels of one filter row) per output pixel through an overlapped-stride tiled map
};
enum UmmaBMode : int {
  B_TILED_K = 0,    // B[N][K] row-major, K contiguous            (weights [Kout][taps*C])
  B_TILED_MN = 2,   // B stored [K][N] row-major, N contiguous    (wgrad 1x1: X[pixels][Cin]; Linear dX)
  B_IM2COL_MN = 3,  // NHWC activations through TMA im2col, N = channels, K = pixels (wgrad kxk)
  B_WINDOW_MN = 4,  // small-C wgrad: the same sliding windows, N = 32 window elements, K = pixels
};
enum UmmaOutMode : int {
  OUT_ROWS = 0,     // D[row][col], row pitch ldd
  OUT_SCATTER = 1,  // row m = (n,p,q) -> NHWC pixel (n, p*osy+oy0, q*osx+ox0) of an [N][OH][OW][ldd] tensor
  OUT_WDGRAD = 3,   // small-C dgrad: tile = (image, input row h); accumulator D[q][s*4+c] is overlap-added over the
                    // filter columns s into the dense row dX[n][h][0..W)[0..C)
  OUT_WINDOW = 2,   // A_WINDOW_K tiles: tile = (image, p-block, q-block), local row l -> (p0 + l / box_q, q0 + l % box_q)
};

struct UmmaParams {
  // GEMM view
  int M, N;              // logical output extent (rows, cols) of one "tap tile group"
  int m_tiles, n_tiles;  // ceil(M/128), ceil(N/BN)
  int tap_tiles;         // wgrad: one output tile group per filter tap; otherwise 1
  int splits;            // split-K factor (K ranges go to separate partial buffers)
  int kb_total;          // number of 32-wide K blocks in the full reduction
  int kb_per_split;
  int a_mode, b_mode, out_mode;
  // im2col geometry (base-pixel grid P x Q per image, lower corner and traversal stride in input pixels)
  int conv_P, conv_Q;
  int lower_w, lower_h, stride_w, stride_h;
  int ntaps, c_chunks;   // A_IM2COL_K: kb -> (tap = kb / c_chunks, c0 = 32 * (kb % c_chunks))
  int b_tap_stride;      // columns of B per tap (padded C) for A_IM2COL_K
  // sliding-window modes: box of win_box_q x win_box_p output pixels per M tile; win_qblocks 32-pixel K blocks per row
  int win_box_q, win_box_p, win_q_tiles, win_p_tiles, win_qblocks;
  // small-C dgrad geometry: per input-row parity class a = (h + pad_h) % stride_h, the filter rows r that reach it
  int dg_H, dg_W, dg_C, dg_S, dg_sw, dg_pw, dg_sh, dg_ph, dg_dh;
  uint8_t dg_cnt[4];
  uint8_t dg_r[4][16];
  uint16_t tap_w[kUmmaMaxTaps];
  uint16_t tap_h[kUmmaMaxTaps];
  // output
  float* D;
  long long ldd;              // row pitch of D in elements
  long long tap_col_stride;   // wgrad: column offset of tap t inside a D row (= Cin)
  long long split_stride;     // elements between split-K partial buffers (0 when splits == 1)
  const float* bias;          // optional [N], added when splits == 1
  float alpha, beta;          // D = alpha*acc + bias + beta*D_old (splits == 1); partials store raw acc
  // beta launches only: D_old is taken through a 1-bit-per-element mask (bit index = element index in D; 0 -> the old value counts
  // as 0).  The fused BN+add+ReLU backward hands the residual branch (dy, ReLU bits) instead of a materialised masked gradient; the
  // dgrad that accumulates into it masks on the fly (zb_conv2d_dgrad_acc_masked).  NULL = plain accumulate.
  const uint32_t* old_bits;
  int scat_OH, scat_OW, scat_sy, scat_oy, scat_sx, scat_ox;  // OUT_SCATTER geometry
  // halo-reuse conv kernel (stride-1 RxS convs): one smem raster of (tp + R - 1) x Wr input pixels per 32-channel chunk
  // serves every filter tap through UMMA descriptors that start tap_w[t] rows (128 bytes each) into the raster
  int halo_slots, halo_slot_bytes, halo_raster_bytes, halo_b_stages, halo_b_resident;
  // stride-2 halo form: the slot holds halo_planes (4) dense rasters, one per (row, column) parity of the input, loaded through ONE
  // tensor map with element strides (1, 2, 2, 1) from shifted start coordinates (halo_dh / halo_dw, relative to stride * first output row /
  // column of the tile); stride 1: one plane, halo_dh[0] / halo_dw[0] = -pad
  int halo_planes, halo_plane_bytes, halo_stride;
  int halo_dh[4], halo_dw[4];
  // small feature maps (7x7): halo_stack images share ONE 128-row tile.  Their rasters ([pad rows, H rows] x [pad columns, W columns]
  // each: the bottom / right padding of one row or image IS the top / left padding of the next, both are zeros) are stacked in the
  // slot, one TMA box per image; window row ps of the tile = image ps / halo_stack_rows, output row ps % halo_stack_rows
  int halo_stack, halo_stack_rows, batch_n;
  // accumulation-chain limit (3xTF32 mode): the tensor core adds into TMEM with truncation, a bias that grows with the number
  // of MMA steps; with chain_kb > 0 the accumulator is flushed through the epilogue (fp32 round-to-nearest adds into D) every
  // chain_kb K blocks (halo kernel: every halo_chain channel chunks) instead of once per tile.  0 = one chain per tile.
  int chain_kb, halo_chain;
  // fused BatchNorm statistics (conv fprop followed by BN): per-column sum(y - shift) and sum((y - shift)^2) of the tile rows
  // this CTA stores, accumulated per epilogue warp in shared memory over all its tiles (the grid is a multiple of n_tiles, so
  // a CTA only ever sees one column block) and written once to stat_partial[(blockIdx.x / n_tiles) * 4 + warp][2][N]
  float* stat_partial;
  const float* stat_shift;
  // Batched GEMM over images (the NCHW pointwise convolutions served without layout staging): the operands are 2-D views
  // [batch * rows][cols] of contiguous [batch][rows][cols] tensors, so a batch index is just an offset on a tensor map's row coordinate.
  //   batch_mode 1: the split index of a tile IS the image (full K range per tile, B rows + split * b_batch_rows, output offset
  //                 split * split_stride, epilogue applies alpha / beta / bias as for an unsplit launch)
  //   batch_mode 2: the reduction runs over (image, K block): K block kb -> image kb / kb_per_batch, A / B row coordinates shifted by
  //                 image * a_batch_rows / b_batch_rows (wgrad: both operands K-major); split-K slices that combined range
  int batch_mode, a_batch_rows, b_batch_rows, kb_per_batch;
  // host only: the K-major B matrix as given to make_map_2d, so that the launcher can re-encode its tensor map with a half-height box
  // when it runs the launch on CTA pairs (umma_kernel CL = 2)
  const float* hb_base;
  long long hb_inner, hb_outer, hb_pitch;
  int* err_flag;              // device word set to 1 on an mbarrier timeout
  double prof_flops;          // host only: algorithmic FLOPs of this launch (profiling)
};

}  // namespace zb
