// umma.cuh — thin inline-PTX wrappers for the Blackwell (sm_100a) tensor path:
// mbarrier, TMA (tiled + im2col), tcgen05.{alloc,mma,commit,ld,fence}, UMMA descriptors.
// No CUTLASS/CuTe dependency; bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor"
// tables (cross-checked against cute/arch/mma_sm100_desc.hpp field comments).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace zb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU box.  On timeout the error word is set and
// the caller bails out of its role loop (the host reports ZB_ERR_TIMEOUT after the launch).
#ifndef ZB_MBAR_TIMEOUT_CYCLES
#define ZB_MBAR_TIMEOUT_CYCLES (4000000000ll)  // ~2 s at 1.9 GHz
#endif
// The timeout clock and the error word are looked at once every 256 failed polls only: the error word is a volatile GLOBAL load
// (LDG.E.STRONG.SYS, several hundred cycles), and with one per poll -- as in round 1 -- a waiter noticed its barrier up to a load
// latency late on every pipeline hand-off (ncu source view of stem_fprop_kernel, profiles/r2_ncu_full_stem.txt: 1.35 M such loads,
// the compare behind them the second-hottest stall of the kernel).  mbarrier.try_wait itself suspends the thread in hardware until the
// phase completes or a short system time limit passes, so re-polling immediately is the cheap part.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* err_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  for (uint32_t spins = 1;; ++spins) {
    if (mbar_try_wait(bar, parity)) return true;
    if ((spins & 255u) == 0u && (clock64() - t0 > ZB_MBAR_TIMEOUT_CYCLES || (err_flag && *err_flag))) {
      if (err_flag) *err_flag = 1;
      return false;
    }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}
// im2col mode, NHWC tensor seen as (C, W, H, N); (off_w, off_h) = filter-tap offset from the base pixel.
__device__ __forceinline__ void tma_load_im2col_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}

// ---------------------------------------------------------------- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope) on purpose: what the waiting side needs
// ordered are this thread's tcgen05 operations, which tcgen05.fence::before_thread_sync / after_thread_sync take care of; a
// .release.cluster arrive compiles to an ERRBAR that waits for all of the warp's outstanding global stores (measured: ~1300 cycles
// per epilogue tile, 17 % of the epilogue warps' time in the 64-channel halo kernel, and the reason short-K GEMMs lost on CTA pairs).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait on a barrier of this CTA that is arrived on from another CTA of the cluster, bounded
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile int* err_flag) {
  long long t0 = 0;
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
    if (t0 == 0) t0 = clock64();
    if ((spins & 255u) == 255u && (clock64() - t0 > ZB_MBAR_TIMEOUT_CYCLES || (err_flag && *err_flag))) {   // (see mbar_wait)
      if (err_flag) *err_flag = 1;
      return false;
    }
  }
}

// 2-D tiled load delivered to the same smem offset (and signalling the same mbarrier offset) in every CTA of cta_mask
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                      uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// im2col-mode load delivered to every CTA of cta_mask (same smem offset, same mbarrier offset)
__device__ __forceinline__ void tma_load_im2col_4d_multicast(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                             int n, uint16_t off_w, uint16_t off_h, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
        "h"(off_h), "h"(cta_mask)
      : "memory");
}

// explicit shared-space 128-bit accesses (a pointer derived through uintptr_t arithmetic compiles to generic LD/ST)
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// Division by a kernel-invariant divisor d (tile counts, row / column extents) as multiply-high + shift (Granlund-Montgomery:
// m = ceil(2^(31 + l) / d), l = ceil(log2 d); exact for 0 <= n < 2^31).  The role loops decode a tile id and a row index per tile;
// a hardware-free 32-bit division is ~25 instructions, and the epilogue warps are bound by their own instruction stream.
struct FastDiv { uint32_t mul, shr; };
__device__ __forceinline__ FastDiv fastdiv_make(int d) {
  FastDiv f;
  if (d <= 1) { f.mul = 0u; f.shr = 0u; return f; }
  const int l = 32 - __clz(d - 1);
  f.mul = static_cast<uint32_t>(((1ull << (31 + l)) + static_cast<unsigned long long>(d) - 1ull) / static_cast<unsigned long long>(d));
  f.shr = static_cast<uint32_t>(l - 1);
  return f;
}
__device__ __forceinline__ int fastdiv(int n, FastDiv f) {
  return f.mul != 0u ? static_cast<int>(__umulhi(static_cast<uint32_t>(n), f.mul) >> f.shr) : n;
}
// packed fp32 pairs (FADD2 / FFMA2 on sm_100): the same IEEE results as the scalar forms, half the instructions
__device__ __forceinline__ float2 add_f32x2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 sub_f32x2(float2 a, float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 fma_f32x2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ void stg128(float* p, float4 v) {   // explicit .global: the pointer may have passed through an opaque asm
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread retire
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Same, arriving on the mbarrier at this smem offset in every CTA of cta_mask (slots filled by multicast TMA are released
// only when all consumers in the cluster are done with them).
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA pair (cta_group::2): one M = 256 MMA spans two CTAs of a cluster; A comes from each CTA's own smem (its 128 rows), B
// is split (each CTA holds N/2 rows at the same smem offset), D lands in each CTA's own TMEM.  Issued by the leader (rank 0) only.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {  // one full warp in EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this smem offset in both CTAs of the pair once all previously issued MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

// TMA loads of a CTA pair: the data lands in this CTA's shared memory, the bytes are counted on an mbarrier that may live in the
// peer CTA (`bar_cluster_addr` is a shared::cluster address, e.g. mapa_shared(smem_u32(bar), 0) = the leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c, int w,
                                                        int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
        "h"(off_h)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- lane base+i).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4      [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4      [46,48) version = 1 (Blackwell)
//   [49,52) base offset = 0          [61,64) layout: 0 none, 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B
//           1 = SWIZZLE_128B_BASE32B (128 B span, 32-byte swizzle atoms)
// K-major, SWIZZLE_128B (rows of 128 B = 32 tf32 along K): SBO = stride between 8-row groups (1024 B when
//   dense), LBO unused (1).
// MN-major tf32 operands must use SWIZZLE_128B_BASE32B (the only layout the tensor core accepts for 32-bit
//   MN-major data; TMA counterpart CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 128 B = 32 elements along
//   M/N, one row per k, swizzle period 4 rows.  LBO = stride between 32-element M/N groups,
//   SBO = stride between 4-k groups (512 B when dense).
constexpr uint32_t kSmemLayoutSw128 = 2;
constexpr uint32_t kSmemLayoutSw128Base32 = 1;
__host__ __device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                             uint32_t layout_type, uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>(base_offset & 7u) << 49;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // version
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;
  return d;
}
// Instruction descriptor (32 bit) for kind::tf32, fp32 accumulate:
//   [4,6) c_format = 1 (F32)   [7,10) a_format = 2 (TF32)   [10,13) b_format = 2 (TF32)
//   [15] a_major (0 = K, 1 = MN)   [16] b_major   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

}  // namespace ptx
}  // namespace zb
