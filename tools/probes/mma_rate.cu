// tcgen05.mma kind::tf32 issue-rate probe: one CTA per SM, operands resident in smem (contents irrelevant: zeros), one thread issues
// `iters` MMAs back to back and commits; we time from first issue to commit arrival with clock64.  Varies N, the operand layout
// (K-major SWIZZLE_128B, the 4 K-steps of a 128-byte row issued consecutively, vs MN-major SWIZZLE_128B_BASE32B), and whether
// consecutive MMAs hit the same TMEM accumulator or rotate over several.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout & 7u) << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int m, int n, int amn, int bmn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(amn) << 15) | (uint32_t(bmn) << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}

// mode: 0 = K-major, 4 K-steps per 128-byte row group, same accumulator; 1 = K-major, rotating over `nacc` accumulators;
//       2 = MN-major, same accumulator; 3 = MN-major rotating; 4 = K-major but every MMA re-reads K-step 0 (same 32 bytes of each row)
//       6 = K-major, 8 MMAs per loop trip with compile-time descriptor offsets (how few cycles can ONE thread spend per MMA?)
//       5 = K-major, the A descriptor starts 0/1/2/58/59/60/116/117/118 rows (128 bytes each) into the tile, changing every 4 MMAs:
//           the shifted-descriptor taps of the halo-reuse conv kernels (3x3 filter over a 58-pixel-wide raster)
__global__ void __launch_bounds__(128, 1) probe(int n, int mode, int nacc, int iters, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    const bool mn = mode == 2 || mode == 3;
    const uint32_t a0 = su32(smem), b0 = su32(smem) + (mode == 5 ? 32768 : 16384);
    const uint32_t taps[9] = {0, 1, 2, 58, 59, 60, 116, 117, 118};
    const uint64_t da0 = mn ? desc(a0, 4096, 512, 1) : desc(a0, 16, 1024, 2);
    const uint64_t db0 = mn ? desc(b0, 4096, 512, 1) : desc(b0, 16, 1024, 2);
    const uint32_t kstep = mn ? 64 : 2;
    const uint32_t id = idesc_tf32(128, n, mn, mn);
    const bool rotate = mode == 1 || mode == 3;
    long long t0 = clock64();
    if (mode == 6) {   // the leanest possible issue stream: 8 MMAs per loop trip, descriptors are compile-time offsets from two registers
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mma(tm, da0 + (j & 3) * 2, db0 + (j & 3) * 2, id, (i + j) > 0 ? 1u : 0u);
      }
    } else
    for (int i = 0; i < iters; ++i) {
      const uint32_t k = mode == 4 ? 0 : (i & 3);
      const uint32_t d = tm + (rotate ? (i % nacc) * n : 0);
      const uint32_t shift = mode == 5 ? taps[(i >> 2) % 9] * 8u : 0u;
      mma(d, da0 + shift + k * kstep, db0 + k * kstep, id, i >= nacc ? 1u : 0u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar)) : "memory");
    asm volatile("{\n.reg .pred P;\nW: mbarrier.try_wait.parity.shared::cta.b64 P, [%0], 0;\n@P bra D;\nbra W;\nD:\n}" ::"r"(su32(&bar)) : "memory");
    cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  CK(cudaMalloc(&d, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  const int iters = 4096;
  const char* names[] = {"K-major SW128, same accumulator", "K-major SW128, rotating accumulators", "MN-major SW128/32B, same accumulator",
                         "MN-major SW128/32B, rotating accumulators", "K-major, every MMA reads K-step 0",
                         "K-major, A start shifted by 3x3 halo taps", "K-major, unrolled x8, constant descriptor offsets"};
  for (int grid : {1, sms}) {
    for (int n : {32, 64, 96, 128, 256}) {
      for (int mode = 0; mode < 7; ++mode) {
        const int nacc = (mode == 1 || mode == 3) ? (512 / n >= 4 ? 4 : 512 / n) : 1;
        probe<<<grid, 128, 80 * 1024>>>(n, mode, nacc, iters, d);
        CK(cudaDeviceSynchronize());
        long long h[256];
        CK(cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per = double(mx) / iters;
        printf("grid %3d  N %3d  %-45s %7.1f clk/MMA   (%5.1f %% of the N/2-cycle tensor rate)\n", grid, n, names[mode], per, 100.0 * (n / 2.0) / per);
      }
    }
  }
  return 0;
}
