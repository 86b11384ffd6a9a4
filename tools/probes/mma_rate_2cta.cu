// tcgen05.mma.cta_group::2 kind::tf32 issue-rate probe (companion of mma_rate.cu): clusters of 2 CTAs, operands resident (zeros) in
// both shared memories, the leader CTA issues `iters` M256 x N x K8 MMAs back to back on one accumulator and commits (multicast)
// to both CTAs.  Prints cycles per MMA: with the same ~77-cycle floor per instruction, a pair does twice the work of a
// cta_group::1 MMA per instruction.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_2cta mma_rate_2cta.cu
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
__device__ __forceinline__ uint32_t idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shifted != 0: the A descriptor starts 0/1/2/58/59/60/116/117/118 rows into the tile, changing every 4 MMAs (halo-kernel taps)
__global__ void __launch_bounds__(128, 1) probe(int n, int iters, long long* cycles, int* status, int shifted) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (64 * 1024) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(su32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  long long t0 = clock64();
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t a0 = su32(smem), b0 = su32(smem) + (shifted ? 32768 : 16384);
    const uint32_t taps[9] = {0, 1, 2, 58, 59, 60, 116, 117, 118};
    const uint64_t da0 = desc(a0, 16, 1024, 2), db0 = desc(b0, 16, 1024, 2);
    const uint32_t id = idesc_tf32(256, n);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t k = i & 3;
      const uint32_t shift = shifted ? taps[(i >> 2) % 9] * 8u : 0u;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tm), "l"(da0 + shift + k * 2),
                   "l"(db0 + k * 2), "r"(id), "r"(i > 0 ? 1u : 0u)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(su32(&bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  }
  if (threadIdx.x == 0) {   // both CTAs wait for the commit (bounded: a protocol mistake must not hang the box)
    bool done = false;
    const long long w0 = clock64();
    while (!done && clock64() - w0 < 2000000000ll) {
      uint32_t ok;
      asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(ok) : "r"(su32(&bar)) : "memory");
      done = ok != 0;
    }
    if (rank == 0) cycles[blockIdx.x / 2] = clock64() - t0;
    if (!done) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  int* st;
  CK(cudaMalloc(&d, sizeof(long long) * sms));
  CK(cudaMalloc(&st, sizeof(int)));
  CK(cudaMemset(st, 0, sizeof(int)));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  const int iters = 4096;
  for (int shifted : {0, 1})
  for (int grid : {2, sms / 2 * 2}) {
    for (int n : {32, 64, 128, 256}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = 80 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, probe, n, iters, d, st, shifted));
      CK(cudaDeviceSynchronize());
      long long h[128];
      int hs = 0;
      CK(cudaMemcpy(h, d, sizeof(long long) * (grid / 2), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&hs, st, sizeof(int), cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (int i = 0; i < grid / 2; ++i) mx = h[i] > mx ? h[i] : mx;
      const double per = double(mx) / iters;
      printf("grid %3d  M 256 (cta_group::2)  N %3d  %s %7.1f clk/MMA   (%5.1f %% of the pair's tensor rate)%s\n", grid, n,
             shifted ? "A shifted by 3x3 halo taps" : "aligned operands          ", per, 100.0 * (n / 2.0) / per, hs ? "  [TIMEOUT]" : "");
    }
  }
  return 0;
}
