// TMA load-throughput probe (no consumer work): one CTA per SM streams boxes into a smem ring and we report bytes/clk/SM
// for several global-memory access shapes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t su32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred P;\nW: mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n@P bra D;\nbra W;\nD:\n}" ::"r"(su32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(su32(dst)), "l"(m), "r"(su32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(su32(dst)), "l"(m), "r"(su32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(dst)), "l"(src), "r"(bytes), "r"(su32(bar)) : "memory");
}

constexpr int STAGES = 4;
// mode 0: 2-D tiled boxes {32, rows}; mode 1: 1-D bulk copies; mode 2: 4-D halo boxes
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, const float* src, int mode, int rows, int iters,
                                                int outer_extent, int box_bytes, int w, int hh, int nimg, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[STAGES];
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    uint32_t ph = 0;
    int s = 0;
    const int stage_bytes = 49152;
    for (int it = 0; it < iters + STAGES; ++it) {
      if (it >= STAGES) wait(&bar[s], ph);   // the load issued STAGES iterations ago
      if (it < iters) {
        expect_tx(&bar[s], box_bytes);
        const long long idx = (static_cast<long long>(it) * gridDim.x + blockIdx.x);
        if (mode == 0) {
          tma2d(smem + s * stage_bytes, &tm, &bar[s], 0, static_cast<int>((idx * rows) % (outer_extent - rows)));
        } else if (mode == 1) {
          bulk1d(smem + s * stage_bytes, reinterpret_cast<const uint8_t*>(src) + (idx * box_bytes) % (static_cast<long long>(outer_extent) * 128 - box_bytes), box_bytes, &bar[s]);
        } else {
          const int n = static_cast<int>(idx % nimg), h0 = static_cast<int>((idx / nimg) * 2 % (hh - 4));
          tma4d(smem + s * stage_bytes, &tm, &bar[s], 0, -1, h0, n);
        }
      }
      if (++s == STAGES) { s = 0; if (it >= STAGES - 1) ph ^= (it >= STAGES) ? 1 : 0; }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncFn enc = reinterpret_cast<EncFn>(fn);
  const size_t bytes = size_t(1) << 30;  // 1 GiB source (larger than L2) and a 64 MiB window (L2 resident) are both tried
  float* src;
  CK(cudaMalloc(&src, bytes));
  CK(cudaMemset(src, 0, bytes));
  long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * 49152 + 1024));
  struct Case { const char* name; int mode; int rows; long long pitch; long long window_bytes; };
  std::vector<Case> cases = {
      {"2D box 32x256, pitch 128 B (contiguous), 1 GiB", 0, 256, 128, bytes},
      {"2D box 32x256, pitch 128 B (contiguous), 64 MiB (L2)", 0, 256, 128, 64 << 20},
      {"2D box 32x256, pitch 4608 B (filter rows), 64 MiB (L2)", 0, 256, 4608, 64 << 20},
      {"2D box 32x256, pitch 256 B, 64 MiB (L2)", 0, 256, 256, 64 << 20},
      {"2D box 32x128, pitch 1024 B, 1 GiB", 0, 128, 1024, bytes},
      {"2D box 32x128, pitch 1024 B, 64 MiB (L2)", 0, 128, 1024, 64 << 20},
      {"1D bulk 32 KiB, 1 GiB", 1, 256, 128, bytes},
      {"1D bulk 32 KiB, 64 MiB (L2)", 1, 256, 128, 64 << 20},
      {"1D bulk 16 KiB, 64 MiB (L2)", 1, 128, 128, 64 << 20},
      {"4D halo box 32x58x4 of NHWC C=64 56x56, 205 MB", 2, 232, 256, 0},
  };
  for (auto& c : cases) {
    CUtensorMap tm;
    int outer = 0, box_bytes = c.rows * 128, w = 0, hh = 0, nimg = 0;
    if (c.mode == 2) {
      const int C = 64, W = 56, H = 56, N = 256;
      cuuint64_t dims[4] = {C, W, H, N};
      cuuint64_t strides[3] = {C * 4ull, W * C * 4ull, H * W * C * 4ull};
      cuuint32_t box[4] = {32, 58, 4, 1}, es[4] = {1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); return 1; }
      w = W; hh = H; nimg = N;
    } else {
      outer = static_cast<int>(c.window_bytes / c.pitch);
      cuuint64_t dims[2] = {32, static_cast<cuuint64_t>(outer)};
      cuuint64_t strides[1] = {static_cast<cuuint64_t>(c.pitch)};
      cuuint32_t box[2] = {32, static_cast<cuuint32_t>(c.rows)}, es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); return 1; }
      if (c.mode == 1) outer = static_cast<int>(c.window_bytes / 128);
    }
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      probe<<<sms, 128, STAGES * 49152 + 1024>>>(tm, src, c.mode, c.rows, iters, outer, box_bytes, w, hh, nimg, cyc);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(sms);
      CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto v : h) mx = v > mx ? v : mx;
      const double total = double(iters) * sms * box_bytes;
      if (rep == 1)
        printf("%-62s %7.1f B/clk/SM  %6.2f TB/s  (%d-byte boxes, %.0f cycles/box, %d in flight)\n", c.name, double(iters) * box_bytes / mx, total / ms / 1e9,
               box_bytes, double(mx) / iters, STAGES);
    }
  }
  return 0;
}
