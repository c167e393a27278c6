// Microbenchmark: the two peaks BASELINE.md section 3 leaves to the builder --
//   (1) vector FP64 DFMA throughput of the whole chip (the co-roofline of the spectral-step kernels),
//   (2) dense tcgen05 throughput for kind::f16 and kind::tf32 (M=128, N=256, both operands in shared memory), whole chip,
// timed with CUDA events after a warm-up launch.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pyqg_generative_b200/csrc/cnn_tc.cuh"
using namespace qgb;

// ---- (1) DFMA: 8 independent chains per thread, 16 x unrolled ----
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;     // keeps the chains alive
}

// ---- (2) tcgen05 ----
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// KIND 0: f16 (K = 16 per instruction), 1: tf32 (K = 8 per instruction; a/b format field = 2)
template <int N, int KIND>
__global__ void __launch_bounds__(128, 1) mma_kernel(int iters) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (ptx::smem_u32(sm) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    const uint32_t a16 = ptx::smem_u32(base) >> 4, b16 = (ptx::smem_u32(base) + 32768) >> 4;
    constexpr uint32_t fmt = KIND == 1 ? 2u : 0u;
    constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = 64u | (1u << 14) | (2u << 29);     // SWIZZLE_128B, SBO 1024 B
    const uint32_t b_hi = 64u | (1u << 14) | (2u << 29);
    if (ptx::elect_one_sync()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | ((a16 + (u & 3) * 2) & 0x3FFF) | (1u << 16);
          const uint64_t bd = ((uint64_t)b_hi << 32) | ((b16 + (u & 3) * 2) & 0x3FFF) | (1u << 16);
          if (KIND == 1) mma_tf32(tm + (u & 1) * N, ad, bd, idesc, 1u);
          else ptx::mma_f16(tm + (u & 1) * N, ad, bd, idesc, 1u);
        }
      }
      ptx::tc_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

template <int N, int KIND>
void run_mma(const char* name, int nsm) {
  auto kern = mma_kernel<N, KIND>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<nsm, 128, 100 * 1024>>>(iters / 10);
  cudaEventRecord(e0);
  kern<<<nsm, 128, 100 * 1024>>>(iters);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double K = KIND == 1 ? 8.0 : 16.0;
  const double flops = 2.0 * 128 * N * K * 8.0 * iters * nsm;
  printf("tcgen05 %-5s M=128 N=%3d K=%2.0f, %d SMs: %8.1f TFLOP/s  (%.3f ms) %s\n", name, N, K, nsm, flops / (ms * 1e-3) / 1e12, ms,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  double* d; cudaMalloc(&d, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks_per_sm : {4, 8}) {
    const int iters = 20000;
    dfma_kernel<<<nsm * blocks_per_sm, 256>>>(d, iters / 10, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<<<nsm * blocks_per_sm, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8 * 16 * (double)iters * 256.0 * nsm * blocks_per_sm;
    printf("DFMA  %d CTAs/SM x 256 threads, 8 chains: %7.2f TFLOP/s fp64  (%.3f ms)\n", blocks_per_sm, flops / (ms * 1e-3) / 1e12, ms);
  }
  run_mma<256, 0>("f16", nsm);
  run_mma<128, 0>("f16", nsm);
  run_mma<256, 1>("tf32", nsm);
  run_mma<128, 1>("tf32", nsm);
  return 0;
}
