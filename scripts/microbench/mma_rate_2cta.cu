// Microbenchmark: cycles per tcgen05.mma with cta_group::2 (M=256 over a CTA pair, each CTA holds 128 rows of A and N/2
// rows of B) against the cta_group::1 figure of mma_rate.cu.  Both operands in shared memory, SWIZZLE_NONE K-major
// canonical layouts, kind::f16 (K=16) and kind::f8f6f4 (K=32: same operand bytes).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pyqg_generative_b200/csrc/cnn_tc.cuh"
using namespace qgb;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_f16(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_f8(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   ptx::smem_u32(bar)), "h"(mask) : "memory");
}

// N = total N of the pair MMA (each CTA stores N/2 rows of B)
template <int N, int F8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(long long* out, int iters) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (ptx::smem_u32(sm) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = ctarank();
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc2(&slot, 512);
  ptx::tc_fence_before(); __syncthreads(); cluster_sync_all(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    const uint32_t a16 = ptx::smem_u32(base) >> 4, b16 = (ptx::smem_u32(base) + 40960) >> 4;
    // instruction descriptor: M = 256 over the pair; e4m3 a/b formats are 0 for f8f6f4, f16 formats 0 for kind::f16
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    // no swizzle: 8 rows x 16 B core matrices, SBO 128 B... A: [K chunk][128 rows][16 B]: LBO = 2048 B, SBO = 128 B
    const uint32_t a_hi = 8u | (1u << 14), a_lbo = 128u;        // 16 B units
    const uint32_t b_hi = 8u | (1u << 14), b_lbo = (N / 2) * 16u / 16u;
    __syncwarp();
    t0 = clock64();
    if (rank == 0 && ptx::elect_one_sync()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | ((a16 + (u & 3) * 256) & 0x3FFF) | (a_lbo << 16);
          const uint64_t bd = ((uint64_t)b_hi << 32) | (b16 + (u & 3) * 2 * b_lbo) | (b_lbo << 16);
          if (F8) mma2_f8(tm + (u % 4) * N, ad, bd, idesc, 1u);
          else mma2_f16(tm + (u % 4) * N, ad, bd, idesc, 1u);
        }
      }
      commit2(&bar, 3);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    t1 = clock64();
  }
  ptx::tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc2(tm, 512);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int N, int F8>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  auto kern = k2<N, F8>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  kern<<<148, 128, 100 * 1024>>>(d, iters);
  kern<<<148, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-26s pair M=256 N=%3d : %7.1f clk/MMA  (math floor %d, 1-CTA equivalent 2 x max(N/2,(128+N)/4) = %d) %s\n", name, N,
         (double)h / (iters * 8.0), N / 2, 2 * ((N / 2) > (128 + N) / 4 ? N / 2 : (128 + N) / 4),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 0>("f16 K=16");
  run<128, 0>("f16 K=16");
  run<32, 0>("f16 K=16");
  run<256, 0>("f16 K=16");
  run<64, 1>("e4m3 K=32");
  run<128, 1>("e4m3 K=32");
  run<32, 1>("e4m3 K=32");
  return 0;
}
