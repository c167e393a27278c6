// Microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) as a function of N, of the A-operand
// shared-memory layout (no swizzle 16 B rows / 64 B swizzle / 128 B swizzle) and of the accumulator pattern.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../pyqg_generative_b200/csrc/cnn_tc.cuh"
using namespace qgb;

template <int N, int LAYOUT, int NACC, int COMMIT_EVERY = 0>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int a_shift16) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* base = sm + ((1024u - (ptx::smem_u32(sm) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bars2[8];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) ptx::mbar_init(&bars2[i], 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc(&slot, 512);
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x < 32) {
    const uint32_t a16 = (ptx::smem_u32(base) >> 4) + a_shift16, b16 = (ptx::smem_u32(base) + 40960) >> 4;
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    uint32_t a_hi, a_lo_extra;
    if (LAYOUT == 0) { a_hi = 36u | (1u << 14); a_lo_extra = (720u << 16); }            // no swizzle: SBO 576 B, LBO 11520 B
    else if (LAYOUT == 1) { a_hi = 144u | (1u << 14) | (4u << 29); a_lo_extra = (1u << 16); }   // SW64: SBO 36*64 B
    else { a_hi = 64u | (1u << 14) | (2u << 29); a_lo_extra = (1u << 16); }            // SW128: SBO 1024 B
    const uint32_t b_hi = 8u | (1u << 14);
    const uint32_t b_lbo = N;  // 16 B units
    __syncwarp();
    t0 = clock64();
    if (ptx::elect_one_sync()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | ((a16 + (u & 3) * 32 + (u >> 2) * 2) & 0x3FFF) | a_lo_extra;
          const uint64_t bd = ((uint64_t)b_hi << 32) | (b16 + (u >> 2) * 2 * b_lbo) | (b_lbo << 16);
          ptx::mma_f16(tm + (u % NACC) * N, ad, bd, idesc, 1u);
          if (COMMIT_EVERY && (u + 1) % COMMIT_EVERY == 0) ptx::tc_commit(&bars2[(it * 8 + u) / COMMIT_EVERY % 8]);
        }
      }
      ptx::tc_commit(&bar);
    }
    __syncwarp();
    ptx::mbar_wait(&bar, 0);
    t1 = clock64();
  }
  ptx::tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int N, int LAYOUT, int NACC, int CE = 0>
void run(const char* name, int shift) {
  long long* d; cudaMalloc(&d, 8);
  auto kern = k<N, LAYOUT, NACC, CE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  kern<<<148, 128, 100 * 1024>>>(d, iters, shift);
  kern<<<148, 128, 100 * 1024>>>(d, iters, shift);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d nacc=%d shift16=%2d : %7.1f clk/MMA  (floor %d) %s\n", name, N, NACC, shift, (double)h / (iters * 8.0), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 0, 4>("noswizzle aligned", 0);
  run<64, 0, 4>("noswizzle shifted 16B", 1);
  run<64, 1, 4>("sw64 aligned", 0);
  run<64, 1, 4>("sw64 shifted 64B", 4);
  run<64, 2, 4>("sw128 aligned", 0);
  run<64, 2, 4>("sw128 shifted 128B", 8);
  run<64, 1, 1>("sw64 same accumulator", 0);
  run<32, 1, 4>("sw64", 0);
  run<128, 1, 4>("sw64", 0);
  run<256, 1, 2>("sw64", 0);
  run<16, 1, 4>("sw64", 0);
  run<64, 1, 4, 8>("sw64 commit every 8", 0);
  run<64, 1, 4, 4>("sw64 commit every 4", 0);
  run<64, 1, 4, 2>("sw64 commit every 2", 0);
  run<64, 1, 4, 1>("sw64 commit every 1", 0);
  run<32, 1, 4, 8>("sw64 commit every 8", 0);
  return 0;
}
