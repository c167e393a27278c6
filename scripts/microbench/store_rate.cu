// Microbenchmark: global-store throughput of ONE persistent CTA per SM as a function of the number of storing warps and of
// the store shape (the epilogue pattern: 32 B per lane at a 64 B pixel stride, two instructions per pixel).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_v8(void* p, uint32_t v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_v4(void* p, uint32_t v) {
  asm volatile("st.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}

// each warp writes "rows" of 32 pixels x 64 B (2 KB per row) round-robin over a large buffer
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(unsigned char* out, long long rows_total, int nwarps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nwarps) return;
  const long long gw = (long long)blockIdx.x * nwarps + warp, nw = (long long)gridDim.x * nwarps;
  for (long long r = gw; r < rows_total; r += nw) {
    unsigned char* p = out + r * 2048 + lane * 64;
    if (MODE == 0) { st_v8(p, (uint32_t)r); st_v8(p + 32, (uint32_t)r); }                       // epilogue pattern
    if (MODE == 1) { st_v4(p, (uint32_t)r); st_v4(p + 16, (uint32_t)r); st_v4(p + 32, (uint32_t)r); st_v4(p + 48, (uint32_t)r); }
    if (MODE == 2) { st_v8(out + r * 2048 + lane * 32, (uint32_t)r); st_v8(out + r * 2048 + 1024 + lane * 32, (uint32_t)r); }   // contiguous lanes
  }
}

template <int MODE>
void run(const char* name, unsigned char* buf, long long bytes, int nwarps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const long long rows = bytes / 2048;
  k<MODE><<<148, 1024>>>(buf, rows, nwarps);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k<MODE><<<148, 1024>>>(buf, rows, nwarps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-34s warps/SM=%2d : %6.2f TB/s\n", name, nwarps, 3.0 * bytes / (ms * 1e-3) / 1e12);
}

int main() {
  const long long bytes = 2ll << 30;
  unsigned char* buf; cudaMalloc(&buf, bytes);
  for (int nw : {4, 8, 16, 32}) {
    run<0>("2 x st.v8 per lane, 64 B stride", buf, bytes, nw);
    run<1>("4 x st.v4 per lane, 64 B stride", buf, bytes, nw);
    run<2>("2 x st.v8, lanes contiguous", buf, bytes, nw);
  }
  return 0;
}
