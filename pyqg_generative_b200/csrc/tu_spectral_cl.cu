// tu_spectral_cl.cu -- translation unit of the cluster (DSMEM) register-FFT spectral step for nx = 128, 256 (spectral_cl.cuh)
#define S64_HELPERS_ONLY
#include "spectral_cl.cuh"

#include "spectral_host.hpp"

#include <cstdlib>

namespace qgb {

namespace {
template <int N, int G, int CL, class K, class... A>
cudaError_t launch_cl_kernel(K kern, int members, cudaStream_t st, A... args);
template <int N, int G, int CL>
cudaError_t launch_cl(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st) {
  if (prog == PROG_BUDGET) return launch_cl_kernel<N, G, CL>(scl::qg_budget_cl_kernel<N, G, CL>, members, st, T, io, members);
  return launch_cl_kernel<N, G, CL>(scl::qg_step_cl_kernel<N, G, CL>, members, st, T, io, prog, members);
}
template <int N, int G, int CL, class K, class... A>
cudaError_t launch_cl_kernel(K kern, int members, cudaStream_t st, A... args) {
  using C = scl::Cfg<N, G, CL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);   // per device: set every time (cheap)
  if (e != cudaSuccess) return e;
  if (CL > 8) {   // 16 CTAs per cluster is the opt-in (non-portable) size
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(members * CL);
  cfg.blockDim = dim3(C::kThreads);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
}  // namespace

bool spectralcl_handles(int N, int prog) {
  return (N == 128 || N == 256) &&
         (prog == PROG_STEP || prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW || prog == PROG_SET_Q || prog == PROG_C2R || prog == PROG_ADVECT || prog == PROG_INVERT || prog == PROG_BUDGET);
}

cudaError_t spectralcl_launch(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st) {
  // CTAs per member.  256^2: 16 CTAs of 256 threads, two per SM (normally of two different members, whose cluster barriers and
  // DRAM waits overlap): +9 % over 8 CTAs of 512 threads on the same box (95.3 k vs 87.6 k member-steps/s, 64 members).  128^2: 2 CTAs
  // of 512 threads (4 x 256 measured 5 % slower at 64 members, equal at 256).  QGB_SCL_ALT=1 selects the other geometry of each.
  static const bool alt = getenv("QGB_SCL_ALT") != nullptr;
  if (T.N == 128) return alt ? launch_cl<128, 8, 4>(T, io, prog, members, st) : launch_cl<128, 8, 2>(T, io, prog, members, st);
  return alt ? launch_cl<256, 16, 8>(T, io, prog, members, st) : launch_cl<256, 16, 16>(T, io, prog, members, st);
}

}  // namespace qgb
