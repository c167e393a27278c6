// tu_spectral_cl.cu -- translation unit of the cluster (DSMEM) register-FFT spectral step for nx = 128, 256 (spectral_cl.cuh)
#define S64_HELPERS_ONLY
#include "spectral_cl.cuh"

#include "spectral_host.hpp"

namespace qgb {

namespace {
template <int N, int G, int CL>
cudaError_t launch_cl(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st) {
  using C = scl::Cfg<N, G, CL>;
  auto kern = scl::qg_step_cl_kernel<N, G, CL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);   // per device: set every time (cheap)
  if (e != cudaSuccess) return e;
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
  return cudaLaunchKernelEx(&cfg, kern, T, io, prog, members);
}
}  // namespace

bool spectralcl_handles(int N, int prog) {
  return (N == 128 || N == 256) &&
         (prog == PROG_STEP || prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW || prog == PROG_SET_Q || prog == PROG_C2R);
}

cudaError_t spectralcl_launch(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st) {
  if (T.N == 128) return launch_cl<128, 8, 2>(T, io, prog, members, st);
  return launch_cl<256, 16, 8>(T, io, prog, members, st);
}

}  // namespace qgb
