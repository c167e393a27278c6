// tu_spectral64.cu -- translation unit of the register-resident 64^2 spectral step (spectral64.cuh)
#include "spectral64.cuh"

#include "spectral_host.hpp"

namespace qgb {

cudaError_t spectral64_configure() {
  cudaError_t e = cudaFuncSetAttribute(s64::qg_budget64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s64::kSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(s64::qg_step64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s64::kSmemBytes);
}

bool spectral64_handles(int prog) {
  return prog == PROG_STEP || prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW || prog == PROG_SET_Q || prog == PROG_C2R || prog == PROG_ADVECT || prog == PROG_INVERT ||
         prog == PROG_BUDGET;
}

cudaError_t spectral64_launch(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st) {
  if (prog == PROG_BUDGET) s64::qg_budget64_kernel<<<members, s64::kThreads, s64::kSmemBytes, st>>>(T, io, members);
  else s64::qg_step64_kernel<<<members, s64::kThreads, s64::kSmemBytes, st>>>(T, io, prog, members);
  return cudaGetLastError();
}

}  // namespace qgb
