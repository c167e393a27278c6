// tu_spectral.cu -- translation unit of the spectral phase-program kernels (spectral.cuh, qg_core.cuh) and their launch code
#include "spectral.cuh"

namespace qgb {

cudaError_t spectral_configure(const SpectralPlan& p) {
  cudaError_t e = cudaSuccess;
#define QGB_ATTR(kern, bytes)                                                                            \
  do {                                                                                                   \
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));          \
    if (e != cudaSuccess) return e;                                                                      \
  } while (0)
  if (p.large) {
    switch (p.N) {
      case 128: QGB_ATTR(qg_program_cluster_kernel<128>, p.large_smem); break;
      case 256: QGB_ATTR(qg_program_cluster_kernel<256>, p.large_smem); break;
      case 512: QGB_ATTR(qg_program_cluster_kernel<512>, p.large_smem); break;
      case 1024: QGB_ATTR(qg_program_cluster_kernel<1024>, p.large_smem); break;
      default: QGB_ATTR(qg_program_cluster_kernel<0>, p.large_smem); break;
    }
    return e;
  }
  // handles of several grid sizes coexist on a device: only ever raise the generic kernel's limit
  static size_t generic_limit[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || p.smem > generic_limit[dev]) {
    QGB_ATTR(qg_program_kernel, p.smem);
    if (dev >= 0 && dev < 64) generic_limit[dev] = p.smem;
  }
  if (p.reg64) {
    e = spectral64_configure();
    if (e != cudaSuccess) return e;
  }
  if (p.fixed) {
    if (p.N == 32) QGB_ATTR((qg_step_fixed_kernel<32, 256>), p.smem);
    if (p.N == 48) QGB_ATTR((qg_step_fixed_kernel<48, 256>), p.smem);
    if (p.N == 64) {
      QGB_ATTR((qg_step_fixed_kernel<64, 256>), p.smem);
      QGB_ATTR((qg_step_fixed_kernel<64, 384>), p.smem);
      QGB_ATTR((qg_step_fixed_kernel<64, 512>), p.smem);
    }
    if (p.N == 96) QGB_ATTR((qg_step_fixed_kernel<96, 512>), p.smem);
    if (p.N == 32) QGB_ATTR((qg_program_fixed_kernel<32, 256>), p.smem);
    if (p.N == 48) QGB_ATTR((qg_program_fixed_kernel<48, 256>), p.smem);
    if (p.N == 96) QGB_ATTR((qg_program_fixed_kernel<96, 512>), p.smem);
    if (p.N == 32) QGB_ATTR((qg_budget_fixed_kernel<32, 256>), p.smem);
    if (p.N == 48) QGB_ATTR((qg_budget_fixed_kernel<48, 256>), p.smem);
    if (p.N == 96) QGB_ATTR((qg_budget_fixed_kernel<96, 512>), p.smem);
  }
#undef QGB_ATTR
  return e;
}

cudaError_t spectral_launch(const SpectralPlan& p, const Tables& TT, const StepIO& io, int prog, cudaStream_t st) {
  if (p.regcl && spectralcl_handles(p.N, prog)) return spectralcl_launch(TT, io, prog, p.members, st);
  if (p.large) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.grid);
    cfg.blockDim = dim3(p.nthreads);
    cfg.dynamicSmemBytes = p.large_smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // compile-time grid size for the power-of-two production sizes (index arithmetic folds to shifts), generic otherwise
    auto kern = p.N == 128 ? qg_program_cluster_kernel<128> : p.N == 256 ? qg_program_cluster_kernel<256>
              : p.N == 512 ? qg_program_cluster_kernel<512> : p.N == 1024 ? qg_program_cluster_kernel<1024>
                                                                            : qg_program_cluster_kernel<0>;
    return cudaLaunchKernelEx(&cfg, kern, TT, io, prog, p.members, p.scratch, p.red_scratch, p.large_lines, p.true_pos);
  }
  if (p.reg64 && spectral64_handles(prog)) return spectral64_launch(TT, io, prog, p.members, st);
  const bool is_step = prog == PROG_STEP || prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW;
  if (is_step && p.fixed) {
    switch (p.N) {
      case 32: qg_step_fixed_kernel<32, 256><<<p.grid, 256, p.smem, st>>>(TT, io, prog, p.members); break;
      case 48: qg_step_fixed_kernel<48, 256><<<p.grid, 256, p.smem, st>>>(TT, io, prog, p.members); break;
      case 64:
        if (p.nt64 == 512) qg_step_fixed_kernel<64, 512><<<p.grid, 512, p.smem, st>>>(TT, io, prog, p.members);
        else if (p.nt64 == 384) qg_step_fixed_kernel<64, 384><<<p.grid, 384, p.smem, st>>>(TT, io, prog, p.members);
        else qg_step_fixed_kernel<64, 256><<<p.grid, 256, p.smem, st>>>(TT, io, prog, p.members);
        break;
      default: qg_step_fixed_kernel<96, 512><<<p.grid, 512, p.smem, st>>>(TT, io, prog, p.members); break;
    }
    return cudaGetLastError();
  }
  if (prog == PROG_BUDGET && p.fixed && p.N != 64) {
    if (p.N == 32) qg_budget_fixed_kernel<32, 256><<<p.grid, 256, p.smem, st>>>(TT, io, p.members);
    else if (p.N == 48) qg_budget_fixed_kernel<48, 256><<<p.grid, 256, p.smem, st>>>(TT, io, p.members);
    else qg_budget_fixed_kernel<96, 512><<<p.grid, 512, p.smem, st>>>(TT, io, p.members);
    return cudaGetLastError();
  }
  if ((prog == PROG_SET_Q || prog == PROG_C2R || prog == PROG_INVERT || prog == PROG_ADVECT) && p.fixed && p.N != 64) {
    if (p.N == 32) qg_program_fixed_kernel<32, 256><<<p.grid, 256, p.smem, st>>>(TT, io, prog, p.members);
    else if (p.N == 48) qg_program_fixed_kernel<48, 256><<<p.grid, 256, p.smem, st>>>(TT, io, prog, p.members);
    else qg_program_fixed_kernel<96, 512><<<p.grid, 512, p.smem, st>>>(TT, io, prog, p.members);
    return cudaGetLastError();
  }
  qg_program_kernel<<<p.grid, p.nthreads, p.smem, st>>>(TT, io, prog, p.members);
  return cudaGetLastError();
}

cudaError_t launch_diag_finish(const double* red, int members, double dt_over_dx, double* ke, double* cfl, int* flags, cudaStream_t st) {
  diag_finish_kernel<<<(members + 127) / 128, 128, 0, st>>>(red, members, dt_over_dx, ke, cfl, flags);
  return cudaGetLastError();
}

cudaError_t launch_spectra(const Tables& T, const cplx* qh, int members, double* kespec, double* ensspec, cudaStream_t st) {
  const int n = 2 * T.N * T.NK;
  spectra_kernel<<<(n + 31) / 32, dim3(32, kSpectraParts), 0, st>>>(T, qh, members, kespec, ensspec);
  return cudaGetLastError();
}

}  // namespace qgb
