// spectral_host.hpp -- host-side interface of the spectral phase-program kernels (spectral.cuh, compiled in tu_spectral.cu).
#pragma once
#include <cuda_runtime.h>

#include "qg_core.cuh"

namespace qgb {

constexpr int kMaxClusterSize = 8;   // portable limit; the size is chosen per handle (4 or 2 for N <= 128, 8 above)

inline size_t program_smem_bytes(int N, int P, int nthreads) {
  size_t b = (size_t)N * P * sizeof(cplx);  // buf
  b += (size_t)N * sizeof(cplx);            // twiddles
  b += ((size_t)N * sizeof(short) + 15) / 16 * 16;
  b += 4 * (size_t)nthreads * sizeof(double);  // reduction scratch
  return b;
}

// Launch geometry of the phase-program kernels, chosen per handle in qgb_create (api.cu)
struct SpectralPlan {
  int N = 0, members = 0, grid = 0, nthreads = 256;
  size_t smem = 0;
  bool fixed = false;       // compile-time specialised step kernel available for this nx
  int nt64 = 384;
  bool regcl = false;       // nx = 128, 256: cluster register-FFT kernel (spectral_cl.cuh) for the same programs
  bool reg64 = false;       // nx = 64: register-resident kernel (spectral64.cuh) for the step, the q setter and irfft2
  bool large = false;       // thread-block-cluster path (the packed field does not fit one CTA)
  int cluster = 8, large_lines = 0;
  size_t large_smem = 0;
  cplx* scratch = nullptr;
  double* red_scratch = nullptr;
  const short* true_pos = nullptr;
};

// Raise the opt-in dynamic shared-memory limits of the kernels this plan launches, on the CURRENT device (the attribute is
// per device / context, so it is set for every handle).
cudaError_t spectral_configure(const SpectralPlan& p);
cudaError_t spectral_launch(const SpectralPlan& p, const Tables& T, const StepIO& io, int prog, cudaStream_t st);
// register-resident 64^2 kernel (tu_spectral64.cu)
cudaError_t spectral64_configure();
bool spectral64_handles(int prog);
cudaError_t spectral64_launch(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st);
// cluster (DSMEM) register-FFT kernel for nx = 128, 256 (tu_spectral_cl.cu)
bool spectralcl_handles(int N, int prog);
cudaError_t spectralcl_launch(const Tables& T, const StepIO& io, int prog, int members, cudaStream_t st);
cudaError_t launch_diag_finish(const double* red, int members, double dt_over_dx, double* ke, double* cfl, int* flags, cudaStream_t st);
cudaError_t launch_spectra(const Tables& T, const cplx* qh, int members, double* kespec, double* ensspec, cudaStream_t st);

}  // namespace qgb
