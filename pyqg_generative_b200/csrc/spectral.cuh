// spectral.cuh -- sm_100a kernels wrapping the per-member phase programs of qg_core.cuh.
// One CTA owns one ensemble member for the whole program; the packed complex field (N x (N+1) complex128) stays in
// shared memory between phases (N=64: 66.5 KB -> 3 CTAs/SM; N=96: 149 KB -> 1 CTA/SM).
#pragma once
#include <cuda_runtime.h>

#include "qg_core.cuh"
#include "spectral_host.hpp"

namespace qgb {

__global__ void __launch_bounds__(512) qg_program_kernel(const __grid_constant__ Tables T,
                                                         const __grid_constant__ StepIO io, int prog, int members) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + (size_t)T.N * T.P;
  short* pos = reinterpret_cast<short*>(tw + T.N);
  double* red = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(pos) + ((size_t)T.N * sizeof(short) + 15) / 16 * 16);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int m = blockIdx.x; m < members; m += gridDim.x) {
    Ctx c{T, io, buf, tw, pos, red, m};
    const int nph = run_program(c, prog, -1, tid, nt);
    for (int ph = 0; ph < nph; ++ph) {
      run_program(c, prog, ph, tid, nt);
      __syncthreads();
    }
  }
}

// ---- compile-time specialised time step -------------------------------------------------------------------------
// The generic kernel above interprets the phase list with run-time N; ncu showed 89 % of its instructions were integer
// index arithmetic (divisions by N, phase decoding).  For the production sizes the step is unrolled at compile time:
// same phase functions, N / radix plan / stage sequence are constants.
template <int CN> struct FixedPlan;
template <> struct FixedPlan<32> { static constexpr int S = 3; static constexpr int R0 = 4, R1 = 4, R2 = 2, R3 = 1; };
template <> struct FixedPlan<48> { static constexpr int S = 3; static constexpr int R0 = 4, R1 = 4, R2 = 3, R3 = 1; };
template <> struct FixedPlan<64> { static constexpr int S = 3; static constexpr int R0 = 4, R1 = 4, R2 = 4, R3 = 1; };
template <> struct FixedPlan<96> { static constexpr int S = 4; static constexpr int R0 = 4, R1 = 4, R2 = 2, R3 = 3; };

template <int CN, int PASS, int STAGE, bool INV>
__device__ __forceinline__ void fixed_stage(const CtxT<CN>& c, int tid, int nt) {
  using PL = FixedPlan<CN>;
  constexpr int R = STAGE == 0 ? PL::R0 : STAGE == 1 ? PL::R1 : STAGE == 2 ? PL::R2 : PL::R3;
  constexpr int n = STAGE == 0 ? CN : STAGE == 1 ? CN / PL::R0 : STAGE == 2 ? CN / (PL::R0 * PL::R1) : CN / (PL::R0 * PL::R1 * PL::R2);
  constexpr int es = PASS == 0 ? 1 : CN + 1, ls = PASS == 0 ? CN + 1 : 1;
  fft_stage_r<R>(c.buf, c.tw, CN, es, ls, CN, n, INV, tid, nt);
  __syncthreads();
}
template <int CN, int PASS, bool INV>
__device__ __forceinline__ void fixed_pass(const CtxT<CN>& c, int tid, int nt) {
  using PL = FixedPlan<CN>;
  if (!INV) {
    fixed_stage<CN, PASS, 0, INV>(c, tid, nt);
    fixed_stage<CN, PASS, 1, INV>(c, tid, nt);
    fixed_stage<CN, PASS, 2, INV>(c, tid, nt);
    if (PL::S == 4) fixed_stage<CN, PASS, 3, INV>(c, tid, nt);
  } else {
    if (PL::S == 4) fixed_stage<CN, PASS, 3, INV>(c, tid, nt);
    fixed_stage<CN, PASS, 2, INV>(c, tid, nt);
    fixed_stage<CN, PASS, 1, INV>(c, tid, nt);
    fixed_stage<CN, PASS, 0, INV>(c, tid, nt);
  }
}
template <int CN, bool INV>
__device__ __forceinline__ void fixed_fft2d(const CtxT<CN>& c, int tid, int nt) {
  fixed_pass<CN, 0, INV>(c, tid, nt);
  fixed_pass<CN, 1, INV>(c, tid, nt);
}

// prog: PROG_STEP, PROG_STEP_DQ or PROG_STEP_DQ_RAW
template <int CN, int NT>
__global__ void __launch_bounds__(NT) qg_step_fixed_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io,
                                                           int prog, int members) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + (size_t)CN * (CN + 1);
  short* pos = reinterpret_cast<short*>(tw + CN);
  const int tid = threadIdx.x;
  const bool with_dq = prog != PROG_STEP;
  for (int m = blockIdx.x; m < members; m += gridDim.x) {
    CtxT<CN> c{T, io, buf, tw, pos, nullptr, m};
    ph_init(c, tid, NT);
    __syncthreads();
#pragma unroll 1
    for (int z = 0; z < 2; ++z) {
      ph_build_uv(c, z, tid, NT);
      __syncthreads();
      fixed_fft2d<CN, true>(c, tid, NT);
      ph_products(c, z, tid, NT);
      __syncthreads();
      fixed_fft2d<CN, false>(c, tid, NT);
      ph_tendency(c, z, tid, NT);
      __syncthreads();
    }
    if (with_dq) {
      ph_load_pair(c, io.dq, tid, NT);
      __syncthreads();
      fixed_fft2d<CN, false>(c, tid, NT);
    }
    ph_update(c, with_dq, prog == PROG_STEP_DQ, tid, NT);
    __syncthreads();
    ph_build_q(c, tid, NT);
    __syncthreads();
    fixed_fft2d<CN, true>(c, tid, NT);
    ph_emit_q(c, tid, NT);
    __syncthreads();
  }
}

// PROG_BUDGET with the grid size at compile time (the phase list of run_program in qg_core.cuh, same phase functions): the generic
// interpreter spent 0.93 ms per sample of 512 members at 48^2 and 4.3 ms at 96^2 (profiles/r2_diag_overhead.md)
template <int CN, int NT>
__global__ void __launch_bounds__(NT) qg_budget_fixed_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io, int members) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + (size_t)CN * (CN + 1);
  short* pos = reinterpret_cast<short*>(tw + CN);
  const int tid = threadIdx.x;
#define QGB_BP(call) do { call; __syncthreads(); } while (0)
  for (int m = blockIdx.x; m < members; m += gridDim.x) {
    CtxT<CN> c{T, io, buf, tw, pos, nullptr, m};
    QGB_BP(ph_init(c, tid, NT));
    QGB_BP(ph_build_xi(c, tid, NT));
    fixed_fft2d<CN, true>(c, tid, NT);
    QGB_BP(ph_store_scr(c, 0, 2, tid, NT));
#pragma unroll 1
    for (int z = 0; z < 2; ++z) {
      QGB_BP(ph_build_uv(c, z, tid, NT));
      fixed_fft2d<CN, true>(c, tid, NT);
      QGB_BP(ph_products_scr(c, z, tid, NT));
      fixed_fft2d<CN, false>(c, tid, NT);
      QGB_BP(ph_bud_keflux(c, z, tid, NT));
    }
    QGB_BP(ph_build_tau(c, tid, NT));
    fixed_fft2d<CN, true>(c, tid, NT);
    QGB_BP(ph_store_scr(c, 2, 1, tid, NT));
    QGB_BP(ph_build_uvbt(c, tid, NT));
    fixed_fft2d<CN, true>(c, tid, NT);
    QGB_BP(ph_products_scr(c, 2, tid, NT));
    fixed_fft2d<CN, false>(c, tid, NT);
    QGB_BP(ph_bud_apeflux(c, tid, NT));
#pragma unroll 1
    for (int z = 0; z < 2; ++z) {
      QGB_BP(ph_build_uv(c, z, tid, NT));
      fixed_fft2d<CN, true>(c, tid, NT);
      QGB_BP(ph_products_anom(c, z, tid, NT));
      fixed_fft2d<CN, false>(c, tid, NT);
      QGB_BP(ph_bud_ens(c, z, tid, NT));
    }
    if (io.dq) {
      QGB_BP(ph_load_pair(c, io.dq, tid, NT));
      fixed_fft2d<CN, false>(c, tid, NT);
      QGB_BP(ph_bud_param(c, tid, NT));
    }
    QGB_BP(ph_bud_diss(c, io.dq != nullptr, tid, NT));
  }
#undef QGB_BP
}

// The q setter, irfft2, _invert and the advection operator (PROG_SET_Q, PROG_C2R, PROG_INVERT, PROG_ADVECT: snapshots and the
// coarse-graining operators) with the grid size at compile time -- the phase lists of run_program, same phase functions
template <int CN, int NT>
__global__ void __launch_bounds__(NT) qg_program_fixed_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io, int prog,
                                                              int members) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + (size_t)CN * (CN + 1);
  short* pos = reinterpret_cast<short*>(tw + CN);
  const int tid = threadIdx.x;
#define QGB_BP(call) do { call; __syncthreads(); } while (0)
  for (int m = blockIdx.x; m < members; m += gridDim.x) {
    CtxT<CN> c{T, io, buf, tw, pos, nullptr, m};
    QGB_BP(ph_init(c, tid, NT));
    if (prog == PROG_SET_Q) {
      QGB_BP(ph_load_pair(c, io.q, tid, NT));
      fixed_fft2d<CN, false>(c, tid, NT);
      QGB_BP(ph_store_qh(c, tid, NT));
      if (io.cnn_x) QGB_BP(ph_emit_x_only(c, tid, NT));
    } else if (prog == PROG_C2R) {
      QGB_BP(ph_build_q(c, tid, NT));
      fixed_fft2d<CN, true>(c, tid, NT);
      QGB_BP(ph_emit_q(c, tid, NT));
    } else if (prog == PROG_INVERT) {
      if (io.ph_out) QGB_BP(ph_store_ph(c, tid, NT));
#pragma unroll 1
      for (int z = 0; z < 2; ++z) {
        QGB_BP(ph_build_uv(c, z, tid, NT));
        fixed_fft2d<CN, true>(c, tid, NT);
        QGB_BP(ph_store_uv(c, z, tid, NT));
      }
      if (io.p_out) {
        QGB_BP(ph_build_p(c, tid, NT));
        fixed_fft2d<CN, true>(c, tid, NT);
        QGB_BP(ph_store_p(c, tid, NT));
      }
    } else {                                           // PROG_ADVECT
#pragma unroll 1
      for (int z = 0; z < 2; ++z) {
        QGB_BP(ph_build_uv(c, z, tid, NT));
        fixed_fft2d<CN, true>(c, tid, NT);
        QGB_BP(ph_products(c, z, tid, NT));
        fixed_fft2d<CN, false>(c, tid, NT);
        QGB_BP(ph_tendency(c, z, tid, NT));
      }
    }
  }
#undef QGB_BP
}

// Large grids (N = 128, 256: the packed field is 0.26 / 1.0 MB and no longer fits one CTA's shared memory): the SAME phase
// programs run with the working field in a per-member global-memory scratch (L2 resident: 64 members x 1 MB at 256^2)
// and a thread-block CLUSTER of 4 or 8 CTAs per member.  Threads are numbered across the cluster, phases are separated
// by the hardware cluster barrier (release/acquire at cluster scope, which also invalidates L1), tables are read in place.
// The 1-D passes of the transforms run per CTA in shared memory (fft2d_pass_tiled): 2 cluster phases per 2-D transform.

__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CN>
__global__ void __launch_bounds__(512, 2) qg_program_cluster_kernel(const __grid_constant__ Tables T,
                                                                 const __grid_constant__ StepIO io, int prog, int members,
                                                                 cplx* scratch, double* red_scratch, int tile_lines, const short* true_pos) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int rank = (int)cluster_ctarank();
  uint32_t csz;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csz));
  const int ncl = gridDim.x / (int)csz, cl = blockIdx.x / (int)csz;
  const int tid = rank * blockDim.x + threadIdx.x, nt = (int)csz * blockDim.x;
  // digit-reversal map of the stages, behind the tile in shared memory (T.pos is the identity on this path)
  short* tpos = reinterpret_cast<short*>(smem_raw + (size_t)tile_lines * (T.N + 1) * sizeof(cplx));
  for (int i = threadIdx.x; i < T.N; i += blockDim.x) tpos[i] = true_pos[i];
  __syncthreads();
  for (int m = cl; m < members; m += ncl) {
    CtxT<CN> c{T, io, scratch + (size_t)m * T.N * T.P, const_cast<cplx*>(T.tw), const_cast<short*>(T.pos),
          red_scratch + (size_t)m * 4 * nt, m, reinterpret_cast<cplx*>(smem_raw), tile_lines, (int)csz, tpos};
    const int nph = run_program(c, prog, -1, tid, nt);
    for (int ph = 0; ph < nph; ++ph) {
      run_program(c, prog, ph, tid, nt);
      cluster_barrier();
    }
  }
}

// ke / cfl / flags from the per-member reduction record written by PROG_DIAG
__global__ void diag_finish_kernel(const double* __restrict__ red, int members, double dt_over_dx, double* ke,
                                   double* cfl, int* flags) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= members) return;
  const double k = red[4 * m], c = fmax(red[4 * m + 1], red[4 * m + 2]) * dt_over_dx;
  if (ke) ke[m] = k;
  if (cfl) cfl[m] = c;
  if (flags) {
    int f = 0;
    if (red[4 * m + 3] != 0.0 || !(k == k) || !(c == c)) f |= 1;
    if (!(c < 1.0)) f |= 2;
    flags[m] = f;
  }
}

// pyqg diagnostics KEspec = wv2 |ph|^2 / M^2, Ensspec = |qh|^2 / M^2, summed over the local members.  Block = 32 spectral points x
// kSpectraParts member slices (a warp reads 32 consecutive points of one member: 512 contiguous bytes); the slices are added in a fixed
// order, so the sums are deterministic.  (One thread per point looping over all members ran 33 blocks for 0.31 ms at 64^2 x 1024.)
constexpr int kSpectraParts = 8;
__global__ void __launch_bounds__(32 * kSpectraParts) spectra_kernel(const __grid_constant__ Tables T, const cplx* __restrict__ qh, int members,
                                                                 double* __restrict__ kespec, double* __restrict__ ensspec) {
  __shared__ double sk[kSpectraParts][32], se[kSpectraParts][32];
  const int NN = T.N * T.NK;
  const int i = blockIdx.x * 32 + threadIdx.x, part = threadIdx.y;
  double ke = 0.0, en = 0.0;
  if (i < 2 * NN) {
    const int z = i / NN, idx = i - z * NN;
    const int l = idx / T.NK, k = idx - l * T.NK;
    const double wv2 = T.kv[k] * T.kv[k] + T.lv[l] * T.lv[l];
    const double a0 = T.a[(2 * z) * NN + idx], a1 = T.a[(2 * z + 1) * NN + idx];
    for (int m = part; m < members; m += kSpectraParts) {
      const cplx q0 = qh[(long long)m * 2 * NN + idx], q1 = qh[(long long)m * 2 * NN + NN + idx];
      const double px = a0 * q0.x + a1 * q1.x, py = a0 * q0.y + a1 * q1.y;
      ke += wv2 * (px * px + py * py);
      const cplx qq = z == 0 ? q0 : q1;
      en += qq.x * qq.x + qq.y * qq.y;
    }
  }
  sk[part][threadIdx.x] = ke;
  se[part][threadIdx.x] = en;
  __syncthreads();
  if (part == 0 && i < 2 * NN) {
    for (int p = 1; p < kSpectraParts; ++p) { ke += sk[p][threadIdx.x]; en += se[p][threadIdx.x]; }
    const double s = T.inv_M * T.inv_M;
    if (kespec) kespec[i] = ke * s;
    if (ensspec) ensspec[i] = en * s;
  }
}

}  // namespace qgb
