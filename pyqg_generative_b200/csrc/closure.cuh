// closure.cuh -- device code of the CNN subgrid closure, fp32 path:
//   * Philox4x32-10 + Box-Muller latent noise with the AR1 / constant update fused
//     (pyqg_generative/tools/stochastic_pyqg.py:30-72, models/cgan_regression.py:154-155,
//      models/mean_var_model.py:102-103)
//   * direct circular 'same' convolution + bias + ReLU + BatchNorm affine (+softplus) in fp32 FFMA
//     (tools/cnn_tools.py:79-98,125-176; models/mean_var_model.py:14-17)
//   * denormalisation / GZ sampling epilogue (models/cgan_regression.py:157-162, mean_var_model.py:105-109)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_ffma.cuh"
#include "philox.cuh"

namespace qgb {

// latent update  z <- a z + b xi  (first call / constant sampler: a=0,b=1 -> z = xi).
// T = float : z lives in channels 2,3 of the closure input (B,4,N,N);  T = double : separate (B,2,N,N) buffer.
// xi_inj (optional) replaces Philox (parity injection).  One thread per 4 consecutive pixels.
template <typename T>
__global__ void latent_update_kernel(T* z, long long mstride, int npix, int members, int member_offset,
                                     uint64_t seed, const uint32_t* __restrict__ draw_counter, int draw_bias, T a, T b, int replace, const T* xi_inj) {
  const uint32_t draw = *draw_counter + (uint32_t)draw_bias;      // device-resident draw counter: the launch is identical every step (CUDA-graph replay)
  const int quads = (npix + 3) / 4;
  const long long total = (long long)members * 2 * quads;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int quad = (int)(i % quads);
    const int ch = (int)((i / quads) % 2);
    const int m = (int)(i / (2LL * quads));
    float xi[4];
    if (xi_inj == nullptr) philox_normal4(seed, (uint32_t)(m + member_offset), draw, (uint32_t)ch, (uint32_t)quad, xi);
    T* zp = z + (long long)m * mstride + (long long)ch * npix;
    for (int j = 0; j < 4; ++j) {
      const int p = quad * 4 + j;
      if (p >= npix) break;
      const T x = xi_inj ? xi_inj[((long long)m * 2 + ch) * npix + p] : (T)xi[j];
      if (replace) {
        zp[p] = x;
      } else {  // numpy evaluates a*noise and b*xi separately, then adds: no FMA contraction
        if (sizeof(T) == 4) zp[p] = (T)__fadd_rn(__fmul_rn((float)a, (float)zp[p]), __fmul_rn((float)b, (float)x));
        else zp[p] = (T)__dadd_rn(__dmul_rn((double)a, (double)zp[p]), __dmul_rn((double)b, (double)x));
      }
    }
  }
}

// one thread: advances the draw counter after the latent kernel of a step has read it
__global__ void bump_counter_kernel(uint32_t* c) { *c += 1u; }

// ---------------------------------------------------------------- closure epilogues -------------------
// gan / vae / ols:  dq = float64( y * y_std ) * weight * scale      (scale = 1/M for the 'deterministic' mean)
__global__ void finish_plain_kernel(const float* __restrict__ y, double* __restrict__ dq, int npix, long long total,
                                    float ys0, float ys1, double weight, float scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)((i / npix) % 2);
    const float v = __fmul_rn(__fmul_rn(y[i], scale), ch ? ys1 : ys0);
    dq[i] = (double)v * weight;
  }
}

// gz:  dq = ( mean + z * sqrt(var) ) * y_std * weight     (models/mean_var_model.py:105-109, evaluated in float64
// exactly as numpy promotes: sqrt in fp32, product and sum in fp64)
__global__ void finish_gz_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                 const double* __restrict__ z, double* __restrict__ dq, int npix, long long total,
                                 float ys0, float ys1, double weight, int use_noise) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)((i / npix) % 2);
    double v = (double)mean[i];
    if (use_noise) v = __dadd_rn(v, __dmul_rn(z[i], (double)sqrtf(var[i])));
    dq[i] = v * (double)(ch ? ys1 : ys0) * weight;
  }
}

// m.PV_forcing = dq - mean(dq) per member and layer (models/parameterization.py:25); one CTA per (member, layer)
__global__ void demean_kernel(const double* __restrict__ dq, double* __restrict__ out, int npix) {
  __shared__ double s[256];
  const double* p = dq + (long long)blockIdx.x * npix;
  double acc = 0.0;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) acc += p[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  const double mean = s[0] / npix;
  double* q = out + (long long)blockIdx.x * npix;
  for (int i = threadIdx.x; i < npix; i += blockDim.x) q[i] = p[i] - mean;
}

// float accumulate helper for the deterministic (mean over M samples) mode
__global__ void axpy_f32_kernel(float* __restrict__ acc, const float* __restrict__ x, long long n, int first) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc[i] = first ? x[i] : acc[i] + x[i];
}

}  // namespace qgb
