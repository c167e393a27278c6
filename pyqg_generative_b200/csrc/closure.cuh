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

// ---------------------------------------------------------------- fp32 direct convolution -------------
// in  : (batch, Cin, ny, nx) with batch stride in_bs;  out: (batch, Cout, ny, nx) with batch stride out_bs
// wp  : weights repacked to [Cin][KS*KS][CoutPad] (CoutPad multiple of CO_T, zero padded)
// One CTA = 16x16 output pixels x CO_T output channels of one image; input channels streamed 8 at a time.
constexpr int kConvTile = 16;
constexpr int kConvCi = 8;

__device__ __forceinline__ int wrap(int v, int n) {
  v %= n;
  return v < 0 ? v + n : v;
}

template <int KS, int CO_T>
__global__ void __launch_bounds__(256) conv_ffma_kernel(const float* __restrict__ in, long long in_bs,
                                                        float* __restrict__ out, long long out_bs,
                                                        const float* __restrict__ wp, const float* __restrict__ bias,
                                                        const float* __restrict__ bn_s, const float* __restrict__ bn_t,
                                                        int Cin, int Cout, int CoutPad, int ny, int nx, int tiles_x,
                                                        int relu_bn, int softplus, int accumulate) {
  constexpr int PAD = KS / 2;
  constexpr int TW = kConvTile + KS - 1;
  constexpr int TWP = TW + 1;
  __shared__ float s_in[kConvCi][TW][TWP];
  __shared__ __align__(16) float s_w[kConvCi][KS * KS][CO_T];
  const int tid = threadIdx.x;
  const int px = tid % kConvTile, py = tid / kConvTile;
  const int ty0 = (blockIdx.x / tiles_x) * kConvTile, tx0 = (blockIdx.x % tiles_x) * kConvTile;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;
  const float* inb = in + (long long)b * in_bs;
  float acc[CO_T];
#pragma unroll
  for (int j = 0; j < CO_T; ++j) acc[j] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += kConvCi) {
    for (int i = tid; i < kConvCi * TW * TW; i += 256) {
      const int ci = i / (TW * TW), r = (i / TW) % TW, cc = i % TW;
      float v = 0.f;
      if (ci0 + ci < Cin) v = inb[((long long)(ci0 + ci) * ny + wrap(ty0 + r - PAD, ny)) * nx + wrap(tx0 + cc - PAD, nx)];
      s_in[ci][r][cc] = v;
    }
    for (int i = tid; i < kConvCi * KS * KS * CO_T; i += 256) {
      const int ci = i / (KS * KS * CO_T), rem = i % (KS * KS * CO_T);
      float v = 0.f;
      if (ci0 + ci < Cin) v = wp[((long long)(ci0 + ci) * KS * KS) * CoutPad + (long long)(rem / CO_T) * CoutPad + co0 + rem % CO_T];
      (&s_w[ci][0][0])[rem] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < kConvCi; ++ci) {
#pragma unroll
      for (int t = 0; t < KS * KS; ++t) {
        const float v = s_in[ci][py + t / KS][px + t % KS];
        if (CO_T % 4 == 0) {
          const float4* w4 = reinterpret_cast<const float4*>(&s_w[ci][t][0]);
#pragma unroll
          for (int j = 0; j < CO_T / 4; ++j) {
            const float4 w = w4[j];
            acc[4 * j + 0] = fmaf(v, w.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(v, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(v, w.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(v, w.w, acc[4 * j + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CO_T; ++j) acc[j] = fmaf(v, s_w[ci][t][j], acc[j]);
        }
      }
    }
    __syncthreads();
  }
  const int y = ty0 + py, x = tx0 + px;
  if (y < ny && x < nx) {
    float* ob = out + (long long)b * out_bs;
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
      const int co = co0 + j;
      if (co < Cout) {
        float v = acc[j] + bias[co];
        if (relu_bn) v = fmaxf(v, 0.f) * bn_s[co] + bn_t[co];
        if (softplus) v = v > 20.f ? v : log1pf(expf(v));  // torch softplus, beta=1, threshold=20
        float* o = ob + ((long long)co * ny + y) * nx + x;
        *o = accumulate ? *o + v : v;
      }
    }
  }
}

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
