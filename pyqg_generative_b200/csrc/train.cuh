// train.cuh -- device code of the AndrewCNN training step, fp32, sm_100a.
//
// Reference: tools/cnn_tools.py:645-700 ``train`` (Adam, MSELoss through ``compute_loss`` :177-182, BatchNorm2d in training
// mode), used by models/mean_var_model.py:41-66 (GZ two-stage fit: mean network, then VarCNN = softplus(AndrewCNN) on the squared
// residuals) and models/ols_model.py ``fit``.  One step = forward with batch statistics, MSE loss, backward, Adam update:
//
//   forward   z_l = conv(a_{l-1}) + b_l ; r_l = relu(z_l) ; a_l = gamma_l (r_l - mean_l) invstd_l + beta_l   (last layer: y = z_L)
//   backward  dgrad = the forward convolution kernel on the flipped, transposed weights (circular 'same' padding is symmetric);
//             wgrad = wgrad_ffma_kernel below; BatchNorm / ReLU backward fused in one element-wise pass between them
//
// Everything stays on the device; the only host traffic of a step is the minibatch in and one double (the loss) out.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_ffma.cuh"

namespace qgb {
namespace train {

constexpr int kRedSplit = 64;        // partial sums per channel in the two-stage per-channel reductions (deterministic order)
constexpr int kWgTile = 16;          // spatial tile of the weight-gradient kernel

// ---- weight layouts -------------------------------------------------------------------------------------------------------
// torch layout W[co][ci][ky][kx] -> conv_ffma layout wp[ci'][tap][co' (padded)].  flip = 0: the forward convolution
// (ci' = ci, co' = co, tap = ky KS + kx).  flip = 1: the data gradient  da[ci] = sum_co sum_tap W[co][ci][KS-1-ky][KS-1-kx] dz[co]
// (ci' = co, co' = ci).
__global__ void pack_weights_kernel(const float* __restrict__ W, float* __restrict__ wp, int cin, int cout, int ks, int co_pad,
                                    int flip) {
  const int kk = ks * ks;
  const int in_c = flip ? cout : cin, out_c = flip ? cin : cout;
  const long long n = (long long)in_c * kk * co_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(i % co_pad), t = (int)((i / co_pad) % kk), c = (int)(i / ((long long)co_pad * kk));
    float v = 0.f;
    if (o < out_c) v = flip ? W[((long long)c * cin + o) * kk + (kk - 1 - t)] : W[((long long)o * cin + c) * kk + t];
    wp[i] = v;
  }
}

// ---- per-channel reductions over (batch, y, x) ----------------------------------------------------------------------------
// MODE 0: (sum r, sum r^2)                     BatchNorm batch statistics
// MODE 1: (sum da, sum da * xhat)              BatchNorm backward, xhat = (r - mean) invstd
// MODE 2: (sum dz, 0)                          bias gradient
// part[(c * kRedSplit + s) * 2 + {0, 1}] in double; block s of channel c strides over the images.
template <int MODE>
__global__ void __launch_bounds__(256) chan_partial_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           int batch, int C, int hw, double* __restrict__ part) {
  const int c = blockIdx.x, s = blockIdx.y;
  double s0 = 0.0, s1 = 0.0;
  float mu = 0.f, is = 0.f;
  if (MODE == 1) { mu = mean[c]; is = invstd[c]; }
  if (hw % 4 == 0) {                      // 16-byte accesses (every grid of the path has an even side)
    const int hw4 = hw / 4;
    const long long total = (long long)batch * hw4;
    for (long long i = (long long)s * 256 + threadIdx.x; i < total; i += (long long)kRedSplit * 256) {
      const int b = (int)(i / hw4), p = (int)(i - (long long)b * hw4);
      const long long off = ((long long)b * C + c) * hw + 4 * p;
      const float4 v4 = *reinterpret_cast<const float4*>(x + off);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      float rr[4] = {0.f, 0.f, 0.f, 0.f};
      if (MODE == 1) { const float4 r4 = *reinterpret_cast<const float4*>(r + off); rr[0] = r4.x; rr[1] = r4.y; rr[2] = r4.z; rr[3] = r4.w; }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s0 += v[k];
        if (MODE == 0) s1 += (double)v[k] * v[k];
        if (MODE == 1) s1 += (double)v[k] * ((rr[k] - mu) * is);
      }
    }
  } else {
    const long long total = (long long)batch * hw;
    for (long long i = (long long)s * 256 + threadIdx.x; i < total; i += (long long)kRedSplit * 256) {
      const int b = (int)(i / hw), p = (int)(i - (long long)b * hw);
      const long long off = ((long long)b * C + c) * hw + p;
      const float v = x[off];
      s0 += v;
      if (MODE == 0) s1 += (double)v * v;
      if (MODE == 1) s1 += (double)v * ((r[off] - mu) * is);
    }
  }
  __shared__ double sh0[256], sh1[256];
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) { sh0[threadIdx.x] += sh0[threadIdx.x + w]; sh1[threadIdx.x] += sh1[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[((long long)c * kRedSplit + s) * 2] = sh0[0]; part[((long long)c * kRedSplit + s) * 2 + 1] = sh1[0]; }
}

// BatchNorm2d training statistics (torch.nn.BatchNorm2d, momentum 0.1, eps 1e-5): batch mean, biased variance for the
// normalisation, unbiased variance into running_var
__global__ void bn_stats_final_kernel(const double* __restrict__ part, int C, double n, float eps, float momentum,
                                      float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean,
                                      float* __restrict__ run_var, int update_running) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < kRedSplit; ++s) { s0 += part[((long long)c * kRedSplit + s) * 2]; s1 += part[((long long)c * kRedSplit + s) * 2 + 1]; }
  const double mu = s0 / n;
  double var = s1 / n - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (update_running) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mu;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unbiased;
  }
}

// sums of a two-stage reduction -> out0[c] (and out1[c])
__global__ void chan_final_kernel(const double* __restrict__ part, int C, float* __restrict__ out0, float* __restrict__ out1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int s = 0; s < kRedSplit; ++s) { s0 += part[((long long)c * kRedSplit + s) * 2]; s1 += part[((long long)c * kRedSplit + s) * 2 + 1]; }
  if (out0) out0[c] = (float)s0;
  if (out1) out1[c] = (float)s1;
}

// a = gamma (r - mean) invstd + beta
__global__ void bn_apply_kernel(const float* __restrict__ r, float* __restrict__ a, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ invstd,
                                int C, int hw, long long total) {
  if (hw % 4 == 0) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)((4 * i / hw) % C);
      const float g = gamma[c], m = mean[c], is = invstd[c], be = beta[c];
      const float4 v = reinterpret_cast<const float4*>(r)[i];
      reinterpret_cast<float4*>(a)[i] = make_float4(g * ((v.x - m) * is) + be, g * ((v.y - m) * is) + be, g * ((v.z - m) * is) + be,
                                                    g * ((v.w - m) * is) + be);
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / hw) % C);
    a[i] = gamma[c] * ((r[i] - mean[c]) * invstd[c]) + beta[c];
  }
}

// BatchNorm + ReLU backward in place:  dz = [r > 0] gamma invstd (da - dbeta / n - xhat dgamma / n)
__global__ void bn_relu_bwd_kernel(float* __restrict__ d, const float* __restrict__ r, const float* __restrict__ gamma,
                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                   const float* __restrict__ dgamma, const float* __restrict__ dbeta, int C, int hw, float inv_n,
                                   long long total) {
  if (hw % 4 == 0) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total / 4; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)((4 * i / hw) % C);
      const float m = mean[c], is = invstd[c], ga = gamma[c], db = dbeta[c], dg = dgamma[c];
      const float4 r4 = reinterpret_cast<const float4*>(r)[i];
      float4 d4 = reinterpret_cast<float4*>(d)[i];
      const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
      float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xhat = (rv[k] - m) * is;
        const float g = ga * is * (dv[k] - db * inv_n - xhat * dg * inv_n);      // (same expression order as the scalar path)
        dv[k] = rv[k] > 0.f ? g : 0.f;
      }
      reinterpret_cast<float4*>(d)[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / hw) % C);
    const float rv = r[i], xhat = (rv - mean[c]) * invstd[c];
    const float g = gamma[c] * invstd[c] * (d[i] - dbeta[c] * inv_n - xhat * dgamma[c] * inv_n);
    d[i] = rv > 0.f ? g : 0.f;
  }
}

// The same backward pass organised per channel (grid (C, kRedSplit) like chan_partial_kernel) so that it also leaves the per-channel
// sums of dz -- the bias gradient of the convolution below -- in ``part``: one pass over the tensor less than a separate reduction.
__global__ void __launch_bounds__(256) bn_relu_bwd_chan_kernel(float* __restrict__ d, const float* __restrict__ r,
                                                               const float* __restrict__ gamma, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ dgamma,
                                                               const float* __restrict__ dbeta, int batch, int C, int hw, float inv_n,
                                                               double* __restrict__ part) {
  const int c = blockIdx.x, s = blockIdx.y;
  const float m = mean[c], is = invstd[c], ga = gamma[c], db = dbeta[c], dg = dgamma[c];
  double s0 = 0.0;
  if (hw % 4 == 0) {
    const int hw4 = hw / 4;
    const long long total = (long long)batch * hw4;
    for (long long i = (long long)s * 256 + threadIdx.x; i < total; i += (long long)kRedSplit * 256) {
      const int b = (int)(i / hw4), p = (int)(i - (long long)b * hw4);
      const long long off = ((long long)b * C + c) * hw + 4 * p;
      const float4 r4 = *reinterpret_cast<const float4*>(r + off);
      const float4 d4 = *reinterpret_cast<const float4*>(d + off);
      const float rv[4] = {r4.x, r4.y, r4.z, r4.w};
      float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xhat = (rv[k] - m) * is;
        const float g = ga * is * (dv[k] - db * inv_n - xhat * dg * inv_n);
        dv[k] = rv[k] > 0.f ? g : 0.f;
        s0 += dv[k];
      }
      *reinterpret_cast<float4*>(d + off) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
  } else {
    const long long total = (long long)batch * hw;
    for (long long i = (long long)s * 256 + threadIdx.x; i < total; i += (long long)kRedSplit * 256) {
      const int b = (int)(i / hw), p = (int)(i - (long long)b * hw);
      const long long off = ((long long)b * C + c) * hw + p;
      const float rv = r[off], xhat = (rv - m) * is;
      const float g = ga * is * (d[off] - db * inv_n - xhat * dg * inv_n);
      const float dz = rv > 0.f ? g : 0.f;
      d[off] = dz;
      s0 += dz;
    }
  }
  __shared__ double sh0[256];
  sh0[threadIdx.x] = s0;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh0[threadIdx.x] += sh0[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[((long long)c * kRedSplit + s) * 2] = sh0[0]; part[((long long)c * kRedSplit + s) * 2 + 1] = 0.0; }
}

// eval-mode BatchNorm folded into the convolution epilogue: s = gamma / sqrt(running_var + eps), t = beta - running_mean s
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rm,
                               const float* __restrict__ rv, float eps, int C, float* __restrict__ s, float* __restrict__ t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  s[c] = sc;
  t[c] = beta[c] - rm[c] * sc;
}

// ---- loss ----------------------------------------------------------------------------------------------------------------
// MSELoss(mean) of y = z (or softplus(z), VarCNN mean_var_model.py:14-17) against the target, and its gradient with respect to z:
// dz = 2 (y - t) / n [* sigmoid(z)].  Block partial sums in double -> loss_part[blockIdx.x].
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ z, const float* __restrict__ t, float* __restrict__ dz,
                                                       long long n, int softplus, double* __restrict__ loss_part) {
  double acc = 0.0;
  const float scale = (float)(2.0 / (double)n);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float zz = z[i];
    float y = zz, dy = 1.f;
    if (softplus) {                    // torch softplus (beta 1, threshold 20) and its derivative
      if (zz <= 20.f) { y = log1pf(expf(zz)); dy = 1.f / (1.f + expf(-zz)); }
    }
    const float e = y - t[i];
    acc += (double)e * e;
    if (dz) dz[i] = scale * e * dy;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_part[blockIdx.x] = sh[0];
}
__global__ void loss_final_kernel(const double* __restrict__ part, int nparts, double inv_n, double* __restrict__ loss) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[i];
    *loss = s * inv_n;
  }
}

// ---- weight gradient ------------------------------------------------------------------------------------------------------
// dW[co][ci][ky][kx] = sum_{b, y, x} dz[b][co][y][x] a[b][ci][(y + ky - p) mod ny][(x + kx - p) mod nx]
// Block = (tile of CO_B = (256 / CI_T) CO_PER output channels) x (CI_T input channels) x all KS^2 taps, looping over its share
// (blockIdx.y of gridDim.y) of the (image, 16 x 16 spatial tile) work items; thread (ci, co group) keeps CO_PER x KS^2
// accumulators in registers and slides a KS x KS window of the input tile along x.  Partial sums go to
// part[split][co][ci][tap]; wgrad_reduce_kernel adds the splits in order (deterministic).
template <int KS>
struct WgGeom {
  static constexpr int TW = kWgTile + KS - 1;
  static constexpr int CI_STRIDE = ((TW * TW + 31) / 32) * 32 + 1;   // == 1 (mod 32): the 32 ci lanes hit 32 banks
  static constexpr int D_PITCH = kWgTile * kWgTile + 1;
};

// DB: the input tile and the output-gradient tile of the NEXT work item are copied global -> shared with cp.async into the second
// half of a double buffer while the FFMAs of the current item run (the synchronous fill left 42 % of the stall samples waiting on
// those loads, profiles/r2_training.md); needs 2 x the shared memory, so one block per SM.
__device__ __forceinline__ int wrap_near(int v, int n) {      // v within one period of [0, n) in the common case
  if (v < 0) v += n; else if (v >= n) v -= n;
  return (unsigned)v < (unsigned)n ? v : wrap(v, n);
}

template <int KS, int CI_T, int CO_PER, int DB>
__global__ void __launch_bounds__(256, ((CO_PER * KS * KS <= 60 && !DB) ? 2 : 1)) wgrad_ffma_kernel(const float* __restrict__ a, const float* __restrict__ dz,
                                                            float* __restrict__ part, int batch, int Cin, int Cout, int ny, int nx,
                                                            int ci_blocks) {
  using G = WgGeom<KS>;
  constexpr int PAD = KS / 2, TW = G::TW, KK = KS * KS;
  constexpr int CG = 256 / CI_T, CO_B = CG * CO_PER;
  constexpr int A_FLOATS = CI_T * G::CI_STRIDE, BUF = A_FLOATS + CO_B * G::D_PITCH;
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int ci_l = tid % CI_T, cg = tid / CI_T;
  const int ci0 = (blockIdx.x % ci_blocks) * CI_T, co0 = (blockIdx.x / ci_blocks) * CO_B;
  const int tiles_x = (nx + kWgTile - 1) / kWgTile, tiles_y = (ny + kWgTile - 1) / kWgTile;
  const int items = batch * tiles_x * tiles_y;
  float acc[CO_PER][KK];
#pragma unroll
  for (int j = 0; j < CO_PER; ++j)
#pragma unroll
    for (int t = 0; t < KK; ++t) acc[j][t] = 0.f;

  auto fill = [&](int buf, int it) {
    float* s_a = smem + buf * BUF;                  // [CI_T][CI_STRIDE]
    float* s_d = s_a + A_FLOATS;                    // [CO_B][D_PITCH]
    const int b = it / (tiles_x * tiles_y), tt = it % (tiles_x * tiles_y);
    const int ty0 = (tt / tiles_x) * kWgTile, tx0 = (tt % tiles_x) * kWgTile;
    const float* ab = a + (long long)b * Cin * ny * nx;
    const float* db = dz + (long long)b * Cout * ny * nx;
    for (int i = tid; i < CI_T * TW * TW; i += 256) {
      const int ci = i / (TW * TW), rr = (i / TW) % TW, cc = i % TW;
      float* d = s_a + ci * G::CI_STRIDE + rr * TW + cc;
      const bool live = ci0 + ci < Cin;
      const float* src = ab + ((long long)(ci0 + ci) * ny + wrap_near(ty0 + rr - PAD, ny)) * nx + wrap_near(tx0 + cc - PAD, nx);
      if (DB) { if (live) cp_async4(d, src); else *d = 0.f; }
      else *d = live ? *src : 0.f;
    }
    for (int i = tid; i < CO_B * kWgTile * kWgTile; i += 256) {
      const int co = i / (kWgTile * kWgTile), p = i % (kWgTile * kWgTile);
      const int y = ty0 + p / kWgTile, x = tx0 + p % kWgTile;
      float* d = s_d + co * G::D_PITCH + p;
      const bool live = co0 + co < Cout && y < ny && x < nx;
      const float* src = db + ((long long)(co0 + co) * ny + y) * nx + x;
      if (DB) { if (live) cp_async4(d, src); else *d = 0.f; }
      else *d = live ? *src : 0.f;
    }
    if (DB) cp_async_commit();
  };

  int it = blockIdx.y, buf = 0;
  if (DB && it < items) fill(0, it);
  for (; it < items; it += gridDim.y) {
    if (DB) {
      const int next = it + gridDim.y;
      if (next < items) { fill(buf ^ 1, next); cp_async_wait<1>(); } else cp_async_wait<0>();
    } else fill(0, it);
    __syncthreads();
    const float* sa = smem + buf * BUF + ci_l * G::CI_STRIDE;
    const float* sd = smem + buf * BUF + A_FLOATS + (cg * CO_PER) * G::D_PITCH;
#pragma unroll 1
    for (int py = 0; py < kWgTile; ++py) {
      float win[KS][KS];                             // win[ky][kx] = input at (py + ky, px + kx)
#pragma unroll
      for (int ky = 0; ky < KS; ++ky)
#pragma unroll
        for (int kx = 1; kx < KS; ++kx) win[ky][kx] = sa[(py + ky) * TW + kx - 1];
#pragma unroll
      for (int px = 0; px < kWgTile; ++px) {
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
          for (int kx = 0; kx < KS - 1; ++kx) win[ky][kx] = win[ky][kx + 1];
          win[ky][KS - 1] = sa[(py + ky) * TW + px + KS - 1];
        }
#pragma unroll
        for (int j = 0; j < CO_PER; ++j) {
          const float g = sd[j * G::D_PITCH + py * kWgTile + px];
#pragma unroll
          for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) acc[j][ky * KS + kx] = fmaf(g, win[ky][kx], acc[j][ky * KS + kx]);
        }
      }
    }
    __syncthreads();
    if (DB) buf ^= 1;
  }
  const int ci = ci0 + ci_l;
  if (ci < Cin) {
#pragma unroll
    for (int j = 0; j < CO_PER; ++j) {
      const int co = co0 + cg * CO_PER + j;
      if (co < Cout) {
        float* o = part + (((long long)blockIdx.y * Cout + co) * Cin + ci) * KK;
#pragma unroll
        for (int t = 0; t < KK; ++t) o[t] = acc[j][t];
      }
    }
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, long long n, float* __restrict__ dW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[(long long)k * n + i];
  dW[i] = s;
}

// ---- Adam (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) ----------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float beta1, float beta2, float eps, float bc1, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace train
}  // namespace qgb
