// conv_ffma.cuh -- fp32 direct circular 'same' convolution + bias + ReLU + per-channel affine (+ softplus) in FFMA
// (reference tools/cnn_tools.py:79-98,125-176: Conv2d(padding='same', padding_mode='circular') -> ReLU -> BatchNorm2d).
// Shared by the fp32 inference path (closure.cuh / api.cu) and the training path (train.cuh), where the same kernel also
// computes the data gradient (a circular correlation with the flipped, transposed weights).
#pragma once
#include <cuda_runtime.h>

namespace qgb {

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


// Two output rows per thread: one CTA = 16 (x) x 32 (y) output pixels x CO_T output channels of one image, thread (px, py) owns rows
// py and py + 16, input channels streamed 4 at a time.  Every broadcast weight load (LDS.128) now feeds 8 FFMAs instead of 4: the
// one-row kernel above issues 9 shared-memory wavefronts per 32 FFMAs and is bound by the LSU pipe, this one 10 per 64.  The order in
// which the products of ONE output are accumulated (input channel ascending, taps ascending) is the same, so results are bit-identical.
constexpr int kConvCi2 = 4;
constexpr int kConvTileY2 = 2 * kConvTile;

template <int KS, int CO_T>
__global__ void __launch_bounds__(256, 2) conv_ffma2_kernel(const float* __restrict__ in, long long in_bs,
                                                            float* __restrict__ out, long long out_bs,
                                                            const float* __restrict__ wp, const float* __restrict__ bias,
                                                            const float* __restrict__ bn_s, const float* __restrict__ bn_t,
                                                            int Cin, int Cout, int CoutPad, int ny, int nx, int tiles_x,
                                                            int relu_bn, int softplus, int accumulate) {
  static_assert(CO_T % 4 == 0, "conv_ffma2_kernel: CO_T must be a multiple of 4");
  constexpr int PAD = KS / 2;
  constexpr int TW = kConvTile + KS - 1, TH = kConvTileY2 + KS - 1;
  constexpr int TWP = TW + 1;
  __shared__ float s_in[kConvCi2][TH][TWP];
  __shared__ __align__(16) float s_w[kConvCi2][KS * KS][CO_T];
  const int tid = threadIdx.x;
  const int px = tid % kConvTile, py = tid / kConvTile;
  const int ty0 = (blockIdx.x / tiles_x) * kConvTileY2, tx0 = (blockIdx.x % tiles_x) * kConvTile;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;
  const float* inb = in + (long long)b * in_bs;
  float acc0[CO_T], acc1[CO_T];
#pragma unroll
  for (int j = 0; j < CO_T; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  for (int ci0 = 0; ci0 < Cin; ci0 += kConvCi2) {
    for (int i = tid; i < kConvCi2 * TH * TW; i += 256) {
      const int ci = i / (TH * TW), r = (i / TW) % TH, cc = i % TW;
      float v = 0.f;
      if (ci0 + ci < Cin) v = inb[((long long)(ci0 + ci) * ny + wrap(ty0 + r - PAD, ny)) * nx + wrap(tx0 + cc - PAD, nx)];
      s_in[ci][r][cc] = v;
    }
    for (int i = tid; i < kConvCi2 * KS * KS * CO_T; i += 256) {
      const int ci = i / (KS * KS * CO_T), rem = i % (KS * KS * CO_T);
      float v = 0.f;
      if (ci0 + ci < Cin) v = wp[((long long)(ci0 + ci) * KS * KS) * CoutPad + (long long)(rem / CO_T) * CoutPad + co0 + rem % CO_T];
      (&s_w[ci][0][0])[rem] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < kConvCi2; ++ci) {
#pragma unroll
      for (int t = 0; t < KS * KS; ++t) {
        const float v0 = s_in[ci][py + t / KS][px + t % KS];
        const float v1 = s_in[ci][py + kConvTile + t / KS][px + t % KS];
        const float4* w4 = reinterpret_cast<const float4*>(&s_w[ci][t][0]);
#pragma unroll
        for (int j = 0; j < CO_T / 4; ++j) {
          const float4 w = w4[j];
          acc0[4 * j + 0] = fmaf(v0, w.x, acc0[4 * j + 0]);
          acc0[4 * j + 1] = fmaf(v0, w.y, acc0[4 * j + 1]);
          acc0[4 * j + 2] = fmaf(v0, w.z, acc0[4 * j + 2]);
          acc0[4 * j + 3] = fmaf(v0, w.w, acc0[4 * j + 3]);
          acc1[4 * j + 0] = fmaf(v1, w.x, acc1[4 * j + 0]);
          acc1[4 * j + 1] = fmaf(v1, w.y, acc1[4 * j + 1]);
          acc1[4 * j + 2] = fmaf(v1, w.z, acc1[4 * j + 2]);
          acc1[4 * j + 3] = fmaf(v1, w.w, acc1[4 * j + 3]);
        }
      }
    }
    __syncthreads();
  }
  const int x = tx0 + px;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int y = ty0 + py + h * kConvTile;
    if (y < ny && x < nx) {
      float* ob = out + (long long)b * out_bs;
#pragma unroll
      for (int j = 0; j < CO_T; ++j) {
        const int co = co0 + j;
        if (co < Cout) {
          float v = (h ? acc1[j] : acc0[j]) + bias[co];
          if (relu_bn) v = fmaxf(v, 0.f) * bn_s[co] + bn_t[co];
          if (softplus) v = v > 20.f ? v : log1pf(expf(v));  // torch softplus, beta=1, threshold=20
          float* o = ob + ((long long)co * ny + y) * nx + x;
          *o = accumulate ? *o + v : v;
        }
      }
    }
  }
}

}  // namespace qgb
