// conv_ffma.cuh -- fp32 direct circular 'same' convolution + bias + ReLU + per-channel affine (+ softplus) in FFMA
// (reference tools/cnn_tools.py:79-98,125-176: Conv2d(padding='same', padding_mode='circular') -> ReLU -> BatchNorm2d).
// Shared by the fp32 inference path (closure.cuh / api.cu) and the training path (train.cuh), where the same kernel also
// computes the data gradient (a circular correlation with the flipped, transposed weights).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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


// Register-tiled variant for the wide layers: one CTA = 16 (x) x 32 (y) output pixels x CO_T output channels of one image, a thread owns
// 4 adjacent pixels of one row x CO_T / 2 channels, input channels streamed 4 at a time.  Per input channel and tap row the thread reads
// its 8 input values with two aligned LDS.128 and every broadcast weight load (LDS.128) feeds 16 FFMAs: 28 shared-memory wavefronts per
// 320 FFMAs (the one-pixel kernel above: 9 per 32 and bound by the LSU pipe).  The order in which the products of ONE output are
// accumulated (input channel ascending, taps ascending) is the same as in conv_ffma_kernel, so results are bit-identical.
// Staging (round 2, profiles/r2_training.md): the input tile + halo and the weight slab of the NEXT channel chunk are copied
// global -> shared with cp.async into the second half of a double buffer while the FFMAs of the current chunk run, from source
// offsets computed once per thread -- the synchronous fill with its div / mod / wrap index arithmetic and exposed load latency was
// 45 % of the stall samples.  Row pitch 48 (= 16 mod 32): the two 16-pixel rows of a warp hit disjoint banks.
constexpr int kConvCi2 = 4;
constexpr int kConvPitch2 = 48;

// TY = rows of the CTA tile: 32 (256 threads, two CTAs per SM) or 16 (128 threads, four per SM) -- 16 where 32-row tiles would hang over
// the image (ny = 48, the grid of the shipped models: 3 tiles of 16 instead of 2 of 32 with a quarter of the work wasted)
template <int KS, int TY>
struct Conv2Geom {
  static constexpr int NT = 8 * TY;                                   // threads: 4 pixel groups per row x TY rows x 2 channel halves
  static constexpr int TW = kConvTile + KS - 1, TH = TY + KS - 1;
  static constexpr int NPOS = TH * TW, PPT = (NPOS + NT - 1) / NT;
  static constexpr int IN_FLOATS = kConvCi2 * TH * kConvPitch2;
};
template <int KS, int CO_T, int TY>
struct Conv2Smem {
  static constexpr int BUF = Conv2Geom<KS, TY>::IN_FLOATS + kConvCi2 * KS * KS * CO_T;      // floats per half of the double buffer
  static constexpr size_t BYTES = 2 * (size_t)BUF * sizeof(float);
};

__device__ __forceinline__ void cp_async4(float* dst_shared, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst_shared, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_shared)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KS, int CO_T, int TY>
__global__ void __launch_bounds__(8 * TY, 64 / TY) conv_ffma2_kernel(const float* __restrict__ in, long long in_bs,
                                                            float* __restrict__ out, long long out_bs,
                                                            const float* __restrict__ wp, const float* __restrict__ bias,
                                                            const float* __restrict__ bn_s, const float* __restrict__ bn_t,
                                                            int Cin, int Cout, int CoutPad, int ny, int nx, int tiles_x,
                                                            int relu_bn, int softplus, int accumulate) {
  static_assert(CO_T % 4 == 0, "conv_ffma2_kernel: CO_T must be a multiple of 4");
  using G = Conv2Geom<KS, TY>;
  constexpr int NT = G::NT;
  constexpr int PAD = KS / 2, KK = KS * KS;
  constexpr int TW = G::TW, TH = G::TH, TWP = kConvPitch2;
  constexpr int BUF = Conv2Smem<KS, CO_T, TY>::BUF;
  constexpr int NW4 = kConvCi2 * KK * (CO_T / 4);
  extern __shared__ __align__(16) float conv2_smem[];
  const int tid = threadIdx.x;
  const int ty0 = (blockIdx.x / tiles_x) * TY, tx0 = (blockIdx.x % tiles_x) * kConvTile;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;
  const float* inb = in + (long long)b * in_bs;
  const long long plane = (long long)ny * nx;
  // this thread's positions of the (TH x TW) tile + halo: source offset inside a channel plane, destination offset inside s_in[ci]
  int src[G::PPT], dst[G::PPT];
#pragma unroll
  for (int k = 0; k < G::PPT; ++k) {
    const int pos = tid + NT * k;
    const int r = pos / TW, cc = pos % TW;
    src[k] = pos < G::NPOS ? wrap(ty0 + r - PAD, ny) * nx + wrap(tx0 + cc - PAD, nx) : -1;
    dst[k] = r * TWP + cc;
  }
  auto fill = [&](int buf, int ci0) {
    float* s_in = conv2_smem + buf * BUF;
    float* s_w = s_in + G::IN_FLOATS;
#pragma unroll
    for (int ci = 0; ci < kConvCi2; ++ci) {
      const bool live = ci0 + ci < Cin;
      const float* pl = inb + (long long)(ci0 + ci) * plane;
#pragma unroll
      for (int k = 0; k < G::PPT; ++k) {
        if (src[k] < 0) continue;
        float* d = s_in + ci * TH * TWP + dst[k];
        if (live) cp_async4(d, pl + src[k]); else *d = 0.f;
      }
    }
    for (int e = tid; e < NW4; e += NT) {
      const int ci = e / (KK * (CO_T / 4)), rem = e % (KK * (CO_T / 4));
      float* d = s_w + (ci * KK) * CO_T + rem * 4;
      if (ci0 + ci < Cin) cp_async16(d, wp + ((long long)(ci0 + ci) * KK + rem / (CO_T / 4)) * CoutPad + co0 + (rem % (CO_T / 4)) * 4);
      else *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  // thread = 4 adjacent pixels (x) of one row x 16 output channels: pixel group pxg (4 per row), row pr (32 per tile), channel half ch
  constexpr int CH = CO_T / 2;
  static_assert(CH % 4 == 0, "conv_ffma2_kernel: CO_T must be a multiple of 8");
  const int pxg = tid % 4, pr = (tid / 4) % TY, ch = tid / (4 * TY);
  float acc[4][CH];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[p][j] = 0.f;

  fill(0, 0);
  int buf = 0;
  for (int ci0 = 0; ci0 < Cin; ci0 += kConvCi2, buf ^= 1) {
    if (ci0 + kConvCi2 < Cin) { fill(buf ^ 1, ci0 + kConvCi2); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    const float* s_in = conv2_smem + buf * BUF;
    const float* s_w = s_in + G::IN_FLOATS;
#pragma unroll 1
    for (int ci = 0; ci < kConvCi2; ++ci) {
      const float* si = s_in + ci * TH * TWP + pr * TWP + pxg * 4;
      const float* sw = s_w + ci * KK * CO_T + ch * CH;
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        // the 4 + KS - 1 <= 8 input values of this row that the thread's 4 pixels see: two aligned 16-byte loads
        const float4 lo = *reinterpret_cast<const float4*>(si + ky * TWP);
        const float4 hi = *reinterpret_cast<const float4*>(si + ky * TWP + 4);
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float4* w4 = reinterpret_cast<const float4*>(sw + (ky * KS + kx) * CO_T);
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 w = w4[j];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              acc[p][4 * j + 0] = fmaf(v[p + kx], w.x, acc[p][4 * j + 0]);
              acc[p][4 * j + 1] = fmaf(v[p + kx], w.y, acc[p][4 * j + 1]);
              acc[p][4 * j + 2] = fmaf(v[p + kx], w.z, acc[p][4 * j + 2]);
              acc[p][4 * j + 3] = fmaf(v[p + kx], w.w, acc[p][4 * j + 3]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  const int y = ty0 + pr, x0 = tx0 + pxg * 4;
  if (y < ny) {
    float* ob = out + (long long)b * out_bs;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int co = co0 + ch * CH + j;
      if (co < Cout) {
        const float bi = bias[co], sc = relu_bn ? bn_s[co] : 1.f, sh = relu_bn ? bn_t[co] : 0.f;
        float* o = ob + ((long long)co * ny + y) * nx + x0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          if (x0 + p < nx) {
            float v = acc[p][j] + bi;
            if (relu_bn) v = fmaxf(v, 0.f) * sc + sh;
            if (softplus) v = v > 20.f ? v : log1pf(expf(v));  // torch softplus, beta=1, threshold=20
            o[p] = accumulate ? o[p] + v : v;
          }
        }
      }
    }
  }
}

// launch helper: picks the tile height, computes the grid, raises the dynamic shared-memory limit (the double buffer needs more than
// the 48 KB default; the attribute is per device and the call costs nothing next to the kernel)
template <int KS, int CO_T, int TY>
inline cudaError_t launch_conv_ffma2_ty(int nimg, cudaStream_t st, const float* in, long long in_bs, float* out, long long out_bs,
                                        const float* wp, const float* bias, const float* bn_s, const float* bn_t, int Cin, int Cout,
                                        int CoutPad, int ny, int nx, int relu_bn, int softplus, int accumulate) {
  constexpr size_t smem = Conv2Smem<KS, CO_T, TY>::BYTES;
  cudaError_t e = cudaFuncSetAttribute(conv_ffma2_kernel<KS, CO_T, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int tiles_x = (nx + kConvTile - 1) / kConvTile, tiles_y = (ny + TY - 1) / TY;
  dim3 grid(tiles_x * tiles_y, (Cout + CO_T - 1) / CO_T, nimg);
  conv_ffma2_kernel<KS, CO_T, TY><<<grid, 8 * TY, smem, st>>>(in, in_bs, out, out_bs, wp, bias, bn_s, bn_t, Cin, Cout, CoutPad, ny, nx,
                                                              tiles_x, relu_bn, softplus, accumulate);
  return cudaGetLastError();
}
template <int KS, int CO_T>
inline cudaError_t launch_conv_ffma2(int nimg, cudaStream_t st, const float* in, long long in_bs, float* out, long long out_bs,
                                     const float* wp, const float* bias, const float* bn_s, const float* bn_t, int Cin, int Cout,
                                     int CoutPad, int ny, int nx, int relu_bn, int softplus, int accumulate) {
  const int rows32 = (ny + 31) / 32 * 32, rows16 = (ny + 15) / 16 * 16;       // rows computed with either tile height
  if (rows16 < rows32)
    return launch_conv_ffma2_ty<KS, CO_T, 16>(nimg, st, in, in_bs, out, out_bs, wp, bias, bn_s, bn_t, Cin, Cout, CoutPad, ny, nx, relu_bn,
                                              softplus, accumulate);
  return launch_conv_ffma2_ty<KS, CO_T, 32>(nimg, st, in, in_bs, out, out_bs, wp, bias, bn_s, bn_t, Cin, Cout, CoutPad, ny, nx, relu_bn,
                                            softplus, accumulate);
}

}  // namespace qgb
