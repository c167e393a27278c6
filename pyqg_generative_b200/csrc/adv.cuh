// adv.cuh -- device code of the variational (CVAE, ELBO) and adversarial (CGAN, WGAN-GP) training steps, fp32.
//
// Reference: models/cvae_regression.py:165-230 (forward / compute_loss) and :250-300 (train_CVAE);
//            models/cgan_regression.py:173-195 (gradient_penalty) and :227-300 (train_CGAN);
//            tools/cnn_tools.py:212-244 (DCGAN_discriminator, bn='None': four 4x4 stride-2 zero-padded convolutions with
//            LeakyReLU(0.2) and a final nx/16 x nx/16 valid convolution, no biases, no normalisation).
// The generator / encoder / decoder are AndrewCNNs and run through the kernels of train.cuh; this file holds what is specific to
// the two losses and the discriminator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qgb {
namespace adv {

constexpr int kPartBlocks = 256;

template <int NV>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double* __restrict__ out) {
  __shared__ double sh[NV][256];
#pragma unroll
  for (int k = 0; k < NV; ++k) sh[k][threadIdx.x] = v[k];
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) {
#pragma unroll
      for (int k = 0; k < NV; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + w];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = sh[k][0];
  }
}

// out[b][c_off + c][p] = src[b][c][p]   (torch.cat along the channel axis, one source at a time)
__global__ void cat_channels_kernel(const float* __restrict__ src, int C_src, float* __restrict__ out, int C_out, int c_off, int hw,
                                    long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw), c = (int)((i / hw) % C_src);
    const long long b = i / ((long long)hw * C_src);
    out[(b * C_out + c_off + c) * hw + p] = src[i];
  }
}

// ---- CVAE ---------------------------------------------------------------------------------------------------------------
// decoder input = cat[x, z],  z = eps * exp(0.5 logvar) + mu   (cvae_regression.py:168-175; encoder output = [mu(2), logvar(2)])
__global__ void cvae_reparam_kernel(const float* __restrict__ x, const float* __restrict__ encout, const float* __restrict__ eps,
                                    float* __restrict__ decin, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw), c = (int)((i / hw) % 2);
    const long long b = i / (2LL * hw);
    const float mu = encout[(b * 4 + c) * hw + p], logvar = encout[(b * 4 + 2 + c) * hw + p];
    decin[(b * 4 + c) * hw + p] = x[i];
    decin[(b * 4 + 2 + c) * hw + p] = eps[i] * expf(0.5f * logvar) + mu;
  }
}

// partial sums over the (B, 2, ny, nx) elements: (yhat - y)^2, 0.5 (mu^2 + var - 1 - logvar), var, mu, mu^2
__global__ void __launch_bounds__(256) cvae_loss_partial_kernel(const float* __restrict__ yhat, const float* __restrict__ y,
                                                                const float* __restrict__ encout, int hw, long long total,
                                                                double* __restrict__ part) {
  double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw), c = (int)((i / hw) % 2);
    const long long b = i / (2LL * hw);
    const float mu = encout[(b * 4 + c) * hw + p], logvar = encout[(b * 4 + 2 + c) * hw + p];
    const float sd = expf(0.5f * logvar), var = sd * sd;
    const float e = yhat[i] - y[i];
    v[0] += (double)(e * e);
    v[1] += (double)(0.5f * (mu * mu + var - 1.f - logvar));
    v[2] += (double)var;
    v[3] += (double)mu;
    v[4] += (double)mu * mu;
  }
  block_sum_store<5>(v, part + (long long)blockIdx.x * 5);
}

// out[0..5] = loss, loss_recon, loss_KL, MSE, var_latent, var_aggr (compute_loss :177-230); out[6] = 1 / (var_p B), the factor of
// d loss / d yhat = (yhat - y) / (var_p B).  decoder_var < 0: 'adaptive', var_p = MSE of this batch (:209-210).
__global__ void cvae_loss_final_kernel(const double* __restrict__ part, int nparts, double n, double batch, double decoder_var,
                                       double* __restrict__ out) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < nparts; ++i)
    for (int k = 0; k < 5; ++k) s[k] += part[(long long)i * 5 + k];
  const double mse = s[0] / n;
  const double var_p = decoder_var < 0.0 ? (double)(float)mse : decoder_var;
  const double recon = s[0] / batch / (2.0 * var_p), kl = s[1] / batch;
  const double var_latent = s[2] / n, mu_var = (s[4] - s[3] * s[3] / n) / (n - 1.0);
  out[0] = recon + kl; out[1] = recon; out[2] = kl; out[3] = mse; out[4] = var_latent; out[5] = mu_var + var_latent;
  out[6] = 1.0 / (var_p * batch);
}

__global__ void cvae_dyhat_kernel(const float* __restrict__ yhat, const float* __restrict__ y, const double* __restrict__ out,
                                  float* __restrict__ dyhat, long long total) {
  const float scale = (float)out[6];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dyhat[i] = (yhat[i] - y[i]) * scale;
}

// gradient with respect to the encoder output from the gradient with respect to the decoder input (channels 2, 3 = z):
//   d mu = dz + mu / B ;  d logvar = dz eps 0.5 std + 0.5 (var - 1) / B
__global__ void cvae_denc_kernel(const float* __restrict__ ddecin, const float* __restrict__ encout, const float* __restrict__ eps,
                                 float* __restrict__ denc, float inv_b, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw), c = (int)((i / hw) % 2);
    const long long b = i / (2LL * hw);
    const long long im = (b * 4 + c) * hw + p, iv = (b * 4 + 2 + c) * hw + p;
    const float mu = encout[im], logvar = encout[iv], dz = ddecin[iv];
    const float sd = expf(0.5f * logvar);
    denc[im] = dz + mu * inv_b;
    denc[iv] = dz * eps[i] * 0.5f * sd + 0.5f * (sd * sd - 1.f) * inv_b;
  }
}

// y += a x
__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += a * x[i];
}

// ---- discriminator (DCGAN_discriminator, bn = 'None') --------------------------------------------------------------------
// Activations are NHWC, (sample, y, x, channel), so that every convolution is one GEMM on an im2col matrix
//   col[(b, oy, ox)][(ky, kx, ci)]        M = B OH^2 rows, K = 16 C columns (4 x 4 taps, stride 2, zero padding 1)
// against the packed weights Wp[co][(ky, kx, ci)]: forward  z = col Wp^T,  data gradient  dcol = dz Wp,  weight gradient
// dWp = dz^T col.  The last layer (k5 x k5 valid convolution of the k5 x k5 map) is the same GEMM on the activation itself.

constexpr int kGemmM = 128, kGemmK = 8, kGemmPitch = kGemmM + 4;     // (depth 16 spills: 8 x 8 accumulators + staging exceed 128 registers)

// C[i][j] (+ epilogue) = sum_k A(i, k) B(k, j),  A(i, k) = A[i sai + k sak],  B(k, j) = B[k sbk + j sbj]  (one of each pair of
// strides is 1: the loader walks the unit-stride direction).  Block tile 128 x BN (BN = 128 or 64), depth kGemmK per stage, 256 threads
// with an 8 x (BN / 16) register tile each (rows ty 8 .. ty 8 + 7; columns tx 4 .. tx 4 + 3 of every 64-column group, so that the B reads of
// a quarter warp are contiguous): per k one thread issues 2 + BN/64 LDS.128 for 8 BN/16 FFMAs.  The next stage is fetched
// into registers while the current one is multiplied (shared memory double-buffered, one barrier per stage).
// blockIdx.z splits the k range into chunks of ksplit (partial results at C + z c_split); EPI 0: none, 1: LeakyReLU(0.2),
// 2: times the LeakyReLU slope of mask[i ldc + j] (1 where mask > 0, else 0.2).
template <int EPI, int BN>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const float* __restrict__ A, long long sai, long long sak,
                                                       const float* __restrict__ B, long long sbk, long long sbj, float* __restrict__ C,
                                                       long long ldc, int M, int N, int K, int ksplit, long long c_split,
                                                       const float* __restrict__ mask) {
  constexpr int TN = BN / 16;                       // columns per thread
  constexpr int NB = BN * kGemmK / 256;             // B elements fetched per thread and stage (4 or 2)
  __shared__ __align__(16) float As[2][kGemmK][kGemmPitch];
  __shared__ __align__(16) float Bs[2][kGemmK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int i0 = blockIdx.y * kGemmM, j0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * ksplit, kend = min(K, kbeg + ksplit);
  const bool a_kfast = sak == 1, b_kfast = sbk == 1;
  float acc[8][TN];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
  constexpr int NA = kGemmM * kGemmK / 256, KR = 256 / kGemmK;      // A elements per thread and stage; rows per pass of the k-fast loader
  float ra[NA], rb[NB];
  // per-thread element pointers of the first stage (advanced by one stage per fetch) and row / column validity
  int pa[NA], pb[NB];                         // element offsets (every operand of this library is far below 2^31 elements)
  bool va[NA], vb[NB];
  const int ka0 = a_kfast ? tid % kGemmK : tid / 128, dka = a_kfast ? 0 : 2;             // k index of element r of a stage: ka0 + r dka
  const int kb0 = b_kfast ? tid % kGemmK : tid / BN, dkb = b_kfast ? 0 : 256 / BN;
#pragma unroll
  for (int r = 0; r < NA; ++r) {
    const int i = a_kfast ? tid / kGemmK + KR * r : tid % 128;
    va[r] = i0 + i < M;
    pa[r] = (int)((long long)(i0 + i) * sai + (long long)(kbeg + ka0 + r * dka) * sak);
  }
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const int j = b_kfast ? tid / kGemmK + KR * r : tid % BN;
    vb[r] = j0 + j < N;
    pb[r] = (int)((long long)(kbeg + kb0 + r * dkb) * sbk + (long long)(j0 + j) * sbj);
  }
  const int stepa = (int)(kGemmK * sak), stepb = (int)(kGemmK * sbk);
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      ra[r] = (va[r] && k0 + ka0 + r * dka < kend) ? A[pa[r]] : 0.f;
      pa[r] += stepa;
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      rb[r] = (vb[r] && k0 + kb0 + r * dkb < kend) ? B[pb[r]] : 0.f;
      pb[r] += stepb;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int r = 0; r < NA; ++r) {
      const int k = a_kfast ? tid % kGemmK : tid / 128 + 2 * r, i = a_kfast ? tid / kGemmK + KR * r : tid % 128;
      As[buf][k][i] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      const int k = b_kfast ? tid % kGemmK : tid / BN + (256 / BN) * r, j = b_kfast ? tid / kGemmK + KR * r : tid % BN;
      Bs[buf][k][j] = rb[r];
    }
  };
  fetch(kbeg);
  stash(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += kGemmK, buf ^= 1) {
    const bool more = k0 + kGemmK < kend;
    if (more) fetch(k0 + kGemmK);
#pragma unroll
    for (int k = 0; k < kGemmK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
#pragma unroll
      for (int q = 0; q < TN / 4; ++q) {
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][q * 64 + tx * 4]);      // column groups 64 apart: conflict-free
        bv[4 * q] = b4.x; bv[4 * q + 1] = b4.y; bv[4 * q + 2] = b4.z; bv[4 * q + 3] = b4.w;
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (more) { stash(buf ^ 1); __syncthreads(); }
  }
  float* Cz = C + (long long)blockIdx.z * c_split;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int gi = i0 + ty * 8 + a;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < TN; ++b) {
      const int gj = j0 + (b / 4) * 64 + tx * 4 + b % 4;
      if (gj >= N) continue;
      float v = acc[a][b];
      if (EPI == 1) v = v > 0.f ? v : 0.2f * v;
      if (EPI == 2) v *= mask[gi * ldc + gj] > 0.f ? 1.f : 0.2f;
      Cz[gi * ldc + gj] = v;
    }
  }
}

// last layer (one output channel): out[b] = sum_k x[b][k] w[k], one block per sample (a GEMM with N = 1 would run on two blocks)
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out,
                                                     int K) {
  const float* xb = x + (long long)blockIdx.x * K;
  float s = 0.f;
  for (int k = threadIdx.x; k < K; k += 256) s = fmaf(xb[k], w[k], s);
  __shared__ float sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int h = 128; h > 0; h >>= 1) {
    if ((int)threadIdx.x < h) sh[threadIdx.x] += sh[threadIdx.x + h];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// (cout, cin, ks, ks) -> Wp[co][(ky, kx, ci)]  (dir 0)  or the same permutation back (dir 1: packed gradient -> torch layout,
// summing ``splits`` partial results ``stride`` floats apart)
__global__ void disc_pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int cout, int cin, int ks, int dir,
                                 int splits, long long stride) {
  const long long n = (long long)cout * cin * ks * ks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(i % ks), ky = (int)((i / ks) % ks), ci = (int)((i / (ks * ks)) % cin), co = (int)(i / ((long long)ks * ks * cin));
    const long long ip = ((long long)co * ks * ks + ky * ks + kx) * cin + ci;
    if (dir == 0) dst[ip] = src[i];
    else {
      float s = 0.f;
      for (int k = 0; k < splits; ++k) s += src[k * stride + ip];
      dst[i] = s;
    }
  }
}

// im2col of a 4 x 4 / stride 2 / pad 1 convolution, NHWC.  Samples [0, b_split) come from src0, the rest from src1 (sample
// b - b_split): the weight gradient runs one GEMM over the ordinary samples and the linearised ones of the gradient penalty.
template <int VEC>
__global__ void im2col_kernel(const float* __restrict__ src0, const float* __restrict__ src1, int b_split, float* __restrict__ col,
                              int B, int H, int C, int OH) {
  // VEC consecutive channels per thread (VEC = 4: 16-byte accesses, C % 4 == 0; VEC = 2: the 6-channel input layer)
  const int CV = C / VEC;
  const long long total = (long long)B * OH * OH * 16 * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * VEC, kk = (int)((i / CV) % 16);
    const long long m = i / (16LL * CV);
    const int ox = (int)(m % OH), oy = (int)((m / OH) % OH), b = (int)(m / ((long long)OH * OH));
    const int iy = oy * 2 - 1 + kk / 4, ix = ox * 2 - 1 + kk % 4;
    const bool inside = iy >= 0 && iy < H && ix >= 0 && ix < H;
    const float* s = b < b_split ? src0 + (long long)b * H * H * C : src1 + (long long)(b - b_split) * H * H * C;
    float* d = col + (m * 16 + kk) * C + c;
    if (VEC == 4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inside) v = *reinterpret_cast<const float4*>(s + ((long long)iy * H + ix) * C + c);
      *reinterpret_cast<float4*>(d) = v;
    } else if (VEC == 2) {
      float2 v = make_float2(0.f, 0.f);
      if (inside) v = *reinterpret_cast<const float2*>(s + ((long long)iy * H + ix) * C + c);
      *reinterpret_cast<float2*>(d) = v;
    } else {
      *d = inside ? s[((long long)iy * H + ix) * C + c] : 0.f;
    }
  }
}

// transpose of im2col as a gather (deterministic): din[b][iy][ix][c] = sum of the <= 4 col entries that read this input pixel,
// times the LeakyReLU slope of mask (the activation this gradient flows into; nullptr = none)
template <int VEC>
__global__ void col2im_kernel(const float* __restrict__ dcol, float* __restrict__ din, const float* __restrict__ mask, int B, int H,
                              int C, int OH) {
  const int CV = C / VEC;                         // VEC consecutive channels per thread (4: 16-byte accesses, 2: the 6-channel input)
  const long long total = (long long)B * H * H * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % CV) * VEC, ix = (int)((i / CV) % H), iy = (int)((i / ((long long)CV * H)) % H), b = (int)(i / ((long long)CV * H * H));
    float s[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s[v] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ky = ((iy + 1) & 1) + 2 * a, oy = (iy + 1 - ky) / 2;
      if (iy + 1 - ky < 0 || oy >= OH) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kx = ((ix + 1) & 1) + 2 * e, ox = (ix + 1 - kx) / 2;
        if (ix + 1 - kx < 0 || ox >= OH) continue;
        const float* p = dcol + (((long long)b * OH + oy) * OH + ox) * 16 * C + (ky * 4 + kx) * C + c;
        if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(p); s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w; }
        else if (VEC == 2) { const float2 t = *reinterpret_cast<const float2*>(p); s[0] += t.x; s[1] += t.y; }
        else s[0] += *p;
      }
    }
    const long long o = (((long long)b * H + iy) * H + ix) * C + c;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (mask) s[v] *= mask[o + v] > 0.f ? 1.f : 0.2f;
      din[o + v] = s[v];
    }
  }
}

// discriminator input (NHWC, 6 channels) of sample block ``mode``:  0: [x, ytrue, yf2]   1: [x, yf1, ytrue]   2: [x, yf1, yf2]
// 3: [x, eps ytrue_cat + (1 - eps) yfake_cat] with ytrue_cat = (ytrue, yf2) if coin == 0 else (yf1, ytrue), yfake_cat = (yf1, yf2)
// (cgan_regression.py:173-185, 266-268).  x, ytrue, yf1, yf2: (B, 2, hw) NCHW.
__global__ void disc_input_kernel(const float* __restrict__ x, const float* __restrict__ yt, const float* __restrict__ yf1,
                                  const float* __restrict__ yf2, const float* __restrict__ eps, int coin, int mode,
                                  float* __restrict__ out, int B, int hw) {
  const long long total = (long long)B * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw);
    const long long b = i / hw;
    float v[6];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const long long s = (b * 2 + c) * hw + p;
      const float t = yt[s], f1 = yf1[s], f2 = yf2[s];
      v[c] = x[s];
      if (mode == 0) { v[2 + c] = t; v[4 + c] = f2; }
      else if (mode == 1) { v[2 + c] = f1; v[4 + c] = t; }
      else if (mode == 2) { v[2 + c] = f1; v[4 + c] = f2; }
      else {
        const float e = eps[b];
        const float ta = coin == 0 ? t : f1, tb = coin == 0 ? f2 : t;
        v[2 + c] = e * ta + (1.f - e) * f1;
        v[4 + c] = e * tb + (1.f - e) * f2;
      }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) out[i * 6 + c] = v[c];
  }
}

// NCHW (B, C, hw) -> NHWC
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), p = (int)((i / C) % hw);
    const long long b = i / ((long long)C * hw);
    out[i] = in[(b * C + c) * hw + p];
  }
}

// channels [c0, c0 + 2) of an NHWC (B, hw, 6) array -> NCHW (B, 2, hw)
__global__ void nhwc_extract_kernel(const float* __restrict__ in, float* __restrict__ out, int c0, int hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw), c = (int)((i / hw) % 2);
    const long long b = i / (2LL * hw);
    out[i] = in[(b * hw + p) * 6 + c0 + c];
  }
}

// WGAN losses of one iteration (cgan_regression.py:269-272) and the output gradients.  o: (4B) = D(true1), D(true2), D(fake),
// D(interp).  stats[0] = D_loss = -0.5 (mean true1 + mean true2) + mean fake, stats[2] = D_drift = 1e-3 mean true1^2.
// d5: d error / d o for the first 3B samples; 1 for the interpolates (their backward pass gives dD/dy, the penalty's argument).
__global__ void disc_loss_kernel(const float* __restrict__ o, int B, double lambda_drift, float* __restrict__ d5,
                                 double* __restrict__ stats) {
  __shared__ double sh[3][256];
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, sq = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float t1 = o[b];
    s1 += t1; s2 += o[B + b]; s3 += o[2 * B + b]; sq += (double)t1 * t1;
    d5[b] = (float)((-0.5 + 2.0 * lambda_drift * t1) / B);
    d5[B + b] = (float)(-0.5 / B);
    d5[2 * B + b] = (float)(1.0 / B);
    d5[3 * B + b] = 1.f;
  }
  sh[0][threadIdx.x] = -0.5 * (s1 + s2) + s3; sh[1][threadIdx.x] = sq; sh[2][threadIdx.x] = 0.0;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) { sh[0][threadIdx.x] += sh[0][threadIdx.x + w]; sh[1][threadIdx.x] += sh[1][threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { stats[0] = sh[0][0] / B; stats[2] = lambda_drift * sh[1][0] / B; }
}

// G_loss = -mean D(x, yf1, yf2) (:279) and its output gradient -1 / B
__global__ void gen_loss_kernel(const float* __restrict__ o, int B, float* __restrict__ d5, double* __restrict__ stats) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double s = 0.0;
  for (int b = 0; b < B; ++b) { s += o[b]; d5[b] = (float)(-1.0 / B); }
  stats[3] = -s / B;
}

// e4[b][k] = d5[b] Wp5[k] times the LeakyReLU slope of h4 (data gradient of the last layer)
__global__ void disc_last_dgrad_kernel(const float* __restrict__ d5, const float* __restrict__ wp5, const float* __restrict__ h4,
                                       float* __restrict__ out, int K, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long b = i / K;
    out[i] = d5[b] * wp5[k] * (h4[i] > 0.f ? 1.f : 0.2f);
  }
}

// gradient penalty (:173-195): g = dD/dy of the interpolates = channels 2..5 of e0 (B, hw, 6).  One block per sample:
// norm_b = |g_b|_2 ; part[b] = (norm_b - 1)^2 ; coef[b] = lambda 2 (norm_b - 1) / (B norm_b)
__global__ void __launch_bounds__(256) gp_norm_kernel(const float* __restrict__ e0, int hw, int B, double lambda_gp,
                                                      double* __restrict__ part, float* __restrict__ coef) {
  const int b = blockIdx.x;
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < hw * 4; i += blockDim.x) {
    const int p = i / 4, c = 2 + i % 4;
    const float g = e0[((long long)b * hw + p) * 6 + c];
    v[0] += (double)g * g;
  }
  __shared__ double res[1];
  block_sum_store<1>(v, res);
  if (threadIdx.x == 0) {
    const double norm = sqrt(res[0]);
    part[b] = (norm - 1.0) * (norm - 1.0);
    coef[b] = (float)(lambda_gp * 2.0 * (norm - 1.0) / ((double)B * norm));
  }
}
// stats[1] = D_grad = lambda mean_b (norm_b - 1)^2 ;  u0 = d D_grad / d g (zero on the x channels)
__global__ void gp_seed_kernel(const float* __restrict__ e0, const float* __restrict__ coef, const double* __restrict__ part, int B,
                               double lambda_gp, float* __restrict__ u0, int hw, double* __restrict__ stats) {
  const long long total = (long long)B * hw * 6;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % 6);
    const long long b = i / (6LL * hw);
    u0[i] = c >= 2 ? coef[b] * e0[i] : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < B; ++b) s += part[b];
    stats[1] = lambda_gp * s / B;
  }
}

}  // namespace adv
}  // namespace qgb
