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

}  // namespace adv
}  // namespace qgb
