// philox.cuh -- counter-based latent noise: Philox4x32-10 + Box-Muller, keyed by (seed, GLOBAL member, draw, channel, pixel quad) so
// that a run does not depend on how members are sharded over GPUs and a value can be (re)generated anywhere -- by the latent
// kernel (closure.cuh) or inside layer 1 of the generator (cnn_tc.cuh), bit for bit.
// Replaces np.random.randn in generate_latent_noise (models/cgan_regression.py:154-155, models/mean_var_model.py:102-103).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qgb {

// ---------------------------------------------------------------- Philox4x32-10 -----------------------
struct Philox {
  __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ static inline void gen(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                    uint32_t (&out)[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// four N(0,1) samples for counter (member_global, draw, channel, quad): Box-Muller on two uniform pairs
__device__ inline void philox_normal4(uint64_t seed, uint32_t member, uint32_t draw, uint32_t chan, uint32_t quad,
                                      float (&z)[4]) {
  uint32_t r[4];
  Philox::gen(seed, quad, chan, draw, member, r);
  const float two_pi = 6.283185307179586f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float u1 = ((float)(r[2 * j] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    const float u2 = ((float)(r[2 * j + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(two_pi * u2, &s, &c);
    z[2 * j] = rad * c;
    z[2 * j + 1] = rad * s;
  }
}

}  // namespace qgb
