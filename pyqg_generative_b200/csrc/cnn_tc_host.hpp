// cnn_tc_host.hpp -- host-side types and entry points of the tcgen05 AndrewCNN forward (kernels: cnn_tc.cuh, compiled in
// tu_cnn_tc.cu).  api.cu only sees these declarations, so the tensor-core kernels build as their own translation unit.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/qgb200.h"
#include "prof.hpp"

namespace qgb {

// Epilogue: y = relu(acc / wscale + b) * s + t  (eval-mode BatchNorm after the ReLU).  The constants travel as a kernel
// parameter.  Layers with COUT <= 32 keep them in REGISTERS for the whole persistent loop, with the BN scale folded on the
// host into the layer's own weights and bias (relu(z) s = sign(s) relu(|s| z); the sign goes into the next layer's weights)
// so that two constants per channel remain.  A first version read bias / scale / shift from shared memory everywhere: 24
// LDS.128 per 32 channels took 25 % of the shared-memory data pipe that the MMAs' operand reads saturate in the thin
// layers (ncu l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld, profiles/r1_history.md).  The wide layers (1 and 2) spend
// 10-40x more MMA time per epilogue item and keep the three constants in shared memory, unfolded.  The SHIFT is never
// folded forward: storing relu(|s| z) without it costs up to 16x in absolute precision where the ReLU output has a large
// mean (measured 1.2e-3 on the shipped VAE decoder instead of 4e-4).
struct TcEpi {
  float b[128], s[128], t[128];
};

struct TcLayer {
  int cin = 0, cout = 0, ks = 0, relu = 0;      // MMA-level shapes (padded): cin multiple of 32, cout = MMA N
  int real_cout = 0, passes = 3;
  __half* w = nullptr;
  TcEpi epi;
  float inv_wscale = 1.f;
};
struct TcNet {
  bool ready = false;
  int cin0 = 0;          // real input channels of the network (4 or 2)
  int kp = 0;            // padded im2col K of layer 1 (128 or 64)
  std::vector<TcLayer> layers;
  TcLayer l2_fast;       // layer 2 packed for the single-pass variant (QGB_PREC_TC_FAST)
  TcLayer l1_direct;     // layer 1 packed for the im2col-free kernel (conv_l1_direct_kernel)
};
struct TcWorkspace {
  __half* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // a0_hi, a0_lo, ping hi/lo, pong hi/lo
  size_t halves[6] = {0, 0, 0, 0, 0, 0};
  // per-layer timing hook (qgb_profile_*, prof.hpp): set by api.cu around a forward pass; slot = 8 * prof_net + layer
  Profiler* prof = nullptr;
  int prof_net = 0;
  long long last_launches = 0;   // kernels launched by the latest tc_forward
  // latent noise generated inside layer 1 (direct kernel, 4-channel input): set by api.cu for ONE forward pass
  // closure epilogue fused into the last layer (set by api.cu for ONE forward pass): dq = float64(y * y_std) * weight
  double* dq_out = nullptr; float dq_ys[2] = {1.f, 1.f}; double dq_weight = 1.0;
  bool noise_inkernel = false; int noise_member0 = 0; unsigned long long noise_seed = 0; const uint32_t* noise_draw = nullptr;
};


bool tc_l1_direct_enabled();   // the im2col-free layer-1 kernel is in use (QGB_TC_L1=im2col switches back)
void tc_free_net(TcNet& n);
void tc_free_workspace(TcWorkspace& w);
// Accepts exactly the default AndrewCNN architecture; leaves n.ready == false otherwise
int tc_pack_net(TcNet& n, int nlayers, const qgb_cnn_layer* L, std::string* err);
int tc_forward(const TcNet& net, TcWorkspace& ws, const float* x, long long x_bs, float* y, long long y_bs, int batch, int ny,
               int nx, int softplus, int accumulate, int nsm, cudaStream_t st, std::string* err, bool fast_l2 = false);

}  // namespace qgb
