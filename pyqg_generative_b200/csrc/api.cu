// api.cu -- C ABI of libqgb200.so (see include/qgb200.h) and the host-side engine that sequences the sm_100a kernels.
// No torch, no CPU fallback: every numerical entry point launches CUDA kernels and fails loudly otherwise.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qgb200.h"
#include "closure.cuh"
#include "cnn_tc_host.hpp"
#include "operators.cuh"
#include "prof.hpp"
#include "qg_host.hpp"
#include "spectral_host.hpp"

using namespace qgb;

static std::atomic<long long> g_launches{0};
static thread_local std::string g_create_error;
#define QGB_COUNT_LAUNCH() (g_launches.fetch_add(1, std::memory_order_relaxed))

namespace {

struct DevNetLayer {
  int cin = 0, cout = 0, ks = 0, relu_bn = 0, cout_pad = 0;
  float *wp = nullptr, *bias = nullptr, *bn_s = nullptr, *bn_t = nullptr;
};
struct DevNet {
  std::vector<DevNetLayer> layers;
  TcNet tc;  // tcgen05 packing of the same network (cnn_tc.cuh); empty when the architecture is not supported
  bool loaded() const { return !layers.empty(); }
};

}  // namespace

struct qgb_handle {
  qgb_config cfg;
  HostTables ht;
  Tables T;
  std::string err;
  // device tables
  cplx* d_tw = nullptr; short* d_pos = nullptr; short* d_pos_id = nullptr; double *d_kv = nullptr, *d_lv = nullptr, *d_a = nullptr, *d_filtr = nullptr;
  // state
  cplx* qh = nullptr; double* q = nullptr; cplx* hist[3] = {nullptr, nullptr, nullptr};
  cplx* ph = nullptr; double *u = nullptr, *v = nullptr, *p = nullptr, *red = nullptr;
  double *d_ke = nullptr, *d_cfl = nullptr; int* d_flags = nullptr;
  double *d_kespec = nullptr, *d_ensspec = nullptr;
  // spectral energy budget (PROG_BUDGET) and pyqg-style time averages of the spectral diagnostics
  cplx* bud_tend = nullptr;          // tendency of the current state (PROG_BUDGET scratch)
  double *bud = nullptr, *bud_scr = nullptr, *bud_sum = nullptr;     // per member / scratch / summed over members
  const double* last_dq = nullptr;   // forcing used by the latest step (closure output or external), for the budget terms
  double* avg = nullptr; long long avg_n = 0; bool avg_on = false; double tavestart = 0.0, taveint = 86400.0;
  long long tc = 0; double t = 0.0; int ablevel = 0;
  int nthreads = 256; size_t smem = 0; int grid = 0;
  bool fixed = false;   // compile-time specialised step kernel available for this nx
  bool reg64 = false;   // register-resident kernel (spectral64.cuh) for nx = 64
  bool regcl = false;   // cluster register-FFT kernel (spectral_cl.cuh) for nx = 128, 256
  int nt64 = 384;
  bool large = false; cplx* scratch = nullptr; double* red_scratch = nullptr;   // cluster path for nx > 96
  int cluster = 8; int large_lines = 0; size_t large_smem = 0;   // lines of a 1-D FFT pass staged per CTA in shared memory
  // closure
  int kind = QGB_CLOSURE_NONE; int precision = QGB_PREC_FP32;
  DevNet nets[2];
  bool calibrated = false; int auto_precision = QGB_PREC_TC;   // per-network precision choice (QGB_PREC_AUTO)
  float x_std[2] = {1.f, 1.f}, y_std[2] = {1.f, 1.f}; double weight = 1.0;
  int sampler = QGB_SAMPLER_AR1, sampler_nsteps = 1, n_mean = 100;
  bool noise_init = false; long long const_counter = 0; uint64_t seed = 0x5eed5eedULL;
  bool fuse_dq = false;         // the next tensor-core forward of network 0 writes the forcing itself (closure epilogue fused)
  bool noise_pending = false;   // this evaluation's white noise is generated inside layer 1 (no latent kernel ran)
  bool noise_regen = false;     // ... so xin channels 2, 3 must be regenerated before the noise is read back
  uint32_t* d_draw = nullptr;   // Philox draw counter, device resident (read by the latent kernel, bumped after it)
  // CUDA-graph replay of the steady-state step (one graph per position of the tendency-history ring)
  cudaGraphExec_t step_graph[3] = {nullptr, nullptr, nullptr}; unsigned long long graph_key = 0; bool graph_failed = false;
  long long graph_replays = 0, graph_launches[3] = {0, 0, 0};
  bool graph_noise_regen[3] = {false, false, false};
  cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the caller's may be the legacy default stream, which cannot capture)
  float* xin = nullptr;     // (B, cin0, N, N) closure input: normalised q [+ latent z for gan/vae]
  int xin_c = 0; bool x_valid = false;
  double* z64 = nullptr;    // gz latent (B,2,N,N)
  void* xi_inj = nullptr; int xi_dtype = 0; bool xi_set = false;  // injected white noise (device copy)
  float* ynet[2] = {nullptr, nullptr};  // network outputs (B,2,N,N)
  float* yacc = nullptr;    // deterministic-mode accumulator
  double* dq = nullptr;     // closure forcing (B,2,N,N), not demeaned
  double* dq_dm = nullptr;  // demeaned copy served to qgb_get
  bool dq_valid = false;
  double calib_err[4] = {-1.0, -1.0, -1.0, -1.0};   // measured rel-L2 / max-norm error of tc, tc_fast against fp32
  double* dq_ext = nullptr; bool ext_set = false;  // externally supplied forcing (qgb_set_forcing)
  float *stage_x = nullptr, *stage_y = nullptr; size_t stage_x_floats = 0, stage_y_floats = 0;   // qgb_cnn_forward host staging
  float* f32_stage = nullptr;   // float32 copy of a real field on its way to the host (qgb_get_f32)
  float* act[2] = {nullptr, nullptr}; size_t act_floats = 0; int act_chunk = 0;  // fp32 path ping-pong activations
  TcWorkspace tcw;
  int nsm = 148;
  // per-layer profiling (qgb_profile_begin/end)
  Profiler prof;
};

namespace {

int fail(qgb_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                        \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) return fail(h, QGB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T)); }

template <typename T>
cudaError_t upload(T** p, const std::vector<T>& v) {
  cudaError_t e = dalloc(p, v.size());
  if (e != cudaSuccess) return e;
  return cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

inline cudaStream_t S(void* s) { return (cudaStream_t)s; }
inline size_t nreal(const qgb_handle* h) { return (size_t)h->cfg.members * 2 * h->ht.N * h->ht.N; }
inline size_t ncplx(const qgb_handle* h) { return (size_t)h->cfg.members * 2 * h->ht.N * h->ht.NK; }

StepIO base_io(qgb_handle* h) {
  StepIO io;
  std::memset(&io, 0, sizeof(io));
  io.qh = h->qh; io.q = h->q;
  io.Hi_over_H[0] = h->ht.Hi_over_H[0]; io.Hi_over_H[1] = h->ht.Hi_over_H[1];
  io.x_std[0] = h->x_std[0]; io.x_std[1] = h->x_std[1];
  io.x_inv[0] = 1.0f / h->x_std[0]; io.x_inv[1] = 1.0f / h->x_std[1];
  io.bud_F = h->ht.Hi_over_H[0] * h->ht.Hi_over_H[1] / (h->cfg.rd * h->cfg.rd);
  io.bud_U = h->cfg.U1 - h->cfg.U2;
  return io;
}

SpectralPlan make_plan(const qgb_handle* h) {
  SpectralPlan p;
  p.N = h->ht.N; p.members = h->cfg.members; p.grid = h->grid; p.nthreads = h->nthreads; p.smem = h->smem;
  p.fixed = h->fixed; p.nt64 = h->nt64; p.reg64 = h->reg64; p.regcl = h->regcl; p.large = h->large; p.cluster = h->cluster; p.large_lines = h->large_lines;
  p.large_smem = h->large_smem; p.scratch = h->scratch; p.red_scratch = h->red_scratch; p.true_pos = h->d_pos;
  return p;
}

int launch_program(qgb_handle* h, const StepIO& io, int prog, cudaStream_t st, const Tables* Tov = nullptr) {
  CUDA_TRY(h, spectral_launch(make_plan(h), Tov ? *Tov : h->T, io, prog, st));
  QGB_COUNT_LAUNCH();
  return QGB_OK;
}

void set_cnn_io(qgb_handle* h, StepIO& io) {
  if (h->kind != QGB_CLOSURE_NONE && h->xin) {
    io.cnn_x = h->xin;
    io.cnn_mstride = (long long)h->xin_c * h->ht.N * h->ht.N;
  }
}

// ---- fp32 network forward over ``batch`` images resident on the device -------------------------------------
int net_forward_fp32(qgb_handle* h, const DevNet& net, const float* x, long long x_bs, float* y, long long y_bs,
                     int batch, int ny, int nx, int softplus, int accumulate, cudaStream_t st) {
  int maxc = 0;
  for (auto& L : net.layers) maxc = L.cout > maxc ? L.cout : maxc;
  const size_t per_img = (size_t)maxc * ny * nx;
  // ping-pong activations sized for a chunk of images
  int chunk = batch < 64 ? batch : 64;
  if (h->act_floats < per_img * chunk) {
    for (int i = 0; i < 2; ++i) { if (h->act[i]) cudaFree(h->act[i]); h->act[i] = nullptr; }
    CUDA_TRY(h, dalloc(&h->act[0], per_img * chunk));
    CUDA_TRY(h, dalloc(&h->act[1], per_img * chunk));
    h->act_floats = per_img * chunk;
  }
  const int tiles_x = (nx + kConvTile - 1) / kConvTile, tiles_y = (ny + kConvTile - 1) / kConvTile;
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = batch - b0 < chunk ? batch - b0 : chunk;
    const float* in = x + (long long)b0 * x_bs;
    long long in_bs = x_bs;
    for (size_t li = 0; li < net.layers.size(); ++li) {
      const DevNetLayer& L = net.layers[li];
      const bool last = li + 1 == net.layers.size();
      float* out = last ? y + (long long)b0 * y_bs : h->act[li & 1];
      const long long out_bs = last ? y_bs : (long long)L.cout * ny * nx;
      const int sp = last ? softplus : 0, acc = last ? accumulate : 0;
      const bool small = L.cout <= 4;
      const int co_t = small ? 2 : 32;
      dim3 grid(tiles_x * tiles_y, (L.cout + co_t - 1) / co_t, nb);          // (the register-tiled kernel of the wide layers sizes its own grid)
      const int pi = h->prof.start(8 * (int)(&net - h->nets) + (int)li, st);
#define QGB_CONV(KS, CT)                                                                                         \
  conv_ffma_kernel<KS, CT><<<grid, 256, 0, st>>>(in, in_bs, out, out_bs, L.wp, L.bias, L.bn_s, L.bn_t, L.cin, L.cout, \
                                                 L.cout_pad, ny, nx, tiles_x, L.relu_bn, sp, acc)
#define QGB_CONV2(KS, CT)                                                                                        \
  CUDA_TRY(h, (launch_conv_ffma2<KS, CT>(nb, st, in, in_bs, out, out_bs, L.wp, L.bias, L.bn_s, L.bn_t, L.cin, L.cout, \
                                         L.cout_pad, ny, nx, L.relu_bn, sp, acc)))
      if (L.ks == 5 && !small) QGB_CONV2(5, 32);
      else if (L.ks == 5) QGB_CONV(5, 2);
      else if (L.ks == 3 && !small) QGB_CONV2(3, 32);
      else if (L.ks == 3) QGB_CONV(3, 2);
      else if (L.ks == 1 && !small) QGB_CONV(1, 32);
      else if (L.ks == 1) QGB_CONV(1, 2);
      else return fail(h, QGB_EUNSUPPORTED, "kernel size %d not supported (1, 3, 5)", L.ks);
#undef QGB_CONV2
#undef QGB_CONV
      QGB_COUNT_LAUNCH();
      CUDA_TRY(h, cudaGetLastError());
      h->prof.stop(pi, st, nb);
      in = out;
      in_bs = out_bs;
    }
  }
  return QGB_OK;
}

int net_forward(qgb_handle* h, int net, const float* x, long long x_bs, float* y, long long y_bs, int batch, int ny,
                int nx, int softplus, int accumulate, int precision, cudaStream_t st) {
  if (precision == QGB_PREC_TC || precision == QGB_PREC_TC_FAST) {
    if (!h->nets[net].tc.ready) return fail(h, QGB_EUNSUPPORTED, "tcgen05 path: network architecture not supported");
    std::string e;
    h->tcw.prof = &h->prof;
    h->tcw.prof_net = net;
    const bool gen = h->noise_pending && net == 0 && x == h->xin;
    h->tcw.dq_out = (h->fuse_dq && net == 0) ? h->dq : nullptr; h->tcw.dq_ys[0] = h->y_std[0]; h->tcw.dq_ys[1] = h->y_std[1]; h->tcw.dq_weight = h->weight;
    h->tcw.noise_inkernel = gen; h->tcw.noise_member0 = h->cfg.member_offset; h->tcw.noise_seed = h->seed; h->tcw.noise_draw = h->d_draw;
    int rc = tc_forward(h->nets[net].tc, h->tcw, x, x_bs, y, y_bs, batch, ny, nx, softplus, accumulate, h->nsm, st, &e,
                        precision == QGB_PREC_TC_FAST);
    h->tcw.noise_inkernel = false;
    h->tcw.dq_out = nullptr;
    if (rc != 0) return fail(h, rc, "%s", e.c_str());
    g_launches.fetch_add(h->tcw.last_launches, std::memory_order_relaxed);
    if (gen) {                       // the draw has been consumed by layer 1
      h->noise_pending = false;
      bump_counter_kernel<<<1, 1, 0, st>>>(h->d_draw);
      QGB_COUNT_LAUNCH();
    }
    return QGB_OK;
  }
  return net_forward_fp32(h, h->nets[net], x, x_bs, y, y_bs, batch, ny, nx, softplus, accumulate, st);
}

template <typename T>
int launch_latent(qgb_handle* h, T* z, long long mstride, double a, double b, int replace, cudaStream_t st, int draw_bias = 0, bool bump = true) {
  const int npix = h->ht.N * h->ht.N;
  const long long total = (long long)h->cfg.members * 2 * ((npix + 3) / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const T* inj = h->xi_set ? (const T*)h->xi_inj : nullptr;
  const int pi = h->prof.start(PROF_LATENT, st);
  latent_update_kernel<T><<<blocks, 256, 0, st>>>(z, mstride, npix, h->cfg.members, h->cfg.member_offset, h->seed,
                                                   h->d_draw, draw_bias, (T)a, (T)b, replace, inj);
  if (bump) { bump_counter_kernel<<<1, 1, 0, st>>>(h->d_draw); QGB_COUNT_LAUNCH(); }
  h->prof.stop(pi, st, h->cfg.members);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(h, cudaGetLastError());
  return QGB_OK;
}

// white-noise draw (+AR1 blend) into the latent storage of the active closure
int draw_latent(qgb_handle* h, double a, double b, int replace, cudaStream_t st, bool allow_inkernel = false) {
  const long long npix = (long long)h->ht.N * h->ht.N;
  h->noise_pending = false;
  if (allow_inkernel && (replace || (a == 0.0 && b == 1.0)) && !h->xi_set && (h->kind == QGB_CLOSURE_GAN || h->kind == QGB_CLOSURE_VAE) &&
      h->nets[0].tc.ready && tc_l1_direct_enabled() && h->ht.N % 16 == 0 && !getenv("QGB_NOISE_STORED")) {
    const int prec = h->precision == QGB_PREC_AUTO ? (h->calibrated ? h->auto_precision : QGB_PREC_FP32) : h->precision;
    if (prec == QGB_PREC_TC || prec == QGB_PREC_TC_FAST) {
      // white noise, tensor-core generator: layer 1 generates z from the Philox counters while it builds its input window
      // (north_star (4): "Philox latent noise generated in-kernel"); the draw counter is bumped after that forward pass
      h->noise_pending = true;
      h->noise_regen = true;
      return QGB_OK;
    }
  }
  h->noise_regen = false;
  if (h->kind == QGB_CLOSURE_GAN || h->kind == QGB_CLOSURE_VAE) {
    if (h->xi_set && h->xi_dtype != 0) return fail(h, QGB_ESTATE, "injected latent must be float32 for gan/vae");
    return launch_latent<float>(h, h->xin + 2 * npix, 4 * npix, a, b, replace, st);
  }
  if (h->kind == QGB_CLOSURE_GZ) {
    if (h->xi_set && h->xi_dtype != 1) return fail(h, QGB_ESTATE, "injected latent must be float64 for gz");
    return launch_latent<double>(h, h->z64, 2 * npix, a, b, replace, st);
  }
  return QGB_OK;  // ols: generate_latent_noise returns 0 (models/ols_model.py:68-69)
}

// ---- per-network precision calibration (QGB_PREC_AUTO) ---------------------------------------------------------------
// partial sums of one comparison: out[4*block + {0,1,2,3}] = sum (y-ref)^2, sum ref^2, max |y-ref|, max |ref|
__global__ void err_partial_kernel(const float* __restrict__ y, const float* __restrict__ ref, long long n, double* __restrict__ out) {
  __shared__ double sh[4][256];
  double d2 = 0.0, r2 = 0.0, dm = 0.0, rm = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double r = ref[i], d = (double)y[i] - r;
    d2 += d * d; r2 += r * r; dm = fmax(dm, fabs(d)); rm = fmax(rm, fabs(r));
  }
  sh[0][threadIdx.x] = d2; sh[1][threadIdx.x] = r2; sh[2][threadIdx.x] = dm; sh[3][threadIdx.x] = rm;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
      sh[2][threadIdx.x] = fmax(sh[2][threadIdx.x], sh[2][threadIdx.x + o]); sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) for (int j = 0; j < 4; ++j) out[4 * blockIdx.x + j] = sh[j][0];
}

// The split-precision tensor-core plan has two variants of layer 2 (75 % of the flops): two passes (a_hi + a_lo) w_hi ("tc") or one
// pass a_hi w_hi ("tc_fast", ~18 % faster end to end).  Which one keeps the closure output within the 1e-3 relative (L2)
// tolerance depends on the NETWORK (shipped GAN: 4.4e-4 single-pass; shipped VAE decoder / GZ mean net: 1.3-1.8e-3), so with
// precision 'auto' the choice is measured once per loaded network set, at its first evaluation, on up to 16 members of the
// actual closure input: fp32 FFMA forward (the parity reference) against both variants; tc_fast is taken when every network's
// relative L2 error stays below kAutoFastTol (30 % margin), tc when that holds for tc, fp32 otherwise.  One-time cost: three
// small forwards and a stream synchronisation.
constexpr double kAutoFastTol = 7e-4, kAutoTcTol = 1e-3;
int calibrate_precision(qgb_handle* h, cudaStream_t st) {
  const int N = h->ht.N, B = h->cfg.members;
  const long long npix = (long long)N * N, x_bs = (long long)h->xin_c * npix;
  h->calibrated = true;
  h->auto_precision = QGB_PREC_FP32;
  for (int i = 0; i < 4; ++i) h->calib_err[i] = -1.0;
  const int nnets = h->kind == QGB_CLOSURE_GZ ? 2 : 1;
  for (int k = 0; k < nnets; ++k) if (!h->nets[k].tc.ready) return QGB_OK;
  if (N % 16) return QGB_OK;
  const int nb = B < 16 ? B : 16;
  const size_t n = (size_t)nb * 2 * npix;
  float* yb = nullptr; double* part = nullptr;
  constexpr int NBLK = 64;
  CUDA_TRY(h, dalloc(&yb, 3 * n));
  if (cudaError_t ce = dalloc(&part, (size_t)2 * NBLK * 4); ce != cudaSuccess) { cudaFree(yb); return fail(h, QGB_ECUDA, "cudaMalloc failed"); }
  double worst[2][2] = {{0, 0}, {0, 0}};      // [variant tc / tc_fast][rel-L2, max-norm]
  int rc = QGB_OK;
  for (int k = 0; k < nnets && rc == QGB_OK; ++k) {
    const int sp = k == 1 ? 1 : 0;
    rc = net_forward(h, k, h->xin, x_bs, yb, 2 * npix, nb, N, N, sp, 0, QGB_PREC_FP32, st);
    if (rc == QGB_OK) rc = net_forward(h, k, h->xin, x_bs, yb + n, 2 * npix, nb, N, N, sp, 0, QGB_PREC_TC, st);
    if (rc == QGB_OK) rc = net_forward(h, k, h->xin, x_bs, yb + 2 * n, 2 * npix, nb, N, N, sp, 0, QGB_PREC_TC_FAST, st);
    if (rc != QGB_OK) break;
    err_partial_kernel<<<NBLK, 256, 0, st>>>(yb + n, yb, (long long)n, part);
    err_partial_kernel<<<NBLK, 256, 0, st>>>(yb + 2 * n, yb, (long long)n, part + NBLK * 4);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    std::vector<double> hp((size_t)2 * NBLK * 4);
    cudaError_t ce = cudaMemcpyAsync(hp.data(), part, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { rc = fail(h, QGB_ECUDA, "precision calibration failed: %s", cudaGetErrorString(ce)); break; }
    for (int v = 0; v < 2; ++v) {
      double d2 = 0, r2 = 0, dm = 0, rm = 0;
      for (int b = 0; b < NBLK; ++b) {
        const double* q = hp.data() + ((size_t)v * NBLK + b) * 4;
        d2 += q[0]; r2 += q[1]; dm = std::fmax(dm, q[2]); rm = std::fmax(rm, q[3]);
      }
      const double l2 = r2 > 0 ? std::sqrt(d2 / r2) : (d2 > 0 ? 1.0 : 0.0), mx = rm > 0 ? dm / rm : (dm > 0 ? 1.0 : 0.0);
      worst[v][0] = std::fmax(worst[v][0], l2);
      worst[v][1] = std::fmax(worst[v][1], mx);
    }
  }
  cudaFree(yb); cudaFree(part);
  if (rc != QGB_OK) return rc;
  h->calib_err[0] = worst[0][0]; h->calib_err[1] = worst[0][1]; h->calib_err[2] = worst[1][0]; h->calib_err[3] = worst[1][1];
  if (worst[1][0] <= kAutoFastTol) h->auto_precision = QGB_PREC_TC_FAST;
  else if (worst[0][0] <= kAutoTcTol) h->auto_precision = QGB_PREC_TC;
  return QGB_OK;
}

// Parameterization.__call__ body.  Returns via *computed whether a new forcing was produced.
int closure_update(qgb_handle* h, cudaStream_t st) {
  if (h->kind == QGB_CLOSURE_NONE) return fail(h, QGB_ESTATE, "no closure loaded");
  if (!h->nets[0].loaded() || (h->kind == QGB_CLOSURE_GZ && !h->nets[1].loaded()))
    return fail(h, QGB_ESTATE, "closure weights not loaded");
  const int N = h->ht.N, B = h->cfg.members;
  const long long npix = (long long)N * N, total = (long long)B * 2 * npix;
  if (!h->x_valid) {
    StepIO io = base_io(h);
    set_cnn_io(h, io);
    int rc = launch_program(h, io, PROG_EMIT_X, st);
    if (rc) return rc;
    h->x_valid = true;
  }
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const long long x_bs = (long long)h->xin_c * npix;
  bool compute = true;
  int prec = h->precision;       // resolved after the noise update (the calibration evaluates the nets on the current input)
  auto resolve = [&]() -> int {
    if (h->precision != QGB_PREC_AUTO) return QGB_OK;
    if (!h->calibrated) { int rc = calibrate_precision(h, st); if (rc) return rc; }
    prec = h->auto_precision;
    return QGB_OK;
  };
  if (h->sampler == QGB_SAMPLER_DETERMINISTIC) {
    if (h->precision == QGB_PREC_AUTO && !h->calibrated) { int rc0 = draw_latent(h, 0.0, 1.0, 1, st); if (rc0) return rc0; }
    if (int rc0 = resolve()) return rc0;
    // predict_mean_snapshot: mean of M generator samples (gan/vae); the mean net (gz); the net itself (ols)
    if (h->kind == QGB_CLOSURE_GAN || h->kind == QGB_CLOSURE_VAE) {
      if (!h->yacc) CUDA_TRY(h, dalloc(&h->yacc, (size_t)total));
      for (int m = 0; m < h->n_mean; ++m) {
        int rc = draw_latent(h, 0.0, 1.0, 1, st);
        if (rc) return rc;
        rc = net_forward(h, 0, h->xin, x_bs, h->yacc, 2 * npix, B, N, N, 0, m > 0, prec, st);
        if (rc) return rc;
      }
      finish_plain_kernel<<<blocks, 256, 0, st>>>(h->yacc, h->dq, (int)npix, total, h->y_std[0], h->y_std[1],
                                                   h->weight, 1.0f / (float)h->n_mean);
      QGB_COUNT_LAUNCH();
    } else if (h->kind == QGB_CLOSURE_GZ) {
      int rc = net_forward(h, 0, h->xin, x_bs, h->ynet[0], 2 * npix, B, N, N, 0, 0, prec, st);
      if (rc) return rc;
      finish_gz_kernel<<<blocks, 256, 0, st>>>(h->ynet[0], nullptr, nullptr, h->dq, (int)npix, total, h->y_std[0],
                                                h->y_std[1], h->weight, 0);
      QGB_COUNT_LAUNCH();
    } else {
      int rc = net_forward(h, 0, h->xin, x_bs, h->ynet[0], 2 * npix, B, N, N, 0, 0, prec, st);
      if (rc) return rc;
      finish_plain_kernel<<<blocks, 256, 0, st>>>(h->ynet[0], h->dq, (int)npix, total, h->y_std[0], h->y_std[1],
                                                   h->weight, 1.0f);
      QGB_COUNT_LAUNCH();
    }
    CUDA_TRY(h, cudaGetLastError());
    h->dq_valid = true;
    return QGB_OK;
  }
  // ---- noise sampler update (tools/stochastic_pyqg.py:30-72) ----
  if (h->sampler == QGB_SAMPLER_AR1) {
    if (h->noise_init) {
      double a = 1.0, b = 0.0;
      if (h->sampler_nsteps > 0) {
        a = 1.0 - 1.0 / h->sampler_nsteps;
        b = std::sqrt(1.0 / h->sampler_nsteps * (2.0 - 1.0 / h->sampler_nsteps));
      }
      int rc = draw_latent(h, a, b, 0, st, true);
      if (rc) return rc;
    } else {
      int rc = draw_latent(h, 0.0, 1.0, 1, st, true);
      if (rc) return rc;
      h->noise_init = true;
    }
  } else {  // constant sampler
    if (h->noise_init) {
      if (h->const_counter % h->sampler_nsteps == 0) {
        int rc = draw_latent(h, 0.0, 1.0, 1, st, true);
        if (rc) return rc;
        h->const_counter = 1;
      } else {
        h->const_counter += 1;
        compute = false;
      }
    } else {
      int rc = draw_latent(h, 0.0, 1.0, 1, st, true);
      if (rc) return rc;
      h->noise_init = true;
      h->const_counter = 1;
    }
  }
  h->xi_set = false;  // an injected xi is consumed by one sampler update
  if (!compute && h->dq_valid) return QGB_OK;
  // ---- predict_snapshot ----
  if (int rc0 = resolve()) return rc0;
  int pf = -1;
  if (h->kind == QGB_CLOSURE_GZ) {
    int rc = net_forward(h, 0, h->xin, x_bs, h->ynet[0], 2 * npix, B, N, N, 0, 0, prec, st);
    if (rc) return rc;
    rc = net_forward(h, 1, h->xin, x_bs, h->ynet[1], 2 * npix, B, N, N, 1, 0, prec, st);
    if (rc) return rc;
    pf = h->prof.start(PROF_FINISH, st);
    finish_gz_kernel<<<blocks, 256, 0, st>>>(h->ynet[0], h->ynet[1], h->z64, h->dq, (int)npix, total, h->y_std[0],
                                              h->y_std[1], h->weight, 1);
  } else {
    // tensor-core generator: the last layer's epilogue writes dq = float64(y * y_std) * weight itself
    h->fuse_dq = (prec == QGB_PREC_TC || prec == QGB_PREC_TC_FAST) && !getenv("QGB_NO_FUSED_EPILOGUE");
    const bool fused = h->fuse_dq;
    int rc = net_forward(h, 0, h->xin, x_bs, h->ynet[0], 2 * npix, B, N, N, 0, 0, prec, st);
    h->fuse_dq = false;
    if (rc) return rc;
    if (fused) { CUDA_TRY(h, cudaGetLastError()); h->dq_valid = true; return QGB_OK; }
    pf = h->prof.start(PROF_FINISH, st);
    finish_plain_kernel<<<blocks, 256, 0, st>>>(h->ynet[0], h->dq, (int)npix, total, h->y_std[0], h->y_std[1],
                                                 h->weight, 1.0f);
  }
  h->prof.stop(pf, st, B);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(h, cudaGetLastError());
  h->dq_valid = true;
  return QGB_OK;
}

int ensure_closure_buffers(qgb_handle* h) {
  const size_t npix = (size_t)h->ht.N * h->ht.N, B = h->cfg.members;
  const int cin0 = (h->kind == QGB_CLOSURE_GAN || h->kind == QGB_CLOSURE_VAE) ? 4 : 2;
  if (h->xin && h->xin_c != cin0) { cudaFree(h->xin); h->xin = nullptr; }
  if (!h->xin) {
    CUDA_TRY(h, dalloc(&h->xin, B * cin0 * npix));
    CUDA_TRY(h, cudaMemset(h->xin, 0, B * cin0 * npix * sizeof(float)));
    h->xin_c = cin0;
    h->x_valid = false;
  }
  if (h->kind == QGB_CLOSURE_GZ && !h->z64) CUDA_TRY(h, dalloc(&h->z64, B * 2 * npix));
  for (int i = 0; i < 2; ++i)
    if (!h->ynet[i]) CUDA_TRY(h, dalloc(&h->ynet[i], B * 2 * npix));
  if (!h->dq) {
    CUDA_TRY(h, dalloc(&h->dq, B * 2 * npix));
    CUDA_TRY(h, dalloc(&h->dq_dm, B * 2 * npix));
  }
  return QGB_OK;
}

void invalidate_graphs(qgb_handle* h) {
  for (int i = 0; i < 3; ++i)
    if (h->step_graph[i]) { cudaGraphExecDestroy(h->step_graph[i]); h->step_graph[i] = nullptr; }
}

void free_net(DevNet& n) {
  for (auto& L : n.layers) { cudaFree(L.wp); cudaFree(L.bias); cudaFree(L.bn_s); cudaFree(L.bn_t); }
  n.layers.clear();
  tc_free_net(n.tc);
}

}  // namespace

// =============================================================================================== C ABI ====
extern "C" {

void qgb_default_config(qgb_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->nx = 64; c->members = 1; c->member_offset = 0; c->device = 0;
  c->L = 1e6; c->dt = 7200.0; c->rek = 5.787e-7; c->filterfac = 23.6; c->beta = 1.5e-11; c->rd = 15000.0;
  c->delta = 0.25; c->H1 = 500.0; c->U1 = 0.025; c->U2 = 0.0;
}

const char* qgb_version(void) { return "qgb200 0.1 (sm_100a)"; }
int64_t qgb_launch_count(void) { return (int64_t)g_launches.load(); }
const char* qgb_last_error(const qgb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int qgb_create(const qgb_config* cfg, qgb_handle** out) {
  if (!cfg || !out) return fail(nullptr, QGB_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->members < 1) return fail(nullptr, QGB_EINVAL, "members must be >= 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, QGB_ECUDA, "no CUDA device available (%s): libqgb200 has no CPU fallback", cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, QGB_EINVAL, "device %d out of range", cfg->device);
  qgb_handle* h = new qgb_handle();
  h->cfg = *cfg;
  if (!build_host_tables(*cfg, h->ht)) {
    delete h;
    return fail(nullptr, QGB_EINVAL, "nx=%d unsupported: must be even with prime factors 2 and 3", cfg->nx);
  }
#define CR(expr)                                                                     \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      fail(nullptr, QGB_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));      \
      qgb_destroy(h);                                                                \
      return QGB_ECUDA;                                                              \
    }                                                                                \
  } while (0)
  CR(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CR(cudaGetDeviceProperties(&prop, cfg->device));
  h->nsm = prop.multiProcessorCount;
  h->nthreads = cfg->nx >= 96 ? 512 : 256;
  h->smem = program_smem_bytes(h->ht.N, h->ht.P, h->nthreads);
  h->large = h->smem > (size_t)prop.sharedMemPerBlockOptin;
  if (h->large && cfg->nx > 1024) {
    fail(nullptr, QGB_EUNSUPPORTED, "nx=%d unsupported (the cluster path covers nx <= 1024)", cfg->nx);
    qgb_destroy(h);
    return QGB_EUNSUPPORTED;
  }
  h->fixed = !h->large && (cfg->nx == 32 || cfg->nx == 48 || cfg->nx == 64 || cfg->nx == 96) && !getenv("QGB_GENERIC_STEP");
  if (h->fixed && cfg->nx == 64) {
    const char* e = getenv("QGB_STEP_NT");
    h->nt64 = e ? atoi(e) : 384;   // measured on B200: 256 -> 0.370 ms, 384 -> 0.299 ms, 512 -> 0.301 ms per 1024 members
  }
  h->reg64 = !h->large && cfg->nx == 64 && !getenv("QGB_S64_OFF");
  h->regcl = h->large && (cfg->nx == 128 || cfg->nx == 256) && !getenv("QGB_SCL_OFF");
  h->grid = cfg->members;
  if (h->large) {
    // one cluster of CTAs per member, persistent over members when the ensemble exceeds the machine.  Cluster size: small
    // grids are barrier-bound, so fewer CTAs per member and more members in flight win (measured 128^2 x 64: 0.286 ms with
    // 8, 0.228 ms with 4); from 256^2 on the per-member work fills 8 CTAs (256^2: 0.873 vs 0.918 ms, 512^2: 1.76 vs 2.78 ms).
    // With more members than 4-CTA clusters fit (2 CTAs per SM), 128^2 runs best with 2 CTAs per member (256 members:
    // 1.21 ms with 4, 0.86 ms with 2).
    h->cluster = h->ht.N <= 128 ? (4 * cfg->members <= 2 * h->nsm ? 4 : 2) : kMaxClusterSize;
    if (const char* e = getenv("QGB_CLUSTER")) { int v = atoi(e); if (v == 2 || v == 4 || v == 8) h->cluster = v; }
    // the lines of a 1-D pass are dealt evenly to the CTAs of the cluster: the size must divide nx (nx = 162, 324, ... = 2 * 3^k
    // only admit 2)
    while (h->cluster > 1 && h->ht.N % h->cluster != 0) h->cluster /= 2;
    if (h->cluster < 2) {
      fail(nullptr, QGB_EUNSUPPORTED, "nx=%d unsupported on the cluster path (needs an even nx)", cfg->nx);
      qgb_destroy(h);
      return QGB_EUNSUPPORTED;
    }
    h->nthreads = 512;
    // lines of a 1-D transform pass that a CTA stages in shared memory at a time: all it owns (N / cluster size) when that
    // leaves room for two CTAs per SM (<= 110 KB), else the largest power-of-two fraction that does
    const int per_cta = h->ht.N / h->cluster;
    const size_t line_bytes = (size_t)(h->ht.N + 1) * sizeof(cplx);
    int lines = per_cta;
    while (lines > 1 && lines * line_bytes > 110 * 1024) lines /= 2;
    if (const char* e = getenv("QGB_LARGE_LINES")) { int v = atoi(e); if (v >= 1 && v <= per_cta && v * line_bytes <= 200 * 1024) lines = v; }
    h->large_lines = lines;
    h->large_smem = lines * line_bytes + (size_t)h->ht.N * sizeof(short) + 16;
    int clusters = (2 * h->nsm) / h->cluster;
    if (clusters > cfg->members) clusters = cfg->members;
    if (clusters < 1) clusters = 1;
    h->grid = clusters * h->cluster;
    CR(dalloc(&h->scratch, (size_t)cfg->members * h->ht.N * h->ht.P));
    CR(dalloc(&h->red_scratch, (size_t)cfg->members * 4 * h->cluster * h->nthreads));
  }
  CR(spectral_configure(make_plan(h)));
  CR(upload(&h->d_tw, h->ht.tw));
  CR(upload(&h->d_pos, h->ht.pos));
  CR(upload(&h->d_kv, h->ht.kv));
  CR(upload(&h->d_lv, h->ht.lv));
  CR(upload(&h->d_a, h->ht.a));
  CR(upload(&h->d_filtr, h->ht.filtr));
  if (h->large) {   // cluster path: the global field is kept in natural order, the stages' digit reversal stays inside the tiles
    std::vector<short> ident(h->ht.N);
    for (int i = 0; i < h->ht.N; ++i) ident[i] = (short)i;
    CR(upload(&h->d_pos_id, ident));
  }
  fill_tables(h->ht, h->T, h->d_tw, h->large ? h->d_pos_id : h->d_pos, h->d_kv, h->d_lv, h->d_a, h->d_filtr);
  const size_t nr = nreal(h), nc = ncplx(h), B = cfg->members;
  CR(dalloc(&h->qh, nc)); CR(cudaMemset(h->qh, 0, nc * sizeof(cplx)));
  CR(dalloc(&h->q, nr)); CR(cudaMemset(h->q, 0, nr * sizeof(double)));
  for (int i = 0; i < 3; ++i) { CR(dalloc(&h->hist[i], nc)); CR(cudaMemset(h->hist[i], 0, nc * sizeof(cplx))); }
  CR(dalloc(&h->red, B * 4));
  CR(dalloc(&h->d_ke, B)); CR(dalloc(&h->d_cfl, B)); CR(dalloc(&h->d_flags, B));
  CR(dalloc(&h->d_draw, 1)); CR(cudaMemset(h->d_draw, 0, sizeof(uint32_t)));
#undef CR
  *out = h;
  return QGB_OK;
}

void qgb_destroy(qgb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  cudaFree(h->d_tw); cudaFree(h->d_pos); cudaFree(h->d_pos_id); cudaFree(h->d_kv); cudaFree(h->d_lv); cudaFree(h->d_a); cudaFree(h->d_filtr);
  cudaFree(h->qh); cudaFree(h->q); cudaFree(h->scratch); cudaFree(h->red_scratch);
  for (int i = 0; i < 3; ++i) cudaFree(h->hist[i]);
  cudaFree(h->ph); cudaFree(h->u); cudaFree(h->v); cudaFree(h->p); cudaFree(h->red);
  invalidate_graphs(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  cudaFree(h->d_draw);
  cudaFree(h->d_ke); cudaFree(h->d_cfl); cudaFree(h->d_flags); cudaFree(h->d_kespec); cudaFree(h->d_ensspec);
  cudaFree(h->stage_x); cudaFree(h->stage_y); cudaFree(h->f32_stage); cudaFree(h->bud); cudaFree(h->bud_tend); cudaFree(h->bud_scr); cudaFree(h->bud_sum); cudaFree(h->avg);
  cudaFree(h->xin); cudaFree(h->z64); cudaFree(h->xi_inj); cudaFree(h->ynet[0]); cudaFree(h->ynet[1]);
  cudaFree(h->yacc); cudaFree(h->dq_ext); cudaFree(h->dq); cudaFree(h->dq_dm); cudaFree(h->act[0]); cudaFree(h->act[1]);
  free_net(h->nets[0]); free_net(h->nets[1]);
  tc_free_workspace(h->tcw);
  h->prof.destroy();
  delete h;
}

int qgb_reset_time(qgb_handle* h) {
  if (!h) return QGB_EINVAL;
  h->tc = 0; h->t = 0.0; h->ablevel = 0;
  h->noise_init = false; h->const_counter = 0; h->dq_valid = false;
  return QGB_OK;
}

int qgb_get_time(qgb_handle* h, double* t, int64_t* tc) {
  if (!h) return QGB_EINVAL;
  if (t) *t = h->t;
  if (tc) *tc = h->tc;
  return QGB_OK;
}

int qgb_set_q(qgb_handle* h, const double* q, int on_device, void* stream) {
  if (!h || !q) return fail(h, QGB_EINVAL, "null argument");
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaMemcpyAsync(h->q, q, nreal(h) * sizeof(double),
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  StepIO io = base_io(h);
  set_cnn_io(h, io);
  int rc = launch_program(h, io, PROG_SET_Q, st);
  if (rc) return rc;
  h->x_valid = io.cnn_x != nullptr;
  if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(st));  // the host buffer may be reused by the caller
  return QGB_OK;
}

int qgb_invert(qgb_handle* h, void* stream) {
  if (!h) return QGB_EINVAL;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->ph) {
    CUDA_TRY(h, dalloc(&h->ph, ncplx(h)));
    CUDA_TRY(h, dalloc(&h->u, nreal(h)));
    CUDA_TRY(h, dalloc(&h->v, nreal(h)));
    CUDA_TRY(h, dalloc(&h->p, nreal(h)));
  }
  StepIO io = base_io(h);
  io.ph_out = h->ph; io.u_out = h->u; io.v_out = h->v; io.p_out = h->p;
  return launch_program(h, io, PROG_INVERT, S(stream));
}

namespace {
// out[i] (+)= sum over members of per_member[m][i]: eight member slices per point (a warp reads 32 consecutive points of one member),
// added in a fixed order -- deterministic for a given number of local members
constexpr int kReduceParts = 8;
__global__ void __launch_bounds__(32 * kReduceParts) reduce_members_kernel(const double* __restrict__ per_member, int members, long long n,
                                                                       double* __restrict__ out, int accumulate) {
  __shared__ double part_sum[kReduceParts][32];
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  double acc = 0.0;
  if (i < n)
    for (int m = threadIdx.y; m < members; m += kReduceParts) acc += per_member[(long long)m * n + i];
  part_sum[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    for (int p = 1; p < kReduceParts; ++p) acc += part_sum[p][threadIdx.x];
    out[i] = accumulate ? out[i] + acc : acc;
  }
}
__global__ void add_kernel(const double* __restrict__ a, long long n, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += a[i];
}

// budget terms of the current state summed over the local members -> ``out`` (kBudgetTerms, N, NK), device pointer
int budget_sums(qgb_handle* h, const double* dq, double* out, int accumulate, cudaStream_t st) {
  const long long NN = (long long)h->ht.N * h->ht.NK, B = h->cfg.members;
  if (!h->bud) {
    CUDA_TRY(h, dalloc(&h->bud, (size_t)(B * kBudgetTerms * NN)));
    CUDA_TRY(h, dalloc(&h->bud_scr, (size_t)(B * 3 * h->ht.N * h->ht.N)));
    CUDA_TRY(h, dalloc(&h->bud_tend, (size_t)(B * 2 * NN)));
  }
  StepIO io = base_io(h);
  io.bud_out = h->bud; io.bud_scr = h->bud_scr; io.bud_tend = h->bud_tend;
  io.dq = dq;
  // the filter-dissipation spectra see the update the NEXT call of _forward_timestep performs: its Adams-Bashforth level and the
  // tendency history as it stands (pyqg evaluates the diagnostics inside _step_forward, before _forward_timestep)
  io.d_p = h->hist[(h->tc + 2) % 3]; io.d_pp = h->hist[(h->tc + 1) % 3];
  ab_coefficients(h->ablevel, h->cfg.dt, io.dt1, io.dt2, io.dt3);
  io.bud_inv_dt = 1.0 / h->cfg.dt;
  io.bud_demean = (dq != nullptr && dq == h->dq) ? 1 : 0;
  int rc = launch_program(h, io, PROG_BUDGET, st);
  if (rc) return rc;
  const long long n = kBudgetTerms * NN;
  reduce_members_kernel<<<(unsigned)((n + 31) / 32), dim3(32, kReduceParts), 0, st>>>(h->bud, (int)B, n, out, accumulate);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(h, cudaGetLastError());
  return QGB_OK;
}

// pyqg Model._increment_diagnostics: called before the time step when t >= dt, t >= tavestart and tc % ceil(taveint/dt) == 0
int sample_averages(qgb_handle* h, const double* dq, cudaStream_t st) {
  const long long NN = (long long)h->ht.N * h->ht.NK, nspec = 4 * NN;
  if (!h->avg) {
    CUDA_TRY(h, dalloc(&h->avg, (size_t)((4 + kBudgetTerms) * NN)));
    CUDA_TRY(h, cudaMemsetAsync(h->avg, 0, (size_t)((4 + kBudgetTerms) * NN) * sizeof(double), st));
    h->avg_n = 0;
  }
  if (!h->d_kespec) { CUDA_TRY(h, dalloc(&h->d_kespec, (size_t)(2 * NN))); CUDA_TRY(h, dalloc(&h->d_ensspec, (size_t)(2 * NN))); }
  CUDA_TRY(h, launch_spectra(h->T, h->qh, h->cfg.members, h->d_kespec, h->d_ensspec, st));
  QGB_COUNT_LAUNCH();
  add_kernel<<<(unsigned)((2 * NN + 127) / 128), 128, 0, st>>>(h->d_kespec, 2 * NN, h->avg);
  add_kernel<<<(unsigned)((2 * NN + 127) / 128), 128, 0, st>>>(h->d_ensspec, 2 * NN, h->avg + 2 * NN);
  QGB_COUNT_LAUNCH(); QGB_COUNT_LAUNCH();
  int rc = budget_sums(h, dq, h->avg + nspec, 1, st);
  if (rc) return rc;
  h->avg_n += 1;
  return QGB_OK;
}
}  // namespace

int qgb_diag_budget(qgb_handle* h, double* out, int on_device, void* stream) {
  if (!h || !out) return fail(h, QGB_EINVAL, "null argument");
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)kBudgetTerms * h->ht.N * h->ht.NK;
  if (on_device) return budget_sums(h, h->last_dq, out, 0, st);
  if (!h->bud_sum) CUDA_TRY(h, dalloc(&h->bud_sum, n));
  int rc = budget_sums(h, h->last_dq, h->bud_sum, 0, st);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(out, h->bud_sum, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_diag_config(qgb_handle* h, double tavestart, double taveint) {
  if (!h) return QGB_EINVAL;
  if (!(taveint > 0.0)) return fail(h, QGB_EINVAL, "taveint must be positive");
  h->tavestart = tavestart; h->taveint = taveint; h->avg_on = true;
  return QGB_OK;
}

int qgb_diag_averages(qgb_handle* h, double* out, int64_t* nsamples, int reset, int on_device, void* stream) {
  if (!h) return QGB_EINVAL;
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)(4 + kBudgetTerms) * h->ht.N * h->ht.NK;
  if (nsamples) *nsamples = h->avg_n;
  if (out) {
    if (h->avg && h->avg_n > 0) CUDA_TRY(h, cudaMemcpyAsync(out, h->avg, n * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    else if (on_device) CUDA_TRY(h, cudaMemsetAsync(out, 0, n * sizeof(double), st));
    else std::memset(out, 0, n * sizeof(double));
    if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(st));
  }
  if (reset) {
    if (h->avg) CUDA_TRY(h, cudaMemsetAsync(h->avg, 0, n * sizeof(double), st));
    h->avg_n = 0;
  }
  return QGB_OK;
}

namespace {
// one time step enqueued on ``st`` (the body of pyqg Model._step_forward)
int step_once(qgb_handle* h, cudaStream_t st) {
  StepIO io = base_io(h);
  set_cnn_io(h, io);
  int prog = PROG_STEP;
  if (h->ext_set) {
    io.dq = h->dq_ext;
    prog = PROG_STEP_DQ_RAW;
    h->ext_set = false;
  } else if (h->kind != QGB_CLOSURE_NONE) {
    int rc = closure_update(h, st);
    if (rc) return rc;
    io.dq = h->dq;
    prog = PROG_STEP_DQ;
  }
  if (h->avg_on && h->t >= h->cfg.dt && h->t >= h->tavestart &&
      h->tc % (long long)std::ceil(h->taveint / h->cfg.dt) == 0) {
    const int pd = h->prof.start(PROF_DIAG, st);
    int rc = sample_averages(h, io.dq, st);
    if (rc) return rc;
    h->prof.stop(pd, st, h->cfg.members);
  }
  const int cur = (int)(h->tc % 3), prev = (int)((h->tc + 2) % 3), pprev = (int)((h->tc + 1) % 3);
  io.d_cur = h->hist[cur]; io.d_p = h->hist[prev]; io.d_pp = h->hist[pprev];
  ab_coefficients(h->ablevel, h->cfg.dt, io.dt1, io.dt2, io.dt3);
  const int pi = h->prof.start(PROF_SPECTRAL, st);
  int rc = launch_program(h, io, prog, st);
  if (rc) return rc;
  h->prof.stop(pi, st, h->cfg.members);
  h->last_dq = io.dq;
  if (h->ablevel < 2) h->ablevel++;
  h->tc += 1;
  h->t += h->cfg.dt;
  h->x_valid = io.cnn_x != nullptr;
  return QGB_OK;
}

// A step is the SAME sequence of launches with the same arguments whenever: AB3 has started (ablevel 2), the sampler draws every
// step (AR1, or constant with nsteps = 1) or there is no closure, nothing was injected for this step, nobody is timing kernels,
// the precision calibration is done and no diagnostics sample is due.  Only the position of the tendency-history ring (tc % 3)
// changes, so three captured graphs cover the steady state; the Philox draw counter lives in device memory.
bool graph_eligible(const qgb_handle* h) {
  if (h->graph_failed || h->ablevel < 2 || h->ext_set || h->xi_set || h->prof.layer != -1) return false;
  if (h->kind != QGB_CLOSURE_NONE) {
    if (!h->noise_init || !h->dq_valid) return false;
    if (h->sampler == QGB_SAMPLER_DETERMINISTIC) return false;
    if (h->sampler == QGB_SAMPLER_CONSTANT && h->sampler_nsteps != 1) return false;
    if (h->precision == QGB_PREC_AUTO && !h->calibrated) return false;
  }
  return true;
}
bool diag_sample_due(const qgb_handle* h) {
  return h->avg_on && h->t >= h->cfg.dt && h->t >= h->tavestart && h->tc % (long long)std::ceil(h->taveint / h->cfg.dt) == 0;
}
// everything a captured step depends on besides tc % 3
unsigned long long graph_state_key(const qgb_handle* h) {
  unsigned long long k = 1469598103934665603ull;
  auto mix = [&](unsigned long long v) { k = (k ^ v) * 1099511628211ull; };
  mix((unsigned long long)h->kind); mix((unsigned long long)h->precision); mix((unsigned long long)h->auto_precision);
  mix((unsigned long long)h->sampler); mix((unsigned long long)h->sampler_nsteps);
  mix((unsigned long long)(h->sampler == QGB_SAMPLER_CONSTANT ? h->const_counter : 0));
  unsigned long long w; std::memcpy(&w, &h->weight, 8); mix(w);
  unsigned int f; std::memcpy(&f, &h->x_std[0], 4); mix(f); std::memcpy(&f, &h->x_std[1], 4); mix(f);
  std::memcpy(&f, &h->y_std[0], 4); mix(f); std::memcpy(&f, &h->y_std[1], 4); mix(f);
  mix((unsigned long long)(uintptr_t)h->xin); mix((unsigned long long)(uintptr_t)h->dq); mix((unsigned long long)h->seed);
  mix((unsigned long long)(uintptr_t)h->tcw.buf[2]); mix((unsigned long long)(uintptr_t)h->nets[0].tc.layers.data());
  return k;
}
}  // namespace

int qgb_step(qgb_handle* h, int nsteps, void* stream) {
  if (!h || nsteps < 0) return fail(h, QGB_EINVAL, "bad argument");
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  static const bool graphs_off = getenv("QGB_NO_GRAPH") != nullptr;
  for (int s = 0; s < nsteps; ++s) {
    if (graphs_off || !graph_eligible(h) || diag_sample_due(h)) {
      int rc = step_once(h, st);
      if (rc) return rc;
      continue;
    }
    const unsigned long long key = graph_state_key(h);
    if (key != h->graph_key) { invalidate_graphs(h); h->graph_key = key; }
    const int slot = (int)(h->tc % 3);
    if (!h->step_graph[slot]) {
      // capture this step (the launches are recorded, not executed), instantiate, then replay it below
      const long long tc0 = h->tc; const double t0 = h->t; const long long cc0 = h->const_counter; const bool xv0 = h->x_valid;
      const long long l0 = g_launches.load();
      cudaGraph_t graph = nullptr;
      cudaError_t ce = cudaSuccess;
      if (!h->cap_stream) ce = cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking);
      if (ce == cudaSuccess) ce = cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal);
      int rc = QGB_OK;
      if (ce == cudaSuccess) {
        rc = step_once(h, h->cap_stream);
        ce = cudaStreamEndCapture(h->cap_stream, &graph);
      }
      h->tc = tc0; h->t = t0; h->const_counter = cc0; h->x_valid = xv0;   // host-side bookkeeping is redone by the replay
      const long long captured = g_launches.load() - l0;                   // kernels recorded in the graph (not run yet)
      g_launches.store(l0);
      if (ce == cudaSuccess && rc == QGB_OK && graph) ce = cudaGraphInstantiate(&h->step_graph[slot], graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (ce != cudaSuccess || rc != QGB_OK || !h->step_graph[slot]) {
        // capture is an optimisation: fall back to plain launches for the rest of this handle's life
        cudaGetLastError();
        h->step_graph[slot] = nullptr;
        h->graph_failed = true;
        h->err.clear();
        rc = step_once(h, st);
        if (rc) return rc;
        continue;
      }
      h->graph_launches[slot] = captured;
      h->graph_noise_regen[slot] = h->noise_regen;
    }
    CUDA_TRY(h, cudaGraphLaunch(h->step_graph[slot], st));
    g_launches.fetch_add(h->graph_launches[slot], std::memory_order_relaxed);
    h->graph_replays += 1;
    if (h->kind != QGB_CLOSURE_NONE) h->noise_regen = h->graph_noise_regen[slot];
    // host-side bookkeeping of step_once
    if (h->kind != QGB_CLOSURE_NONE) {
      if (h->sampler == QGB_SAMPLER_CONSTANT) h->const_counter = 1;
      h->last_dq = h->dq;
    } else {
      h->last_dq = nullptr;
    }
    h->tc += 1;
    h->t += h->cfg.dt;
    h->x_valid = h->kind != QGB_CLOSURE_NONE && h->xin != nullptr;
  }
  return QGB_OK;
}

int64_t qgb_graph_replays(const qgb_handle* h) { return h ? (int64_t)h->graph_replays : 0; }

int qgb_step_host(qgb_handle* h, const double* q_in, double* q_out, int nsteps, void* stream) {
  if (!h) return QGB_EINVAL;
  int rc = QGB_OK;
  if (q_in) rc = qgb_set_q(h, q_in, 0, stream);
  if (rc) return rc;
  rc = qgb_step(h, nsteps, stream);
  if (rc) return rc;
  if (q_out) rc = qgb_get(h, QGB_F_Q, q_out, 0, stream);
  return rc;
}

int qgb_step_host_async(qgb_handle* h, const double* q_in, double* q_out, int nsteps, void* stream) {
  if (!h) return QGB_EINVAL;
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (q_in) {
    CUDA_TRY(h, cudaMemcpyAsync(h->q, q_in, nreal(h) * sizeof(double), cudaMemcpyHostToDevice, st));
    StepIO io = base_io(h);
    set_cnn_io(h, io);
    int rc = launch_program(h, io, PROG_SET_Q, st);
    if (rc) return rc;
    h->x_valid = io.cnn_x != nullptr;
  }
  int rc = qgb_step(h, nsteps, stream);
  if (rc) return rc;
  if (q_out) CUDA_TRY(h, cudaMemcpyAsync(q_out, h->q, nreal(h) * sizeof(double), cudaMemcpyDeviceToHost, st));
  return QGB_OK;
}

int qgb_get(qgb_handle* h, int field, void* out, int on_device, void* stream) {
  if (!h || !out) return fail(h, QGB_EINVAL, "null argument");
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const void* src = nullptr;
  size_t bytes = 0;
  switch (field) {
    case QGB_F_Q: src = h->q; bytes = nreal(h) * sizeof(double); break;
    case QGB_F_QH: src = h->qh; bytes = ncplx(h) * sizeof(cplx); break;
    case QGB_F_PH: src = h->ph; bytes = ncplx(h) * sizeof(cplx); break;
    case QGB_F_U: src = h->u; bytes = nreal(h) * sizeof(double); break;
    case QGB_F_V: src = h->v; bytes = nreal(h) * sizeof(double); break;
    case QGB_F_P: src = h->p; bytes = nreal(h) * sizeof(double); break;
    case QGB_F_DQHDT: src = h->hist[(h->tc + 2) % 3]; bytes = ncplx(h) * sizeof(cplx); break;
    case QGB_F_FORCING: {
      if (!h->dq_valid) return fail(h, QGB_ESTATE, "no closure forcing has been computed yet");
      demean_kernel<<<h->cfg.members * 2, 256, 0, st>>>(h->dq, h->dq_dm, h->ht.N * h->ht.N);
      QGB_COUNT_LAUNCH();
      CUDA_TRY(h, cudaGetLastError());
      src = h->dq_dm; bytes = nreal(h) * sizeof(double);
      break;
    }
    case QGB_F_NOISE: {
      if (!h->noise_init) return fail(h, QGB_ESTATE, "latent noise not initialised");
      const size_t npix = (size_t)h->ht.N * h->ht.N;
      if (h->kind == QGB_CLOSURE_GZ) { src = h->z64; bytes = nreal(h) * sizeof(double); break; }
      if (h->kind == QGB_CLOSURE_GAN || h->kind == QGB_CLOSURE_VAE) {
        if (h->noise_regen) {        // the latest draw only ever existed inside layer 1: regenerate it from the same counters
          const long long np2 = (long long)h->ht.N * h->ht.N;
          const bool inj = h->xi_set;
          h->xi_set = false;
          int rc = launch_latent<float>(h, h->xin + 2 * np2, 4 * np2, 0.0, 1.0, 1, st, -1, false);
          h->xi_set = inj;
          if (rc) return rc;
          h->noise_regen = false;
        }
        // strided gather of channels 2,3 of the closure input
        CUDA_TRY(h, cudaMemcpy2DAsync(out, 2 * npix * sizeof(float), h->xin + 2 * npix, 4 * npix * sizeof(float),
                                      2 * npix * sizeof(float), h->cfg.members,
                                      on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(st));
        return QGB_OK;
      }
      return fail(h, QGB_ESTATE, "closure has no latent noise");
    }
    default: return fail(h, QGB_EINVAL, "unknown field %d", field);
  }
  if (!src) return fail(h, QGB_ESTATE, "field %d not available (call qgb_invert first)", field);
  CUDA_TRY(h, cudaMemcpyAsync(out, src, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(st));
  return QGB_OK;
}

namespace {
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = (float)in[i];
}
}  // namespace

int qgb_get_f32(qgb_handle* h, int field, float* out, int on_device, int async, void* stream) {
  if (!h || !out) return fail(h, QGB_EINVAL, "null argument");
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const double* src = nullptr;
  switch (field) {
    case QGB_F_Q: src = h->q; break;
    case QGB_F_U: src = h->u; break;
    case QGB_F_V: src = h->v; break;
    case QGB_F_P: src = h->p; break;
    default: return fail(h, QGB_EINVAL, "qgb_get_f32: field %d is not a real (B,2,N,N) field", field);
  }
  if (!src) return fail(h, QGB_ESTATE, "field %d not available (call qgb_invert first)", field);
  const long long n = (long long)nreal(h);
  float* dst = out;
  if (!on_device) {
    if (!h->f32_stage) CUDA_TRY(h, dalloc(&h->f32_stage, (size_t)n));
    dst = h->f32_stage;
  }
  f64_to_f32_kernel<<<h->nsm * 8, 256, 0, st>>>(src, dst, n);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(h, cudaGetLastError());
  if (!on_device) {
    CUDA_TRY(h, cudaMemcpyAsync(out, dst, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (!async) CUDA_TRY(h, cudaStreamSynchronize(st));
  }
  return QGB_OK;
}

// ---- closure ------------------------------------------------------------------------------------------------
int qgb_cnn_load(qgb_handle* h, int kind, int net, int nlayers, const qgb_cnn_layer* layers) {
  if (!h || !layers || nlayers < 1 || net < 0 || net > 1) return fail(h, QGB_EINVAL, "bad argument");
  if (kind < QGB_CLOSURE_GAN || kind > QGB_CLOSURE_RAW) return fail(h, QGB_EINVAL, "unknown closure kind %d", kind);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  for (int i = 1; i < nlayers; ++i)
    if (layers[i].cin != layers[i - 1].cout) return fail(h, QGB_EINVAL, "layer %d: cin does not match previous cout", i);
  // validate everything before anything is freed or replaced: a failed load leaves the handle as it was
  for (int i = 0; i < nlayers; ++i) {
    const qgb_cnn_layer& L = layers[i];
    if (L.cin < 1 || L.cout < 1) return fail(h, QGB_EINVAL, "layer %d: bad channel counts", i);
    if (L.ksize != 1 && L.ksize != 3 && L.ksize != 5) return fail(h, QGB_EUNSUPPORTED, "layer %d: kernel size %d not supported (1, 3, 5)", i, L.ksize);
    if (!L.weight || !L.bias) return fail(h, QGB_EINVAL, "layer %d: null weight/bias", i);
    if (L.relu_bn && (!L.bn_scale || !L.bn_shift)) return fail(h, QGB_EINVAL, "layer %d: null batch-norm affine", i);
  }
  if (kind != QGB_CLOSURE_RAW) {
    const int cin0 = (kind == QGB_CLOSURE_GAN || kind == QGB_CLOSURE_VAE) ? 4 : 2;
    if (layers[0].cin != cin0) return fail(h, QGB_EINVAL, "first layer must have %d input channels, got %d", cin0, layers[0].cin);
    if (layers[nlayers - 1].cout != 2) return fail(h, QGB_EINVAL, "last layer must have 2 output channels");
  }
  DevNet fresh;      // built aside, swapped in on success
  for (int i = 0; i < nlayers; ++i) {
    const qgb_cnn_layer& L = layers[i];
    DevNetLayer D;
    D.cin = L.cin; D.cout = L.cout; D.ks = L.ksize; D.relu_bn = L.relu_bn;
    const int co_t = L.cout <= 4 ? 2 : 32;
    D.cout_pad = (L.cout + co_t - 1) / co_t * co_t;
    const int kk = L.ksize * L.ksize;
    std::vector<float> wp((size_t)L.cin * kk * D.cout_pad, 0.f);
    for (int co = 0; co < L.cout; ++co)
      for (int ci = 0; ci < L.cin; ++ci)
        for (int t = 0; t < kk; ++t) wp[((size_t)ci * kk + t) * D.cout_pad + co] = L.weight[((size_t)co * L.cin + ci) * kk + t];
    std::vector<float> bias(L.bias, L.bias + L.cout), s(L.cout, 1.f), tt(L.cout, 0.f);
    if (L.relu_bn) { s.assign(L.bn_scale, L.bn_scale + L.cout); tt.assign(L.bn_shift, L.bn_shift + L.cout); }
    cudaError_t ce = upload(&D.wp, wp);
    if (ce == cudaSuccess) ce = upload(&D.bias, bias);
    if (ce == cudaSuccess) ce = upload(&D.bn_s, s);
    if (ce == cudaSuccess) ce = upload(&D.bn_t, tt);
    fresh.layers.push_back(D);        // (pushed first so that free_net releases a partial upload too)
    if (ce != cudaSuccess) {
      free_net(fresh);
      return fail(h, QGB_ECUDA, "qgb_cnn_load: upload of layer %d failed: %s", i, cudaGetErrorString(ce));
    }
  }
  std::string e;
  if (tc_pack_net(fresh.tc, nlayers, layers, &e) == QGB_ECUDA) {   // (other failures just leave fresh.tc.ready == false: fp32 only)
    free_net(fresh);
    return fail(h, QGB_ECUDA, "qgb_cnn_load: packing the tensor-core weights failed");
  }
  if (kind != QGB_CLOSURE_RAW) {
    if (h->kind != kind) { free_net(h->nets[0]); free_net(h->nets[1]); h->dq_valid = false; h->noise_init = false; }
    h->kind = kind;
  }
  invalidate_graphs(h);
  free_net(h->nets[net]);
  h->nets[net] = std::move(fresh);
  h->calibrated = false;
  return kind == QGB_CLOSURE_RAW ? QGB_OK : ensure_closure_buffers(h);
}

int qgb_closure_config(qgb_handle* h, const float x_std[2], const float y_std[2], double weight, int precision) {
  if (!h || !x_std || !y_std) return fail(h, QGB_EINVAL, "null argument");
  if (precision < QGB_PREC_FP32 || precision > QGB_PREC_AUTO) return fail(h, QGB_EINVAL, "unknown precision %d", precision);
  h->x_std[0] = x_std[0]; h->x_std[1] = x_std[1];
  h->y_std[0] = y_std[0]; h->y_std[1] = y_std[1];
  h->weight = weight;
  invalidate_graphs(h);
  if (precision != h->precision) h->calibrated = false;
  h->precision = precision;
  h->x_valid = false;
  return QGB_OK;
}

int qgb_closure_precision(qgb_handle* h, int* precision, double err[4]) {
  if (!h) return QGB_EINVAL;
  if (precision) *precision = h->precision == QGB_PREC_AUTO ? (h->calibrated ? h->auto_precision : QGB_PREC_AUTO) : h->precision;
  if (err) for (int i = 0; i < 4; ++i) err[i] = h->calib_err[i];
  return QGB_OK;
}

int qgb_set_sampler(qgb_handle* h, int kind, int nsteps, int n_mean) {
  if (!h) return QGB_EINVAL;
  if (kind < QGB_SAMPLER_AR1 || kind > QGB_SAMPLER_DETERMINISTIC) return fail(h, QGB_EINVAL, "Unknown sampling type");
  if (kind == QGB_SAMPLER_CONSTANT && nsteps < 1) return fail(h, QGB_EINVAL, "constant sampler needs nsteps >= 1");
  invalidate_graphs(h);
  h->sampler = kind; h->sampler_nsteps = nsteps; h->n_mean = n_mean > 0 ? n_mean : 100;
  h->noise_init = false; h->const_counter = 0;
  return QGB_OK;
}

int qgb_seed(qgb_handle* h, uint64_t seed) {
  if (!h) return QGB_EINVAL;
  h->seed = seed;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  CUDA_TRY(h, cudaMemset(h->d_draw, 0, sizeof(uint32_t)));
  invalidate_graphs(h);
  return QGB_OK;
}

int qgb_set_latent(qgb_handle* h, const void* xi, int dtype, int on_device, void* stream) {
  if (!h) return QGB_EINVAL;
  if (!xi) { h->xi_set = false; return QGB_OK; }
  if (dtype != 0 && dtype != 1) return fail(h, QGB_EINVAL, "dtype must be 0 (float) or 1 (double)");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const size_t bytes = nreal(h) * (dtype ? 8 : 4);
  if (!h->xi_inj) CUDA_TRY(h, cudaMalloc(&h->xi_inj, nreal(h) * 8));
  CUDA_TRY(h, cudaMemcpyAsync(h->xi_inj, xi, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, S(stream)));
  if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(S(stream)));
  h->xi_dtype = dtype; h->xi_set = true;
  return QGB_OK;
}

int qgb_set_forcing(qgb_handle* h, const double* dq, int on_device, void* stream) {
  if (!h) return QGB_EINVAL;
  if (!dq) { h->ext_set = false; return QGB_OK; }
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  if (!h->dq_ext) CUDA_TRY(h, dalloc(&h->dq_ext, nreal(h)));
  CUDA_TRY(h, cudaMemcpyAsync(h->dq_ext, dq, nreal(h) * sizeof(double),
                              on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, S(stream)));
  if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(S(stream)));
  h->ext_set = true;
  return QGB_OK;
}

int qgb_profile_begin(qgb_handle* h, int net, int layer) {
  if (!h || net < 0 || net > 1 || layer < 0 || layer > 7) return fail(h, QGB_EINVAL, "bad argument");
  h->prof.reset();
  h->prof.net = net; h->prof.layer = layer;
  return QGB_OK;
}

int qgb_profile_end(qgb_handle* h, double* total_ms, int64_t* launches, int64_t* images) {
  if (!h) return QGB_EINVAL;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  double tot = 0.0;
  long long units = 0;
  for (auto& r : h->prof.recs) {
    CUDA_TRY(h, cudaEventSynchronize(r.b));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, r.a, r.b));
    tot += ms;
    units += r.units;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)h->prof.recs.size();
  if (images) *images = (int64_t)units;
  h->prof.reset();
  return QGB_OK;
}

int qgb_profile_all_begin(qgb_handle* h) {
  if (!h) return QGB_EINVAL;
  h->prof.reset();
  h->prof.layer = -2;
  return QGB_OK;
}

int qgb_profile_all_end(qgb_handle* h, double ms[QGB_PROF_SLOTS], int64_t launches[QGB_PROF_SLOTS], int64_t units[QGB_PROF_SLOTS]) {
  if (!h) return QGB_EINVAL;
  static_assert(QGB_PROF_SLOTS == PROF_SLOTS, "header / prof.hpp slot counts differ");
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  for (int i = 0; i < QGB_PROF_SLOTS; ++i) { if (ms) ms[i] = 0.0; if (launches) launches[i] = 0; if (units) units[i] = 0; }
  for (auto& r : h->prof.recs) {
    CUDA_TRY(h, cudaEventSynchronize(r.b));
    float t = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&t, r.a, r.b));
    if (ms) ms[r.slot] += t;
    if (launches) launches[r.slot] += 1;
    if (units) units[r.slot] += r.units;
  }
  h->prof.reset();
  return QGB_OK;
}

int qgb_closure_eval(qgb_handle* h, void* stream) {
  if (!h) return QGB_EINVAL;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  return closure_update(h, S(stream));
}

int qgb_cnn_forward(qgb_handle* h, int net, const float* x, float* y, int batch, int ny, int nx, int softplus,
                    int precision, int on_device, void* stream) {
  if (!h || !x || !y || batch < 1 || ny < 1 || nx < 1 || net < 0 || net > 1) return fail(h, QGB_EINVAL, "bad argument");
  if (!h->nets[net].loaded()) return fail(h, QGB_ESTATE, "network %d not loaded", net);
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const int cin = h->nets[net].layers.front().cin, cout = h->nets[net].layers.back().cout;
  if (precision == QGB_PREC_AUTO)     // the calibrated choice of the coupled closure if there is one, else the safe tensor-core plan
    precision = h->calibrated ? h->auto_precision : ((h->nets[net].tc.ready && ny % 16 == 0 && nx % 16 == 0) ? QGB_PREC_TC : QGB_PREC_FP32);
  const long long x_bs = (long long)cin * ny * nx, y_bs = (long long)cout * ny * nx;
  if (on_device) return net_forward(h, net, x, x_bs, y, y_bs, batch, ny, nx, softplus, 0, precision, st);
  // host caller: staging buffers live in the handle and only grow (no allocation per call on the plugin path)
  const size_t need_x = (size_t)batch * x_bs, need_y = (size_t)batch * y_bs;
  if (h->stage_x_floats < need_x) {
    cudaFree(h->stage_x); h->stage_x = nullptr; h->stage_x_floats = 0;
    CUDA_TRY(h, dalloc(&h->stage_x, need_x));
    h->stage_x_floats = need_x;
  }
  if (h->stage_y_floats < need_y) {
    cudaFree(h->stage_y); h->stage_y = nullptr; h->stage_y_floats = 0;
    CUDA_TRY(h, dalloc(&h->stage_y, need_y));
    h->stage_y_floats = need_y;
  }
  float *dx = h->stage_x, *dy = h->stage_y;
  int rc = QGB_OK;
  cudaError_t e = cudaMemcpyAsync(dx, x, need_x * 4, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) rc = net_forward(h, net, dx, x_bs, dy, y_bs, batch, ny, nx, softplus, 0, precision, st);
  if (e == cudaSuccess && rc == QGB_OK) e = cudaMemcpyAsync(y, dy, need_y * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return fail(h, QGB_ECUDA, "cnn_forward: %s", cudaGetErrorString(e));
  return rc;
}

// ---- diagnostics ----------------------------------------------------------------------------------------------
int qgb_diag(qgb_handle* h, double* ke, double* cfl, int32_t* flags, int on_device, void* stream) {
  if (!h) return QGB_EINVAL;
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  StepIO io = base_io(h);
  io.red_out = h->red;
  int rc = launch_program(h, io, PROG_DIAG, st);
  if (rc) return rc;
  const int B = h->cfg.members;
  CUDA_TRY(h, launch_diag_finish(h->red, B, h->cfg.dt / h->ht.dx, h->d_ke, h->d_cfl, h->d_flags, st));
  QGB_COUNT_LAUNCH();
  CUDA_TRY(h, cudaGetLastError());
  const cudaMemcpyKind kd = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (ke) CUDA_TRY(h, cudaMemcpyAsync(ke, h->d_ke, B * sizeof(double), kd, st));
  if (cfl) CUDA_TRY(h, cudaMemcpyAsync(cfl, h->d_cfl, B * sizeof(double), kd, st));
  if (flags) CUDA_TRY(h, cudaMemcpyAsync(flags, h->d_flags, B * sizeof(int), kd, st));
  if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_diag_spectra(qgb_handle* h, double* kespec, double* ensspec, int on_device, void* stream) {
  if (!h) return QGB_EINVAL;
  cudaStream_t st = S(stream);
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  const int n = 2 * h->ht.N * h->ht.NK;
  if (!h->d_kespec) { CUDA_TRY(h, dalloc(&h->d_kespec, (size_t)n)); CUDA_TRY(h, dalloc(&h->d_ensspec, (size_t)n)); }
  CUDA_TRY(h, launch_spectra(h->T, h->qh, h->cfg.members, h->d_kespec, h->d_ensspec, st));
  QGB_COUNT_LAUNCH();
  CUDA_TRY(h, cudaGetLastError());
  const cudaMemcpyKind kd = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (kespec) CUDA_TRY(h, cudaMemcpyAsync(kespec, h->d_kespec, n * sizeof(double), kd, st));
  if (ensspec) CUDA_TRY(h, cudaMemcpyAsync(ensspec, h->d_ensspec, n * sizeof(double), kd, st));
  if (!on_device) CUDA_TRY(h, cudaStreamSynchronize(st));
  return QGB_OK;
}

// ---- coarse-graining (operators.cuh + the phase programs) --------------------------------------------------------
namespace {
// Work handles of the stateless coarse-graining entry points.  Creating a handle costs milliseconds (tables built on the
// host and uploaded, ~15 allocations, function attributes) against microseconds of kernel time per snapshot, and forcing
// datasets call these entries per snapshot, operator and resolution, so handles are kept in a small process-wide cache
// keyed by the full configuration and handed out exclusively (busy flag) -- every entry point synchronises its stream
// before it returns, so a released handle has no work in flight.
struct HandleCache {
  struct Entry { qgb_handle* h; bool busy; unsigned long long stamp; };
  std::mutex mu;
  std::vector<Entry> entries;
  unsigned long long clock = 0;
  static bool same(const qgb_config& a, const qgb_config& b) {
    return a.nx == b.nx && a.members == b.members && a.member_offset == b.member_offset && a.device == b.device && a.L == b.L &&
           a.dt == b.dt && a.rek == b.rek && a.filterfac == b.filterfac && a.beta == b.beta && a.rd == b.rd && a.delta == b.delta &&
           a.H1 == b.H1 && a.U1 == b.U1 && a.U2 == b.U2;
  }
  int acquire(const qgb_config& c, qgb_handle** out) {
    if (cudaSetDevice(c.device) != cudaSuccess) return fail(nullptr, QGB_ECUDA, "cudaSetDevice(%d) failed", c.device);
    {
      std::lock_guard<std::mutex> lk(mu);
      for (auto& e : entries)
        if (!e.busy && same(e.h->cfg, c)) { e.busy = true; e.stamp = ++clock; *out = e.h; return QGB_OK; }
    }
    qgb_handle* h = nullptr;
    int rc = qgb_create(&c, &h);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(mu);
    if (entries.size() >= 32) {      // evict the least recently used idle handle
      int victim = -1;
      for (int i = 0; i < (int)entries.size(); ++i)
        if (!entries[i].busy && (victim < 0 || entries[i].stamp < entries[victim].stamp)) victim = i;
      if (victim >= 0) { qgb_destroy(entries[victim].h); entries.erase(entries.begin() + victim); }
    }
    entries.push_back({h, true, ++clock});
    *out = h;
    return QGB_OK;
  }
  void release(qgb_handle* h) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& e : entries)
      if (e.h == h) { e.busy = false; return; }
  }
};
HandleCache& handle_cache() { static HandleCache c; return c; }
struct HandleGuard {
  qgb_handle* h = nullptr;
  int create(const qgb_config& c) { return handle_cache().acquire(c, &h); }
  ~HandleGuard() { if (h) handle_cache().release(h); }
};
int grid_for(long long n) { long long b = (n + 255) / 256; return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }
Tables no_background(const Tables& T) {   // advect(var, u, v) uses anomaly velocities and no beta / drag terms
  Tables t = T;
  t.Ubg[0] = t.Ubg[1] = 0.0; t.Qy[0] = t.Qy[1] = 0.0; t.rek = 0.0;
  return t;
}
}  // namespace

namespace {
// advect(var, u, v, '3/2-rule') (tools/operators.py:258-266) for B members on the n-grid, all fields device-resident:
// adv_h (B,2,n,n/2+1) = ik * F(I_n(I_N(q) I_N(u))) + il * F(I_n(I_N(q) I_N(v))),  N = 3n/2, I = fft_interpolate.
int advect_dealiased(const qgb_config& base, int n, int B, const double* q, const double* u, const double* v, cplx* adv_h,
                     cudaStream_t st) {
  const int N = (3 * n) / 2;
  qgb_config c3n = base, c3N = base, c2N = base;
  c3n.nx = n; c3n.members = 3 * B;
  c3N.nx = N; c3N.members = 3 * B;
  c2N.nx = N; c2N.members = 2 * B;
  HandleGuard g3n, g3N, g2N;
  int rc;
  if ((rc = g3n.create(c3n))) return rc;
  if ((rc = g3N.create(c3N))) return rc;
  if ((rc = g2N.create(c2N))) return rc;
  qgb_handle *h3n = g3n.h, *h3N = g3N.h, *h2N = g2N.h;
  const size_t fn = (size_t)B * 2 * n * n, fN = (size_t)B * 2 * N * N;
  CUDA_TRY(nullptr, cudaMemcpyAsync(h3n->q, q, fn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(nullptr, cudaMemcpyAsync(h3n->q + fn, u, fn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(nullptr, cudaMemcpyAsync(h3n->q + 2 * fn, v, fn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  StepIO io = base_io(h3n);
  if ((rc = launch_program(h3n, io, PROG_SET_Q, st))) return fail(nullptr, rc, "%s", h3n->err.c_str());
  const double up = ((double)N / n) * ((double)N / n);
  resample_kernel<<<grid_for((long long)6 * B * N * (N / 2 + 1)), 256, 0, st>>>(h3n->qh, h3N->qh, 6 * B, n, N, up);
  StepIO ioN = base_io(h3N);
  if ((rc = launch_program(h3N, ioN, PROG_C2R, st))) return fail(nullptr, rc, "%s", h3N->err.c_str());
  rmul_kernel<<<grid_for((long long)fN), 256, 0, st>>>(h3N->q, h3N->q + fN, h2N->q, (long long)fN);            // q u
  rmul_kernel<<<grid_for((long long)fN), 256, 0, st>>>(h3N->q, h3N->q + 2 * fN, h2N->q + fN, (long long)fN);   // q v
  StepIO io2 = base_io(h2N);
  if ((rc = launch_program(h2N, io2, PROG_SET_Q, st))) return fail(nullptr, rc, "%s", h2N->err.c_str());
  // back to the n grid (reuse h3n->qh as scratch: first 2B members), then the spectral divergence
  resample_kernel<<<grid_for((long long)4 * B * n * (n / 2 + 1)), 256, 0, st>>>(h2N->qh, h3n->qh, 4 * B, N, n, 1.0 / up);
  const size_t cn = (size_t)B * 2 * n * (n / 2 + 1);
  spectral_div_kernel<<<grid_for((long long)cn), 256, 0, st>>>(h3n->qh, h3n->qh + cn, adv_h, 2 * B, n, base.L);
  g_launches.fetch_add(5, std::memory_order_relaxed);
  CUDA_TRY(nullptr, cudaGetLastError());
  CUDA_TRY(nullptr, cudaStreamSynchronize(st));     // the temporary handles are released on return
  return QGB_OK;
}

// advect(var, u, v, '2/3-rule') (tools/operators.py:253-257): q, u, v low-passed with the sharp filter of
// pyqg.QGModel(nx, filterfac=1e+20), products on the same grid, spectral divergence, low-passed again.
// adv_h (B,2,n,n/2+1) = filtr (ik F(q~ u~) + il F(q~ v~)).
int advect_23(const qgb_config& base, int n, int B, const double* q, const double* u, const double* v, cplx* adv_h, cudaStream_t st) {
  qgb_config c3 = base, c2 = base;
  c3.nx = n; c3.members = 3 * B;
  c2.nx = n; c2.members = 2 * B;
  HandleGuard g3, g2;
  int rc;
  if ((rc = g3.create(c3))) return rc;
  if ((rc = g2.create(c2))) return rc;
  qgb_handle *h3 = g3.h, *h2 = g2.h;
  const size_t fn = (size_t)B * 2 * n * n;
  const double sharp = 1e20;
  CUDA_TRY(nullptr, cudaMemcpyAsync(h3->q, q, fn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(nullptr, cudaMemcpyAsync(h3->q + fn, u, fn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(nullptr, cudaMemcpyAsync(h3->q + 2 * fn, v, fn * sizeof(double), cudaMemcpyDeviceToDevice, st));
  StepIO io = base_io(h3);
  if ((rc = launch_program(h3, io, PROG_SET_Q, st))) return fail(nullptr, rc, "%s", h3->err.c_str());
  spectral_filter_kernel<<<grid_for((long long)6 * B * n * (n / 2 + 1)), 256, 0, st>>>(h3->qh, 6 * B, n, base.L, sharp);
  if ((rc = launch_program(h3, io, PROG_C2R, st))) return fail(nullptr, rc, "%s", h3->err.c_str());
  rmul_kernel<<<grid_for((long long)fn), 256, 0, st>>>(h3->q, h3->q + fn, h2->q, (long long)fn);            // q~ u~
  rmul_kernel<<<grid_for((long long)fn), 256, 0, st>>>(h3->q, h3->q + 2 * fn, h2->q + fn, (long long)fn);   // q~ v~
  StepIO io2 = base_io(h2);
  if ((rc = launch_program(h2, io2, PROG_SET_Q, st))) return fail(nullptr, rc, "%s", h2->err.c_str());
  const size_t cn = (size_t)B * 2 * n * (n / 2 + 1);
  spectral_div_kernel<<<grid_for((long long)cn), 256, 0, st>>>(h2->qh, h2->qh + cn, adv_h, 2 * B, n, base.L);
  spectral_filter_kernel<<<grid_for((long long)cn), 256, 0, st>>>(adv_h, 2 * B, n, base.L, sharp);
  g_launches.fetch_add(5, std::memory_order_relaxed);
  CUDA_TRY(nullptr, cudaGetLastError());
  CUDA_TRY(nullptr, cudaStreamSynchronize(st));     // the temporary handles are released on return
  return QGB_OK;
}
}  // namespace

int qgb_fft_interpolate(int device, int n, int N, int batch, const double* in, double* out, int on_device, void* stream) {
  if (!in || !out || batch < 1) return fail(nullptr, QGB_EINVAL, "bad argument");
  if (n % 2 != 0 || N % 2 != 0) return fail(nullptr, QGB_EINVAL, "Grid sizes (n,N) must be even");
  cudaStream_t st = S(stream);
  const int pairs = (batch + 1) / 2;
  qgb_config ca, cb;
  qgb_default_config(&ca);
  ca.device = device; ca.members = pairs; ca.nx = n;
  cb = ca; cb.nx = N;
  HandleGuard ga, gb;
  int rc;
  if ((rc = ga.create(ca))) return rc;
  if ((rc = gb.create(cb))) return rc;
  const cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind kout = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CUDA_TRY(nullptr, cudaMemsetAsync(ga.h->q, 0, nreal(ga.h) * sizeof(double), st));
  CUDA_TRY(nullptr, cudaMemcpyAsync(ga.h->q, in, (size_t)batch * n * n * sizeof(double), kin, st));
  StepIO io = base_io(ga.h);
  if ((rc = launch_program(ga.h, io, PROG_SET_Q, st))) return fail(nullptr, rc, "%s", ga.h->err.c_str());
  const double sc = ((double)N / n) * ((double)N / n);
  resample_kernel<<<grid_for((long long)2 * pairs * N * (N / 2 + 1)), 256, 0, st>>>(ga.h->qh, gb.h->qh, 2 * pairs, n, N, sc);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(nullptr, cudaGetLastError());
  StepIO iob = base_io(gb.h);
  if ((rc = launch_program(gb.h, iob, PROG_C2R, st))) return fail(nullptr, rc, "%s", gb.h->err.c_str());
  CUDA_TRY(nullptr, cudaMemcpyAsync(out, gb.h->q, (size_t)batch * N * N * sizeof(double), kout, st));
  CUDA_TRY(nullptr, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_operator(int device, int op, int n, int nc, int batch, const double* in, double* out, int on_device,
                 void* stream) {
  if (!in || !out || batch < 1) return fail(nullptr, QGB_EINVAL, "bad argument");
  if (op != 1 && op != 2 && op != 4 && op != 5) return fail(nullptr, QGB_EINVAL, "operator %d not supported (1, 2, 4, 5)", op);
  if (nc % 2 != 0) return fail(nullptr, QGB_EINVAL, "nc must be even");
  if (nc > n || nc < 4) return fail(nullptr, QGB_EINVAL, "nc must satisfy 4 <= nc <= n");
  cudaStream_t st = S(stream);
  const int pairs = (batch + 1) / 2;
  qgb_config cf, cc;
  qgb_default_config(&cf);
  cf.device = device; cf.members = pairs; cf.nx = n;
  cc = cf; cc.nx = nc;
  HandleGuard gf, gc;
  int rc = gf.create(cf);
  if (rc) return rc;
  rc = gc.create(cc);
  if (rc) return rc;
  qgb_handle *hf = gf.h, *hc = gc.h;
  const cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind kout = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  CUDA_TRY(nullptr, cudaMemsetAsync(hf->q, 0, nreal(hf) * sizeof(double), st));
  CUDA_TRY(nullptr, cudaMemcpyAsync(hf->q, in, (size_t)batch * n * n * sizeof(double), kin, st));
  StepIO io = base_io(hf);
  rc = launch_program(hf, io, PROG_SET_Q, st);                       // rfft2 of the field pairs
  if (rc) return fail(nullptr, rc, "%s", hf->err.c_str());
  const long long tot = (long long)2 * pairs * nc * (nc / 2 + 1);
  trunc_filter_kernel<<<grid_for(tot), 256, 0, st>>>(hf->qh, hc->qh, 2 * pairs, n, nc, op, cf.L, 1.0);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(nullptr, cudaGetLastError());
  StepIO ioc = base_io(hc);
  rc = launch_program(hc, ioc, PROG_C2R, st);                        // irfft2 on the coarse grid
  if (rc) return fail(nullptr, rc, "%s", hc->err.c_str());
  CUDA_TRY(nullptr, cudaMemcpyAsync(out, hc->q, (size_t)batch * nc * nc * sizeof(double), kout, st));
  CUDA_TRY(nullptr, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_subgrid_forcing(const qgb_config* cfg, int op, int nc, int dealias, int batch, const double* q, double* forcing,
                        double* qf, double* uf, double* vf, double* pf, int on_device, void* stream) {
  if (!cfg || !q || batch < 1) return fail(nullptr, QGB_EINVAL, "bad argument");
  if (dealias != 0 && dealias != 1 && dealias != 2) return fail(nullptr, QGB_EINVAL, "dealias should be none or 2/3-rule or 3/2-rule");
  if (op != 1 && op != 2 && op != 4 && op != 5) return fail(nullptr, QGB_EINVAL, "operator %d not supported (1, 2, 4, 5)", op);
  const int n = cfg->nx;
  if (nc % 2 != 0) return fail(nullptr, QGB_EINVAL, "nc must be even");
  if (nc > n || nc < 4) return fail(nullptr, QGB_EINVAL, "nc must satisfy 4 <= nc <= n");
  cudaStream_t st = S(stream);
  qgb_config cf = *cfg, cc = *cfg;
  cf.members = batch; cc.members = batch; cc.nx = nc;
  HandleGuard gf, gc;
  int rc = gf.create(cf);
  if (rc) return rc;
  rc = gc.create(cc);
  if (rc) return rc;
  qgb_handle *hf = gf.h, *hc = gc.h;
  const cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind kout = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  const Tables T0f = no_background(hf->T), T0c = no_background(hc->T);
  const long long totc = (long long)2 * batch * nc * (nc / 2 + 1);
#define OPRUN(h, io, prog, tov)                                              \
  do {                                                                       \
    int _rc = launch_program(h, io, prog, st, tov);                          \
    if (_rc) return fail(nullptr, _rc, "%s", (h)->err.c_str());              \
  } while (0)
  // fine grid: qh = rfft2(q); adv_f_h = ik (uq)_h + il (vq)_h with u, v from the fine inversion (apply_operator_to_model(q, 1, id))
  CUDA_TRY(nullptr, cudaMemcpyAsync(hf->q, q, nreal(hf) * sizeof(double), kin, st));
  StepIO io = base_io(hf);
  OPRUN(hf, io, PROG_SET_Q, nullptr);
  io.d_cur = hf->hist[0];
  double fine_sign = -1.0;
  if (dealias == 0) {
    OPRUN(hf, io, PROG_ADVECT, &T0f);                                 // hist[0] = -adv_f_h
  } else {
    rc = qgb_invert(hf, stream);                                      // u, v of the fine model
    if (rc) return fail(nullptr, rc, "%s", hf->err.c_str());
    rc = dealias == 1 ? advect_23(cf, n, batch, hf->q, hf->u, hf->v, hf->hist[0], st)
                      : advect_dealiased(cf, n, batch, hf->q, hf->u, hf->v, hf->hist[0], st);   // hist[0] = +adv_f_h
    if (rc) return rc;
    fine_sign = 1.0;
  }
  // coarse grid: qf_h = op(q)_h ; S_f = op(adv_f)_h
  trunc_filter_kernel<<<grid_for(totc), 256, 0, st>>>(hf->qh, hc->qh, 2 * batch, n, nc, op, cf.L, 1.0);
  trunc_filter_kernel<<<grid_for(totc), 256, 0, st>>>(hf->hist[0], hc->hist[1], 2 * batch, n, nc, op, cf.L, fine_sign);
  g_launches.fetch_add(2, std::memory_order_relaxed);
  CUDA_TRY(nullptr, cudaGetLastError());
  StepIO ioc = base_io(hc);
  OPRUN(hc, ioc, PROG_C2R, nullptr);                                  // qf = irfft2(qf_h)
  rc = qgb_invert(hc, stream);                                        // psi_f, u_f, v_f (apply_operator_to_model)
  if (rc) return fail(nullptr, rc, "%s", hc->err.c_str());
  ioc.d_cur = hc->hist[0];
  double coarse_sign = -1.0;
  if (dealias == 0) {
    OPRUN(hc, ioc, PROG_ADVECT, &T0c);                                // hist[0] = -adv_c_h
  } else {
    rc = dealias == 1 ? advect_23(cc, nc, batch, hc->q, hc->u, hc->v, hc->hist[0], st)
                      : advect_dealiased(cc, nc, batch, hc->q, hc->u, hc->v, hc->hist[0], st);
    if (rc) return rc;
    coarse_sign = 1.0;
  }
  caxpby_kernel<<<grid_for(totc), 256, 0, st>>>(hc->hist[0], hc->hist[1], hc->hist[2], totc, coarse_sign, -1.0);
  QGB_COUNT_LAUNCH();
  CUDA_TRY(nullptr, cudaGetLastError());
  if (qf) CUDA_TRY(nullptr, cudaMemcpyAsync(qf, hc->q, nreal(hc) * sizeof(double), kout, st));
  if (uf) CUDA_TRY(nullptr, cudaMemcpyAsync(uf, hc->u, nreal(hc) * sizeof(double), kout, st));
  if (vf) CUDA_TRY(nullptr, cudaMemcpyAsync(vf, hc->v, nreal(hc) * sizeof(double), kout, st));
  if (pf) CUDA_TRY(nullptr, cudaMemcpyAsync(pf, hc->p, nreal(hc) * sizeof(double), kout, st));
  if (forcing) {
    StepIO iof = base_io(hc);
    iof.qh = hc->hist[2];
    iof.q = hc->u;                                                    // reuse a real buffer for irfft2(forcing_h)
    OPRUN(hc, iof, PROG_C2R, nullptr);
    CUDA_TRY(nullptr, cudaMemcpyAsync(forcing, hc->u, nreal(hc) * sizeof(double), kout, st));
  }
#undef OPRUN
  CUDA_TRY(nullptr, cudaStreamSynchronize(st));
  return QGB_OK;
}

}  // extern "C"
