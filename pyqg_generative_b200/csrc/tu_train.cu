// tu_train.cu -- C ABI of the on-device AndrewCNN training step (include/qgb200.h "training"; kernels in train.cuh).
// Reference: tools/cnn_tools.py:645-700 (train), :177-182 (compute_loss), models/mean_var_model.py:41-66 (two-stage GZ fit).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qgb200.h"
#include "adv.cuh"
#include "tgemm.cuh"
#include "train.cuh"

using namespace qgb;
using namespace qgb::train;
using namespace qgb::adv;

namespace {
std::string g_train_create_error;
}

struct qgb_trainer {
  int device = 0, ny = 0, nx = 0, max_batch = 0, softplus = 0, nlayers = 0;
  struct Layer {
    int cin, cout, ks, bn;                 // bn: ReLU + BatchNorm follow the convolution (every layer but the last)
    size_t w, b, g, be;                    // offsets into the flat parameter vector: conv weight, conv bias, BN weight, BN bias
    size_t rm;                             // offset into the flat buffer vector: running_mean (running_var follows at rm + cout)
    size_t stat;                           // offset into mean / invstd / fold scratch (per BN channel)
  };
  std::vector<Layer> L;
  size_t nparams = 0, nbuffers = 0, nstat = 0;
  float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr, *BUF = nullptr;
  float *mean = nullptr, *invstd = nullptr, *fold_s = nullptr, *fold_t = nullptr, *ones = nullptr, *zeros = nullptr;
  float *x = nullptr, *t = nullptr;        // staged minibatch (host callers)
  std::vector<float*> r, a;                // per (slot, layer): r_l = relu(conv) [last layer: z_L], a_l = BatchNorm output
  int nslots = 1, slot = 0;                // activation sets (the GAN step keeps two generator passes alive); slot = the one in use
  float* Gacc = nullptr;                   // gradient accumulator over several backward passes
  float* scr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // scratch of the CVAE / CGAN steps (grown on demand)
  size_t scr_floats[6] = {0, 0, 0, 0, 0, 0};
  double* stats = nullptr;                 // loss partial sums and results of those steps
  float* d[2] = {nullptr, nullptr};        // gradient ping-pong, max channels x batch x pixels
  float* wp = nullptr;                     // packed weights of the layer at hand
  float* wg_part = nullptr; size_t wg_part_floats = 0;
  double *red_part = nullptr, *loss_part = nullptr, *loss = nullptr;
  long long adam_t = 0;
  float beta1 = 0.9f, beta2 = 0.999f, adam_eps = 1e-8f, bn_eps = 1e-5f, bn_momentum = 0.1f;
  long long launches = 0;
  std::string err;
};

namespace {

int tfail(qgb_trainer* t, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (t) t->err = buf; else g_train_create_error = buf;
  return code;
}
#define TR_TRY(t, expr)                                                                                         \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return tfail(t, QGB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

template <typename T>
cudaError_t talloc(T** p, size_t n) { return cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)); }

constexpr int kLossBlocks = 256;
constexpr int kMaxSlots = 2;

inline float*& R(qgb_trainer* t, int l) { return t->r[(size_t)t->slot * t->nlayers + l]; }
inline float*& A(qgb_trainer* t, int l) { return t->a[(size_t)t->slot * t->nlayers + l]; }
inline float* MEAN(qgb_trainer* t, const qgb_trainer::Layer& L) { return t->mean + (size_t)t->slot * t->nstat + L.stat; }
inline float* INVSTD(qgb_trainer* t, const qgb_trainer::Layer& L) { return t->invstd + (size_t)t->slot * t->nstat + L.stat; }

inline int co_pad_of(int cout) { const int ct = cout <= 4 ? 2 : 32; return (cout + ct - 1) / ct * ct; }

// y = conv(x) + bias, optionally relu and a per-channel affine (the FFMA kernel of the fp32 inference path)
int conv(qgb_trainer* t, const float* in, float* out, const float* wp, const float* bias, const float* s, const float* sh, int cin,
         int cout, int ks, int relu_affine, int batch, cudaStream_t st) {
  const int ny = t->ny, nx = t->nx;
  const int tiles_x = (nx + kConvTile - 1) / kConvTile, tiles_y = (ny + kConvTile - 1) / kConvTile;
  const bool small = cout <= 4;
  const int co_t = small ? 2 : 32, cpad = co_pad_of(cout);
  dim3 grid(tiles_x * tiles_y, (cout + co_t - 1) / co_t, batch);          // (the register-tiled kernel of the wide layers sizes its own grid)
  const long long ibs = (long long)cin * ny * nx, obs = (long long)cout * ny * nx;
#define QGB_TCONV(KS, CT) conv_ffma_kernel<KS, CT><<<grid, 256, 0, st>>>(in, ibs, out, obs, wp, bias, s, sh, cin, cout, cpad, ny, nx, tiles_x, relu_affine, 0, 0)
#define QGB_TCONV2(KS, CT) TR_TRY(t, (launch_conv_ffma2<KS, CT>(batch, st, in, ibs, out, obs, wp, bias, s, sh, cin, cout, cpad, ny, nx, relu_affine, 0, 0)))
  if (ks == 5 && !small) QGB_TCONV2(5, 32);
  else if (ks == 5) QGB_TCONV(5, 2);
  else if (ks == 3 && !small) QGB_TCONV2(3, 32);
  else if (ks == 3) QGB_TCONV(3, 2);
  else if (ks == 1 && !small) QGB_TCONV(1, 32);
  else if (ks == 1) QGB_TCONV(1, 2);
  else return tfail(t, QGB_EUNSUPPORTED, "kernel size %d not supported (1, 3, 5)", ks);
#undef QGB_TCONV2
#undef QGB_TCONV
  t->launches++;
  TR_TRY(t, cudaGetLastError());
  return QGB_OK;
}

int pack(qgb_trainer* t, const qgb_trainer::Layer& L, int flip, cudaStream_t st) {
  const int out_c = flip ? L.cin : L.cout, in_c = flip ? L.cout : L.cin;
  const long long n = (long long)in_c * L.ks * L.ks * co_pad_of(out_c);
  pack_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->P + L.w, t->wp, L.cin, L.cout, L.ks, co_pad_of(out_c), flip);
  t->launches++;
  TR_TRY(t, cudaGetLastError());
  return QGB_OK;
}

template <int MODE>
int chan_reduce(qgb_trainer* t, const float* x, const float* r, const float* mean, const float* invstd, int batch, int C,
                cudaStream_t st) {
  chan_partial_kernel<MODE><<<dim3(C, kRedSplit), 256, 0, st>>>(x, r, mean, invstd, batch, C, t->ny * t->nx, t->red_part);
  t->launches++;
  TR_TRY(t, cudaGetLastError());
  return QGB_OK;
}

template <int KS, int CI_T, int CO_PER, int DB = 0>
int wgrad_launch(qgb_trainer* t, const float* a, const float* dz, float* dW, int cin, int cout, int batch, cudaStream_t st) {
  using G = WgGeom<KS>;
  constexpr int CO_B = (256 / CI_T) * CO_PER;
  const size_t smem = (size_t)(DB ? 2 : 1) * (CI_T * G::CI_STRIDE + CO_B * G::D_PITCH) * sizeof(float);
  auto kern = wgrad_ffma_kernel<KS, CI_T, CO_PER, DB>;
  TR_TRY(t, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int ci_blocks = (cin + CI_T - 1) / CI_T, co_blocks = (cout + CO_B - 1) / CO_B;
  const int tiles = ((t->nx + kWgTile - 1) / kWgTile) * ((t->ny + kWgTile - 1) / kWgTile);
  const int items = batch * tiles;
  int splits = (2 * 148 + ci_blocks * co_blocks - 1) / (ci_blocks * co_blocks);    // about two waves of blocks
  if (splits > items) splits = items;
  if (splits < 1) splits = 1;
  const size_t n = (size_t)cout * cin * KS * KS;
  if (t->wg_part_floats < n * splits) {
    if (t->wg_part) cudaFree(t->wg_part);
    t->wg_part = nullptr; t->wg_part_floats = 0;
    TR_TRY(t, talloc(&t->wg_part, n * splits));
    t->wg_part_floats = n * splits;
  }
  kern<<<dim3(ci_blocks * co_blocks, splits), 256, smem, st>>>(a, dz, t->wg_part, batch, cin, cout, t->ny, t->nx, ci_blocks);
  TR_TRY(t, cudaGetLastError());
  wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t->wg_part, splits, (long long)n, dW);
  TR_TRY(t, cudaGetLastError());
  t->launches += 2;
  return QGB_OK;
}

int wgrad(qgb_trainer* t, const float* a, const float* dz, float* dW, int cin, int cout, int ks, int batch, cudaStream_t st) {
  const bool thin_in = cin <= 4, thin_out = cout <= 8;
  if (ks == 5) {
    if (thin_in) return wgrad_launch<5, 4, 1, 1>(t, a, dz, dW, cin, cout, batch, st);
    if (thin_out) return wgrad_launch<5, 32, 1>(t, a, dz, dW, cin, cout, batch, st);
    // 4 output channels per thread, one block per SM: 4.3 ms for the 128 -> 64 layer (64 images, 64^2); the 2-channel variant with two
    // blocks per SM measured 6.2 ms
    return wgrad_launch<5, 32, 4, 1>(t, a, dz, dW, cin, cout, batch, st);     // cp.async double buffer (one block per SM anyway)
  }
  if (ks == 3) {
    if (thin_in) return wgrad_launch<3, 4, 2>(t, a, dz, dW, cin, cout, batch, st);
    if (thin_out) return wgrad_launch<3, 32, 1>(t, a, dz, dW, cin, cout, batch, st);
    return wgrad_launch<3, 32, 4, 1>(t, a, dz, dW, cin, cout, batch, st);
  }
  if (ks == 1) {
    if (thin_in) return wgrad_launch<1, 4, 2>(t, a, dz, dW, cin, cout, batch, st);
    if (thin_out) return wgrad_launch<1, 32, 1>(t, a, dz, dW, cin, cout, batch, st);
    return wgrad_launch<1, 32, 4>(t, a, dz, dW, cin, cout, batch, st);
  }
  return tfail(t, QGB_EUNSUPPORTED, "kernel size %d not supported (1, 3, 5)", ks);
}

inline unsigned ew_blocks(long long n) { long long b = (n + 255) / 256; return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

// stage a host minibatch (or borrow device pointers)
int stage(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, const float** xd, const float** yd, cudaStream_t st) {
  if (!x || !y) return tfail(t, QGB_EINVAL, "null minibatch");
  if (batch < 1 || batch > t->max_batch) return tfail(t, QGB_EINVAL, "batch %d outside 1..%d", batch, t->max_batch);
  const size_t hw = (size_t)t->ny * t->nx;
  if (on_device) { *xd = x; *yd = y; return QGB_OK; }
  TR_TRY(t, cudaMemcpyAsync(t->x, x, batch * t->L.front().cin * hw * sizeof(float), cudaMemcpyHostToDevice, st));
  TR_TRY(t, cudaMemcpyAsync(t->t, y, batch * t->L.back().cout * hw * sizeof(float), cudaMemcpyHostToDevice, st));
  *xd = t->x; *yd = t->t;
  return QGB_OK;
}

// a second set of saved activations (the GAN step differentiates through two generator passes of the same minibatch)
int ensure_slots(qgb_trainer* t, int n) {
  if (n > kMaxSlots) return tfail(t, QGB_EINVAL, "at most %d activation slots", kMaxSlots);
  const size_t hw = (size_t)t->ny * t->nx, B = t->max_batch;
  while (t->nslots < n) {
    for (int l = 0; l < t->nlayers; ++l) {
      float *r = nullptr, *a = nullptr;
      TR_TRY(t, talloc(&r, B * t->L[l].cout * hw));
      t->r.push_back(r);
      if (t->L[l].bn) TR_TRY(t, talloc(&a, B * t->L[l].cout * hw));
      t->a.push_back(a);
    }
    t->nslots++;
  }
  return QGB_OK;
}

// forward pass; training = batch statistics (and running-statistics update), else running statistics folded into the epilogue
int forward(qgb_trainer* t, const float* xd, int batch, bool training, cudaStream_t st) {
  const int hw = t->ny * t->nx;
  const float* in = xd;
  for (int l = 0; l < t->nlayers; ++l) {
    const auto& L = t->L[l];
    int rc = pack(t, L, 0, st);
    if (rc) return rc;
    if (!L.bn) {
      rc = conv(t, in, R(t, l), t->wp, t->P + L.b, t->ones, t->zeros, L.cin, L.cout, L.ks, 0, batch, st);
      if (rc) return rc;
      in = R(t, l);
    } else if (training) {
      rc = conv(t, in, R(t, l), t->wp, t->P + L.b, t->ones, t->zeros, L.cin, L.cout, L.ks, 1, batch, st);
      if (rc) return rc;
      rc = chan_reduce<0>(t, R(t, l), nullptr, nullptr, nullptr, batch, L.cout, st);
      if (rc) return rc;
      bn_stats_final_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(t->red_part, L.cout, (double)batch * hw, t->bn_eps, t->bn_momentum,
                                                                  MEAN(t, L), INVSTD(t, L), t->BUF + L.rm,
                                                                  t->BUF + L.rm + L.cout, 1);
      const long long total = (long long)batch * L.cout * hw;
      bn_apply_kernel<<<ew_blocks(total), 256, 0, st>>>(R(t, l), A(t, l), t->P + L.g, t->P + L.be, MEAN(t, L), INVSTD(t, L),
                                                        L.cout, hw, total);
      t->launches += 2;
      TR_TRY(t, cudaGetLastError());
      in = A(t, l);
    } else {
      bn_fold_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(t->P + L.g, t->P + L.be, t->BUF + L.rm, t->BUF + L.rm + L.cout, t->bn_eps,
                                                           L.cout, t->fold_s + L.stat, t->fold_t + L.stat);
      t->launches++;
      TR_TRY(t, cudaGetLastError());
      rc = conv(t, in, A(t, l), t->wp, t->P + L.b, t->fold_s + L.stat, t->fold_t + L.stat, L.cin, L.cout, L.ks, 1, batch, st);
      if (rc) return rc;
      in = A(t, l);
    }
  }
  return QGB_OK;
}

int loss_and_grad(qgb_trainer* t, const float* yd, int batch, bool want_grad, cudaStream_t st) {
  const auto& L = t->L.back();
  const long long n = (long long)batch * L.cout * t->ny * t->nx;
  mse_loss_kernel<<<kLossBlocks, 256, 0, st>>>(R(t, t->nlayers - 1), yd, want_grad ? t->d[0] : nullptr, n, t->softplus, t->loss_part);
  loss_final_kernel<<<1, 32, 0, st>>>(t->loss_part, kLossBlocks, 1.0 / (double)n, t->loss);
  t->launches += 2;
  TR_TRY(t, cudaGetLastError());
  return QGB_OK;
}

// gradients of every parameter into t->G (same flat layout as the parameters); d[0] holds dL/dz of the last layer on entry.
// dx (optional): gradient with respect to the network input, (batch, cin, ny, nx).
int backward(qgb_trainer* t, const float* xd, int batch, cudaStream_t st, float* dx = nullptr) {
  const int hw = t->ny * t->nx;
  int cur = 0;
  for (int l = t->nlayers - 1; l >= 0; --l) {
    const auto& L = t->L[l];
    float* dz = t->d[cur];
    const float* a_in = l == 0 ? xd : A(t, l - 1);
    // bias gradient
    // (bias gradient: for the last layer a reduction of its own; below it the BatchNorm / ReLU backward pass of the previous iteration of
    // this loop has left the per-channel sums of dz with the rest of its work)
    int rc = QGB_OK;
    if (l == t->nlayers - 1) {
      rc = chan_reduce<2>(t, dz, nullptr, nullptr, nullptr, batch, L.cout, st);
      if (rc) return rc;
      chan_final_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(t->red_part, L.cout, t->G + L.b, nullptr);
      t->launches++;
      TR_TRY(t, cudaGetLastError());
    }
    // weight gradient
    rc = wgrad(t, a_in, dz, t->G + L.w, L.cin, L.cout, L.ks, batch, st);
    if (rc) return rc;
    if (l == 0 && !dx) break;
    // data gradient: the forward kernel on the flipped, transposed weights, no bias
    rc = pack(t, L, 1, st);
    if (rc) return rc;
    float* da = l == 0 ? dx : t->d[cur ^ 1];
    rc = conv(t, dz, da, t->wp, t->zeros, t->ones, t->zeros, L.cout, L.cin, L.ks, 0, batch, st);
    if (rc) return rc;
    if (l == 0) break;
    // BatchNorm + ReLU backward of layer l - 1
    const auto& Lp = t->L[l - 1];
    rc = chan_reduce<1>(t, da, R(t, l - 1), MEAN(t, Lp), INVSTD(t, Lp), batch, Lp.cout, st);
    if (rc) return rc;
    chan_final_kernel<<<(Lp.cout + 127) / 128, 128, 0, st>>>(t->red_part, Lp.cout, t->G + Lp.be, t->G + Lp.g);
    bn_relu_bwd_chan_kernel<<<dim3(Lp.cout, kRedSplit), 256, 0, st>>>(da, R(t, l - 1), t->P + Lp.g, MEAN(t, Lp), INVSTD(t, Lp), t->G + Lp.g,
                                                                      t->G + Lp.be, batch, Lp.cout, hw, 1.f / ((float)batch * hw), t->red_part);
    chan_final_kernel<<<(Lp.cout + 127) / 128, 128, 0, st>>>(t->red_part, Lp.cout, t->G + Lp.b, nullptr);       // bias gradient of layer l - 1
    t->launches += 3;
    TR_TRY(t, cudaGetLastError());
    cur ^= 1;
  }
  return QGB_OK;
}

int ensure_scratch(qgb_trainer* t, int i, size_t floats) {
  if (t->scr_floats[i] >= floats) return QGB_OK;
  if (t->scr[i]) cudaFree(t->scr[i]);
  t->scr[i] = nullptr; t->scr_floats[i] = 0;
  TR_TRY(t, talloc(&t->scr[i], floats));
  t->scr_floats[i] = floats;
  return QGB_OK;
}
int ensure_stats(qgb_trainer* t) {
  if (!t->stats) TR_TRY(t, talloc(&t->stats, (size_t)kPartBlocks * 8 + 64));
  return QGB_OK;
}
// host array -> device scratch i (or borrow the device pointer)
int stage_into(qgb_trainer* t, int i, const float* src, size_t floats, int on_device, const float** out, cudaStream_t st) {
  if (on_device) { *out = src; return QGB_OK; }
  int rc = ensure_scratch(t, i, floats);
  if (rc) return rc;
  TR_TRY(t, cudaMemcpyAsync(t->scr[i], src, floats * sizeof(float), cudaMemcpyHostToDevice, st));
  *out = t->scr[i];
  return QGB_OK;
}

// one Adam update of every parameter from t->G
int adam_update(qgb_trainer* t, double lr, cudaStream_t st) {
  t->adam_t += 1;
  const double bc1 = 1.0 - std::pow((double)t->beta1, (double)t->adam_t), bc2 = 1.0 - std::pow((double)t->beta2, (double)t->adam_t);
  adam_kernel<<<ew_blocks((long long)t->nparams), 256, 0, st>>>(t->P, t->G, t->M, t->V, (long long)t->nparams, (float)lr, t->beta1,
                                                                t->beta2, t->adam_eps, (float)bc1, (float)std::sqrt(bc2));
  t->launches++;
  TR_TRY(t, cudaGetLastError());
  return QGB_OK;
}

}  // namespace

extern "C" {

int qgb_train_create(int device, int nlayers, const int32_t* channels, const int32_t* ksizes, int ny, int nx, int max_batch,
                     int softplus, qgb_trainer** out) {
  if (!out || !channels || !ksizes || nlayers < 1 || ny < 1 || nx < 1 || max_batch < 1)
    return tfail(nullptr, QGB_EINVAL, "bad argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return tfail(nullptr, QGB_ECUDA, "no CUDA device available: libqgb200 has no CPU fallback");
  if (device < 0 || device >= ndev) return tfail(nullptr, QGB_EINVAL, "device %d out of range", device);
  qgb_trainer* t = new qgb_trainer();
  auto bail = [&](int rc) { g_train_create_error = t->err; qgb_train_destroy(t); return rc; };
  t->device = device; t->ny = ny; t->nx = nx; t->max_batch = max_batch; t->softplus = softplus; t->nlayers = nlayers;
#define CR(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { t->err = std::string(#expr) + ": " + cudaGetErrorString(_e); return bail(QGB_ECUDA); } } while (0)
  CR(cudaSetDevice(device));
  int maxc = channels[0];
  size_t max_wp = 0;
  for (int l = 0; l < nlayers; ++l) {
    qgb_trainer::Layer L;
    L.cin = channels[l]; L.cout = channels[l + 1]; L.ks = ksizes[l]; L.bn = l + 1 < nlayers;
    if (L.cin < 1 || L.cout < 1 || !(L.ks == 1 || L.ks == 3 || L.ks == 5)) { t->err = "unsupported layer (kernel sizes 1, 3, 5)"; return bail(QGB_EUNSUPPORTED); }
    L.w = t->nparams; t->nparams += (size_t)L.cout * L.cin * L.ks * L.ks;
    L.b = t->nparams; t->nparams += L.cout;
    L.g = L.be = L.rm = L.stat = 0;
    if (L.bn) {
      L.g = t->nparams; t->nparams += L.cout;
      L.be = t->nparams; t->nparams += L.cout;
      L.rm = t->nbuffers; t->nbuffers += 2 * (size_t)L.cout;
      L.stat = t->nstat; t->nstat += L.cout;
    }
    maxc = L.cout > maxc ? L.cout : maxc;
    const size_t f = (size_t)L.cin * L.ks * L.ks * co_pad_of(L.cout), g = (size_t)L.cout * L.ks * L.ks * co_pad_of(L.cin);
    max_wp = f > max_wp ? f : max_wp;
    max_wp = g > max_wp ? g : max_wp;
    t->L.push_back(L);
  }
  const size_t hw = (size_t)ny * nx, B = max_batch;
  CR(talloc(&t->P, t->nparams)); CR(talloc(&t->G, t->nparams)); CR(talloc(&t->M, t->nparams)); CR(talloc(&t->V, t->nparams));
  CR(talloc(&t->BUF, t->nbuffers));
  CR(talloc(&t->mean, t->nstat * kMaxSlots)); CR(talloc(&t->invstd, t->nstat * kMaxSlots)); CR(talloc(&t->fold_s, t->nstat)); CR(talloc(&t->fold_t, t->nstat));
  CR(talloc(&t->ones, (size_t)maxc)); CR(talloc(&t->zeros, (size_t)maxc));
  CR(talloc(&t->x, B * channels[0] * hw)); CR(talloc(&t->t, B * channels[nlayers] * hw));
  for (int l = 0; l < nlayers; ++l) {
    float *r = nullptr, *a = nullptr;
    CR(talloc(&r, B * t->L[l].cout * hw));
    t->r.push_back(r);
    if (t->L[l].bn) CR(talloc(&a, B * t->L[l].cout * hw));
    t->a.push_back(a);
  }
  CR(talloc(&t->d[0], B * maxc * hw)); CR(talloc(&t->d[1], B * maxc * hw));
  CR(talloc(&t->wp, max_wp));
  CR(talloc(&t->red_part, (size_t)maxc * kRedSplit * 2)); CR(talloc(&t->loss_part, (size_t)kLossBlocks)); CR(talloc(&t->loss, (size_t)1));
  CR(cudaMemset(t->M, 0, t->nparams * sizeof(float))); CR(cudaMemset(t->V, 0, t->nparams * sizeof(float)));
  CR(cudaMemset(t->P, 0, t->nparams * sizeof(float))); CR(cudaMemset(t->G, 0, t->nparams * sizeof(float)));
  CR(cudaMemset(t->zeros, 0, (size_t)maxc * sizeof(float)));
  fill_kernel<<<1, 256>>>(t->ones, maxc, 1.f);
  CR(cudaGetLastError());
  CR(cudaDeviceSynchronize());
#undef CR
  *out = t;
  return QGB_OK;
}

void qgb_train_destroy(qgb_trainer* t) {
  if (!t) return;
  cudaSetDevice(t->device);
  for (float* p : {t->Gacc, t->P, t->G, t->M, t->V, t->BUF, t->mean, t->invstd, t->fold_s, t->fold_t, t->ones, t->zeros, t->x, t->t, t->d[0],
                   t->d[1], t->wp, t->wg_part})
    if (p) cudaFree(p);
  for (float* p : t->scr) if (p) cudaFree(p);
  if (t->stats) cudaFree(t->stats);
  for (float* p : t->r) if (p) cudaFree(p);
  for (float* p : t->a) if (p) cudaFree(p);
  if (t->red_part) cudaFree(t->red_part);
  if (t->loss_part) cudaFree(t->loss_part);
  if (t->loss) cudaFree(t->loss);
  delete t;
}

const char* qgb_train_last_error(const qgb_trainer* t) { return t ? t->err.c_str() : g_train_create_error.c_str(); }
int64_t qgb_train_num_params(const qgb_trainer* t) { return t ? (int64_t)t->nparams : 0; }
int64_t qgb_train_num_buffers(const qgb_trainer* t) { return t ? (int64_t)t->nbuffers : 0; }
int64_t qgb_train_launch_count(const qgb_trainer* t) { return t ? (int64_t)t->launches : 0; }

int qgb_train_set_params(qgb_trainer* t, const float* params, const float* buffers, int reset_optimizer) {
  if (!t) return QGB_EINVAL;
  TR_TRY(t, cudaSetDevice(t->device));
  if (params) TR_TRY(t, cudaMemcpy(t->P, params, t->nparams * sizeof(float), cudaMemcpyHostToDevice));
  if (buffers && t->nbuffers) TR_TRY(t, cudaMemcpy(t->BUF, buffers, t->nbuffers * sizeof(float), cudaMemcpyHostToDevice));
  if (reset_optimizer) {
    TR_TRY(t, cudaMemset(t->M, 0, t->nparams * sizeof(float)));
    TR_TRY(t, cudaMemset(t->V, 0, t->nparams * sizeof(float)));
    t->adam_t = 0;
  }
  return QGB_OK;
}

int qgb_train_get_params(qgb_trainer* t, float* params, float* buffers) {
  if (!t) return QGB_EINVAL;
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaDeviceSynchronize());
  if (params) TR_TRY(t, cudaMemcpy(params, t->P, t->nparams * sizeof(float), cudaMemcpyDeviceToHost));
  if (buffers && t->nbuffers) TR_TRY(t, cudaMemcpy(buffers, t->BUF, t->nbuffers * sizeof(float), cudaMemcpyDeviceToHost));
  return QGB_OK;
}

int qgb_train_step(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, double lr, double* loss, void* stream) {
  if (!t) return QGB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  TR_TRY(t, cudaSetDevice(t->device));
  const float *xd, *yd;
  int rc = stage(t, x, y, batch, on_device, &xd, &yd, st);
  if (rc) return rc;
  if ((rc = forward(t, xd, batch, true, st))) return rc;
  if ((rc = loss_and_grad(t, yd, batch, true, st))) return rc;
  if ((rc = backward(t, xd, batch, st))) return rc;
  if ((rc = adam_update(t, lr, st))) return rc;
  if (loss) {
    TR_TRY(t, cudaMemcpyAsync(loss, t->loss, sizeof(double), cudaMemcpyDeviceToHost, st));
    TR_TRY(t, cudaStreamSynchronize(st));
  }
  return QGB_OK;
}

int qgb_train_grads(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, float* grads, double* loss,
                    int update_running, void* stream) {
  if (!t) return QGB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  TR_TRY(t, cudaSetDevice(t->device));
  const float *xd, *yd;
  int rc = stage(t, x, y, batch, on_device, &xd, &yd, st);
  if (rc) return rc;
  std::vector<float> keep;
  if (!update_running && t->nbuffers) {      // a pure gradient evaluation leaves the running statistics alone
    keep.resize(t->nbuffers);
    TR_TRY(t, cudaMemcpyAsync(keep.data(), t->BUF, t->nbuffers * sizeof(float), cudaMemcpyDeviceToHost, st));
    TR_TRY(t, cudaStreamSynchronize(st));
  }
  if ((rc = forward(t, xd, batch, true, st))) return rc;
  if ((rc = loss_and_grad(t, yd, batch, true, st))) return rc;
  if ((rc = backward(t, xd, batch, st))) return rc;
  if (!keep.empty()) TR_TRY(t, cudaMemcpyAsync(t->BUF, keep.data(), t->nbuffers * sizeof(float), cudaMemcpyHostToDevice, st));
  if (grads) TR_TRY(t, cudaMemcpyAsync(grads, t->G, t->nparams * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (loss) TR_TRY(t, cudaMemcpyAsync(loss, t->loss, sizeof(double), cudaMemcpyDeviceToHost, st));
  TR_TRY(t, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_train_eval_loss(qgb_trainer* t, const float* x, const float* y, int batch, int on_device, double* loss, void* stream) {
  if (!t || !loss) return QGB_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  TR_TRY(t, cudaSetDevice(t->device));
  const float *xd, *yd;
  int rc = stage(t, x, y, batch, on_device, &xd, &yd, st);
  if (rc) return rc;
  if ((rc = forward(t, xd, batch, false, st))) return rc;
  if ((rc = loss_and_grad(t, yd, batch, false, st))) return rc;
  TR_TRY(t, cudaMemcpyAsync(loss, t->loss, sizeof(double), cudaMemcpyDeviceToHost, st));
  TR_TRY(t, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_train_get_grads(qgb_trainer* t, float* grads) {
  if (!t || !grads) return QGB_EINVAL;
  TR_TRY(t, cudaSetDevice(t->device));
  TR_TRY(t, cudaDeviceSynchronize());
  TR_TRY(t, cudaMemcpy(grads, t->G, t->nparams * sizeof(float), cudaMemcpyDeviceToHost));
  return QGB_OK;
}

int qgb_train_set_adam(qgb_trainer* t, double beta1, double beta2) {
  if (!t || !(beta1 >= 0.0 && beta1 < 1.0) || !(beta2 >= 0.0 && beta2 < 1.0)) return tfail(t, QGB_EINVAL, "betas must lie in [0, 1)");
  t->beta1 = (float)beta1; t->beta2 = (float)beta2;
  return QGB_OK;
}

int qgb_train_cvae_step(qgb_trainer* enc, qgb_trainer* dec, const float* x, const float* y, const float* eps, int batch,
                        int on_device, double lr, double decoder_var, int update, double* losses, void* stream) {
  if (!enc || !dec) return QGB_EINVAL;
  qgb_trainer* t = enc;
  if (!x || !y || !eps) return tfail(t, QGB_EINVAL, "null minibatch");
  if (enc->L.front().cin != 4 || enc->L.back().cout != 4 || dec->L.front().cin != 4 || dec->L.back().cout != 2)
    return tfail(t, QGB_EINVAL, "expected an encoder 4 -> 4 ([x, y] -> [mu, logvar]) and a decoder 4 -> 2 ([x, z] -> y)");
  if (enc->ny != dec->ny || enc->nx != dec->nx || enc->device != dec->device) return tfail(t, QGB_EINVAL, "encoder and decoder differ in grid or device");
  if (batch < 1 || batch > enc->max_batch || batch > dec->max_batch) return tfail(t, QGB_EINVAL, "batch %d outside 1..max_batch", batch);
  if (enc->softplus || dec->softplus) return tfail(t, QGB_EINVAL, "softplus heads are not part of the CVAE");
  cudaStream_t st = (cudaStream_t)stream;
  TR_TRY(t, cudaSetDevice(t->device));
  const int hw = t->ny * t->nx;
  const size_t f2 = (size_t)batch * 2 * hw, f4 = (size_t)batch * 4 * hw;
  const float *xd, *yd, *ed;
  int rc;
  if ((rc = stage_into(t, 0, x, f2, on_device, &xd, st))) return rc;
  if ((rc = stage_into(t, 1, y, f2, on_device, &yd, st))) return rc;
  if ((rc = stage_into(t, 2, eps, f2, on_device, &ed, st))) return rc;
  if ((rc = ensure_scratch(t, 3, f4)) || (rc = ensure_scratch(t, 4, f4)) || (rc = ensure_scratch(t, 5, f4)) || (rc = ensure_stats(t))) return rc;
  float *encin = t->scr[3], *decin = t->scr[4], *ddecin = t->scr[5];
  enc->slot = dec->slot = 0;
  // encoder on cat[x, y]
  cat_channels_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(xd, 2, encin, 4, 0, hw, (long long)f2);
  cat_channels_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(yd, 2, encin, 4, 2, hw, (long long)f2);
  t->launches += 2;
  if ((rc = forward(enc, encin, batch, true, st))) return rc;
  const float* encout = R(enc, enc->nlayers - 1);
  // z = eps std + mu ; decoder on cat[x, z]
  cvae_reparam_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(xd, encout, ed, decin, hw, (long long)f2);
  t->launches++;
  if ((rc = forward(dec, decin, batch, true, st))) { enc->err = dec->err; return rc; }
  const float* yhat = R(dec, dec->nlayers - 1);
  // losses and d loss / d yhat
  cvae_loss_partial_kernel<<<kPartBlocks, 256, 0, st>>>(yhat, yd, encout, hw, (long long)f2, t->stats + 64);
  cvae_loss_final_kernel<<<1, 32, 0, st>>>(t->stats + 64, kPartBlocks, (double)f2, (double)batch, decoder_var, t->stats);
  cvae_dyhat_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(yhat, yd, t->stats, dec->d[0], (long long)f2);
  t->launches += 3;
  TR_TRY(t, cudaGetLastError());
  // decoder backward down to its input, encoder backward
  if ((rc = backward(dec, decin, batch, st, ddecin))) { enc->err = dec->err; return rc; }
  cvae_denc_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(ddecin, encout, ed, enc->d[0], 1.f / (float)batch, hw, (long long)f2);
  t->launches++;
  TR_TRY(t, cudaGetLastError());
  if ((rc = backward(enc, encin, batch, st))) return rc;
  if (update) {
    if ((rc = adam_update(enc, lr, st))) return rc;
    if ((rc = adam_update(dec, lr, st))) { enc->err = dec->err; return rc; }
  }
  if (losses) TR_TRY(t, cudaMemcpyAsync(losses, t->stats, 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TR_TRY(t, cudaStreamSynchronize(st));
  return QGB_OK;
}

}  // extern "C"

// =====================================================================================================================
// Discriminator and the WGAN-GP iteration (models/cgan_regression.py:173-195, 227-300; tools/cnn_tools.py:212-244)
// =====================================================================================================================
struct qgb_disc {
  int device = 0, nx = 0, B = 0, cin = 6, ndf = 64;
  struct Layer { int cin, cout, ks, H, OH; size_t w, K; };   // H: input side, OH: output side, K = ks^2 cin (GEMM depth)
  Layer L[5];
  size_t nparams = 0;
  float *P = nullptr, *G = nullptr, *M = nullptr, *V = nullptr, *Wp = nullptr;
  float* h[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};    // NHWC activations of up to 4B samples: input, 4 x LeakyReLU(conv)
  float* dl[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // gradient with respect to the output of layer k (before LeakyReLU)
  float* u[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};    // linearised pass of the gradient penalty (B samples)
  float *o = nullptr, *e0 = nullptr, *col = nullptr, *part = nullptr, *coef = nullptr, *eps = nullptr;
  float* colL[4] = {nullptr, nullptr, nullptr, nullptr};   // im2col matrices of the forward pass, kept per layer: the weight gradient
                                                           // reuses them (the linearised pass writes its rows over the interpolates')
  float* dyf[2] = {nullptr, nullptr};
  size_t col_floats = 0, part_floats = 0;
  double* stats = nullptr;
  long long adam_t = 0, launches = 0;
  float beta1 = 0.5f, beta2 = 0.999f, adam_eps = 1e-8f;
  std::string err;
  size_t act(int k) const { return k == 0 ? (size_t)nx * nx * cin : (size_t)L[k - 1].OH * L[k - 1].OH * L[k - 1].cout; }   // per sample
};

namespace {
std::string g_disc_create_error;
int dfail(qgb_disc* d, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (d) d->err = buf; else g_disc_create_error = buf;
  return code;
}
#define D_TRY(d, expr)                                                                                          \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return dfail(d, QGB_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// tensor-core path (tgemm.cuh: tcgen05 kind::tf32, 3-term split, fp32-accurate) unless QGB_DISC_GEMM=ffma asks for the FFMA kernel
inline bool disc_gemm_tc() {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("QGB_DISC_GEMM"); mode = (e && std::strcmp(e, "ffma") == 0) ? 0 : 1; }
  return mode == 1;
}

template <int EPI>
int gemm(qgb_disc* d, const float* A, long long sai, long long sak, const float* B, long long sbk, long long sbj, float* C,
         long long ldc, int M, int N, int K, int splits, long long c_split, const float* mask, cudaStream_t st) {
  if (disc_gemm_tc()) {
    int ksplit = (K + splits - 1) / splits;
    ksplit = (ksplit + tg::BK - 1) / tg::BK * tg::BK;
    splits = (K + ksplit - 1) / ksplit;
    const int a_vec = sak == 1 && sai % 4 == 0 && K % 4 == 0 && ((uintptr_t)A & 15) == 0;
    const int b_vec = sbk == 1 && sbj % 4 == 0 && K % 4 == 0 && ((uintptr_t)B & 15) == 0;
    D_TRY(d, cudaFuncSetAttribute(tg::tgemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::SMEM_BYTES));
    dim3 grid((N + tg::BN - 1) / tg::BN, (M + tg::BM - 1) / tg::BM, splits);
    tg::tgemm_kernel<EPI><<<grid, tg::kThreads, tg::SMEM_BYTES, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, ksplit, c_split, mask, a_vec, b_vec);
    d->launches++;
    D_TRY(d, cudaGetLastError());
    return splits;
  }
  int ksplit = (K + splits - 1) / splits;
  ksplit = (ksplit + kGemmK - 1) / kGemmK * kGemmK;
  splits = (K + ksplit - 1) / ksplit;
  if (N > 64) {
    dim3 grid((N + 127) / 128, (M + kGemmM - 1) / kGemmM, splits);
    sgemm_kernel<EPI, 128><<<grid, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, ksplit, c_split, mask);
  } else {
    dim3 grid((N + 63) / 64, (M + kGemmM - 1) / kGemmM, splits);
    sgemm_kernel<EPI, 64><<<grid, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, M, N, K, ksplit, c_split, mask);
  }
  d->launches++;
  D_TRY(d, cudaGetLastError());
  return splits;       // (>= 1; errors are negative QGB codes)
}

int disc_pack(qgb_disc* d, cudaStream_t st) {
  for (int k = 0; k < 5; ++k) {
    const auto& L = d->L[k];
    const long long n = (long long)L.cout * L.K;
    disc_pack_kernel<<<ew_blocks(n), 256, 0, st>>>(d->P + L.w, d->Wp + L.w, L.cout, L.cin, L.ks, 0, 1, 0);
    d->launches++;
  }
  D_TRY(d, cudaGetLastError());
  return QGB_OK;
}

// im2col of samples taken from src0 (the first b_split) and src1 (the rest); 16-byte accesses where the channel count allows
int im2col(qgb_disc* d, const float* src0, const float* src1, int b_split, float* dst, int nb, int H, int C, int OH, cudaStream_t st) {
  const long long tot = (long long)nb * OH * OH * 16 * C;
  if (C % 4 == 0) im2col_kernel<4><<<ew_blocks(tot / 4), 256, 0, st>>>(src0, src1, b_split, dst, nb, H, C, OH);
  else if (C % 2 == 0) im2col_kernel<2><<<ew_blocks(tot / 2), 256, 0, st>>>(src0, src1, b_split, dst, nb, H, C, OH);
  else im2col_kernel<1><<<ew_blocks(tot), 256, 0, st>>>(src0, src1, b_split, dst, nb, H, C, OH);
  d->launches++;
  D_TRY(d, cudaGetLastError());
  return QGB_OK;
}

// D on samples [b0, b0 + nb) of h[0]: activations into h[1..4], outputs into o
int disc_forward(qgb_disc* d, int b0, int nb, cudaStream_t st) {
  for (int k = 0; k < 4; ++k) {
    const auto& L = d->L[k];
    float* colk = d->colL[k] + (size_t)b0 * L.OH * L.OH * L.K;
    int rc = im2col(d, d->h[k] + b0 * d->act(k), nullptr, nb, colk, nb, L.H, L.cin, L.OH, st);
    if (rc) return rc;
    rc = gemm<1>(d, colk, (long long)L.K, 1, d->Wp + L.w, 1, (long long)L.K, d->h[k + 1] + b0 * d->act(k + 1), L.cout,
                     nb * L.OH * L.OH, L.cout, (int)L.K, 1, 0, nullptr, st);
    if (rc < 0) return rc;
  }
  const auto& L = d->L[4];
  rowdot_kernel<<<nb, 256, 0, st>>>(d->h[4] + b0 * d->act(4), d->Wp + L.w, d->o + b0, (int)L.K);
  d->launches++;
  D_TRY(d, cudaGetLastError());
  return QGB_OK;
}

// data gradients of samples [b0, b0 + nb) from dl[4] down to dl[0]; the gradient with respect to the input (e0, unmasked) for the
// samples [ib0, ib0 + inb) only
int disc_backward_data(qgb_disc* d, int b0, int nb, int ib0, int inb, cudaStream_t st) {
  {
    const auto& L = d->L[4];
    const long long tot = (long long)nb * L.K;
    disc_last_dgrad_kernel<<<ew_blocks(tot), 256, 0, st>>>(d->dl[4] + b0, d->Wp + L.w, d->h[4] + b0 * d->act(4), d->dl[3] + b0 * d->act(4),
                                                           (int)L.K, tot);
    d->launches++;
  }
  for (int k = 3; k >= 0; --k) {
    const auto& L = d->L[k];
    const int s0 = k == 0 ? ib0 : b0, sn = k == 0 ? inb : nb;
    if (sn <= 0) break;
    const int M = sn * L.OH * L.OH;
    int rc = gemm<0>(d, d->dl[k] + s0 * d->act(k + 1), L.cout, 1, d->Wp + L.w, (long long)L.K, 1, d->col, (long long)L.K, M, (int)L.K,
                     L.cout, 1, 0, nullptr, st);
    if (rc < 0) return rc;
    const long long tot = (long long)sn * d->act(k);
    float* dst = k > 0 ? d->dl[k - 1] + s0 * d->act(k) : d->e0;
    const float* msk = k > 0 ? d->h[k] + s0 * d->act(k) : nullptr;
    if (L.cin % 4 == 0) col2im_kernel<4><<<ew_blocks(tot / 4), 256, 0, st>>>(d->col, dst, msk, sn, L.H, L.cin, L.OH);
    else if (L.cin % 2 == 0) col2im_kernel<2><<<ew_blocks(tot / 2), 256, 0, st>>>(d->col, dst, msk, sn, L.H, L.cin, L.OH);
    else col2im_kernel<1><<<ew_blocks(tot), 256, 0, st>>>(d->col, dst, msk, sn, L.H, L.cin, L.OH);
    d->launches++;
  }
  D_TRY(d, cudaGetLastError());
  return QGB_OK;
}

// u[k + 1] = slope(h[k + 1]) (W_k * u[k]) for the nu samples whose activations start at sample m0 of h
int disc_linearised(qgb_disc* d, int m0, int nu, cudaStream_t st) {
  for (int k = 0; k < 4; ++k) {
    const auto& L = d->L[k];
    float* colk = d->colL[k] + (size_t)m0 * L.OH * L.OH * L.K;       // over the rows of the interpolates: their forward pass is done
    int rc = im2col(d, d->u[k], nullptr, nu, colk, nu, L.H, L.cin, L.OH, st);
    if (rc) return rc;
    rc = gemm<2>(d, colk, (long long)L.K, 1, d->Wp + L.w, 1, (long long)L.K, d->u[k + 1], L.cout, nu * L.OH * L.OH, L.cout,
                     (int)L.K, 1, 0, d->h[k + 1] + m0 * d->act(k + 1), st);
    if (rc < 0) return rc;
  }
  return QGB_OK;
}

// weight gradients into d->G: sum over the nA ordinary samples (inputs h[k]) and the nU linearised ones (inputs u[k]; their
// output gradients follow the ordinary ones in dl[k])
int disc_wgrad(qgb_disc* d, int nA, int nU, cudaStream_t st) {
  const int nb = nA + nU;
  for (int k = 0; k < 5; ++k) {
    const auto& L = d->L[k];
    const int OH2 = k < 4 ? L.OH * L.OH : 1;
    const int Mred = nb * OH2;
    const float* colk = d->col;
    if (k < 4) {
      colk = d->colL[k];                  // rows [0, nA OH^2): the forward pass; rows from nA OH^2: the linearised pass (disc_linearised(nA, nU))
    } else {
      D_TRY(d, cudaMemcpyAsync(d->col, d->h[4], (size_t)nA * L.K * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if (nU) D_TRY(d, cudaMemcpyAsync(d->col + (size_t)nA * L.K, d->u[4], (size_t)nU * L.K * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    const int bn = (disc_gemm_tc() || (int)L.K > 64) ? 128 : 64;
    const int tiles = ((L.cout + kGemmM - 1) / kGemmM) * (((int)L.K + bn - 1) / bn);
    int splits = (2 * 148 + tiles - 1) / tiles;
    const int kq = disc_gemm_tc() ? tg::BK : kGemmK;
    if (splits > (Mred + kq - 1) / kq) splits = (Mred + kq - 1) / kq;
    if (splits < 1) splits = 1;
    const size_t n = (size_t)L.cout * L.K;
    if (d->part_floats < n * splits) {
      if (d->part) cudaFree(d->part);
      d->part = nullptr; d->part_floats = 0;
      D_TRY(d, talloc(&d->part, n * splits));
      d->part_floats = n * splits;
    }
    int rc = gemm<0>(d, d->dl[k], 1, L.cout, colk, (long long)L.K, 1, d->part, (long long)L.K, L.cout, (int)L.K, Mred, splits,
                     (long long)n, nullptr, st);
    if (rc < 0) return rc;
    disc_pack_kernel<<<ew_blocks((long long)n), 256, 0, st>>>(d->part, d->G + L.w, L.cout, L.cin, L.ks, 1, rc, (long long)n);
    d->launches++;
  }
  D_TRY(d, cudaGetLastError());
  return QGB_OK;
}

int disc_adam(qgb_disc* d, double lr, cudaStream_t st) {
  d->adam_t += 1;
  const double bc1 = 1.0 - std::pow((double)d->beta1, (double)d->adam_t), bc2 = 1.0 - std::pow((double)d->beta2, (double)d->adam_t);
  adam_kernel<<<ew_blocks((long long)d->nparams), 256, 0, st>>>(d->P, d->G, d->M, d->V, (long long)d->nparams, (float)lr, d->beta1,
                                                                d->beta2, d->adam_eps, (float)bc1, (float)std::sqrt(bc2));
  d->launches++;
  D_TRY(d, cudaGetLastError());
  return QGB_OK;
}
}  // namespace

extern "C" {

int qgb_disc_create(int device, int in_channels, int ndf, int nx, int max_batch, qgb_disc** out) {
  if (!out || in_channels != 6 || ndf < 1 || max_batch < 1)
    return dfail(nullptr, QGB_EINVAL, "bad argument (the CGAN discriminator sees 6 channels: x, y1, y2)");
  *out = nullptr;
  if (nx < 16 || nx % 16 != 0) return dfail(nullptr, QGB_EUNSUPPORTED, "nx must be a multiple of 16 (four stride-2 layers, then an nx/16 kernel)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return dfail(nullptr, QGB_ECUDA, "no CUDA device available: libqgb200 has no CPU fallback");
  if (device < 0 || device >= ndev) return dfail(nullptr, QGB_EINVAL, "device %d out of range", device);
  qgb_disc* d = new qgb_disc();
  auto bail = [&](int rc) { g_disc_create_error = d->err; qgb_disc_destroy(d); return rc; };
#define CR(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { d->err = std::string(#expr) + ": " + cudaGetErrorString(_e); return bail(QGB_ECUDA); } } while (0)
  CR(cudaSetDevice(device));
  d->device = device; d->nx = nx; d->B = max_batch; d->cin = in_channels; d->ndf = ndf;
  int c = in_channels, H = nx;
  for (int k = 0; k < 5; ++k) {
    auto& L = d->L[k];
    L.cin = c; L.H = H;
    if (k < 4) { L.cout = ndf << k; L.ks = 4; L.OH = H / 2; } else { L.cout = 1; L.ks = H; L.OH = 1; }
    L.K = (size_t)L.ks * L.ks * L.cin;
    L.w = d->nparams; d->nparams += (size_t)L.cout * L.K;
    c = L.cout; H = L.OH;
  }
  const size_t B4 = 4 * (size_t)max_batch, B1 = max_batch;
  CR(talloc(&d->P, d->nparams)); CR(talloc(&d->G, d->nparams)); CR(talloc(&d->M, d->nparams)); CR(talloc(&d->V, d->nparams));
  CR(talloc(&d->Wp, d->nparams));
  for (int k = 0; k < 5; ++k) {
    CR(talloc(&d->h[k], B4 * d->act(k)));
    CR(talloc(&d->u[k], B1 * d->act(k)));
    CR(talloc(&d->dl[k], B4 * (k < 4 ? d->act(k + 1) : 1)));
    if (k < 4) {
      const size_t cf = B4 * d->L[k].OH * d->L[k].OH * d->L[k].K;
      d->col_floats = cf > d->col_floats ? cf : d->col_floats;
      CR(talloc(&d->colL[k], cf));
    }
  }
  if (d->col_floats < B4 * d->L[4].K) d->col_floats = B4 * d->L[4].K;
  CR(talloc(&d->col, d->col_floats));
  CR(talloc(&d->o, B4)); CR(talloc(&d->e0, B1 * d->act(0))); CR(talloc(&d->coef, B1)); CR(talloc(&d->eps, B1));
  CR(talloc(&d->dyf[0], B1 * 2 * nx * nx)); CR(talloc(&d->dyf[1], B1 * 2 * nx * nx));
  CR(talloc(&d->stats, (size_t)16 + B1));
  CR(cudaMemset(d->P, 0, d->nparams * sizeof(float))); CR(cudaMemset(d->G, 0, d->nparams * sizeof(float)));
  CR(cudaMemset(d->M, 0, d->nparams * sizeof(float))); CR(cudaMemset(d->V, 0, d->nparams * sizeof(float)));
  CR(cudaMemset(d->stats, 0, (16 + B1) * sizeof(double)));
  CR(cudaDeviceSynchronize());
#undef CR
  *out = d;
  return QGB_OK;
}

void qgb_disc_destroy(qgb_disc* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  for (float* p : {d->P, d->G, d->M, d->V, d->Wp, d->o, d->e0, d->col, d->part, d->coef, d->eps, d->dyf[0], d->dyf[1]})
    if (p) cudaFree(p);
  for (int k = 0; k < 5; ++k) { if (d->h[k]) cudaFree(d->h[k]); if (d->u[k]) cudaFree(d->u[k]); if (d->dl[k]) cudaFree(d->dl[k]); }
  for (int k = 0; k < 4; ++k) if (d->colL[k]) cudaFree(d->colL[k]);
  if (d->stats) cudaFree(d->stats);
  delete d;
}

const char* qgb_disc_last_error(const qgb_disc* d) { return d ? d->err.c_str() : g_disc_create_error.c_str(); }
int64_t qgb_disc_num_params(const qgb_disc* d) { return d ? (int64_t)d->nparams : 0; }
int64_t qgb_disc_launch_count(const qgb_disc* d) { return d ? (int64_t)d->launches : 0; }

int qgb_disc_set_params(qgb_disc* d, const float* params, int reset_optimizer) {
  if (!d) return QGB_EINVAL;
  D_TRY(d, cudaSetDevice(d->device));
  if (params) D_TRY(d, cudaMemcpy(d->P, params, d->nparams * sizeof(float), cudaMemcpyHostToDevice));
  if (reset_optimizer) {
    D_TRY(d, cudaMemset(d->M, 0, d->nparams * sizeof(float)));
    D_TRY(d, cudaMemset(d->V, 0, d->nparams * sizeof(float)));
    d->adam_t = 0;
  }
  return QGB_OK;
}

int qgb_disc_get_params(qgb_disc* d, float* params, float* grads) {
  if (!d) return QGB_EINVAL;
  D_TRY(d, cudaSetDevice(d->device));
  D_TRY(d, cudaDeviceSynchronize());
  if (params) D_TRY(d, cudaMemcpy(params, d->P, d->nparams * sizeof(float), cudaMemcpyDeviceToHost));
  if (grads) D_TRY(d, cudaMemcpy(grads, d->G, d->nparams * sizeof(float), cudaMemcpyDeviceToHost));
  return QGB_OK;
}

int qgb_disc_forward(qgb_disc* d, const float* x, int batch, int on_device, float* out, void* stream) {
  if (!d || !x || !out) return QGB_EINVAL;
  if (batch < 1 || batch > 4 * d->B) return dfail(d, QGB_EINVAL, "batch %d outside 1..%d", batch, 4 * d->B);
  cudaStream_t st = (cudaStream_t)stream;
  D_TRY(d, cudaSetDevice(d->device));
  const size_t n = (size_t)batch * d->act(0);
  const float* xd = x;
  if (!on_device) {                                   // staged through the im2col scratch (>= 16 x the input)
    D_TRY(d, cudaMemcpyAsync(d->col, x, n * sizeof(float), cudaMemcpyHostToDevice, st));
    xd = d->col;
  }
  nchw_to_nhwc_kernel<<<ew_blocks((long long)n), 256, 0, st>>>(xd, d->h[0], d->cin, d->nx * d->nx, (long long)n);
  d->launches++;
  int rc;
  if ((rc = disc_pack(d, st)) || (rc = disc_forward(d, 0, batch, st))) return rc;
  D_TRY(d, cudaMemcpyAsync(out, d->o, batch * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  D_TRY(d, cudaStreamSynchronize(st));
  return QGB_OK;
}

int qgb_train_cgan_step(qgb_trainer* Gt, qgb_disc* D, const float* x, const float* y, const float* z1, const float* z2,
                        const float* eps, int coin, int batch, int on_device, double lr_d, double lr_g, int update_d, int g_mode,
                        double* losses, void* stream) {
  if (!Gt || !D) return QGB_EINVAL;
  qgb_trainer* t = Gt;
  if (!x || !y || !z1 || !z2 || !eps) return tfail(t, QGB_EINVAL, "null minibatch");
  if (Gt->L.front().cin != 4 || Gt->L.back().cout != 2 || Gt->softplus) return tfail(t, QGB_EINVAL, "expected a generator 4 -> 2 ([x, z] -> y)");
  if (Gt->ny != D->nx || Gt->nx != D->nx || Gt->device != D->device) return tfail(t, QGB_EINVAL, "generator and discriminator differ in grid or device");
  if (batch < 1 || batch > Gt->max_batch || batch > D->B) return tfail(t, QGB_EINVAL, "batch %d outside 1..max_batch", batch);
  if (g_mode < 0 || g_mode > 2) return tfail(t, QGB_EINVAL, "g_mode must be 0 (skip), 1 (gradients) or 2 (gradients + Adam)");
  cudaStream_t st = (cudaStream_t)stream;
  TR_TRY(t, cudaSetDevice(t->device));
  const int hw = t->ny * t->nx, B = batch;
  const size_t f2 = (size_t)B * 2 * hw, f4 = (size_t)B * 4 * hw;
  const double lambda_gp = 10.0, lambda_drift = 1e-3;            // LAMBDA_GP, LAMBDA_DRIFT (cgan_regression.py:18-19)
  const float *xd, *yd, *z1d, *z2d;
  int rc;
  if ((rc = stage_into(t, 0, x, f2, on_device, &xd, st)) || (rc = stage_into(t, 1, y, f2, on_device, &yd, st)) ||
      (rc = stage_into(t, 2, z1, f2, on_device, &z1d, st)) || (rc = stage_into(t, 3, z2, f2, on_device, &z2d, st)))
    return rc;
  if ((rc = ensure_scratch(t, 4, f4)) || (rc = ensure_scratch(t, 5, f4)) || (rc = ensure_slots(t, 2))) return rc;
  if (!t->Gacc) TR_TRY(t, talloc(&t->Gacc, t->nparams));
  TR_TRY(t, cudaMemcpyAsync(D->eps, eps, B * sizeof(float), cudaMemcpyHostToDevice, st));     // eps: always a host array (B floats)
  float* gin[2] = {t->scr[4], t->scr[5]};
  const float* zz[2] = {z1d, z2d};
  const float* yf[2];
  // yfake1 = G(x, z1), yfake2 = G(x, z2) in training mode (:262-263)
  for (int s = 0; s < 2; ++s) {
    t->slot = s;
    cat_channels_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(xd, 2, gin[s], 4, 0, hw, (long long)f2);
    cat_channels_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(zz[s], 2, gin[s], 4, 2, hw, (long long)f2);
    t->launches += 2;
    if ((rc = forward(t, gin[s], B, true, st))) { t->slot = 0; return rc; }
    yf[s] = R(t, t->nlayers - 1);
  }
  t->slot = 0;
#define DCALL(expr) do { int _rc = (expr); if (_rc) { t->err = D->err; return _rc; } } while (0)
  // discriminator inputs: true1, true2, fake, interpolates (:266-268, 173-185)
  for (int mode = 0; mode < 4; ++mode) {
    disc_input_kernel<<<ew_blocks((long long)B * hw), 256, 0, st>>>(xd, yd, yf[0], yf[1], D->eps, coin, mode,
                                                                    D->h[0] + (size_t)mode * B * D->act(0), B, hw);
    D->launches++;
  }
  DCALL(disc_pack(D, st));
  DCALL(disc_forward(D, 0, 4 * B, st));
  // the four blocks of d5 must be B apart: the loss kernel writes them at stride ``batch``
  disc_loss_kernel<<<1, 256, 0, st>>>(D->o, B, lambda_drift, D->dl[4], D->stats);
  D->launches++;
  DCALL(disc_backward_data(D, 0, 4 * B, 3 * B, B, st));
  gp_norm_kernel<<<B, 256, 0, st>>>(D->e0, hw, B, lambda_gp, D->stats + 16, D->coef);
  gp_seed_kernel<<<ew_blocks((long long)B * hw * 6), 256, 0, st>>>(D->e0, D->coef, D->stats + 16, B, lambda_gp, D->u[0], hw, D->stats);
  D->launches += 2;
  TR_TRY(t, cudaGetLastError());
  DCALL(disc_linearised(D, 3 * B, B, st));
  DCALL(disc_wgrad(D, 3 * B, B, st));
  if (update_d) DCALL(disc_adam(D, lr_d, st));
  if (g_mode) {
    // G_loss = -mean D(x, yfake1, yfake2) with the updated discriminator (:277-282); the fake block of h[0] is still in place
    if (update_d) DCALL(disc_pack(D, st));
    TR_TRY(t, cudaMemcpyAsync(D->h[0], D->h[0] + (size_t)2 * B * D->act(0), (size_t)B * D->act(0) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DCALL(disc_forward(D, 0, B, st));
    gen_loss_kernel<<<1, 32, 0, st>>>(D->o, B, D->dl[4], D->stats);
    D->launches++;
    DCALL(disc_backward_data(D, 0, B, 0, B, st));
    nhwc_extract_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(D->e0, D->dyf[0], 2, hw, (long long)f2);
    nhwc_extract_kernel<<<ew_blocks((long long)f2), 256, 0, st>>>(D->e0, D->dyf[1], 4, hw, (long long)f2);
    D->launches += 2;
    for (int s = 0; s < 2; ++s) {
      t->slot = s;
      TR_TRY(t, cudaMemcpyAsync(t->d[0], D->dyf[s], f2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
      if ((rc = backward(t, gin[s], B, st))) { t->slot = 0; return rc; }
      if (s == 0) TR_TRY(t, cudaMemcpyAsync(t->Gacc, t->G, t->nparams * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    t->slot = 0;
    axpy_kernel<<<ew_blocks((long long)t->nparams), 256, 0, st>>>(t->G, t->Gacc, 1.f, (long long)t->nparams);
    t->launches++;
    TR_TRY(t, cudaGetLastError());
    if (g_mode == 2 && (rc = adam_update(t, lr_g, st))) return rc;
  }
#undef DCALL
  if (losses) TR_TRY(t, cudaMemcpyAsync(losses, D->stats, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  TR_TRY(t, cudaStreamSynchronize(st));
  return QGB_OK;
}

}  // extern "C"
