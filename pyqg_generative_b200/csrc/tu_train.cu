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
  const int tiles_y2 = (ny + kConvTileY2 - 1) / kConvTileY2;
  const bool small = cout <= 4;
  const int co_t = small ? 2 : 32, cpad = co_pad_of(cout);
  dim3 grid(tiles_x * ((small || ks != 5) ? tiles_y : tiles_y2), (cout + co_t - 1) / co_t, batch);
  const long long ibs = (long long)cin * ny * nx, obs = (long long)cout * ny * nx;
#define QGB_TCONV(KS, CT) conv_ffma_kernel<KS, CT><<<grid, 256, 0, st>>>(in, ibs, out, obs, wp, bias, s, sh, cin, cout, cpad, ny, nx, tiles_x, relu_affine, 0, 0)
#define QGB_TCONV2(KS, CT) conv_ffma2_kernel<KS, CT><<<grid, 256, 0, st>>>(in, ibs, out, obs, wp, bias, s, sh, cin, cout, cpad, ny, nx, tiles_x, relu_affine, 0, 0)
  if (ks == 5 && !small) QGB_TCONV2(5, 32);
  else if (ks == 5) QGB_TCONV(5, 2);
  else if (ks == 3 && !small) QGB_TCONV(3, 32);
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

template <int KS, int CI_T, int CO_PER>
int wgrad_launch(qgb_trainer* t, const float* a, const float* dz, float* dW, int cin, int cout, int batch, cudaStream_t st) {
  using G = WgGeom<KS>;
  constexpr int CO_B = (256 / CI_T) * CO_PER;
  const size_t smem = (size_t)(CI_T * G::CI_STRIDE + CO_B * G::D_PITCH) * sizeof(float);
  auto kern = wgrad_ffma_kernel<KS, CI_T, CO_PER>;
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
    if (thin_in) return wgrad_launch<5, 4, 2>(t, a, dz, dW, cin, cout, batch, st);
    if (thin_out) return wgrad_launch<5, 32, 1>(t, a, dz, dW, cin, cout, batch, st);
    // 4 output channels per thread, one block per SM: 4.3 ms for the 128 -> 64 layer (64 images, 64^2); the 2-channel variant with two
    // blocks per SM measured 6.2 ms
    return wgrad_launch<5, 32, 4>(t, a, dz, dW, cin, cout, batch, st);
  }
  if (ks == 3) {
    if (thin_in) return wgrad_launch<3, 4, 2>(t, a, dz, dW, cin, cout, batch, st);
    if (thin_out) return wgrad_launch<3, 32, 1>(t, a, dz, dW, cin, cout, batch, st);
    return wgrad_launch<3, 32, 4>(t, a, dz, dW, cin, cout, batch, st);
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
    int rc = chan_reduce<2>(t, dz, nullptr, nullptr, nullptr, batch, L.cout, st);
    if (rc) return rc;
    chan_final_kernel<<<(L.cout + 127) / 128, 128, 0, st>>>(t->red_part, L.cout, t->G + L.b, nullptr);
    t->launches++;
    TR_TRY(t, cudaGetLastError());
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
    const long long total = (long long)batch * Lp.cout * hw;
    bn_relu_bwd_kernel<<<ew_blocks(total), 256, 0, st>>>(da, R(t, l - 1), t->P + Lp.g, MEAN(t, Lp), INVSTD(t, Lp),
                                                         t->G + Lp.g, t->G + Lp.be, Lp.cout, hw, 1.f / ((float)batch * hw), total);
    t->launches += 2;
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
