// operators.cuh -- pointwise kernels of the coarse-graining path (pyqg_generative/tools/operators.py).
// The FFTs, inversions and Jacobians are the phase programs of qg_core.cuh (PROG_SET_Q = rfft2 of a field pair,
// PROG_C2R = irfft2, PROG_INVERT, PROG_ADVECT); this file adds the spectral truncation + filter between the fine and the
// coarse grid and small complex-array helpers.  Host sequencing lives in api.cu (qgb_operator, qgb_subgrid_forcing).
#pragma once
#include <cuda_runtime.h>

#include "qg_core.cuh"

namespace qgb {

// cut_off (operators.py:117-132) followed by the operator's filter, in spectral space:
//   out[b][z][lc][kc] = in[b][z][lf][kc] / ratio^2 * filt(lc,kc),   lf = lc (lc < n) or lc + N - nc (lc >= n), n = nc/2
//   with trunc[n,0] = 0 and trunc[:,n] = 0 (FILTER_2h_HARMONICS, :126-130)
// op 1: pyqg exponential filter of a DEFAULT coarse model (model_filter :92-99 builds pyqg.QGModel(nx=nc): filterfac 23.6)
// op 2: gauss_filter(X, nc//2) :84-90  -> exp(-wv^2 (2 dx_c)^2 / 24);   op 5: no filter
// op 4: Operator4 = model_filter(Operator2(X, nc)) :213-214 -> the product of the two filters
__global__ void trunc_filter_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int fields, int N, int nc, int op,
                                    double L, double sign) {
  const int NK = N / 2 + 1, nkc = nc / 2 + 1, n = nc / 2;
  const long long total = (long long)fields * nc * nkc;
  const double pi = 3.14159265358979323846;
  const double dk = 2.0 * pi / L, dxc = L / nc, r2 = ((double)N / nc) * ((double)N / nc);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % nkc), lc = (int)((i / nkc) % nc);
    const long long f = i / ((long long)nkc * nc);
    cplx v = cmake(0.0, 0.0);
    if (!(kc == n) && !(lc == n && kc == 0)) {
      const int lf = lc < n ? lc : lc + N - nc;
      v = in[(f * N + lf) * NK + kc];
      const double kk = dk * kc, ll = dk * (lc < n ? lc : lc - nc);
      double filt = 1.0;
      if (op == 2 || op == 4) filt = exp(-(kk * kk + ll * ll) * (2.0 * dxc) * (2.0 * dxc) / 24.0);
      if (op == 1 || op == 4) {
        const double wvx = sqrt((kk * dxc) * (kk * dxc) + (ll * dxc) * (ll * dxc));
        if (wvx > 0.65 * pi) { const double d = wvx - 0.65 * pi; filt *= exp(-23.6 * d * d * d * d); }
      }
      const double s = sign * filt / r2;
      v = cmake(v.x * s, v.y * s);
    }
    out[i] = v;
  }
}

// fft_interpolate (operators.py:134-190) in spectral space between half-plane arrays of two grid sizes (either direction):
//   out[f][lo][ko] = scale * in[f][ls][ko]  for the nn = min(n_src,n_dst)/2 lowest wavenumbers of each sign,
//   zero elsewhere, with the 2h harmonics removed (truncate_2h): column ko = nn and the entry (l = -nn, k = 0).
__global__ void resample_kernel(const cplx* __restrict__ in, cplx* __restrict__ out, int fields, int n_src, int n_dst,
                                double scale) {
  const int nk_src = n_src / 2 + 1, nk_dst = n_dst / 2 + 1, nn = (n_src < n_dst ? n_src : n_dst) / 2;
  const long long total = (long long)fields * n_dst * nk_dst;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ko = (int)(i % nk_dst), lo = (int)((i / nk_dst) % n_dst);
    const long long f = i / ((long long)nk_dst * n_dst);
    cplx v = cmake(0.0, 0.0);
    if (ko < nn) {
      int ls = -1;
      if (lo < nn) ls = lo;
      else if (lo >= n_dst - nn) ls = lo - n_dst + n_src;
      if (ls >= 0 && !(ko == 0 && lo == n_dst - nn)) {      // (l = -nn, k = 0) has no phase: removed
        const cplx a = in[(f * n_src + ls) * nk_src + ko];
        v = cmake(a.x * scale, a.y * scale);
      }
    }
    out[i] = v;
  }
}

// in-place pyqg exponential filter with an arbitrary filterfac on half-plane arrays of an n-grid (pyqg _initialize_filter:
// filtr = exp(-filterfac (wv dx - 0.65 pi)^4) above the cut-off, 1 below).  filterfac = 1e20 is the sharp cut-off of
// advect(..., '2/3-rule') (operators.py:253-257: pyqg.QGModel(nx, filterfac=1e+20)).
__global__ void spectral_filter_kernel(cplx* __restrict__ a, int fields, int n, double L, double filterfac) {
  const int nk = n / 2 + 1;
  const double pi = 3.14159265358979323846;
  const double dk = 2.0 * pi / L, dx = L / n;
  const long long total = (long long)fields * n * nk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % nk), l = (int)((i / nk) % n);
    const double kk = dk * k, ll = dk * (l < n / 2 ? l : l - n);
    const double wvx = sqrt((kk * dx) * (kk * dx) + (ll * dx) * (ll * dx));
    if (wvx > 0.65 * pi) {
      const double d = wvx - 0.65 * pi, f = exp(-filterfac * d * d * d * d);
      a[i] = cmake(a[i].x * f, a[i].y * f);
    }
  }
}

// out[b] = a[b] * c[b] elementwise on real arrays (products on the 3/2 grid)
__global__ void rmul_kernel(const double* __restrict__ a, const double* __restrict__ c, double* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] * c[i];
}

// spectral divergence (operators.py:241-247): out[f][l][k] = i k fx[f][l][k] + i l fy[f][l][k] on a half-plane grid of size n
__global__ void spectral_div_kernel(const cplx* __restrict__ fx, const cplx* __restrict__ fy, cplx* __restrict__ out,
                                    int fields, int n, double L) {
  const int nk = n / 2 + 1;
  const double dk = 2.0 * 3.14159265358979323846 / L;
  const long long total = (long long)fields * n * nk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % nk), l = (int)((i / nk) % n);
    const double kv = dk * k, lv = dk * (l < n / 2 ? l : l - n);
    const cplx a = fx[i], b = fy[i];
    out[i] = cmake(-(kv * a.y + lv * b.y), kv * a.x + lv * b.x);
  }
}

// out = sa*a + sb*b on complex arrays (forcing_h = adv_coarse_h - op(adv_fine)_h)
__global__ void caxpby_kernel(const cplx* __restrict__ a, const cplx* __restrict__ b, cplx* __restrict__ out, long long n,
                              double sa, double sb) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = cmake(sa * a[i].x + sb * b[i].x, sa * a[i].y + sb * b[i].y);
}

}  // namespace qgb
