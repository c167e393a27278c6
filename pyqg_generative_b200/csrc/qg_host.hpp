// qg_host.hpp -- host-side construction of the member-independent tables (grid, inversion matrix, filter,
// FFT plan).  Mirrors pyqg 0.7.2 Model._initialize_grid / _initialize_filter and
// QGModel._initialize_background / _initialize_inversion_matrix.
#pragma once
#include <cmath>
#include <vector>

#include "../../include/qgb200.h"
#include "qg_core.cuh"

namespace qgb {

struct HostTables {
  int N = 0, NK = 0, P = 0, nstages = 0;
  int radix[kMaxStages] = {0};
  std::vector<cplx> tw;
  std::vector<short> pos;
  std::vector<double> kv, lv, a, filtr;
  double Ubg[2], Qy[2], rek, inv_M, Hi_over_H[2], dx, F1, F2;
};

inline bool build_host_tables(const qgb_config& cfg, HostTables& t) {
  const int N = cfg.nx;
  if (N < 4 || (N & 1)) return false;
  if (!make_radix_plan(N, t.radix, &t.nstages)) return false;
  t.N = N;
  t.NK = N / 2 + 1;
  t.P = N + 1;
  const double pi = 3.14159265358979323846;
  t.tw.resize(N);
  for (int i = 0; i < N; ++i) {
    // exact-ish twiddles: reduce the angle to the first octant before calling cos/sin
    const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)i / (long double)N;
    t.tw[i] = cmake((double)cosl(ang), (double)sinl(ang));
  }
  t.pos.resize(N);
  for (int f = 0; f < N; ++f) t.pos[f] = (short)pos_of_freq(N, t.radix, t.nstages, f);
  const double dk = 2.0 * pi / cfg.L;
  t.kv.resize(t.NK);
  t.lv.resize(N);
  for (int k = 0; k < t.NK; ++k) t.kv[k] = dk * (double)k;
  for (int l = 0; l < N; ++l) t.lv[l] = dk * (double)(l < N / 2 ? l : l - N);
  t.dx = cfg.L / N;
  // background (QGModel._initialize_background)
  t.F1 = 1.0 / (cfg.rd * cfg.rd) / (1.0 + cfg.delta);
  t.F2 = cfg.delta * t.F1;
  t.Ubg[0] = cfg.U1;
  t.Ubg[1] = cfg.U2;
  t.Qy[0] = cfg.beta + t.F1 * (cfg.U1 - cfg.U2);
  t.Qy[1] = cfg.beta - t.F2 * (cfg.U1 - cfg.U2);
  const double H1 = cfg.H1, H2 = cfg.H1 / cfg.delta;
  t.Hi_over_H[0] = H1 / (H1 + H2);
  t.Hi_over_H[1] = H2 / (H1 + H2);
  t.rek = cfg.rek;
  t.inv_M = 1.0 / ((double)N * (double)N);
  // inversion matrix + filter
  const int NN = N * t.NK;
  t.a.assign(4 * (size_t)NN, 0.0);
  t.filtr.assign(NN, 1.0);
  const double cphi = 0.65 * pi;
  for (int l = 0; l < N; ++l)
    for (int k = 0; k < t.NK; ++k) {
      const int i = l * t.NK + k;
      const double wv2 = t.kv[k] * t.kv[k] + t.lv[l] * t.lv[l];
      const double det = wv2 * (wv2 + t.F1 + t.F2);
      if (det != 0.0) {
        const double di = 1.0 / det;  // numpy: masked_equal(det,0)**-1 then multiply
        t.a[0 * (size_t)NN + i] = -(wv2 + t.F2) * di;
        t.a[1 * (size_t)NN + i] = -t.F1 * di;
        t.a[2 * (size_t)NN + i] = -t.F2 * di;
        t.a[3 * (size_t)NN + i] = -(wv2 + t.F1) * di;
      }
      const double kx = t.kv[k] * t.dx, ly = t.lv[l] * t.dx;
      const double wvx = std::sqrt(kx * kx + ly * ly);
      if (wvx > cphi) {
        const double d = wvx - cphi;
        t.filtr[i] = std::exp(-cfg.filterfac * (d * d * d * d));
      }
    }
  return true;
}

inline void fill_tables(const HostTables& h, Tables& T, const cplx* tw, const short* pos, const double* kv,
                        const double* lv, const double* a, const double* filtr) {
  T.N = h.N; T.NK = h.NK; T.P = h.P; T.nstages = h.nstages;
  for (int i = 0; i < kMaxStages; ++i) T.radix[i] = h.radix[i];
  T.tw = tw; T.pos = pos; T.kv = kv; T.lv = lv; T.a = a; T.filtr = filtr;
  T.Ubg[0] = h.Ubg[0]; T.Ubg[1] = h.Ubg[1];
  T.Qy[0] = h.Qy[0]; T.Qy[1] = h.Qy[1];
  T.rek = h.rek; T.inv_M = h.inv_M;
}

// Adams-Bashforth coefficients (pyqg kernel.pyx _forward_timestep): Euler, AB2, then AB3
inline void ab_coefficients(int ablevel, double dt, double& dt1, double& dt2, double& dt3) {
  if (ablevel == 0) { dt1 = dt; dt2 = 0.0; dt3 = 0.0; }
  else if (ablevel == 1) { dt1 = 1.5 * dt; dt2 = -0.5 * dt; dt3 = 0.0; }
  else { dt1 = 23.0 / 12.0 * dt; dt2 = -16.0 / 12.0 * dt; dt3 = 5.0 / 12.0 * dt; }
}

}  // namespace qgb
