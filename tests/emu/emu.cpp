// Host emulation of the per-member phase programs of pyqg_generative_b200/csrc/qg_core.cuh.
// TEST INFRASTRUCTURE ONLY: lets the CPU test-suite check the FFT / Hermitian-packing / AB3 index arithmetic of
// the CUDA kernels against the oracle without a GPU.  The product package never loads this library.
#include <cstring>
#include <vector>

#include "../../pyqg_generative_b200/csrc/qg_host.hpp"

using namespace qgb;

extern "C" int qgbemu_run(const qgb_config* cfg, int prog, int nthreads, double* qh, double* q, double* d_cur,
                          const double* d_p, const double* d_pp, const double* dq, float* cnn_x, float xstd0,
                          float xstd1, int ablevel, double* ph_out, double* u_out, double* v_out, double* p_out,
                          double* red_out, double* bud_out, double* bud_scr, int bud_demean) {
  HostTables h;
  if (!build_host_tables(*cfg, h)) return -1;
  Tables T;
  fill_tables(h, T, h.tw.data(), h.pos.data(), h.kv.data(), h.lv.data(), h.a.data(), h.filtr.data());
  StepIO io;
  std::memset(&io, 0, sizeof(io));
  io.qh = (cplx*)qh; io.q = q; io.d_cur = (cplx*)d_cur; io.d_p = (const cplx*)d_p; io.d_pp = (const cplx*)d_pp;
  io.dq = dq; io.cnn_x = cnn_x; io.cnn_mstride = 2LL * h.N * h.N;
  io.x_std[0] = xstd0; io.x_std[1] = xstd1;
  ab_coefficients(ablevel, cfg->dt, io.dt1, io.dt2, io.dt3);
  io.ph_out = (cplx*)ph_out; io.u_out = u_out; io.v_out = v_out; io.p_out = p_out; io.red_out = red_out;
  io.Hi_over_H[0] = h.Hi_over_H[0]; io.Hi_over_H[1] = h.Hi_over_H[1];
  io.bud_out = bud_out; io.bud_scr = bud_scr;
  std::vector<cplx> bud_tend(bud_out ? (size_t)cfg->members * 2 * h.N * h.NK : 0);
  io.bud_tend = bud_tend.data(); io.bud_inv_dt = 1.0 / cfg->dt; io.bud_demean = bud_demean;
  io.bud_F = h.Hi_over_H[0] * h.Hi_over_H[1] / (cfg->rd * cfg->rd); io.bud_U = cfg->U1 - cfg->U2;
  std::vector<cplx> buf((size_t)h.N * h.P), tw(h.N);
  std::vector<short> pos(h.N);
  std::vector<double> red(4 * (size_t)nthreads);
  for (int m = 0; m < cfg->members; ++m) {
    Ctx c{T, io, buf.data(), tw.data(), pos.data(), red.data(), m};
    const int nph = run_program(c, prog, -1, 0, nthreads);
    for (int ph = 0; ph < nph; ++ph)
      for (int tid = 0; tid < nthreads; ++tid) run_program(c, prog, ph, tid, nthreads);
  }
  return 0;
}
