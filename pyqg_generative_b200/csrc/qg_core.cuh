// qg_core.cuh -- per-member spectral time step of the two-layer QG model, written as a sequence of
// barrier-separated PHASES over one member's fields held in shared memory.
//
// The same phase functions are compiled twice:
//   * by nvcc into the sm_100a kernels of spectral.cu (one CTA per ensemble member; phases separated by
//     __syncthreads()), which is the product path;
//   * by g++ into tests/emu (a thread-by-thread host emulation used ONLY by the CPU test-suite to check the
//     index arithmetic against the oracle before any GPU time is spent).
//
// What is computed (pyqg 0.7.2 PseudoSpectralKernel; call sites in the reference:
// pyqg_generative/tools/simulate.py:132,137,168 and tools/operators.py:229-234,303-305,323-326):
//   _invert                ph = A qh ; uh = -il ph ; vh = ik ph ; u,v = irfft2
//   _do_advection          uq=(u+Ubg)q ; vq=v q ; dqhdt = -(ik uqh + il vqh + ikQy ph)
//   _do_friction           dqhdt[1] += rek wv2 ph[1]
//   _do_q_subgrid_param.   dqhdt += rfft2(dq)          (dq demeaned == zeroing its (0,0) mode,
//                                                        models/parameterization.py:25)
//   _forward_timestep      qh = filtr (qh + dt1 dqhdt + dt2 dqhdt_p + dt3 dqhdt_pp) ; q = irfft2(qh)
//
// FFT design: two REAL fields are packed as one COMPLEX N x N transform (u+iv, uq+ivq, q1+iq2, dq1+idq2), so a
// step costs 6 complex 2-D FFTs.  Transforms are in-place mixed-radix (4,2,3) decimation-in-frequency forward
// (natural in -> digit-reversed out) and decimation-in-time inverse (digit-reversed in -> natural out); spectral
// data therefore lives in shared memory at permuted positions (pos_of_freq) and is never reordered.  The row pitch
// is N+1 complex numbers so both row and column passes are bank-conflict free for 16-byte accesses.
#pragma once
#include <math.h>
#include <stdint.h>
#if defined(__CUDACC__)
#include <type_traits>
#endif

#if defined(__CUDACC__)
#define QGB_HD __host__ __device__ __forceinline__
#else
#define QGB_HD inline
#endif

namespace qgb {

struct alignas(16) cplx {
  double x, y;
};

QGB_HD cplx cmake(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
QGB_HD cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
QGB_HD cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
QGB_HD cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QGB_HD cplx cmulc(cplx a, cplx b) { return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
QGB_HD cplx cconj(cplx a) { return cmake(a.x, -a.y); }
QGB_HD cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
QGB_HD cplx cmuli(cplx a, double c) { return cmake(-c * a.y, c * a.x); }  // (i c) * a

constexpr int kMaxStages = 8;

// Member-independent tables (device pointers in the kernels, host pointers in the emulation)
struct Tables {
  int N, NK, P;  // grid size, N/2+1, shared-memory row pitch (complex elements)
  int nstages;
  int radix[kMaxStages];
  const cplx* tw;          // [N]      exp(-2 pi i t / N)
  const short* pos;        // [N]      position of frequency index f after the forward transform
  const double* kv;        // [NK]     zonal wavenumbers  kk
  const double* lv;        // [N]      meridional wavenumbers ll (FFT order)
  const double* a;         // [4][N*NK] inversion matrix a00,a01,a10,a11 (pyqg _initialize_inversion_matrix)
  const double* filtr;     // [N*NK]   exponential filter (pyqg _initialize_filter)
  double Ubg[2], Qy[2], rek, inv_M;
};

struct StepIO {
  cplx* qh;             // (B,2,N,NK)  in/out
  double* q;            // (B,2,N,N)   in/out
  cplx* d_cur;          // (B,2,N,NK)  tendency of this step (becomes dqhdt_p)
  const cplx* d_p;      // dqhdt_p
  const cplx* d_pp;     // dqhdt_pp
  const double* dq;     // (B,2,N,N) closure forcing or nullptr
  float* cnn_x;         // closure input: member stride cnn_mstride floats, channel stride N*N; or nullptr
  long long cnn_mstride;
  float x_std[2];
  float x_inv[2];       // 1 / x_std (rounded): the register-FFT kernels divide with one reciprocal multiply + one residual correction
  double dt1, dt2, dt3;
  // invert / diagnostics outputs (may be null)
  cplx* ph_out;
  double* u_out;
  double* v_out;
  double* p_out;
  double* red_out;      // (B,4): ke, max|u+U|, max|v|, nonfinite count
  double Hi_over_H[2];
  // spectral energy budget (PROG_BUDGET): per-member terms (B, kBudgetTerms, N, NK) and real scratch (B, 3, N, N)
  double* bud_out;
  double* bud_scr;
  double bud_F;         // rd^-2 del1 del2
  double bud_U;         // U1 - U2
  cplx* bud_tend;       // (B, 2, N, NK): tendency of the current state without the forcing (for the filter-dissipation spectra)
  double bud_inv_dt;    // 1 / dt
  int bud_demean;       // the forcing io.dq is a closure output: its mean is removed before it is used (models/parameterization.py:25)
};

// PROG_BUDGET terms (pyqg QGModel._initialize_model_diagnostics / Model._initialize_diagnostics; all / M^2)
enum BudgetTerm { BUD_KEFLUX = 0, BUD_APEFLUX = 1, BUD_APEGEN = 2, BUD_KEFRIC = 3, BUD_ENTSPEC = 4, BUD_PARAM_KE = 5,
                  BUD_PARAM_APE = 6,
                  // enstrophy budget and filter dissipation (Model._initialize_core_diagnostics: ENSflux, ENSgenspec,
                  // ENSfrictionspec, Dissspec, ENSDissspec, ENSparamspec; consumers tools/comparison_tools.py:222-225)
                  BUD_ENSFLUX = 7, BUD_ENSGEN = 8, BUD_ENSFRIC = 9, BUD_DISS = 10, BUD_ENSDISS = 11, BUD_ENSPARAM = 12 };
constexpr int kBudgetTerms = 13;

// Per-CTA context: shared-memory views + which member this CTA owns.  CN > 0 fixes the grid size at compile time
// (the specialised step kernels of spectral.cuh: all index divisions become shifts / multiplies); CN = 0 reads it from T.
template <int CN>
struct CtxT {
  static constexpr int kN = CN;
  const Tables& T;   // lives in kernel parameter (constant) space on the device
  const StepIO& io;
  cplx* buf;      // [N*P]
  cplx* tw;       // [N]   shared copy of the twiddles
  short* pos;     // [N]   shared copy of the position map
  double* red;    // [4*nthreads] reduction scratch (diagnostic program only)
  int member;
  // thread-block-cluster path (N >= 128): the field lives in global memory; a 1-D pass of the 2-D transform stages the
  // lines owned by this CTA through ``tile`` (shared memory, tile_lines x (N+1)) and runs all its stages locally
  cplx* tile = nullptr;
  int tile_lines = 0, ncta = 1;
  const short* tpos = nullptr;   // digit-reversal map of the transform stages (the global field is kept in NATURAL order: ``pos`` = identity)
  QGB_HD int N() const { return CN ? CN : T.N; }
  QGB_HD int NK() const { return CN ? CN / 2 + 1 : T.NK; }
  QGB_HD int P() const { return CN ? CN + 1 : T.P; }
};
using Ctx = CtxT<0>;

// ------------------------------------------------------------------------------------------------------
// 1-D FFT stage over ``nlines`` lines of the shared buffer.  es = element stride, ls = line stride.
// Forward stage (sub-FFT length n, radix r, m = n/r):  y[k1] = w_n^{j2 k1} * sum_j1 w_r^{j1 k1} x[j1 m + j2]
// stored in place at k1 m + j2.  The inverse stage is its exact adjoint (conjugate twiddles, then conj DFT_r).
// ------------------------------------------------------------------------------------------------------
template <int r>
QGB_HD void fft_stage_r(cplx* buf, const cplx* tw, int N, int es, int ls, int nlines, int n, bool inverse,
                        int tid, int nt) {
  const int m = n / r;
  const int nb = N / r;
  const int tws = N / n;
  const int items = nlines * nb;
  const int step = m * es;
  const double sgn = inverse ? 1.0 : -1.0;
  for (int i = tid; i < items; i += nt) {
    const int b = i / nlines;
    const int line = i - b * nlines;
    const int blk = b / m;
    const int j2 = b - blk * m;
    cplx* p = buf + line * ls + (blk * n + j2) * es;
    if (r == 4) {
      cplx x0 = p[0], x1 = p[step], x2 = p[2 * step], x3 = p[3 * step];
      if (m == 1) {                 // last forward / first inverse stage: all twiddles are 1 (folds away when n is constant)
        cplx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), d = csub(x1, x3);
        cplx t3 = cmake(-sgn * d.y, sgn * d.x);
        p[0] = cadd(t0, t2); p[step] = cadd(t1, t3); p[2 * step] = csub(t0, t2); p[3 * step] = csub(t1, t3);
        continue;
      }
      if (inverse) {
        x1 = cmulc(x1, tw[j2 * tws]);
        x2 = cmulc(x2, tw[2 * j2 * tws]);
        x3 = cmulc(x3, tw[3 * j2 * tws]);
      }
      cplx t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), d = csub(x1, x3);
      cplx t3 = cmake(-sgn * d.y, sgn * d.x);  // (sgn*i) * d   (forward: -i, inverse: +i)
      cplx y0 = cadd(t0, t2), y2 = csub(t0, t2), y1 = cadd(t1, t3), y3 = csub(t1, t3);
      if (!inverse) {
        y1 = cmul(y1, tw[j2 * tws]);
        y2 = cmul(y2, tw[2 * j2 * tws]);
        y3 = cmul(y3, tw[3 * j2 * tws]);
      }
      p[0] = y0; p[step] = y1; p[2 * step] = y2; p[3 * step] = y3;
    } else if (r == 2) {
      cplx x0 = p[0], x1 = p[step];
      if (inverse) x1 = cmulc(x1, tw[j2 * tws]);
      cplx y0 = cadd(x0, x1), y1 = csub(x0, x1);
      if (!inverse) y1 = cmul(y1, tw[j2 * tws]);
      p[0] = y0; p[step] = y1;
    } else {  // r == 3
      cplx x0 = p[0], x1 = p[step], x2 = p[2 * step];
      if (inverse) {
        x1 = cmulc(x1, tw[j2 * tws]);
        x2 = cmulc(x2, tw[2 * j2 * tws]);
      }
      const double h = 0.86602540378443864676;  // sqrt(3)/2
      cplx s = cadd(x1, x2), d = csub(x1, x2);
      cplx m1 = cmake(x0.x - 0.5 * s.x, x0.y - 0.5 * s.y);
      cplx m2 = cmake(-sgn * h * d.y, sgn * h * d.x);  // (sgn*i*h) * d
      cplx y0 = cadd(x0, s), y1 = cadd(m1, m2), y2 = csub(m1, m2);
      if (!inverse) {
        y1 = cmul(y1, tw[j2 * tws]);
        y2 = cmul(y2, tw[2 * j2 * tws]);
      }
      p[0] = y0; p[step] = y1; p[2 * step] = y2;
    }
  }
}

QGB_HD void fft_stage(cplx* buf, const cplx* tw, int N, int es, int ls, int nlines, int r, int n, bool inverse,
                      int tid, int nt) {
  if (r == 4) fft_stage_r<4>(buf, tw, N, es, ls, nlines, n, inverse, tid, nt);
  else if (r == 2) fft_stage_r<2>(buf, tw, N, es, ls, nlines, n, inverse, tid, nt);
  else fft_stage_r<3>(buf, tw, N, es, ls, nlines, n, inverse, tid, nt);
}

// number of barrier-separated phases of one 2-D transform
QGB_HD int fft2d_phases(const Tables& T) { return 2 * T.nstages; }

#ifdef __CUDACC__
// One whole 1-D pass (all radix stages) of the 2-D transform for a field in global memory, cluster path.  The lines of
// the pass are dealt to the CTAs of the cluster (rows for the x pass, columns for the y pass); each CTA copies its lines
// into shared memory (coalesced: contiguous rows, or 16-byte elements of adjacent columns), runs the stages with block
// barriers only and writes the lines back.  The digit reversal of the in-place stages is undone while the lines move
// between global and shared memory, so the global field is in natural order (``pos`` = identity) and the pointwise phases
// read and write it coalesced.  A 2-D transform is 2 cluster phases instead of 2 * nstages, and the stages
// run at shared-memory speed instead of one L2 round trip per butterfly.
// ``load(row, col)`` / ``store(row, col, v)`` replace the buffer read / write (fusing the pointwise phase before / after
// the pass); mode 0 = forward stages, 1 = inverse stages, 2 = inverse stages, ``mid(row, col, v)`` on every element,
// forward stages (a physical-space product between two transforms along the same direction).  TileNone = use the buffer.
struct TileNone {};
// compile-time radix plan of a power-of-two size (the order make_radix_plan produces: 4s first, then one 2)
__host__ __device__ constexpr int pow2_stages(int n) { int s = 0; while (n % 4 == 0 && n > 1) { n /= 4; ++s; } while (n % 2 == 0 && n > 1) { n /= 2; ++s; } return s; }
__host__ __device__ constexpr int pow2_radix(int n, int s) { for (int i = 0; i < s; ++i) n /= (n % 4 == 0 ? 4 : 2); return n % 4 == 0 ? 4 : 2; }
__host__ __device__ constexpr int pow2_len(int n, int s) { for (int i = 0; i < s; ++i) n /= (n % 4 == 0 ? 4 : 2); return n; }
template <int CN, int S, bool INV>
__device__ __forceinline__ void tile_stages_fixed(cplx* tile, const cplx* tw, int L, int lt, int ntc) {
  constexpr int NS = pow2_stages(CN);
  if constexpr (S < NS) {
    constexpr int s = INV ? NS - 1 - S : S;
    fft_stage_r<pow2_radix(CN, s)>(tile, tw, CN, 1, CN + 1, L, pow2_len(CN, s), INV, lt, ntc);
    __syncthreads();
    tile_stages_fixed<CN, S + 1, INV>(tile, tw, L, lt, ntc);
  }
}
template <class C, class Load, class Mid, class Store>
__device__ void tiled_pass(const C& c, int pass, int mode, const Load& load, const Mid& mid, const Store& store, int tid, int nt) {
  const Tables& T = c.T;
  const int N = c.N(), P = c.P(), TP = N + 1;
  const int ntc = nt / c.ncta, rank = tid / ntc, lt = tid - rank * ntc;
  const int per_cta = N / c.ncta;                 // lines owned by this CTA
  cplx* tile = c.tile;
  for (int l0 = 0; l0 < per_cta; l0 += c.tile_lines) {
    const int L = per_cta - l0 < c.tile_lines ? per_cta - l0 : c.tile_lines;
    const int first = rank * per_cta + l0;
    for (int i = lt; i < L * N; i += ntc) {
      int line, e;
      if (pass == 0) { line = i / N; e = i - line * N; } else { e = i / L; line = i - e * L; }
      const int row = pass == 0 ? first + line : e, col = pass == 0 ? e : first + line;
      const int te = mode != 0 ? c.tpos[e] : e;          // spectral input: frequency e sits at its digit-reversed slot
      if constexpr (std::is_same<Load, TileNone>::value) tile[line * TP + te] = c.buf[row * P + col];
      else tile[line * TP + te] = load(row, col);
    }
    __syncthreads();
    for (int part = 0; part < (mode == 2 ? 2 : 1); ++part) {
      const bool inverse = mode == 1 || (mode == 2 && part == 0);
      if (part == 1) {
        if constexpr (!std::is_same<Mid, TileNone>::value) {
          for (int i = lt; i < L * N; i += ntc) {
            int line, e;
            if (pass == 0) { line = i / N; e = i - line * N; } else { e = i / L; line = i - e * L; }
            const int row = pass == 0 ? first + line : e, col = pass == 0 ? e : first + line;
            tile[line * TP + e] = mid(row, col, tile[line * TP + e]);
          }
          __syncthreads();
        }
      }
      if constexpr (C::kN > 0) {      // compile-time size: the stage sequence is unrolled, index arithmetic folds to shifts
        if (inverse) tile_stages_fixed<C::kN, 0, true>(tile, c.tw, L, lt, ntc);
        else tile_stages_fixed<C::kN, 0, false>(tile, c.tw, L, lt, ntc);
      } else {
        for (int si = 0; si < T.nstages; ++si) {
          const int s = inverse ? T.nstages - 1 - si : si;
          int n = N;
          for (int j = 0; j < s; ++j) n /= T.radix[j];
          fft_stage(tile, c.tw, N, 1, TP, L, T.radix[s], n, inverse, lt, ntc);
          __syncthreads();
        }
      }
    }
    for (int i = lt; i < L * N; i += ntc) {
      int line, e;
      if (pass == 0) { line = i / N; e = i - line * N; } else { e = i / L; line = i - e * L; }
      const int row = pass == 0 ? first + line : e, col = pass == 0 ? e : first + line;
      const int te = mode != 1 ? c.tpos[e] : e;          // spectral output leaves in natural frequency order
      if constexpr (std::is_same<Store, TileNone>::value) c.buf[row * P + col] = tile[line * TP + te];
      else store(row, col, tile[line * TP + te]);
    }
    __syncthreads();
  }
}
template <class C>
__device__ void fft2d_pass_tiled(const C& c, int pass, bool inverse, int tid, int nt) {
  tiled_pass(c, pass, inverse ? 1 : 0, TileNone(), TileNone(), TileNone(), tid, nt);
}
#endif

// phase ``ph`` (0 .. 2*nstages-1) of the 2-D transform of the whole N x N buffer
template <class C>
QGB_HD void fft2d_phase(const C& c, int ph, bool inverse, int tid, int nt) {
  const Tables& T = c.T;
#ifdef __CUDA_ARCH__
  if (c.tile) { fft2d_pass_tiled(c, ph, inverse, tid, nt); return; }
#endif
  const int S = T.nstages;
  const int pass = ph / S;  // 0: along x (rows), 1: along y (columns)
  int s = ph - pass * S;
  if (inverse) s = S - 1 - s;
  int n = c.N();
  for (int i = 0; i < s; ++i) n /= T.radix[i];
  const int es = pass == 0 ? 1 : c.P();
  const int ls = pass == 0 ? c.P() : 1;
  fft_stage(c.buf, c.tw, c.N(), es, ls, c.N(), T.radix[s], n, inverse, tid, nt);
}

// ------------------------------------------------------------------------------------------------------
// pointwise phases
// ------------------------------------------------------------------------------------------------------
template <class C>
QGB_HD void ph_init(const C& c, int tid, int nt) {
  if (c.tw == c.T.tw) return;   // large-N path: tables are used in place (global memory), nothing to stage
  for (int i = tid; i < c.N(); i += nt) {
    c.tw[i] = c.T.tw[i];
    c.pos[i] = c.T.pos[i];
  }
}

// streamfunction of layer z at half-plane entry (l,k):  ph = a[z][0] qh0 + a[z][1] qh1   (pyqg _invert)
template <class C>
QGB_HD cplx half_ph(const C& c, const cplx* qh, int z, int idx) {
  const int NN = c.N() * c.NK();
  const cplx q0 = qh[idx], q1 = qh[NN + idx];
  const double a0 = c.T.a[(2 * z) * NN + idx], a1 = c.T.a[(2 * z + 1) * NN + idx];
  return cmake(a0 * q0.x + a1 * q1.x, a0 * q0.y + a1 * q1.y);
}

// uh = -il ph, vh = ik ph packed as uh + i vh at a half-plane entry
template <class C>
QGB_HD void half_uv(const C& c, const cplx* qh, int z, int l, int k, cplx& uh, cplx& vh) {
  const cplx ph = half_ph(c, qh, z, l * c.NK() + k);
  const double lv = c.T.lv[l], kv = c.T.kv[k];
  uh = cmake(lv * ph.y, -lv * ph.x);
  vh = cmake(-kv * ph.y, kv * ph.x);
}

// Fill the buffer with E(A') + i E(B') where A,B are two half-plane spectra given by ``get(l,k,A,B)``,
// ' symmetrises the self-conjugate columns k=0,N/2 (what a c2r transform does implicitly by dropping the
// imaginary part there) and E is the Hermitian extension to the full plane.
template <class C, class Get>
QGB_HD cplx packed_value(const C& c, const Get& get, int l, int k) {
  const int N = c.N(), H = N / 2;
  const int lm = l == 0 ? 0 : N - l;
  cplx A, B;
  if (k <= H) {
    get(l, k, A, B);
    if (k == 0 || k == H) {
      cplx A2, B2;
      get(lm, k, A2, B2);
      A = cmake(0.5 * (A.x + A2.x), 0.5 * (A.y - A2.y));
      B = cmake(0.5 * (B.x + B2.x), 0.5 * (B.y - B2.y));
    }
  } else {
    get(lm, N - k, A, B);
    A = cconj(A);
    B = cconj(B);
  }
  return cmake(A.x - B.y, A.y + B.x);
}
template <class C, class Get>
QGB_HD void build_packed(const C& c, Get get, int tid, int nt) {
  const int N = c.N(), P = c.P();
  for (int i = tid; i < N * N; i += nt) {
    const int l = i / N, k = i - l * N;
    c.buf[c.pos[l] * P + c.pos[k]] = packed_value(c, get, l, k);
  }
}

template <class C>
struct GetUV {
  const C& c; const cplx* qh; int z;
  QGB_HD void operator()(int l, int k, cplx& A, cplx& B) const { half_uv(c, qh, z, l, k, A, B); }
};
template <class C>
struct GetQ {
  const C& c; const cplx* qh;
  QGB_HD void operator()(int l, int k, cplx& A, cplx& B) const {
    const int idx = l * c.NK() + k;
    A = qh[idx];
    B = qh[c.N() * c.NK() + idx];
  }
};
template <class C>
struct GetP {
  const C& c; const cplx* qh;
  QGB_HD void operator()(int l, int k, cplx& A, cplx& B) const {
    const int idx = l * c.NK() + k;
    A = half_ph(c, qh, 0, idx);
    B = half_ph(c, qh, 1, idx);
  }
};

template <class C>
QGB_HD const cplx* member_qh(const C& c) { return c.io.qh + (long long)c.member * 2 * c.N() * c.NK(); }

template <class C>
QGB_HD void ph_build_uv(const C& c, int z, int tid, int nt) {
  GetUV<C> g{c, member_qh(c), z};
  build_packed(c, g, tid, nt);
}
template <class C>
QGB_HD void ph_build_q(const C& c, int tid, int nt) {
  GetQ<C> g{c, member_qh(c)};
  build_packed(c, g, tid, nt);
}
template <class C>
QGB_HD void ph_build_p(const C& c, int tid, int nt) {
  GetP<C> g{c, member_qh(c)};
  build_packed(c, g, tid, nt);
}

// buf = (u + Ubg) q + i v q        (pyqg _do_advection, physical-space products)
template <class C>
QGB_HD void ph_products(const C& c, int z, int tid, int nt) {
  const int N = c.N(), P = c.P();
  const double* q = c.io.q + ((long long)c.member * 2 + z) * N * N;
  const double s = c.T.inv_M, U = c.T.Ubg[z];
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    cplx w = c.buf[y * P + x];
    const double qq = q[i];
    c.buf[y * P + x] = cmake((w.x * s + U) * qq, (w.y * s) * qq);
  }
}

// split a packed forward transform W = FFT(a + i b) into the two half-plane spectra
template <class C>
QGB_HD void unpack_pair(const C& c, int l, int k, cplx& A, cplx& B) {
  const int N = c.N(), P = c.P();
  const int lm = l == 0 ? 0 : N - l, km = k == 0 ? 0 : N - k;
  const cplx w1 = c.buf[c.pos[l] * P + c.pos[k]];
  const cplx w2 = c.buf[c.pos[lm] * P + c.pos[km]];
  A = cmake(0.5 * (w1.x + w2.x), 0.5 * (w1.y - w2.y));
  B = cmake(0.5 * (w1.y + w2.y), -0.5 * (w1.x - w2.x));
}

// dqhdt_z = -(ik uqh + il vqh + ikQy ph) (+ rek wv2 ph for the bottom layer) -> d_cur
template <class C>
QGB_HD void ph_tendency(const C& c, int z, int tid, int nt) {
  const int N = c.N(), NK = c.NK();
  const cplx* qh = member_qh(c);
  cplx* d = c.io.d_cur + ((long long)c.member * 2 + z) * N * NK;
  for (int i = tid; i < N * NK; i += nt) {
    const int l = i / NK, k = i - l * NK;
    cplx uqh, vqh;
    unpack_pair(c, l, k, uqh, vqh);
    const cplx ph = half_ph(c, qh, z, i);
    const double kv = c.T.kv[k], lv = c.T.lv[l];
    const cplx t1 = cmuli(uqh, kv), t2 = cmuli(vqh, lv), t3 = cmuli(ph, kv * c.T.Qy[z]);
    cplx r = cmake(-(t1.x + t2.x + t3.x), -(t1.y + t2.y + t3.y));
    if (z == 1 && c.T.rek != 0.0) {
      const double f = c.T.rek * (kv * kv + lv * lv);
      r.x += f * ph.x;
      r.y += f * ph.y;
    }
    d[i] = r;
  }
}

// buf = dq1 + i dq2
template <class C>
QGB_HD void ph_load_pair(const C& c, const double* f, int tid, int nt) {
  const int N = c.N(), P = c.P();
  const double* f0 = f + (long long)c.member * 2 * N * N;
  const double* f1 = f0 + N * N;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    c.buf[y * P + x] = cmake(f0[i], f1[i]);
  }
}

// (+ rfft2(dq) with its mean removed) ; Adams-Bashforth update with the exponential filter
template <class C>
QGB_HD void ph_update(const C& c, bool has_dq, bool demean, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const long long mo = (long long)c.member * 2 * NN;
  cplx* qh = c.io.qh + mo;
  cplx* d = c.io.d_cur + mo;
  const cplx* dp = c.io.d_p + mo;
  const cplx* dpp = c.io.d_pp + mo;
  const double dt1 = c.io.dt1, dt2 = c.io.dt2, dt3 = c.io.dt3;
  for (int i = tid; i < NN; i += nt) {
    cplx f0 = cmake(0, 0), f1 = cmake(0, 0);
    if (has_dq && !(demean && i == 0)) {
      const int l = i / NK, k = i - l * NK;
      unpack_pair(c, l, k, f0, f1);
    }
    const double fl = c.T.filtr[i];
    for (int z = 0; z < 2; ++z) {
      const cplx f = z == 0 ? f0 : f1;
      const int j = z * NN + i;
      const cplx dd = cadd(d[j], f);
      d[j] = dd;
      const cplx q0 = qh[j], a = dp[j], b = dpp[j];
      qh[j] = cmake(fl * (q0.x + dt1 * dd.x + dt2 * a.x + dt3 * b.x), fl * (q0.y + dt1 * dd.y + dt2 * a.y + dt3 * b.y));
    }
  }
}

// q = irfft2(qh) (real/imag of the packed inverse), also emits the fp32 normalised closure input
template <class C>
QGB_HD void ph_emit_q(const C& c, int tid, int nt) {
  const int N = c.N(), P = c.P();
  double* q = c.io.q + (long long)c.member * 2 * N * N;
  const double s = c.T.inv_M;
  float* x = c.io.cnn_x ? c.io.cnn_x + (long long)c.member * c.io.cnn_mstride : nullptr;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, xx = i - y * N;
    const cplx w = c.buf[y * P + xx];
    const double q0 = w.x * s, q1 = w.y * s;
    q[i] = q0;
    q[N * N + i] = q1;
    if (x) {  // x_scale.normalize(m.q.astype('float32'))  (models/cgan_regression.py:158)
      x[i] = (float)q0 / c.io.x_std[0];
      x[N * N + i] = (float)q1 / c.io.x_std[1];
    }
  }
}

// closure input from the current q without touching the spectral state
template <class C>
QGB_HD void ph_emit_x_only(const C& c, int tid, int nt) {
  const int N = c.N();
  const double* q = c.io.q + (long long)c.member * 2 * N * N;
  float* x = c.io.cnn_x + (long long)c.member * c.io.cnn_mstride;
  for (int i = tid; i < 2 * N * N; i += nt) x[i] = (float)q[i] / c.io.x_std[i / (N * N)];
}

// qh = rfft2(q) from the packed forward transform (pyqg ``q`` setter)
template <class C>
QGB_HD void ph_store_qh(const C& c, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  cplx* qh = c.io.qh + (long long)c.member * 2 * NN;
  for (int i = tid; i < NN; i += nt) {
    const int l = i / NK, k = i - l * NK;
    cplx a, b;
    unpack_pair(c, l, k, a, b);
    qh[i] = a;
    qh[NN + i] = b;
  }
}

template <class C>
QGB_HD void ph_store_uv(const C& c, int z, int tid, int nt) {
  const int N = c.N(), P = c.P();
  const long long o = ((long long)c.member * 2 + z) * N * N;
  const double s = c.T.inv_M;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    const cplx w = c.buf[y * P + x];
    if (c.io.u_out) c.io.u_out[o + i] = w.x * s;
    if (c.io.v_out) c.io.v_out[o + i] = w.y * s;
  }
}

template <class C>
QGB_HD void ph_store_ph(const C& c, int tid, int nt) {
  const int NN = c.N() * c.NK();
  const cplx* qh = member_qh(c);
  cplx* out = c.io.ph_out + (long long)c.member * 2 * NN;
  for (int i = tid; i < NN; i += nt) {
    out[i] = half_ph(c, qh, 0, i);
    out[NN + i] = half_ph(c, qh, 1, i);
  }
}

template <class C>
QGB_HD void ph_store_p(const C& c, int tid, int nt) {
  const int N = c.N(), P = c.P();
  double* p = c.io.p_out + (long long)c.member * 2 * N * N;
  const double s = c.T.inv_M;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    const cplx w = c.buf[y * P + x];
    p[i] = w.x * s;
    p[N * N + i] = w.y * s;
  }
}

// ---- spectral energy budget (pyqg _calc_derived_fields + the add_diagnostic lambdas of qg_model.py) -------------------
template <class C>
struct GetXi {   // xi_h = -wv2 ph  (relative vorticity), both layers as one packed pair
  const C& c; const cplx* qh;
  QGB_HD void operator()(int l, int k, cplx& A, cplx& B) const {
    const int idx = l * c.NK() + k;
    const double w = -(c.T.kv[k] * c.T.kv[k] + c.T.lv[l] * c.T.lv[l]);
    A = cscale(half_ph(c, qh, 0, idx), w);
    B = cscale(half_ph(c, qh, 1, idx), w);
  }
};
template <class C>
struct GetTau {  // tau_h = ph0 - ph1 (paired with zero)
  const C& c; const cplx* qh;
  QGB_HD void operator()(int l, int k, cplx& A, cplx& B) const {
    const int idx = l * c.NK() + k;
    A = csub(half_ph(c, qh, 0, idx), half_ph(c, qh, 1, idx));
    B = cmake(0.0, 0.0);
  }
};
template <class C>
struct GetUVbt {  // barotropic velocities del1 u0 + del2 u1, del1 v0 + del2 v1
  const C& c; const cplx* qh;
  QGB_HD void operator()(int l, int k, cplx& A, cplx& B) const {
    cplx u0, v0, u1, v1;
    half_uv(c, qh, 0, l, k, u0, v0);
    half_uv(c, qh, 1, l, k, u1, v1);
    const double d1 = c.io.Hi_over_H[0], d2 = c.io.Hi_over_H[1];
    A = cmake(d1 * u0.x + d2 * u1.x, d1 * u0.y + d2 * u1.y);
    B = cmake(d1 * v0.x + d2 * v1.x, d1 * v0.y + d2 * v1.y);
  }
};
template <class C>
QGB_HD void ph_build_xi(const C& c, int tid, int nt) { GetXi<C> g{c, member_qh(c)}; build_packed(c, g, tid, nt); }
template <class C>
QGB_HD void ph_build_tau(const C& c, int tid, int nt) { GetTau<C> g{c, member_qh(c)}; build_packed(c, g, tid, nt); }
template <class C>
QGB_HD void ph_build_uvbt(const C& c, int tid, int nt) { GetUVbt<C> g{c, member_qh(c)}; build_packed(c, g, tid, nt); }

// scratch fields slot, slot+1 = real / imaginary part of the packed inverse transform
template <class C>
QGB_HD void ph_store_scr(const C& c, int slot, int nfields, int tid, int nt) {
  const int N = c.N(), P = c.P();
  double* f = c.io.bud_scr + ((long long)c.member * 3 + slot) * N * N;
  const double s = c.T.inv_M;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    const cplx w = c.buf[y * P + x];
    f[i] = w.x * s;
    if (nfields == 2) f[N * N + i] = w.y * s;
  }
}
// buf = u f + i v f  with (u, v) the packed inverse transform in the buffer and f a scratch field
template <class C>
QGB_HD void ph_products_scr(const C& c, int slot, int tid, int nt) {
  const int N = c.N(), P = c.P();
  const double* f = c.io.bud_scr + ((long long)c.member * 3 + slot) * N * N;
  const double s = c.T.inv_M;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    const cplx w = c.buf[y * P + x];
    const double ff = f[i] * s;
    c.buf[y * P + x] = cmake(w.x * ff, w.y * ff);
  }
}
// KEflux (+)= del_z Re(ph_z conj(Jpxi_z)) / M^2,  Jpxi_z = ik F(u xi) + il F(v xi)
template <class C>
QGB_HD void ph_bud_keflux(const C& c, int z, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const cplx* qh = member_qh(c);
  double* out = c.io.bud_out + ((long long)c.member * kBudgetTerms + BUD_KEFLUX) * NN;
  const double s = c.T.inv_M * c.T.inv_M * c.io.Hi_over_H[z];
  for (int i = tid; i < NN; i += nt) {
    const int l = i / NK, k = i - l * NK;
    cplx a, b;
    unpack_pair(c, l, k, a, b);
    const cplx J = cadd(cmuli(a, c.T.kv[k]), cmuli(b, c.T.lv[l]));
    const cplx ph = half_ph(c, qh, z, i);
    const double v = s * (ph.x * J.x + ph.y * J.y);
    out[i] = z == 0 ? v : out[i] + v;
  }
}
// APEflux = F Re((ph0 - ph1) conj(Jptpc)) / M^2 with Jptpc = -(ik F(ubt tau) + il F(vbt tau)); plus the pointwise terms
template <class C>
QGB_HD void ph_bud_apeflux(const C& c, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const cplx* qh = member_qh(c);
  double* out = c.io.bud_out + (long long)c.member * kBudgetTerms * NN;
  const double m2 = c.T.inv_M * c.T.inv_M, d1 = c.io.Hi_over_H[0], d2 = c.io.Hi_over_H[1], F = c.io.bud_F;
  for (int i = tid; i < NN; i += nt) {
    const int l = i / NK, k = i - l * NK;
    cplx a, b;
    unpack_pair(c, l, k, a, b);
    const cplx J = cadd(cmuli(a, c.T.kv[k]), cmuli(b, c.T.lv[l]));      // = -Jptpc
    const cplx p0 = half_ph(c, qh, 0, i), p1 = half_ph(c, qh, 1, i);
    const cplx t = csub(p0, p1);
    out[BUD_APEFLUX * NN + i] = -F * m2 * (t.x * J.x + t.y * J.y);
    // APEgenspec = U F Re(ik (del1 ph0 + del2 ph1) conj(ph0 - ph1)) / M^2
    const cplx bt = cmake(d1 * p0.x + d2 * p1.x, d1 * p0.y + d2 * p1.y);
    const cplx ikbt = cmuli(bt, c.T.kv[k]);
    out[BUD_APEGEN * NN + i] = c.io.bud_U * F * m2 * (ikbt.x * t.x + ikbt.y * t.y);
    const double wv2 = c.T.kv[k] * c.T.kv[k] + c.T.lv[l] * c.T.lv[l];
    out[BUD_KEFRIC * NN + i] = -c.T.rek * d2 * wv2 * m2 * (p1.x * p1.x + p1.y * p1.y);
    const cplx q0 = qh[i], q1 = qh[NN + i];
    const cplx e = cmake(d1 * q0.x + d2 * q1.x, d1 * q0.y + d2 * q1.y);
    out[BUD_ENTSPEC * NN + i] = m2 * (e.x * e.x + e.y * e.y);
    out[BUD_PARAM_KE * NN + i] = 0.0;
    out[BUD_PARAM_APE * NN + i] = 0.0;
    // ENSgenspec = -sum_z del_z Re(ikQy_z conj(qh_z) ph_z) / M^2 ;  ENSfrictionspec = rek del2 wv2 Re(conj(qh_1) ph_1) / M^2
    const double kk = c.T.kv[k];
    out[BUD_ENSGEN * NN + i] = m2 * kk * (d1 * c.T.Qy[0] * (q0.x * p0.y - q0.y * p0.x) + d2 * c.T.Qy[1] * (q1.x * p1.y - q1.y * p1.x));
    out[BUD_ENSFRIC * NN + i] = c.T.rek * d2 * wv2 * m2 * (q1.x * p1.x + q1.y * p1.y);
    out[BUD_ENSPARAM * NN + i] = 0.0;
  }
}
// buf = u q + i v q with the ANOMALY velocities (pyqg _calc_derived_fields: Jq = _advect(q, u, v))
template <class C>
QGB_HD void ph_products_anom(const C& c, int z, int tid, int nt) {
  const int N = c.N(), P = c.P();
  const double* q = c.io.q + ((long long)c.member * 2 + z) * N * N;
  const double s = c.T.inv_M;
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    const cplx w = c.buf[y * P + x];
    const double qq = q[i] * s;
    c.buf[y * P + x] = cmake(w.x * qq, w.y * qq);
  }
}
// ENSflux (+)= -del_z Re(conj(qh_z) Jq_z) / M^2 with Jq_z = ik F(u q) + il F(v q); the tendency of the current state follows from
// the same Jacobian: dqhdt_z = -(Jq_z + ik U_z qh_z + ikQy_z ph_z) (+ rek wv2 ph_1), kept for the dissipation spectra
template <class C>
QGB_HD void ph_bud_ens(const C& c, int z, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const cplx* qh = member_qh(c);
  double* out = c.io.bud_out + ((long long)c.member * kBudgetTerms + BUD_ENSFLUX) * NN;
  cplx* tend = c.io.bud_tend + ((long long)c.member * 2 + z) * NN;
  const double s = c.T.inv_M * c.T.inv_M * c.io.Hi_over_H[z];
  for (int i = tid; i < NN; i += nt) {
    const int l = i / NK, k = i - l * NK;
    cplx a, b;
    unpack_pair(c, l, k, a, b);
    const double kv = c.T.kv[k], lv = c.T.lv[l];
    const cplx J = cadd(cmuli(a, kv), cmuli(b, lv));
    const cplx q = qh[z * NN + i], ph = half_ph(c, qh, z, i);
    const double v = -s * (q.x * J.x + q.y * J.y);
    out[i] = z == 0 ? v : out[i] + v;
    const cplx t1 = cmuli(q, kv * c.T.Ubg[z]), t3 = cmuli(ph, kv * c.T.Qy[z]);
    cplx r = cmake(-(J.x + t1.x + t3.x), -(J.y + t1.y + t3.y));
    if (z == 1 && c.T.rek != 0.0) {
      const double f = c.T.rek * (kv * kv + lv * lv);
      r.x += f * ph.x;
      r.y += f * ph.y;
    }
    tend[i] = r;
  }
}
// Dissspec = -sum_z del_z Re(conj(ph_z) D_z) / (dt M^2), ENSDissspec = sum_z del_z Re(conj(qh_z) D_z) / (dt M^2) with
// D_z = (filtr - 1) (qh_z + dt1 dqhdt_z + dt2 dqhdt_p_z + dt3 dqhdt_pp_z)  (pyqg Model: dissipation_spectrum), and
// ENSparamspec = sum_z del_z Re(conj(qh_z) dqh_z) / M^2.  With a forcing the buffer holds its packed forward transform.
template <class C>
QGB_HD void ph_bud_diss(const C& c, bool has_dq, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const long long mo = (long long)c.member * 2 * NN;
  const cplx* qh = member_qh(c);
  const cplx* tend = c.io.bud_tend + mo;
  const cplx* dp = c.io.d_p + mo;
  const cplx* dpp = c.io.d_pp + mo;
  double* out = c.io.bud_out + (long long)c.member * kBudgetTerms * NN;
  const double m2 = c.T.inv_M * c.T.inv_M, dt1 = c.io.dt1, dt2 = c.io.dt2, dt3 = c.io.dt3;
  const double inv_dt = c.io.bud_inv_dt;
  for (int i = tid; i < NN; i += nt) {
    cplx f[2] = {cmake(0, 0), cmake(0, 0)};
    if (has_dq && !(c.io.bud_demean && i == 0)) {
      const int l = i / NK, k = i - l * NK;
      unpack_pair(c, l, k, f[0], f[1]);
    }
    const double fm = c.T.filtr[i] - 1.0;
    double diss = 0.0, ensdiss = 0.0, ensparam = 0.0;
    for (int z = 0; z < 2; ++z) {
      const int j = z * NN + i;
      const cplx q = qh[j], ph = half_ph(c, qh, z, i);
      const cplx dd = cadd(tend[j], f[z]), a = dp[j], b = dpp[j];
      const cplx D = cmake(fm * (q.x + dt1 * dd.x + dt2 * a.x + dt3 * b.x), fm * (q.y + dt1 * dd.y + dt2 * a.y + dt3 * b.y));
      const double w = c.io.Hi_over_H[z];
      diss -= w * (ph.x * D.x + ph.y * D.y);
      ensdiss += w * (q.x * D.x + q.y * D.y);
      ensparam += w * (q.x * f[z].x + q.y * f[z].y);
    }
    out[BUD_DISS * NN + i] = diss * inv_dt * m2;
    out[BUD_ENSDISS * NN + i] = ensdiss * inv_dt * m2;
    out[BUD_ENSPARAM * NN + i] = ensparam * m2;
  }
}
// parameterization terms from dqh = rfft2(dq) (the buffer holds the packed forward transform of the forcing pair)
template <class C>
QGB_HD void ph_bud_param(const C& c, int tid, int nt) {
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const cplx* qh = member_qh(c);
  double* out = c.io.bud_out + (long long)c.member * kBudgetTerms * NN;
  const double m2 = c.T.inv_M * c.T.inv_M, d1 = c.io.Hi_over_H[0], d2 = c.io.Hi_over_H[1], F = c.io.bud_F;
  for (int i = tid; i < NN; i += nt) {
    const int l = i / NK, k = i - l * NK;
    cplx f0, f1;
    unpack_pair(c, l, k, f0, f1);
    const double a00 = c.T.a[i], a01 = c.T.a[NN + i], a10 = c.T.a[2 * NN + i], a11 = c.T.a[3 * NN + i];
    const cplx dp0 = cmake(a00 * f0.x + a01 * f1.x, a00 * f0.y + a01 * f1.y);
    const cplx dp1 = cmake(a10 * f0.x + a11 * f1.x, a10 * f0.y + a11 * f1.y);
    const cplx p0 = half_ph(c, qh, 0, i), p1 = half_ph(c, qh, 1, i);
    const double wv2 = c.T.kv[k] * c.T.kv[k] + c.T.lv[l] * c.T.lv[l];
    out[BUD_PARAM_KE * NN + i] = wv2 * m2 * (d1 * (p0.x * dp0.x + p0.y * dp0.y) + d2 * (p1.x * dp1.x + p1.y * dp1.y));
    const cplx t = csub(p0, p1), dt = csub(dp0, dp1);
    out[BUD_PARAM_APE * NN + i] = F * m2 * (t.x * dt.x + t.y * dt.y);
  }
}

// diagnostics partials: red[0*nt+tid] ke, [1] max|u+U|, [2] max|v|, [3] non-finite count
template <class C>
QGB_HD void ph_red_clear(const C& c, int tid, int nt) {
  for (int j = 0; j < 4; ++j) c.red[j * nt + tid] = 0.0;
}
template <class C>
QGB_HD void ph_red_ke(const C& c, int tid, int nt) {  // pyqg _calc_ke via spec_var(wv*ph)
  const int N = c.N(), NK = c.NK(), NN = N * NK;
  const cplx* qh = member_qh(c);
  double acc = 0.0, bad = 0.0;
  for (int i = tid; i < NN; i += nt) {
    const int l = i / NK, k = i - l * NK;
    const double wv2 = c.T.kv[k] * c.T.kv[k] + c.T.lv[l] * c.T.lv[l];
    const double wgt = (k == 0 || k == NK - 1) ? 1.0 : 2.0;
    for (int z = 0; z < 2; ++z) {
      const cplx ph = half_ph(c, qh, z, i);
      acc += 0.5 * c.io.Hi_over_H[z] * wgt * wv2 * (ph.x * ph.x + ph.y * ph.y);
      const cplx qq = qh[z * NN + i];
      if (!(fabs(qq.x) < 1e300) || !(fabs(qq.y) < 1e300)) bad += 1.0;
    }
  }
  c.red[0 * nt + tid] += acc * c.T.inv_M * c.T.inv_M;
  c.red[3 * nt + tid] += bad;
}
template <class C>
QGB_HD void ph_red_uv(const C& c, int z, int tid, int nt) {  // pyqg _calc_cfl numerator
  const int N = c.N(), P = c.P();
  const double s = c.T.inv_M, U = c.T.Ubg[z];
  double mu = c.red[1 * nt + tid], mv = c.red[2 * nt + tid];
  for (int i = tid; i < N * N; i += nt) {
    const int y = i / N, x = i - y * N;
    const cplx w = c.buf[y * P + x];
    mu = fmax(mu, fabs(w.x * s + U));
    mv = fmax(mv, fabs(w.y * s));
  }
  c.red[1 * nt + tid] = mu;
  c.red[2 * nt + tid] = mv;
}
template <class C>
QGB_HD void ph_red_final(const C& c, int tid, int nt) {
  if (tid != 0) return;
  double ke = 0, mu = 0, mv = 0, bad = 0;
  for (int t = 0; t < nt; ++t) {
    ke += c.red[t];
    mu = fmax(mu, c.red[nt + t]);
    mv = fmax(mv, c.red[2 * nt + t]);
    bad += c.red[3 * nt + t];
  }
  double* o = c.io.red_out + (long long)c.member * 4;
  o[0] = ke; o[1] = mu; o[2] = mv; o[3] = bad;
}

// ------------------------------------------------------------------------------------------------------
// programs: ordered phase lists.  ``phase`` selects which phase runs; returns the number of phases.
// ------------------------------------------------------------------------------------------------------
// PROG_STEP_DQ removes the mean of dq (closure output, models/parameterization.py:25); PROG_STEP_DQ_RAW adds dq as given
// PROG_ADVECT: d_cur = -(ik uqh + il vqh + ikQy ph) [+ friction] without time stepping (tools/operators.py:249-252 ``advect`` when
// the tables carry Ubg = Qy = rek = 0);  PROG_C2R: q = irfft2(qh) for arbitrary half-plane spectra (tools/operators.py:132)
// PROG_BUDGET: spectral energy budget terms of the current state (and of the closure forcing io.dq when given) -> bud_out
enum Program { PROG_STEP = 0, PROG_STEP_DQ = 1, PROG_SET_Q = 2, PROG_INVERT = 3, PROG_DIAG = 4, PROG_EMIT_X = 5, PROG_STEP_DQ_RAW = 6,
               PROG_ADVECT = 7, PROG_C2R = 8, PROG_BUDGET = 9 };

#ifdef __CUDACC__
// ---- cluster path: the time step with the pointwise phases fused into the tiled transform passes ---------------------
// Per layer z: build (u,v) spectra | inverse x pass | [inverse y pass -> (u+U) q, v q -> forward y pass] | forward x pass |
// tendency;  then [dq -> forward y pass] | forward x pass | AB3 update | build q spectra | inverse x pass |
// [inverse y pass -> q, closure input].  14 (16 with a forcing) cluster phases instead of 23 (26).  The forward transform
// runs its y pass first here (the two 1-D transforms commute and the digit-reversed layout is per dimension).
template <class C>
struct TileProducts {   // ph_products on one element: the tile holds the unnormalised (u, v) at physical (y, x)
  const C& c; const double* q; double s, U;
  __device__ cplx operator()(int y, int x, cplx w) const {
    const double qq = q[y * c.N() + x];
    return cmake((w.x * s + U) * qq, (w.y * s) * qq);
  }
};
template <class C>
struct TileLoadPair {   // ph_load_pair
  const C& c; const double* f0; const double* f1;
  __device__ cplx operator()(int y, int x) const { return cmake(f0[y * c.N() + x], f1[y * c.N() + x]); }
};
template <class C>
struct TileEmitQ {      // ph_emit_q
  const C& c; double* q; float* x; double s;
  __device__ void operator()(int y, int xx, cplx w) const {
    const int N = c.N(), i = y * N + xx;
    const double q0 = w.x * s, q1 = w.y * s;
    q[i] = q0;
    q[N * N + i] = q1;
    if (x) {
      x[i] = (float)q0 / c.io.x_std[0];
      x[N * N + i] = (float)q1 / c.io.x_std[1];
    }
  }
};

template <class C>
__device__ int run_step_tiled(const C& c, int prog, int phase, int tid, int nt) {
  int _n = 0;
  const int N = c.N();
  const bool with_dq = prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW;
#define QGB_TRUN(stmt) do { if (_n == phase) { stmt; } ++_n; } while (0)
  for (int z = 0; z < 2; ++z) {
    QGB_TRUN(ph_build_uv(c, z, tid, nt));     // (building inside the pass load would turn the scatter into uncoalesced reads: slower)
    QGB_TRUN((tiled_pass(c, 0, 1, TileNone(), TileNone(), TileNone(), tid, nt)));
    QGB_TRUN((tiled_pass(c, 1, 2, TileNone(),
                         TileProducts<C>{c, c.io.q + ((long long)c.member * 2 + z) * N * N, c.T.inv_M, c.T.Ubg[z]}, TileNone(), tid, nt)));
    QGB_TRUN((tiled_pass(c, 0, 0, TileNone(), TileNone(), TileNone(), tid, nt)));
    QGB_TRUN(ph_tendency(c, z, tid, nt));
  }
  if (with_dq) {
    const double* f0 = c.io.dq + (long long)c.member * 2 * N * N;
    QGB_TRUN((tiled_pass(c, 1, 0, TileLoadPair<C>{c, f0, f0 + N * N}, TileNone(), TileNone(), tid, nt)));
    QGB_TRUN((tiled_pass(c, 0, 0, TileNone(), TileNone(), TileNone(), tid, nt)));
  }
  QGB_TRUN(ph_update(c, with_dq, prog == PROG_STEP_DQ, tid, nt));
  QGB_TRUN(ph_build_q(c, tid, nt));
  QGB_TRUN((tiled_pass(c, 0, 1, TileNone(), TileNone(), TileNone(), tid, nt)));
  QGB_TRUN((tiled_pass(c, 1, 1, TileNone(), TileNone(),
                       TileEmitQ<C>{c, c.io.q + (long long)c.member * 2 * N * N,
                                    c.io.cnn_x ? c.io.cnn_x + (long long)c.member * c.io.cnn_mstride : nullptr, c.T.inv_M}, tid, nt)));
#undef QGB_TRUN
  return _n;
}
#endif

#define QGB_RUN(stmt)        \
  do {                       \
    if (_n == phase) { stmt; } \
    ++_n;                    \
  } while (0)
#define QGB_FFT(inv)                                                    \
  do {                                                                  \
    for (int _f = 0; _f < F; ++_f) QGB_RUN(fft2d_phase(c, _f, inv, tid, nt)); \
  } while (0)

template <class C>
QGB_HD int run_program(const C& c, int prog, int phase, int tid, int nt) {
#ifdef __CUDA_ARCH__
  if (c.tile && (prog == PROG_STEP || prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW))
    return run_step_tiled(c, prog, phase, tid, nt);
#endif
  int _n = 0;
  const int F = c.tile ? 2 : fft2d_phases(c.T);
  QGB_RUN(ph_init(c, tid, nt));
  const bool with_dq = prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW;
  if (prog == PROG_STEP || with_dq) {
    for (int z = 0; z < 2; ++z) {
      QGB_RUN(ph_build_uv(c, z, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_products(c, z, tid, nt));
      QGB_FFT(false);
      QGB_RUN(ph_tendency(c, z, tid, nt));
    }
    if (with_dq) {
      QGB_RUN(ph_load_pair(c, c.io.dq, tid, nt));
      QGB_FFT(false);
    }
    QGB_RUN(ph_update(c, with_dq, prog == PROG_STEP_DQ, tid, nt));
    QGB_RUN(ph_build_q(c, tid, nt));
    QGB_FFT(true);
    QGB_RUN(ph_emit_q(c, tid, nt));
  } else if (prog == PROG_ADVECT) {
    for (int z = 0; z < 2; ++z) {
      QGB_RUN(ph_build_uv(c, z, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_products(c, z, tid, nt));
      QGB_FFT(false);
      QGB_RUN(ph_tendency(c, z, tid, nt));
    }
  } else if (prog == PROG_C2R) {
    QGB_RUN(ph_build_q(c, tid, nt));
    QGB_FFT(true);
    QGB_RUN(ph_emit_q(c, tid, nt));
  } else if (prog == PROG_SET_Q) {
    QGB_RUN(ph_load_pair(c, c.io.q, tid, nt));
    QGB_FFT(false);
    QGB_RUN(ph_store_qh(c, tid, nt));
    if (c.io.cnn_x) QGB_RUN(ph_emit_x_only(c, tid, nt));
  } else if (prog == PROG_INVERT) {
    if (c.io.ph_out) QGB_RUN(ph_store_ph(c, tid, nt));
    for (int z = 0; z < 2; ++z) {
      QGB_RUN(ph_build_uv(c, z, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_store_uv(c, z, tid, nt));
    }
    if (c.io.p_out) {
      QGB_RUN(ph_build_p(c, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_store_p(c, tid, nt));
    }
  } else if (prog == PROG_DIAG) {
    QGB_RUN(ph_red_clear(c, tid, nt));
    QGB_RUN(ph_red_ke(c, tid, nt));
    for (int z = 0; z < 2; ++z) {
      QGB_RUN(ph_build_uv(c, z, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_red_uv(c, z, tid, nt));
    }
    QGB_RUN(ph_red_final(c, tid, nt));
  } else if (prog == PROG_EMIT_X) {
    QGB_RUN(ph_emit_x_only(c, tid, nt));
  } else if (prog == PROG_BUDGET) {
    QGB_RUN(ph_build_xi(c, tid, nt));
    QGB_FFT(true);
    QGB_RUN(ph_store_scr(c, 0, 2, tid, nt));
    for (int z = 0; z < 2; ++z) {
      QGB_RUN(ph_build_uv(c, z, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_products_scr(c, z, tid, nt));
      QGB_FFT(false);
      QGB_RUN(ph_bud_keflux(c, z, tid, nt));
    }
    QGB_RUN(ph_build_tau(c, tid, nt));
    QGB_FFT(true);
    QGB_RUN(ph_store_scr(c, 2, 1, tid, nt));
    QGB_RUN(ph_build_uvbt(c, tid, nt));
    QGB_FFT(true);
    QGB_RUN(ph_products_scr(c, 2, tid, nt));
    QGB_FFT(false);
    QGB_RUN(ph_bud_apeflux(c, tid, nt));
    for (int z = 0; z < 2; ++z) {
      QGB_RUN(ph_build_uv(c, z, tid, nt));
      QGB_FFT(true);
      QGB_RUN(ph_products_anom(c, z, tid, nt));
      QGB_FFT(false);
      QGB_RUN(ph_bud_ens(c, z, tid, nt));
    }
    if (c.io.dq) {
      QGB_RUN(ph_load_pair(c, c.io.dq, tid, nt));
      QGB_FFT(false);
      QGB_RUN(ph_bud_param(c, tid, nt));
    }
    QGB_RUN(ph_bud_diss(c, c.io.dq != nullptr, tid, nt));
  }
  return _n;
}

// ---- host-side plan construction (shared by the library and the emulation) ---------------------------
inline bool make_radix_plan(int N, int* radix, int* nstages) {
  int n = N, s = 0;
  while (n % 4 == 0 && s < kMaxStages) { radix[s++] = 4; n /= 4; }
  while (n % 2 == 0 && s < kMaxStages) { radix[s++] = 2; n /= 2; }
  while (n % 3 == 0 && s < kMaxStages) { radix[s++] = 3; n /= 3; }
  *nstages = s;
  return n == 1;
}

// position of frequency f after the in-place forward transform
inline int pos_of_freq(int N, const int* radix, int nstages, int f) {
  int pos = 0, span = N;
  for (int s = 0; s < nstages; ++s) {
    span /= radix[s];
    pos += (f % radix[s]) * span;
    f /= radix[s];
  }
  return pos;
}

}  // namespace qgb
