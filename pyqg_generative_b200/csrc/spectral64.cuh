// spectral64.cuh -- register-FFT fused spectral time step for nx = 64 (the headline grid), sm_100a.
//
// What it computes is the same pyqg step as qg_core.cuh (Model._step_forward: _invert, _do_advection, _do_friction,
// _do_q_subgrid_parameterization, _forward_timestep; reference call sites tools/simulate.py:132,137,168).  What changes is
// where the work happens.  The first-generation kernel kept a packed 64 x 64 complex field in shared memory and ran every
// radix-4 stage of every 1-D pass as a shared-memory round trip (36 passes + barriers per step: 64 % of the shared-memory
// pipe, 0.21 of the HBM roofline).  Here one CTA of 256 threads owns one member and
//
//   * a 64-point line transform is done in REGISTERS by four lanes: a 16-point FFT per lane, a 4 x 4 block transpose across
//     the four lanes with warp shuffles (2 xor rounds), twiddles, four radix-4 butterflies (fft64);
//   * shared memory is touched once per 2-D transform for the transposition between the x pass (lane quad = one row) and the
//     y pass (lane quad = one column), and once for handing the half-plane spectra to / from the pointwise phases;
//   * two REAL fields ride one complex transform (u + i v, uq + i vq, q1 + i q2, dq1 + i dq2).  The split into the two
//     half-plane spectra happens between the passes (A_y(k) = (W_y(k) + conj W_y(-k)) / 2 ...), so the y pass transforms the
//     31 + 31 columns k = 1..31 of A and B plus TWO packed columns A(0) + i A(32), B(0) + i B(32) (those four sequences are
//     real in y): 64 column transforms for 256 threads, and no Hermitian extension is ever built;
//   * the spectral pointwise phases (inversion, tendency, AB3 + filter) run over the 64 x 33 half-plane points in array
//     order -- consecutive lanes = consecutive k: every global access is a coalesced 16-byte access -- and the physical-space
//     products are fused in registers between an inverse and a forward x pass;
//   * the inverse transform is the forward routine on conjugated data, and the whole step is ONE loop over rounds with one
//     copy of each stage in the instruction stream (the first fully unrolled version was 217 KB of SASS and stalled on
//     instruction fetch).
#pragma once
#include <cuda_runtime.h>

#include "qg_core.cuh"

namespace qgb {

// pointwise spectral work, one bit per stage; a phase runs the selected stages on every half-plane point in array order
enum { PW_TEND0 = 1, PW_TEND1 = 2, PW_UPDATE = 4, PW_UV0 = 8, PW_UV1 = 16, PW_STORE_QH = 32, PW_LOAD_QH = 64, PW_FORCING = 128,
       PW_PSI = 256 };   // PW_PSI: streamfunctions of both layers as one pair (and io.ph_out)                    (pyqg _invert)
// physical-space stage between the inverse and the forward x pass of a round
enum { PH_PRODUCTS0 = 0, PH_PRODUCTS1 = 1, PH_LOAD_DQ = 2, PH_LOAD_Q = 3, PH_EMIT = 4,
       PH_STORE_UV0 = 5, PH_STORE_UV1 = 6, PH_STORE_P = 7,     // PROG_INVERT: u, v of a layer / psi of both layers to io.*_out
       // PROG_BUDGET (round<true>): products of the inverse transform with scratch field 0..2 / with q_0, q_1 (values < PH_STORE_UV0 are
       // product stages), inverse transform stored to the scratch fields 0, 1 / to scratch field 2
       PH_SCR_STORE01 = 8, PH_SCR_STORE2 = 9, PH_SCR_PROD0 = -8, PH_SCR_PROD1 = -7, PH_SCR_PROD2 = -6, PH_ANOM0 = -4, PH_ANOM1 = -3 };

namespace s64 {

constexpr int N = 64, NK = 33, NN = N * NK, NPIX = N * N;
constexpr int PITCH = 68;                       // complex elements per row of the transposition buffer (== 4 mod 8: conflict free)
constexpr int B_OFF = 36;                       // B plane starts at column 36 of a row (A: 0..31, B: 36..67)
constexpr int kThreads = 256;
constexpr size_t kSmemBytes = (size_t)N * PITCH * sizeof(cplx) + N * sizeof(cplx);   // buffer + twiddle table

#define S64_INL __device__ __forceinline__
#ifdef S64_NOINLINE
#define S64_PHASE __device__ __noinline__
#else
#define S64_PHASE __device__ __forceinline__
#endif
#ifndef S64_PREFETCH
#define S64_PREFETCH 0
#endif
#ifndef S64_NB_LIGHT
#define S64_NB_LIGHT 2     // half-plane points in flight per thread in the light pointwise phases
#endif
#ifndef S64_NB_UV
#define S64_NB_UV S64_NB_LIGHT   // ... in the phases that only read or only write the spectra (few registers per point)
#endif
#ifndef S64_NB_UPD
#define S64_NB_UPD 1       // ... in the Adams-Bashforth update (8 complex operands per point)
#endif

// ---- small complex helpers ---------------------------------------------------------------------------------------------
template <bool INV>
S64_INL void dft4(cplx& a, cplx& b, cplx& c, cplx& d) {
  const cplx t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = csub(b, d);
  a = cadd(t0, t2);
  c = csub(t0, t2);
  if (!INV) {            // forward: y1 = t1 - i t3, y3 = t1 + i t3
    b = cmake(t1.x + t3.y, t1.y - t3.x);
    d = cmake(t1.x - t3.y, t1.y + t3.x);
  } else {
    b = cmake(t1.x - t3.y, t1.y + t3.x);
    d = cmake(t1.x + t3.y, t1.y - t3.x);
  }
}

// a * w16^E (forward: w16 = exp(-2 pi i / 16); INV: its conjugate), E a compile-time constant
template <int E, bool INV>
S64_INL cplx mulw16(cplx a) {
  constexpr int e = ((E % 16) + 16) % 16;
  constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178, H = 0.70710678118654752440;
  if (e == 0) return a;
  if (e == 8) return cmake(-a.x, -a.y);
  if (e == 4) return INV ? cmake(-a.y, a.x) : cmake(a.y, -a.x);        // -i (forward), +i (inverse)
  if (e == 12) return INV ? cmake(a.y, -a.x) : cmake(-a.y, a.x);
  // general: w = (c, -s) forward, (c, s) inverse with c = cos(2 pi e / 16), s = sin(2 pi e / 16)
  constexpr double c = e == 1 ? C1 : e == 2 ? H : e == 3 ? S1 : e == 5 ? -S1 : e == 6 ? -H : e == 7 ? -C1 : e == 9 ? -C1
                     : e == 10 ? -H : e == 11 ? -S1 : e == 13 ? S1 : e == 14 ? H : C1;
  constexpr double s = e == 1 ? S1 : e == 2 ? H : e == 3 ? C1 : e == 5 ? C1 : e == 6 ? H : e == 7 ? S1 : e == 9 ? -S1
                     : e == 10 ? -H : e == 11 ? -C1 : e == 13 ? -C1 : e == 14 ? -H : -S1;
  constexpr double wy = INV ? s : -s;
  return cmake(a.x * c - a.y * wy, a.x * wy + a.y * c);
}

// 16-point DFT in registers, natural order in and out:  v[k] <- sum_j v[j] w16^{jk}   (INV: conjugate kernel, unnormalised)
template <bool INV>
S64_INL void fft16(cplx (&v)[16]) {
  // n = 4a + b, k = c + 4d:  radix-4 over a for each b, twiddle w16^{bc}, radix-4 over b for each c
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);     // v[4c + b] = T[b][c]
  v[5] = mulw16<1, INV>(v[5]);   v[6] = mulw16<2, INV>(v[6]);    v[7] = mulw16<3, INV>(v[7]);
  v[9] = mulw16<2, INV>(v[9]);   v[10] = mulw16<4, INV>(v[10]);  v[11] = mulw16<6, INV>(v[11]);
  v[13] = mulw16<3, INV>(v[13]); v[14] = mulw16<6, INV>(v[14]);  v[15] = mulw16<9, INV>(v[15]);
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // v[4c + d] = Y[c + 4d]
  // transpose the 4 x 4 register naming to natural order (free: static indices)
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) { const cplx tmp = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = tmp; }
}

S64_INL cplx shfl_xor_c(cplx a, int mask, unsigned lanes = 0xffffffffu) {
  return cmake(__shfl_xor_sync(lanes, a.x, mask), __shfl_xor_sync(lanes, a.y, mask));
}
S64_INL cplx shfl_c(cplx a, int src, unsigned lanes) {
  return cmake(__shfl_sync(lanes, a.x, src), __shfl_sync(lanes, a.y, src));
}
// a / b in fp32 with b's rounded reciprocal at hand: quotient estimate + one residual correction (the sequence the compiler's
// IEEE division runs, without its range checks: |a / b| is O(1) here) -- x_scale.normalize(q.astype('float32')), cnn_tools.py:524-528
S64_INL float div_by(float a, float b, float inv) {
  const float q = a * inv;
  return fmaf(fmaf(-q, b, a), inv, q);
}
S64_INL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
S64_INL cplx sel(bool p, cplx a, cplx b) { return cmake(p ? a.x : b.x, p ? a.y : b.y); }

// One xor round of the 4 x 4 block transpose across the four lanes of a line: lane bit ``bit`` decides which element of
// each register pair (e0 = slot with that index bit clear, e1 = set) stays and which one travels.
S64_INL void exchange_pair(cplx& e0, cplx& e1, bool bit, int mask) {
  const cplx send = sel(bit, e0, e1);
  const cplx recv = shfl_xor_c(send, mask);
  e0 = sel(bit, recv, e0);
  e1 = sel(bit, e1, recv);
}
// LS = lane stride of the line index t (1: lanes 4r + t, x pass; 8: lanes 8t + s, y pass)
template <int LS>
S64_INL void round_hi(cplx (&v)[16], int t) {     // index bit 1 of the block number / of t
  const bool bit = (t >> 1) & 1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int be = 0; be < 2; ++be) exchange_pair(v[4 * i + be], v[4 * i + 2 + be], bit, 2 * LS);
}
template <int LS>
S64_INL void round_lo(cplx (&v)[16], int t) {     // index bit 0
  const bool bit = t & 1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int s1 = 0; s1 < 2; ++s1) exchange_pair(v[4 * i + 2 * s1], v[4 * i + 2 * s1 + 1], bit, LS);
}

// 64-point DFT of a line distributed over four lanes t = 0..3 (forward kernel exp(-2 pi i n k / 64)):
//   in  v[j] = x[4 j + t]            out v[4 i + k2] = X[t + 4 i + 16 k2]
// m_lo / m_hi: xor masks of the lanes that differ in bit 0 / bit 1 of t;  om[n] = w64^{n t} (n = 1, 2, 3)
S64_INL void fft64(cplx (&v)[16], int t, int m_lo, int m_hi, const cplx* tw) {
  const cplx om1 = tw[t], om2 = tw[2 * t], om3 = tw[3 * t];
  fft16<false>(v);                        // v[k1] = sum_j x[4j + t] w16^{j k1}
  {                                       // 4 x 4 block transpose: afterwards v[4 i + n2] = Z_{n2}[t + 4 i]
    const bool bit = (t >> 1) & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int be = 0; be < 2; ++be) exchange_pair(v[4 * i + be], v[4 * i + 2 + be], bit, m_hi);
  }
  {
    const bool bit = t & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int s1 = 0; s1 < 2; ++s1) exchange_pair(v[4 * i + 2 * s1], v[4 * i + 2 * s1 + 1], bit, m_lo);
  }
  // twiddle w64^{n2 (t + 4 i)} = om[n2] * w16^{n2 i}
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[4 * i + 1] = cmul(v[4 * i + 1], om1);
    v[4 * i + 2] = cmul(v[4 * i + 2], om2);
    v[4 * i + 3] = cmul(v[4 * i + 3], om3);
  }
  v[5] = mulw16<1, false>(v[5]);   v[6] = mulw16<2, false>(v[6]);    v[7] = mulw16<3, false>(v[7]);
  v[9] = mulw16<2, false>(v[9]);   v[10] = mulw16<4, false>(v[10]);  v[11] = mulw16<6, false>(v[11]);
  v[13] = mulw16<3, false>(v[13]); v[14] = mulw16<6, false>(v[14]);  v[15] = mulw16<9, false>(v[15]);
#pragma unroll
  for (int i = 0; i < 4; ++i) dft4<false>(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
S64_INL void conj16(cplx (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j].y = -v[j].y;
}
// rename v[4 i + k2] (index t + 4 i + 16 k2 = 4 (i + 4 k2) + t) to the input naming v[j], j = i + 4 k2 (static: free)
S64_INL void out_to_in_naming(cplx (&v)[16]) {
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) { const cplx tmp = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = tmp; }
}

// ---- thread geometry ---------------------------------------------------------------------------------------------------
// x pass: lane = 4 r + t, row y = 8 warp + r.   y pass: lane = 8 t + s, slot = 4 warp + (s & 3), isB = (s >> 2) & 1;
// slot 0 = the packed columns (k = 0 and k = 32), slot c = column k = c.
// Shared buffer (rows of PITCH complex numbers), used in turn as
//   T  transposition, physical -> spectral: T[y][kx] = W_y(kx), kx = 0..63
//   T' transposition, spectral -> physical: T'[y][c] = A_y(c), T'[y][36 + c] = B_y(c)  (c = 0: packed C0 / C1)
//   S  half-plane spectra: S[l][k] = Ahat(l, k), S[l][36 + k] = Bhat(l, k) for k = 1..31; columns 0 / 36 hold the packed
//      transforms C0(l) / C1(l) after a forward transform; before an inverse one the raw k = 32 and k = 0 spectra sit in the pad
//      columns 32 / 33 (A / B at k = 32) and 34 / 35 (k = 0).
struct Geo {
  int warp, lane, xt, y, yt, slot, isB, col;
  bool packed;
  __device__ Geo() {
    warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
    xt = lane & 3; y = 8 * warp + (lane >> 2);
    yt = lane >> 3; slot = 4 * warp + (lane & 3); isB = (lane >> 2) & 1;
    packed = slot == 0;
    col = slot + (isB ? B_OFF : 0);
  }
};

// ---- half-plane spectra in the buffer --------------------------------------------------------------------------------------
// after a forward transform: (Ahat, Bhat) at (l, k).  k = 0 / 32 come out of the packed columns:
// e0(l) = (C(l) + conj C(-l)) / 2 (k = 0), e1(l) = (C(l) - conj C(-l)) / (2 i) (k = 32)
S64_INL void spec_read(const cplx* S, int l, int k, cplx& A, cplx& B) {
  if (k != 0 && k != 32) { A = S[l * PITCH + k]; B = S[l * PITCH + B_OFF + k]; return; }
  const int ln = (N - l) & (N - 1);
  const cplx c0 = S[l * PITCH], n0 = S[ln * PITCH], c1 = S[l * PITCH + B_OFF], n1 = S[ln * PITCH + B_OFF];
  if (k == 0) {
    A = cmake(0.5 * (c0.x + n0.x), 0.5 * (c0.y - n0.y));
    B = cmake(0.5 * (c1.x + n1.x), 0.5 * (c1.y - n1.y));
  } else {
    A = cmake(0.5 * (c0.y + n0.y), 0.5 * (n0.x - c0.x));
    B = cmake(0.5 * (c1.y + n1.y), 0.5 * (n1.x - c1.x));
  }
}
// before an inverse transform: raw spectra; k = 32 and k = 0 go to the pad columns 32 / 33 and 34 / 35 (the packed lanes combine
// them; columns 0 / 36 may still be read as C(-l) by the point (64 - l, 0) of the same phase)
S64_INL void spec_write(cplx* S, int l, int k, cplx A, cplx B) {
  const int ca = k == 32 ? 32 : (k == 0 ? 34 : k), cb = k == 32 ? 33 : (k == 0 ? 35 : B_OFF + k);
  S[l * PITCH + ca] = A;
  S[l * PITCH + cb] = B;
}

// per-member base pointers, rebuilt from (io, member) inside every phase instead of being carried across the whole step
// (registers are the scarce resource: 128 per thread for two CTAs per SM)
struct MemberPtrs {
  cplx* qh; double* q; cplx* d_cur; const cplx* d_p; const cplx* d_pp; const double* dq; float* cnn_x;
  __device__ MemberPtrs(const StepIO& io, int m) {
    qh = io.qh + (long long)m * 2 * NN;
    q = io.q + (long long)m * 2 * NPIX;
    d_cur = io.d_cur ? io.d_cur + (long long)m * 2 * NN : nullptr;
    d_p = io.d_p ? io.d_p + (long long)m * 2 * NN : nullptr;
    d_pp = io.d_pp ? io.d_pp + (long long)m * 2 * NN : nullptr;
    dq = io.dq ? io.dq + (long long)m * 2 * NPIX : nullptr;
    cnn_x = io.cnn_x ? io.cnn_x + (long long)m * io.cnn_mstride : nullptr;
  }
};


// ---- spectral energy / enstrophy budgets (PROG_BUDGET) at one half-plane point ------------------------------------------------
// The arithmetic of qg_core.cuh ph_bud_keflux / ph_bud_apeflux / ph_bud_ens / ph_bud_param / ph_bud_diss and of its spectral
// builders GetXi / GetUV / GetTau / GetUVbt (pyqg _calc_derived_fields + the add_diagnostic lambdas), one point at a time, for the
// register-FFT kernels: a budget phase first consumes the pair of half-plane spectra (A, B) the last forward transform left in S
// (BR_*), then leaves the next pair to be transformed back in S (BWR_*).
enum { BR_NONE = 0, BR_KEFLUX0, BR_KEFLUX1, BR_APEFLUX, BR_ENS0, BR_ENS1, BR_DISS, BR_PARAM_DISS };
enum { BWR_NONE = 0, BWR_XI, BWR_UV0, BWR_UV1, BWR_TAU, BWR_UVBT };

template <int WR>
S64_INL void bud_build(cplx p0, cplx p1, double kv, double lv, double d1, double d2, cplx& A, cplx& B) {
  if (WR == BWR_XI) {                                  // xi_h = -wv2 ph (relative vorticity), both layers
    const double w = -(kv * kv + lv * lv);
    A = cscale(p0, w); B = cscale(p1, w);
  } else if (WR == BWR_UV0 || WR == BWR_UV1) {         // uh = -il ph, vh = ik ph
    const cplx ph = WR == BWR_UV0 ? p0 : p1;
    A = cmake(lv * ph.y, -lv * ph.x); B = cmake(-kv * ph.y, kv * ph.x);
  } else if (WR == BWR_TAU) {                          // tau_h = ph0 - ph1, paired with zero
    A = csub(p0, p1); B = cmake(0.0, 0.0);
  } else {                                             // barotropic velocities del1 u0 + del2 u1, del1 v0 + del2 v1
    const cplx u0 = cmake(lv * p0.y, -lv * p0.x), v0 = cmake(-kv * p0.y, kv * p0.x);
    const cplx u1 = cmake(lv * p1.y, -lv * p1.x), v1 = cmake(-kv * p1.y, kv * p1.x);
    A = cmake(d1 * u0.x + d2 * u1.x, d1 * u0.y + d2 * u1.y);
    B = cmake(d1 * v0.x + d2 * v1.x, d1 * v0.y + d2 * v1.y);
  }
}

// out / tend / dp / dpp: this member's bud_out (kBudgetTerms, NN), bud_tend, dqhdt_p, dqhdt_pp; a**: inversion coefficients at idx
template <int RD>
S64_INL void bud_read(const Tables& T, const StepIO& io, double* out, cplx* tend, const cplx* dp, const cplx* dpp, int NN, int idx,
                      double kv, double lv, cplx q0, cplx q1, cplx p0, cplx p1, double a00, double a01, double a10, double a11,
                      cplx A, cplx B) {
  const double m2 = T.inv_M * T.inv_M, d1 = io.Hi_over_H[0], d2 = io.Hi_over_H[1];
  if (RD == BR_KEFLUX0 || RD == BR_KEFLUX1) {
    // KEflux (+)= del_z Re(ph_z conj(Jpxi_z)) / M^2,  Jpxi_z = ik F(u xi) + il F(v xi)
    constexpr int z = RD == BR_KEFLUX1 ? 1 : 0;
    const cplx J = cadd(cmuli(A, kv), cmuli(B, lv));
    const cplx ph = z ? p1 : p0;
    const double v = (m2 * io.Hi_over_H[z]) * (ph.x * J.x + ph.y * J.y);
    out[BUD_KEFLUX * NN + idx] = z == 0 ? v : out[BUD_KEFLUX * NN + idx] + v;
  } else if (RD == BR_APEFLUX) {
    const double F = io.bud_F;
    const cplx J = cadd(cmuli(A, kv), cmuli(B, lv));                    // = -Jptpc
    const cplx t = csub(p0, p1);
    out[BUD_APEFLUX * NN + idx] = -F * m2 * (t.x * J.x + t.y * J.y);
    const cplx bt = cmake(d1 * p0.x + d2 * p1.x, d1 * p0.y + d2 * p1.y);
    const cplx ikbt = cmuli(bt, kv);
    out[BUD_APEGEN * NN + idx] = io.bud_U * F * m2 * (ikbt.x * t.x + ikbt.y * t.y);
    const double wv2 = kv * kv + lv * lv;
    out[BUD_KEFRIC * NN + idx] = -T.rek * d2 * wv2 * m2 * (p1.x * p1.x + p1.y * p1.y);
    const cplx e = cmake(d1 * q0.x + d2 * q1.x, d1 * q0.y + d2 * q1.y);
    out[BUD_ENTSPEC * NN + idx] = m2 * (e.x * e.x + e.y * e.y);
    out[BUD_PARAM_KE * NN + idx] = 0.0;
    out[BUD_PARAM_APE * NN + idx] = 0.0;
    out[BUD_ENSGEN * NN + idx] = m2 * kv * (d1 * T.Qy[0] * (q0.x * p0.y - q0.y * p0.x) + d2 * T.Qy[1] * (q1.x * p1.y - q1.y * p1.x));
    out[BUD_ENSFRIC * NN + idx] = T.rek * d2 * wv2 * m2 * (q1.x * p1.x + q1.y * p1.y);
    out[BUD_ENSPARAM * NN + idx] = 0.0;
  } else if (RD == BR_ENS0 || RD == BR_ENS1) {
    // ENSflux (+)= -del_z Re(conj(qh_z) Jq_z) / M^2; the tendency of the current state from the same Jacobian (for the dissipation spectra)
    constexpr int z = RD == BR_ENS1 ? 1 : 0;
    const cplx J = cadd(cmuli(A, kv), cmuli(B, lv));
    const cplx q = z ? q1 : q0, ph = z ? p1 : p0;
    const double v = -(m2 * io.Hi_over_H[z]) * (q.x * J.x + q.y * J.y);
    out[BUD_ENSFLUX * NN + idx] = z == 0 ? v : out[BUD_ENSFLUX * NN + idx] + v;
    const cplx t1 = cmuli(q, kv * T.Ubg[z]), t3 = cmuli(ph, kv * T.Qy[z]);
    cplx r = cmake(-(J.x + t1.x + t3.x), -(J.y + t1.y + t3.y));
    if (z == 1 && T.rek != 0.0) {
      const double f = T.rek * (kv * kv + lv * lv);
      r.x += f * ph.x;
      r.y += f * ph.y;
    }
    tend[z * NN + idx] = r;
  } else if (RD == BR_DISS || RD == BR_PARAM_DISS) {
    cplx f[2] = {cmake(0.0, 0.0), cmake(0.0, 0.0)};
    if (RD == BR_PARAM_DISS) {
      // parameterization terms from dqh = rfft2(dq) = (A, B)
      const cplx dp0 = cmake(a00 * A.x + a01 * B.x, a00 * A.y + a01 * B.y), dp1 = cmake(a10 * A.x + a11 * B.x, a10 * A.y + a11 * B.y);
      const double wv2 = kv * kv + lv * lv, F = io.bud_F;
      out[BUD_PARAM_KE * NN + idx] = wv2 * m2 * (d1 * (p0.x * dp0.x + p0.y * dp0.y) + d2 * (p1.x * dp1.x + p1.y * dp1.y));
      const cplx t = csub(p0, p1), dtau = csub(dp0, dp1);
      out[BUD_PARAM_APE * NN + idx] = F * m2 * (t.x * dtau.x + t.y * dtau.y);
      if (!(io.bud_demean && idx == 0)) { f[0] = A; f[1] = B; }
    }
    // Dissspec / ENSDissspec from D_z = (filtr - 1)(qh_z + dt1 dqhdt_z + dt2 dqhdt_p_z + dt3 dqhdt_pp_z); ENSparamspec
    const double fm = T.filtr[idx] - 1.0, dt1 = io.dt1, dt2 = io.dt2, dt3 = io.dt3;
    double diss = 0.0, ensdiss = 0.0, ensparam = 0.0;
#pragma unroll
    for (int z = 0; z < 2; ++z) {
      const int j = z * NN + idx;
      const cplx q = z ? q1 : q0, ph = z ? p1 : p0;
      const cplx dd = cadd(tend[j], f[z]), a = dp[j], b = dpp[j];
      const cplx D = cmake(fm * (q.x + dt1 * dd.x + dt2 * a.x + dt3 * b.x), fm * (q.y + dt1 * dd.y + dt2 * a.y + dt3 * b.y));
      const double w = io.Hi_over_H[z];
      diss -= w * (ph.x * D.x + ph.y * D.y);
      ensdiss += w * (q.x * D.x + q.y * D.y);
      ensparam += w * (q.x * f[z].x + q.y * f[z].y);
    }
    out[BUD_DISS * NN + idx] = diss * io.bud_inv_dt * m2;
    out[BUD_ENSDISS * NN + idx] = ensdiss * io.bud_inv_dt * m2;
    out[BUD_ENSPARAM * NN + idx] = ensparam * m2;
  }
}

// ST = the stages (compile time: straight-line code), NB = points in flight per thread: all global / shared loads of a batch
// are issued before the first dependent instruction (the first version, one point at a time behind run-time stage tests, spent
// 46 % of its stall samples waiting for these loads).
template <int ST, int NB>
S64_PHASE void pointwise_phase(const Tables& T, const StepIO& io, int member, cplx* S, bool demean) {
  const MemberPtrs P(io, member);
  const double dkw = T.kv[1];                          // wavenumber spacing 2 pi / L: kv[k] = dk k, lv[l] = dk (l < 32 ? l : l - 64)
  constexpr bool kRead = (ST & (PW_TEND0 | PW_TEND1 | PW_FORCING | PW_STORE_QH)) != 0;
  constexpr bool kQh = (ST & PW_STORE_QH) == 0;
  constexpr bool kTend = (ST & (PW_TEND0 | PW_TEND1)) != 0, kUv = (ST & (PW_UV0 | PW_UV1)) != 0, kUpd = (ST & PW_UPDATE) != 0;
  constexpr bool kPsi = (ST & PW_PSI) != 0;            // inversion coefficients of layer 0 ride in aT, of layer 1 in aU
  constexpr int zT = (ST & PW_TEND1) ? 1 : 0, zU = (ST & (PW_UV1 | PW_PSI)) ? 1 : 0;
  static_assert(!kPsi || ST == PW_PSI, "PW_PSI runs alone");
  constexpr int NIT = (9 + NB - 1) / NB;
  const double dt1 = io.dt1, dt2 = io.dt2, dt3 = io.dt3;
#pragma unroll 1
  for (int b = 0; b < NIT; ++b) {
    int idx[NB], l[NB], k[NB];
    bool ok[NB];
    cplx A[NB], B[NB], q0[NB], q1[NB], dc0[NB], dc1[NB], dp0[NB], dp1[NB], dpp0[NB], dpp1[NB];
    double aT0[NB], aT1[NB], aU0[NB], aU1[NB], fl[NB], kv[NB], lv[NB];
    // ---- loads ----
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int i = threadIdx.x + kThreads * (b * NB + u);
      ok[u] = i < NN;
      idx[u] = ok[u] ? i : NN - 1;
      l[u] = idx[u] / NK;
      k[u] = idx[u] - l[u] * NK;
      if (kQh) { q0[u] = P.qh[idx[u]]; q1[u] = P.qh[NN + idx[u]]; }
      if (kTend || kPsi) { aT0[u] = T.a[(2 * zT) * NN + idx[u]]; aT1[u] = T.a[(2 * zT + 1) * NN + idx[u]]; }
      if (kUv || kPsi) { aU0[u] = T.a[(2 * zU) * NN + idx[u]]; aU1[u] = T.a[(2 * zU + 1) * NN + idx[u]]; }
      if (kUpd) {
        fl[u] = T.filtr[idx[u]];
        dc0[u] = P.d_cur[idx[u]];
        if (!(ST & PW_TEND1)) dc1[u] = P.d_cur[NN + idx[u]];
        dp0[u] = P.d_p[idx[u]]; dp1[u] = P.d_p[NN + idx[u]];
        dpp0[u] = P.d_pp[idx[u]]; dpp1[u] = P.d_pp[NN + idx[u]];
      }
      if (kQh) { kv[u] = dkw * (double)k[u]; lv[u] = dkw * (double)(l[u] < N / 2 ? l[u] : l[u] - N); }
    }
    // ---- arithmetic and stores ----
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (kRead) spec_read(S, l[u], k[u], A[u], B[u]);      // (shared memory: short latency, read where it is used)
      if (ST & PW_STORE_QH) {
        if (ok[u]) { P.qh[idx[u]] = A[u]; P.qh[NN + idx[u]] = B[u]; }
        continue;
      }
      cplx r = cmake(0.0, 0.0);
      if (kTend) {
        // dqhdt_z = -(ik uqh + il vqh + ikQy ph) (+ rek wv2 ph for the bottom layer)       (pyqg _do_advection / _do_friction)
        const cplx ph = cmake(aT0[u] * q0[u].x + aT1[u] * q1[u].x, aT0[u] * q0[u].y + aT1[u] * q1[u].y);
        const cplx t1 = cmuli(A[u], kv[u]), t2 = cmuli(B[u], lv[u]), t3 = cmuli(ph, kv[u] * T.Qy[zT]);
        r = cmake(-(t1.x + t2.x + t3.x), -(t1.y + t2.y + t3.y));
        if (zT == 1 && T.rek != 0.0) {
          const double f = T.rek * (kv[u] * kv[u] + lv[u] * lv[u]);
          r.x += f * ph.x;
          r.y += f * ph.y;
        }
        if (!kUpd && ok[u]) P.d_cur[zT * NN + idx[u]] = r;
      }
      if (kUpd) {
        // (+ rfft2(dq), mean removed for closure output) ; Adams-Bashforth update with the exponential filter (_forward_timestep)
        cplx f0 = cmake(0.0, 0.0), f1 = cmake(0.0, 0.0);
        if ((ST & PW_FORCING) && !(demean && idx[u] == 0)) { f0 = A[u]; f1 = B[u]; }
        const cplx dd0 = cadd(dc0[u], f0), dd1 = cadd((ST & PW_TEND1) ? r : dc1[u], f1);
        const cplx n0 = cmake(fl[u] * (q0[u].x + dt1 * dd0.x + dt2 * dp0[u].x + dt3 * dpp0[u].x),
                              fl[u] * (q0[u].y + dt1 * dd0.y + dt2 * dp0[u].y + dt3 * dpp0[u].y));
        const cplx n1 = cmake(fl[u] * (q1[u].x + dt1 * dd1.x + dt2 * dp1[u].x + dt3 * dpp1[u].x),
                              fl[u] * (q1[u].y + dt1 * dd1.y + dt2 * dp1[u].y + dt3 * dpp1[u].y));
        if (ok[u]) {
          P.d_cur[idx[u]] = dd0; P.d_cur[NN + idx[u]] = dd1;
          P.qh[idx[u]] = n0; P.qh[NN + idx[u]] = n1;
          spec_write(S, l[u], k[u], n0, n1);           // input of the final inverse transform: q = irfft2(qh)
        }
      }
      if (kUv && ok[u]) {
        // uh = -il ph, vh = ik ph                                                                       (pyqg _invert)
        const cplx ph = cmake(aU0[u] * q0[u].x + aU1[u] * q1[u].x, aU0[u] * q0[u].y + aU1[u] * q1[u].y);
        spec_write(S, l[u], k[u], cmake(lv[u] * ph.y, -lv[u] * ph.x), cmake(-kv[u] * ph.y, kv[u] * ph.x));
      }
      if ((ST & PW_LOAD_QH) && ok[u]) spec_write(S, l[u], k[u], q0[u], q1[u]);
      if (kPsi && ok[u]) {
        // ph_z = a[z][0] qh_0 + a[z][1] qh_1                                                           (pyqg _invert)
        const cplx p0 = cmake(aT0[u] * q0[u].x + aT1[u] * q1[u].x, aT0[u] * q0[u].y + aT1[u] * q1[u].y);
        const cplx p1 = cmake(aU0[u] * q0[u].x + aU1[u] * q1[u].x, aU0[u] * q0[u].y + aU1[u] * q1[u].y);
        spec_write(S, l[u], k[u], p0, p1);
        if (io.ph_out) { cplx* o = io.ph_out + (long long)member * 2 * NN; o[idx[u]] = p0; o[NN + idx[u]] = p1; }
      }
    }
  }
}

// One round: [inverse 2-D transform of the spectra in S] -> physical stage in registers -> [forward 2-D transform into S].
template <bool BUD = false>      // BUD: the physical stages of the budget program (PH_SCR_*, PH_ANOM*) are compiled in
S64_PHASE void round(bool has_inv, int phys, bool has_fwd, const Tables& T, const StepIO& io, int member, cplx* buf, const cplx* tw) {
  const Geo g;
  const MemberPtrs P(io, member);
  cplx v[16];
  const double s = T.inv_M;
#pragma unroll 1
  for (int hp = has_inv ? 0 : 2; hp < (has_fwd ? 4 : 2); ++hp) {
    const bool ypass = hp == 0 || hp == 3;
    // ---- stage in ----
    if (hp == 0) {                                     // columns of S, l = 4 j + t  (packed lanes: build C with the c2r convention)
      if (!g.packed) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = buf[(4 * j + g.yt) * PITCH + g.col];
      } else {
        // C(l) = sym(X0)(l) + i sym(X1)(l), sym(X)(l) = (X(l) + conj X(-l)) / 2: the imaginary parts of the k = 0, 32 columns
        // are dropped after the l transform, exactly what a c2r transform does
        const int c0 = g.isB ? 35 : 34, c1 = g.isB ? 33 : 32;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int l = 4 * j + g.yt, ln = (N - l) & (N - 1);
          const cplx x0 = buf[l * PITCH + c0], n0 = buf[ln * PITCH + c0], x1 = buf[l * PITCH + c1], n1 = buf[ln * PITCH + c1];
          v[j] = cmake(0.5 * (x0.x + n0.x - x1.y + n1.y), 0.5 * (x0.y - n0.y + x1.x + n1.x));
        }
      }
      conj16(v);
    } else if (hp == 1) {                              // rows of T': W_y(kx) from A_y, B_y; kx = 4 j + t
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int kx = 4 * j + g.xt;                   // j < 8 <=> kx < 32
        const bool dc = j == 0 && g.xt == 0, nyq = j == 8 && g.xt == 0;
        const int c = nyq ? 0 : (j < 8 ? kx : 64 - kx);
        const cplx A = buf[g.y * PITCH + c], B = buf[g.y * PITCH + B_OFF + c];
        cplx w = j < 8 ? cmake(A.x - B.y, A.y + B.x) : cmake(A.x + B.y, B.x - A.y);   // A + i B ; conj A + i conj B
        if (j == 0) w = sel(dc, cmake(A.x, B.x), w);    // kx = 0:  Re C0 + i Re C1
        if (j == 8) w = sel(nyq, cmake(A.y, B.y), w);   // kx = 32: Im C0 + i Im C1
        v[j] = cmake(w.x, -w.y);                        // conjugated: the inverse transform is conj(FFT(conj .))
      }
    } else if (hp == 3) {                              // columns of T: split W_y into A_y / B_y (packed lanes: C0 / C1), y = 4 j + t
      const int c1 = g.packed ? 0 : g.slot, c2 = g.packed ? 32 : 64 - g.slot;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int yy = 4 * j + g.yt;
        const cplx Pw = buf[yy * PITCH + c1], Q = buf[yy * PITCH + c2];
        const cplx nA = cmake(0.5 * (Pw.x + Q.x), 0.5 * (Pw.y - Q.y)), nB = cmake(0.5 * (Pw.y + Q.y), 0.5 * (Q.x - Pw.x));
        const cplx pA = cmake(Pw.x, Q.x), pB = cmake(Pw.y, Q.y);
        v[j] = g.packed ? (g.isB ? pB : pA) : (g.isB ? nB : nA);
      }
    } else if (!has_inv) {                             // hp == 2 of a forward-only round: load the real pair, x = 4 j + t
      const double* f0 = phys == PH_LOAD_DQ ? P.dq : P.q;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = g.y * N + 4 * j + g.xt;
        v[j] = cmake(f0[i], f0[NPIX + i]);
        if (phys == PH_LOAD_Q && P.cnn_x) { P.cnn_x[i] = div_by((float)v[j].x, io.x_std[0], io.x_inv[0]); P.cnn_x[NPIX + i] = div_by((float)v[j].y, io.x_std[1], io.x_inv[1]); }
      }
    }                                                  // (else hp == 2: v comes from the physical stage of hp == 1)
    // ---- the line transform ----
    fft64(v, ypass ? g.yt : g.xt, ypass ? 8 : 1, ypass ? 16 : 2, tw);
    // ---- stage out ----
    if (hp == 0) {                                     // A_y / B_y at y = t + 4 i + 16 k2 -> T'
      __syncthreads();                                 // every column of S has been read
#pragma unroll
      for (int m = 0; m < 16; ++m) buf[(g.yt + 4 * (m >> 2) + 16 * (m & 3)) * PITCH + g.col] = cmake(v[m].x, -v[m].y);
      __syncthreads();
    } else if (hp == 1) {                              // physical row y, x = t + 4 i + 16 k2 (conjugate back, scale)
      conj16(v);
      if (phys >= PH_STORE_UV0) {
        // PROG_INVERT: (u, v) of layer z, or (psi_0, psi_1), scaled and stored (null outputs are skipped)
        const long long mo = (long long)member * 2 * NPIX;
        double* d0 = phys == PH_STORE_P ? io.p_out : io.u_out;
        double* d1 = phys == PH_STORE_P ? io.p_out : io.v_out;
        if (d0) d0 += mo + (phys == PH_STORE_UV1 ? NPIX : 0);
        if (d1) d1 += mo + (phys == PH_STORE_UV0 ? 0 : NPIX);
        if (BUD && phys >= PH_SCR_STORE01) {           // xi_0, xi_1 -> scratch fields 0, 1;  tau -> scratch field 2
          d0 = io.bud_scr + ((long long)member * 3 + (phys == PH_SCR_STORE2 ? 2 : 0)) * NPIX;
          d1 = phys == PH_SCR_STORE2 ? nullptr : d0 + NPIX;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int i = g.y * N + g.xt + 4 * (m >> 2) + 16 * (m & 3);
          if (d0) d0[i] = v[m].x * s;
          if (d1) d1[i] = v[m].y * s;
        }
      } else if (phys == PH_EMIT) {
        // q = irfft2(qh) (+ the fp32 normalised closure input  x_scale.normalize(m.q.astype('float32')))
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int i = g.y * N + g.xt + 4 * (m >> 2) + 16 * (m & 3);
          const double q0 = v[m].x * s, q1 = v[m].y * s;
          P.q[i] = q0;
          P.q[NPIX + i] = q1;
          if (P.cnn_x) { P.cnn_x[i] = div_by((float)q0, io.x_std[0], io.x_inv[0]); P.cnn_x[NPIX + i] = div_by((float)q1, io.x_std[1], io.x_inv[1]); }
        }
      } else {
        // (u + Ubg) q + i v q                                            (pyqg _do_advection, physical-space products)
        const int z = phys == PH_PRODUCTS1 ? 1 : 0;
        const double* qz = P.q + z * NPIX + g.y * N + g.xt;
        double U = T.Ubg[z];
        if (BUD) {                                     // budget products use the ANOMALY velocities: u f + i v f
          U = 0.0;
          if (phys == PH_ANOM0 || phys == PH_ANOM1) qz = P.q + (phys - PH_ANOM0) * NPIX + g.y * N + g.xt;
          else qz = io.bud_scr + ((long long)member * 3 + (phys - PH_SCR_PROD0)) * NPIX + g.y * N + g.xt;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double qq = qz[4 * (m >> 2) + 16 * (m & 3)];
          v[m] = cmake((v[m].x * s + U) * qq, (v[m].y * s) * qq);
        }
        out_to_in_naming(v);
      }
    } else if (hp == 2) {                              // W_y(kx), kx = t + 4 i + 16 k2 -> T
      __syncthreads();                                 // (T' fully read by every x quad)
#pragma unroll
      for (int m = 0; m < 16; ++m) buf[g.y * PITCH + g.xt + 4 * (m >> 2) + 16 * (m & 3)] = v[m];
      __syncthreads();
    } else {                                           // hp == 3: Ahat / Bhat / packed C at l = t + 4 i + 16 k2 -> S
      __syncthreads();
#pragma unroll
      for (int m = 0; m < 16; ++m) buf[(g.yt + 4 * (m >> 2) + 16 * (m & 3)) * PITCH + g.col] = v[m];
      __syncthreads();
    }
  }
}


#ifndef S64_HELPERS_ONLY      // (spectral_cl.cuh reuses the building blocks above)
// prog: PROG_STEP, PROG_STEP_DQ, PROG_STEP_DQ_RAW, PROG_SET_Q, PROG_C2R, PROG_ADVECT (the tendencies of both layers into d_cur, no
// update), PROG_INVERT (u, v of both layers, psi of both layers, the spectral psi -> io.u_out, v_out, p_out, ph_out)
__global__ void __launch_bounds__(kThreads, 2) qg_step64_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io,
                                                                int prog, int members) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + N * PITCH;
  for (int i = threadIdx.x; i < N; i += kThreads) tw[i] = T.tw[i];
  const bool with_dq = prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW;
  const bool demean = prog == PROG_STEP_DQ;
  const int nrounds = prog == PROG_C2R ? 1 : prog == PROG_SET_Q ? 2 : (with_dq ? 4 : 3);
  for (int m = blockIdx.x; m < members; m += gridDim.x) {
    if (S64_PREFETCH && prog != PROG_INVERT && prog != PROG_ADVECT) {   // pull everything this member's step reads from HBM into L2 now: the pointwise phases then see L2 latency, not DRAM's
      const MemberPtrs P(io, m);
      const int nspec = 2 * NN * (int)sizeof(cplx) / 128, nphys = 2 * NPIX * (int)sizeof(double) / 128;
      for (int i = threadIdx.x; i < nspec; i += kThreads) {
        prefetch_l2(reinterpret_cast<const char*>(P.qh) + 128 * i);
        if (prog != PROG_SET_Q && prog != PROG_C2R) {
          prefetch_l2(reinterpret_cast<const char*>(P.d_p) + 128 * i);
          prefetch_l2(reinterpret_cast<const char*>(P.d_pp) + 128 * i);
        }
      }
      if (prog != PROG_C2R)
        for (int i = threadIdx.x; i < nphys; i += kThreads) {
          prefetch_l2(reinterpret_cast<const char*>(P.q) + 128 * i);
          if (P.dq) prefetch_l2(reinterpret_cast<const char*>(P.dq) + 128 * i);
        }
    }
    __syncthreads();                                   // twiddles staged / the previous member's last round has left the buffer
#pragma unroll 1
    for (int r = 0; r < nrounds; ++r) {
      // the program: which pointwise stages run before round r, and what round r is
      int pw = 0, phys = PH_EMIT;
      bool inv = true, fwd = true, rnd = true;
      if (prog == PROG_C2R) { pw = PW_LOAD_QH; fwd = false; }
      else if (prog == PROG_SET_Q) {
        if (r == 0) { inv = false; phys = PH_LOAD_Q; } else { pw = PW_STORE_QH; rnd = false; }
      } else if (prog == PROG_INVERT) {
        fwd = false;
        if (r == 0) { pw = PW_UV0; phys = PH_STORE_UV0; }
        else if (r == 1) { pw = PW_UV1; phys = PH_STORE_UV1; }
        else { pw = PW_PSI; phys = PH_STORE_P; rnd = io.p_out != nullptr; if (!rnd && !io.ph_out) pw = 0; }
      } else if (r == 0) { pw = PW_UV0; phys = PH_PRODUCTS0; }
      else if (r == 1) { pw = PW_TEND0 | PW_UV1; phys = PH_PRODUCTS1; }
      else if (r == 2 && prog == PROG_ADVECT) { pw = PW_TEND1; rnd = false; }
      else if (r == 2 && with_dq) { pw = PW_TEND1; inv = false; phys = PH_LOAD_DQ; }
      else if (r == 2) { pw = PW_TEND1 | PW_UPDATE; fwd = false; }
      else { pw = PW_FORCING | PW_UPDATE; fwd = false; }
      if (pw) {
        switch (pw) {
          case PW_UV0: pointwise_phase<PW_UV0, S64_NB_UV>(T, io, m, buf, demean); break;
          case PW_UV1: pointwise_phase<PW_UV1, S64_NB_UV>(T, io, m, buf, demean); break;
          case PW_PSI: pointwise_phase<PW_PSI, S64_NB_UV>(T, io, m, buf, demean); break;
          case PW_TEND0 | PW_UV1: pointwise_phase<PW_TEND0 | PW_UV1, S64_NB_LIGHT>(T, io, m, buf, demean); break;
          case PW_TEND1: pointwise_phase<PW_TEND1, S64_NB_LIGHT>(T, io, m, buf, demean); break;
          case PW_TEND1 | PW_UPDATE: pointwise_phase<PW_TEND1 | PW_UPDATE, S64_NB_UPD>(T, io, m, buf, demean); break;
          case PW_FORCING | PW_UPDATE: pointwise_phase<PW_FORCING | PW_UPDATE, S64_NB_UPD>(T, io, m, buf, demean); break;
          case PW_STORE_QH: pointwise_phase<PW_STORE_QH, S64_NB_UV>(T, io, m, buf, demean); break;
          default: pointwise_phase<PW_LOAD_QH, S64_NB_UV>(T, io, m, buf, demean); break;
        }
        __syncthreads();
      }
      if (rnd) round(inv, phys, fwd, T, io, m, buf, tw);
      if (prog == PROG_INVERT) __syncthreads();        // an inverse-only round ends with reads of T': the next phase writes S over it
    }
  }
}

// ---- PROG_BUDGET (qg_core.cuh run_program: 7 inverse + 5 forward packed transforms, +1 with a forcing) ---------------------------
// RD: what the phase does with the spectra of the last forward transform; WR: the pair it leaves in S for the next inverse one.
template <int RD, int WR, int NB>
S64_PHASE void budget_phase(const Tables& T, const StepIO& io, int member, cplx* S) {
  const long long mo = (long long)member * 2 * NN;
  const cplx* qh = io.qh + mo;
  double* out = io.bud_out + (long long)member * kBudgetTerms * NN;
  cplx* tend = io.bud_tend + mo;
  const cplx *dp = io.d_p + mo, *dpp = io.d_pp + mo;
  const double dkw = T.kv[1], d1 = io.Hi_over_H[0], d2 = io.Hi_over_H[1];
  constexpr bool kRead = RD != BR_NONE && RD != BR_DISS;
  constexpr int NIT = (9 + NB - 1) / NB;
#pragma unroll 1
  for (int b = 0; b < NIT; ++b) {
    int idx[NB], l[NB], k[NB];
    bool ok[NB];
    cplx q0[NB], q1[NB];
    double a00[NB], a01[NB], a10[NB], a11[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int i = threadIdx.x + kThreads * (b * NB + u);
      ok[u] = i < NN;
      idx[u] = ok[u] ? i : NN - 1;
      l[u] = idx[u] / NK;
      k[u] = idx[u] - l[u] * NK;
      q0[u] = qh[idx[u]]; q1[u] = qh[NN + idx[u]];
      a00[u] = T.a[idx[u]]; a01[u] = T.a[NN + idx[u]]; a10[u] = T.a[2 * NN + idx[u]]; a11[u] = T.a[3 * NN + idx[u]];
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const double kv = dkw * (double)k[u], lv = dkw * (double)(l[u] < N / 2 ? l[u] : l[u] - N);
      const cplx p0 = cmake(a00[u] * q0[u].x + a01[u] * q1[u].x, a00[u] * q0[u].y + a01[u] * q1[u].y);
      const cplx p1 = cmake(a10[u] * q0[u].x + a11[u] * q1[u].x, a10[u] * q0[u].y + a11[u] * q1[u].y);
      cplx A = cmake(0.0, 0.0), B = cmake(0.0, 0.0);
      if (kRead) spec_read(S, l[u], k[u], A, B);
      if (RD != BR_NONE && ok[u])
        bud_read<RD>(T, io, out, tend, dp, dpp, NN, idx[u], kv, lv, q0[u], q1[u], p0, p1, a00[u], a01[u], a10[u], a11[u], A, B);
      if (WR != BWR_NONE && ok[u]) {
        cplx oA, oB;
        bud_build<WR>(p0, p1, kv, lv, d1, d2, oA, oB);
        spec_write(S, l[u], k[u], oA, oB);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2) qg_budget64_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io, int members) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + N * PITCH;
  for (int i = threadIdx.x; i < N; i += kThreads) tw[i] = T.tw[i];
  const bool has_dq = io.dq != nullptr;
  for (int m = blockIdx.x; m < members; m += gridDim.x) {
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
      // phase r (consume the last forward transform, build the next pair), then round r: inverse -> physical stage -> [forward]
      int phys = PH_EMIT;
      bool inv = true, fwd = true, rnd = true;
      switch (r) {
        case 0: budget_phase<BR_NONE, BWR_XI, 2>(T, io, m, buf); phys = PH_SCR_STORE01; fwd = false; break;      // xi_0, xi_1
        case 1: budget_phase<BR_NONE, BWR_UV0, 2>(T, io, m, buf); phys = PH_SCR_PROD0; break;                    // u_0 xi_0 + i v_0 xi_0
        case 2: budget_phase<BR_KEFLUX0, BWR_UV1, 2>(T, io, m, buf); phys = PH_SCR_PROD1; break;
        case 3: budget_phase<BR_KEFLUX1, BWR_TAU, 2>(T, io, m, buf); phys = PH_SCR_STORE2; fwd = false; break;   // tau
        case 4: budget_phase<BR_NONE, BWR_UVBT, 2>(T, io, m, buf); phys = PH_SCR_PROD2; break;                   // u_bt tau + i v_bt tau
        case 5: budget_phase<BR_APEFLUX, BWR_UV0, 2>(T, io, m, buf); phys = PH_ANOM0; break;                     // u_0 q_0 + i v_0 q_0
        case 6: budget_phase<BR_ENS0, BWR_UV1, 2>(T, io, m, buf); phys = PH_ANOM1; break;
        case 7: budget_phase<BR_ENS1, BWR_NONE, 2>(T, io, m, buf); inv = false; phys = PH_LOAD_DQ; rnd = has_dq; break;   // rfft2(dq)
        default:
          if (has_dq) budget_phase<BR_PARAM_DISS, BWR_NONE, 1>(T, io, m, buf);
          else budget_phase<BR_DISS, BWR_NONE, 1>(T, io, m, buf);
          rnd = false;
          break;
      }
      __syncthreads();
      if (rnd) round<true>(inv, phys, fwd, T, io, m, buf, tw);
      if (!fwd) __syncthreads();                       // an inverse-only round ends with reads of T': the next phase writes S over it
    }
  }
}

#endif

}  // namespace s64
}  // namespace qgb
