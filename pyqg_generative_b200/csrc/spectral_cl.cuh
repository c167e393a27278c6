// spectral_cl.cuh -- register-FFT fused spectral time step for nx = 128 and 256 on a thread-block CLUSTER, sm_100a.
//
// Same step as spectral64.cuh (pyqg Model._step_forward, see qg_core.cuh), same building blocks (line transforms in registers,
// 16 complex numbers per thread, the lanes of a line exchanging data with warp shuffles; two real fields per complex transform,
// split into half-plane spectra between the passes; pointwise phases over half-plane points), scaled out:
//
//   nx = 128: a line is transformed by G = 8 lanes (16-point FFT, 3 xor rounds, two radix-8 butterflies), CL = 2 CTAs per member
//   nx = 256: G = 16 lanes (16-point FFT, 4 xor rounds, 16-point FFT),                                   CL = 8 CTAs per member
//
// Every CTA has (nx / CL) * G = 512 threads.  In the x pass it owns nx / CL rows; in the y pass nx / (2 CL) columns of each of the
// two half-plane spectra; in the pointwise phases the wavenumbers k of those columns (CTA 0 also k = nx / 2).  The transposition
// between the passes goes through DISTRIBUTED SHARED MEMORY: the x-pass lanes store W_y(kx) with st.shared::cluster straight into
// the buffer of the CTA that owns column kx (and back for the inverse transform), separated by cluster barriers -- the first-
// generation cluster kernel (qg_core.cuh tiled_pass) bounced the whole field through an L2 scratch four times per transform and
// ran every radix stage as a shared-memory pass (0.07 of the HBM roofline at 256^2, 2.5x the algorithmic DRAM traffic).
// The half-plane spectra never leave the CTA: no Hermitian extension is built and the pointwise phases read and write global
// memory in runs of nx / (2 CL) consecutive wavenumbers.
#pragma once
#include <cuda_runtime.h>

#include "spectral64.cuh"

namespace qgb {
namespace scl {

using s64::conj16;
using s64::dft4;
using s64::exchange_pair;
using s64::fft16;
using s64::mulw16;
using s64::sel;
using s64::bud_build;
using s64::bud_read;
using s64::BR_NONE; using s64::BR_KEFLUX0; using s64::BR_KEFLUX1; using s64::BR_APEFLUX; using s64::BR_ENS0; using s64::BR_ENS1;
using s64::BR_DISS; using s64::BR_PARAM_DISS;
using s64::BWR_NONE; using s64::BWR_XI; using s64::BWR_UV0; using s64::BWR_UV1; using s64::BWR_TAU; using s64::BWR_UVBT;

template <int N_, int G_, int CL_>
struct Cfg {
  static constexpr int N = N_, G = G_, CL = CL_;
  static constexpr int NK = N / 2 + 1, NN = N * NK, NPIX = N * N, H = N / 2;
  static constexpr int RPC = N / CL;            // rows of the x pass per CTA
  static constexpr int SPC = N / (2 * CL);      // columns (wavenumbers k) of each spectrum per CTA
  static constexpr int LPW = 32 / G;            // lines per warp
  static constexpr int I = 16 / G;              // k1 values per lane after the exchange
  static constexpr int kThreads = RPC * G;
  static constexpr int PS = 2 * SPC + 5;        // row pitch of T / S: plus | minus (A | B) columns + 4 pad columns; odd
  static constexpr int PT = N + 1;              // row pitch of T' (RPC rows): A columns 0..N/2-1, B columns N/2..N-1; odd
  static constexpr int kBuf = (N * PS > RPC * PT) ? N * PS : RPC * PT;
  static constexpr size_t kSmemBytes = (size_t)(kBuf + N + 1) * sizeof(cplx);   // buffer, twiddles, hand-over mbarrier
  // bytes that land in one CTA's buffer per transposition (T' = RPC rows x N columns, T = N rows x 2 SPC columns: N^2 / CL numbers)
  static constexpr uint32_t kHandoverBytes = (uint32_t)(N * (N / CL) * sizeof(cplx));
  // read-after-write hand-over of a transposition: plain remote stores and a release / acquire cluster barrier (shipped), or
  // -DSCL_ASYNC_HANDOVER: st.async + the destination's mbarrier (complete_tx).  The second removes the barrier's MEMBAR.ALL.GPU from
  // the loop but every 16-byte store then updates the transaction count of ONE mbarrier per destination CTA: measured on one box
  // (scripts/ab_cluster_sync.sh) 3 % slower at 256^2 and 30 % slower at 128^2 (profiles/r2_cluster256_ncu.md).
#ifdef SCL_ASYNC_HANDOVER
  static constexpr bool kAsyncHandover = true;
#else
  static constexpr bool kAsyncHandover = false;
#endif
  static_assert(kHandoverBytes < (1u << 20), "mbarrier tx-count range");
  static constexpr int kCtasPerSm = kThreads <= 256 ? 2 : 1;   // 128 registers per thread: 512 threads per SM either way
  // split the write-after-read cluster barriers around the line transform (cluster_arrive_exec / cluster_wait_exec): measured on
  // one box (scripts/ab_cluster_sync.sh) +4.3 % at 256^2 (8 / 16 CTAs per member), -4.2 % at 128^2 (2 CTAs: the barrier is
  // cheap there and the pinned arrive costs scheduling freedom)
#ifdef SCL_JOINT_SYNC
  static constexpr bool kSplitSync = false;
#else
  static constexpr bool kSplitSync = CL >= 8;
#endif
  static_assert((kThreads == 512 || kThreads == 256) && (G == 8 || G == 16) && 16 * G == N, "unsupported geometry");
};

#define SCL_INL __device__ __forceinline__
// half-plane points whose global loads are in flight together, per thread and pointwise phase (DRAM latency is the limiter of
// these kernels: one CTA per SM, 16 warps)
#ifndef SCL_PREFETCH
#define SCL_PREFETCH 0
#endif
#ifndef SCL_NB_UV
#define SCL_NB_UV 2
#endif
#ifndef SCL_NB_TEND
#define SCL_NB_TEND 2
#endif
#ifndef SCL_NB_UPD
#define SCL_NB_UPD 1
#endif


SCL_INL uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// Execution-only cluster barrier for the write-after-read hazards ("every CTA has finished READING its buffer, peers may now
// overwrite it"): the loads were consumed by arithmetic that precedes the barrier in program order, so no release fence is
// needed -- the fence of the full barrier waits for every outstanding global store of the pointwise phases (15 % of the stall
// samples of the 256^2 kernel were membar stalls, profiles/r2_cluster256_ncu.md).
SCL_INL void cluster_sync_exec() {
#ifdef SCL_STRICT_SYNC
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
#endif
}
SCL_INL void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The same execution-only barrier SPLIT around the line transform: a CTA arrives as soon as its own reads of the buffer have
// been consumed (the first 16-point FFT of fftN has every loaded value as an operand) and waits only before its first remote
// store, so barrier latency and the skew between the CTAs of a member hide behind the transform (and, for the x pass, behind the
// physical-space stage and a second transform) instead of following it.  ``anchor`` is a value computed from every loaded
// number: the empty volatile asm pins that arithmetic (and so the completed loads) ahead of the arrive in program order.
SCL_INL void cluster_arrive_exec(double& anchor) {
  asm volatile("" : "+d"(anchor));
#ifdef SCL_STRICT_SYNC
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
}
SCL_INL void cluster_arrive_exec() {       // (after a __syncthreads: the CTA's shared-memory reads are complete)
#ifdef SCL_STRICT_SYNC
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
}
SCL_INL void cluster_wait_exec() {
#ifdef SCL_STRICT_SYNC
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#else
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
#endif
}
// store a complex number into the shared memory of CTA ``rank`` of this cluster at the same offset as the local address
SCL_INL void st_cluster(const cplx* local, uint32_t rank, cplx v) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(ra), "d"(v.x), "d"(v.y) : "memory");
}

// Hand-over of a transposition through the DESTINATION's mbarrier: every remote store is an st.async that completes 16 bytes of
// the transaction count of the mbarrier at the same offset in the destination CTA; the destination arms its mbarrier with the
// N^2 / CL numbers it is about to receive and waits for the phase locally -- no cluster-wide release / acquire barrier.
SCL_INL void st_cluster_async(const cplx* local, const uint64_t* local_bar, uint32_t rank, cplx v) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local), lb = (uint32_t)__cvta_generic_to_shared(local_bar);
  uint32_t ra, rb;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(lb), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(ra), "d"(v.x), "d"(v.y), "r"(rb)
               : "memory");
}
SCL_INL void handover_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SCL_INL void handover_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
SCL_INL void handover_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "SCL_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra SCL_WAIT_DONE;\n\t"
      "bra SCL_WAIT_LOOP;\n\t"
      "SCL_WAIT_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity)
      : "memory");
}

// 8-point DFT, natural order in and out (forward kernel)
SCL_INL void dft8(cplx& x0, cplx& x1, cplx& x2, cplx& x3, cplx& x4, cplx& x5, cplx& x6, cplx& x7) {
  cplx e0 = x0, e1 = x2, e2 = x4, e3 = x6, o0 = x1, o1 = x3, o2 = x5, o3 = x7;
  dft4<false>(e0, e1, e2, e3);
  dft4<false>(o0, o1, o2, o3);
  o1 = mulw16<2, false>(o1);
  o2 = mulw16<4, false>(o2);
  o3 = mulw16<6, false>(o3);
  x0 = cadd(e0, o0); x4 = csub(e0, o0);
  x1 = cadd(e1, o1); x5 = csub(e1, o1);
  x2 = cadd(e2, o2); x6 = csub(e2, o2);
  x3 = cadd(e3, o3); x7 = csub(e3, o3);
}

// N = 16 G point DFT of a line distributed over G lanes t = 0..G-1 (lanes lane0 + t of a warp), forward kernel:
//   in  v[j] = x[G j + t]            out v[G i + k2] = X[t + G i + 16 k2]   (i < 16 / G, k2 < G)
template <class C>
SCL_INL void fftN(cplx (&v)[16], int t, const cplx* tw, bool arrive = false) {
  constexpr int G = C::G, I = C::I, N = C::N;
  fft16<false>(v);                        // v[k1] = sum_j x[G j + t] w16^{j k1};  k1 = dest + G i sits at index G i + dest
  if (C::kSplitSync && arrive) {          // (uniform) v[0] = sum of all 16 inputs: every value staged in has arrived in registers
    cluster_arrive_exec(v[0].x);
    asm volatile("" : "+d"(v[0].y));
  }
  // G x G block transpose across the lanes of the line: one xor round per bit of the lane index
#pragma unroll
  for (int p = (G == 16 ? 3 : 2); p >= 0; --p) {
    const bool bit = (t >> p) & 1;
#pragma unroll
    for (int i = 0; i < I; ++i)
#pragma unroll
      for (int q = 0; q < G; ++q)
        if (!(q & (1 << p))) exchange_pair(v[G * i + q], v[G * i + (q | (1 << p))], bit, 1 << p);
  }
  // afterwards v[G i + n2] = Z_{n2}[t + G i];  twiddle w_N^{n2 (t + G i)}
#pragma unroll
  for (int i = 0; i < I; ++i)
#pragma unroll
    for (int n2 = 1; n2 < G; ++n2) v[G * i + n2] = cmul(v[G * i + n2], tw[(n2 * (t + G * i)) & (N - 1)]);
  if (G == 16) fft16<false>(v);
  else {
#pragma unroll
    for (int i = 0; i < I; ++i) dft8(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7]);
  }
}
// rename the output naming v[G i + k2] (index t + G i + 16 k2 = G (i + I k2) + t) to the input naming v[j], j = i + I k2
template <class C>
SCL_INL void out_to_in_naming(cplx (&v)[16]) {
  if (C::I == 1) return;
  cplx tmp[16];
#pragma unroll
  for (int i = 0; i < C::I; ++i)
#pragma unroll
    for (int k2 = 0; k2 < C::G; ++k2) tmp[i + C::I * k2] = v[C::G * i + k2];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = tmp[j];
}
// index carried by register m of the output naming, for lane t
template <class C>
SCL_INL int out_index(int m, int t) { return t + C::G * (m / C::G) + 16 * (m % C::G); }

// ---- geometry -----------------------------------------------------------------------------------------------------------
// lane = G * (line within the warp) + t.  x pass: row y = rank * RPC + line.  y pass: line = isB * SPC + sl, column slot
// rank * SPC + sl of spectrum A (isB = 0) or B; slot 0 (CTA 0) = the packed columns k = 0 and k = N / 2.
// Shared buffer, used in turn as
//   T   [N rows y][PS]      W_y(c) at column sl, W_y(-c) at column SPC + sl for the CTA's slots c        (physical -> spectral)
//   S   [N rows l][PS]      Ahat(l, c) at sl, Bhat(l, c) at SPC + sl; pad columns 2 SPC .. 2 SPC + 3 = raw A / B at k = N/2, A / B at k = 0
//   T'  [RPC rows y][PT]    A_y(c) at column c, B_y(c) at N / 2 + c for ALL slots c                      (spectral -> physical)
template <class C>
struct Geo {
  int t, line, rank, y, sl, isB, slot, colS;
  bool packed;
  __device__ Geo() {
    const int lane = threadIdx.x & 31;
    t = lane % C::G;
    line = (threadIdx.x >> 5) * C::LPW + lane / C::G;
    rank = (int)cluster_rank();
    y = rank * C::RPC + line;
    sl = line % C::SPC;
    isB = line / C::SPC;
    slot = rank * C::SPC + sl;
    packed = slot == 0;
    colS = sl + (isB ? C::SPC : 0);
  }
};

template <class C>
struct MemberPtrs {
  cplx* qh; double* q; cplx* d_cur; const cplx* d_p; const cplx* d_pp; const double* dq; float* cnn_x;
  __device__ MemberPtrs(const StepIO& io, int m) {
    qh = io.qh + (long long)m * 2 * C::NN;
    q = io.q + (long long)m * 2 * C::NPIX;
    d_cur = io.d_cur ? io.d_cur + (long long)m * 2 * C::NN : nullptr;
    d_p = io.d_p ? io.d_p + (long long)m * 2 * C::NN : nullptr;
    d_pp = io.d_pp ? io.d_pp + (long long)m * 2 * C::NN : nullptr;
    dq = io.dq ? io.dq + (long long)m * 2 * C::NPIX : nullptr;
    cnn_x = io.cnn_x ? io.cnn_x + (long long)m * io.cnn_mstride : nullptr;
  }
};

// half-plane spectra of this CTA's wavenumbers in S (kk = k - rank * SPC; k = N / 2 belongs to CTA 0)
template <class C>
SCL_INL void spec_read(const cplx* S, int l, int k, int kk, cplx& A, cplx& B) {
  if (k != 0 && k != C::H) { A = S[l * C::PS + kk]; B = S[l * C::PS + C::SPC + kk]; return; }
  const int ln = (C::N - l) & (C::N - 1);
  const cplx c0 = S[l * C::PS], n0 = S[ln * C::PS], c1 = S[l * C::PS + C::SPC], n1 = S[ln * C::PS + C::SPC];
  if (k == 0) {
    A = cmake(0.5 * (c0.x + n0.x), 0.5 * (c0.y - n0.y));
    B = cmake(0.5 * (c1.x + n1.x), 0.5 * (c1.y - n1.y));
  } else {
    A = cmake(0.5 * (c0.y + n0.y), 0.5 * (n0.x - c0.x));
    B = cmake(0.5 * (c1.y + n1.y), 0.5 * (n1.x - c1.x));
  }
}
template <class C>
SCL_INL void spec_write(cplx* S, int l, int k, int kk, cplx A, cplx B) {
  const int ca = k == C::H ? 2 * C::SPC : (k == 0 ? 2 * C::SPC + 2 : kk);
  const int cb = k == C::H ? 2 * C::SPC + 1 : (k == 0 ? 2 * C::SPC + 3 : C::SPC + kk);
  S[l * C::PS + ca] = A;
  S[l * C::PS + cb] = B;
}

// pointwise stages on the CTA's half-plane points (see spectral64.cuh pointwise_phase); ST compile time, NB points in flight
template <class C, int ST, int NB>
SCL_INL void pointwise_phase(const Tables& T, const StepIO& io, int member, int rank, cplx* S, bool demean) {
  constexpr bool kRead = (ST & (PW_TEND0 | PW_TEND1 | PW_FORCING | PW_STORE_QH)) != 0;
  constexpr bool kQh = (ST & PW_STORE_QH) == 0;
  constexpr bool kTend = (ST & (PW_TEND0 | PW_TEND1)) != 0, kUv = (ST & (PW_UV0 | PW_UV1)) != 0, kUpd = (ST & PW_UPDATE) != 0;
  constexpr bool kPsi = (ST & PW_PSI) != 0;            // inversion coefficients of layer 0 ride in aT, of layer 1 in aU
  constexpr int zT = (ST & PW_TEND1) ? 1 : 0, zU = (ST & (PW_UV1 | PW_PSI)) ? 1 : 0;
  static_assert(!kPsi || ST == PW_PSI, "PW_PSI runs alone");
  constexpr int NN = C::NN, NK = C::NK;
  constexpr int NMAIN = C::SPC * C::N / C::kThreads;         // 8 points per thread; CTA 0 adds the column k = N / 2
  constexpr int NIT = (NMAIN + 1 + NB - 1) / NB;
  const MemberPtrs<C> P(io, member);
  const double dkw = T.kv[1];
  const double dt1 = io.dt1, dt2 = io.dt2, dt3 = io.dt3;
#pragma unroll 1
  for (int b = 0; b < NIT; ++b) {
    int idx[NB], l[NB], k[NB], kk[NB];
    bool ok[NB];
    cplx A[NB], B[NB], q0[NB], q1[NB], dc0[NB], dc1[NB], dp0[NB], dp1[NB], dpp0[NB], dpp1[NB];
    double aT0[NB], aT1[NB], aU0[NB], aU1[NB], fl[NB], kv[NB], lv[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int n = b * NB + u;
      if (n < NMAIN) {
        const int i = threadIdx.x + C::kThreads * n;
        ok[u] = true;
        kk[u] = i % C::SPC;
        l[u] = i / C::SPC;
        k[u] = rank * C::SPC + kk[u];
      } else {                                    // the Nyquist column k = N / 2 (CTA 0, one point for the first N threads)
        ok[u] = n == NMAIN && rank == 0 && (int)threadIdx.x < C::N;
        l[u] = ok[u] ? (int)threadIdx.x : 0;
        k[u] = ok[u] ? C::H : rank * C::SPC + 1;
        kk[u] = ok[u] ? 0 : 1;
      }
      idx[u] = l[u] * NK + k[u];
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
      if (kQh) { kv[u] = dkw * (double)k[u]; lv[u] = dkw * (double)(l[u] < C::H ? l[u] : l[u] - C::N); }
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      if (kRead) spec_read<C>(S, l[u], k[u], kk[u], A[u], B[u]);      // (shared memory: short latency, read where it is used)
      if (ST & PW_STORE_QH) {
        if (ok[u]) { P.qh[idx[u]] = A[u]; P.qh[NN + idx[u]] = B[u]; }
        continue;
      }
      cplx r = cmake(0.0, 0.0);
      if (kTend) {
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
          spec_write<C>(S, l[u], k[u], kk[u], n0, n1);
        }
      }
      if (kUv && ok[u]) {
        const cplx ph = cmake(aU0[u] * q0[u].x + aU1[u] * q1[u].x, aU0[u] * q0[u].y + aU1[u] * q1[u].y);
        spec_write<C>(S, l[u], k[u], kk[u], cmake(lv[u] * ph.y, -lv[u] * ph.x), cmake(-kv[u] * ph.y, kv[u] * ph.x));
      }
      if ((ST & PW_LOAD_QH) && ok[u]) spec_write<C>(S, l[u], k[u], kk[u], q0[u], q1[u]);
      if (kPsi && ok[u]) {                               // ph_z = a[z][0] qh_0 + a[z][1] qh_1            (pyqg _invert)
        const cplx p0 = cmake(aT0[u] * q0[u].x + aT1[u] * q1[u].x, aT0[u] * q0[u].y + aT1[u] * q1[u].y);
        const cplx p1 = cmake(aU0[u] * q0[u].x + aU1[u] * q1[u].x, aU0[u] * q0[u].y + aU1[u] * q1[u].y);
        spec_write<C>(S, l[u], k[u], kk[u], p0, p1);
        if (io.ph_out) { cplx* o = io.ph_out + (long long)member * 2 * NN; o[idx[u]] = p0; o[NN + idx[u]] = p1; }
      }
    }
  }
}

// One round: [inverse 2-D transform of the spectra in S] -> physical stage in registers -> [forward 2-D transform into S]
template <class C, bool BUD = false>      // BUD: the physical stages of the budget program (PH_SCR_*, PH_ANOM*) are compiled in
SCL_INL void round(bool has_inv, int phys, bool has_fwd, const Tables& T, const StepIO& io, int member, cplx* buf, const cplx* tw,
                   uint64_t* hbar, uint32_t& hphase) {
  constexpr int N = C::N, G = C::G, H = C::H, SPC = C::SPC, PS = C::PS, PT = C::PT, RPC = C::RPC, NPIX = C::NPIX;
  const Geo<C> g;
  const MemberPtrs<C> P(io, member);
  cplx v[16];
  const double s = T.inv_M;
#pragma unroll 1
  for (int hp = has_inv ? 0 : 2; hp < (has_fwd ? 4 : 2); ++hp) {
    // ---- stage in ----
    if (hp == 0) {                                     // columns of S, l = G j + t  (packed lanes: C with the c2r convention)
      if (!g.packed) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = buf[(G * j + g.t) * PS + g.colS];
      } else {
        const int c0 = 2 * SPC + (g.isB ? 3 : 2), c1 = 2 * SPC + (g.isB ? 1 : 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int l = G * j + g.t, ln = (N - l) & (N - 1);
          const cplx x0 = buf[l * PS + c0], n0 = buf[ln * PS + c0], x1 = buf[l * PS + c1], n1 = buf[ln * PS + c1];
          v[j] = cmake(0.5 * (x0.x + n0.x - x1.y + n1.y), 0.5 * (x0.y - n0.y + x1.x + n1.x));
        }
      }
      conj16(v);
    } else if (hp == 1) {                              // row of T': W_y(kx) from A_y, B_y; kx = G j + t (j < 8 <=> kx < N / 2)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int kx = G * j + g.t;
        const bool dc = j == 0 && g.t == 0, nyq = j == 8 && g.t == 0;
        const int c = nyq ? 0 : (j < 8 ? kx : N - kx);
        const cplx A = buf[g.line * PT + c], B = buf[g.line * PT + H + c];
        cplx w = j < 8 ? cmake(A.x - B.y, A.y + B.x) : cmake(A.x + B.y, B.x - A.y);
        if (j == 0) w = sel(dc, cmake(A.x, B.x), w);
        if (j == 8) w = sel(nyq, cmake(A.y, B.y), w);
        v[j] = cmake(w.x, -w.y);
      }
    } else if (hp == 3) {                              // columns of T: split W_y into A_y / B_y (packed: C0 / C1), y = G j + t
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int yy = G * j + g.t;
        const cplx Pw = buf[yy * PS + g.sl], Q = buf[yy * PS + SPC + g.sl];     // W(c), W(-c)   (slot 0: W(0), W(N/2))
        const cplx nA = cmake(0.5 * (Pw.x + Q.x), 0.5 * (Pw.y - Q.y)), nB = cmake(0.5 * (Pw.y + Q.y), 0.5 * (Q.x - Pw.x));
        const cplx pA = cmake(Pw.x, Q.x), pB = cmake(Pw.y, Q.y);
        v[j] = g.packed ? (g.isB ? pB : pA) : (g.isB ? nB : nA);
      }
    } else if (!has_inv) {                             // hp == 2 of a forward-only round: load the real pair, x = G j + t
      const double* f0 = phys == PH_LOAD_DQ ? P.dq : P.q;
      if (C::kSplitSync) cluster_arrive_exec();        // (nothing of the buffer is read here: arrive before the global loads)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = g.y * N + G * j + g.t;
        v[j] = cmake(f0[i], f0[NPIX + i]);
        if (phys == PH_LOAD_Q && P.cnn_x) { P.cnn_x[i] = s64::div_by((float)v[j].x, io.x_std[0], io.x_inv[0]); P.cnn_x[NPIX + i] = s64::div_by((float)v[j].y, io.x_std[1], io.x_inv[1]); }
      }
    }
    // ---- the line transform ----
    // write-after-read barriers, split (cluster_arrive_exec): hp 0 arrives once its S columns are in registers and waits before
    // storing into the peers' T'; the x pass arrives once its T' row has been read (hp 1; a forward-only round arrives
    // at the top of hp 2, the pointwise phase before it ended with __syncthreads) and waits in hp 2 before storing into the peers' T
    fftN<C>(v, g.t, tw, hp == 0 || (hp == 1 && has_fwd));
    // ---- stage out ----
    if (hp == 0) {                                     // A_y / B_y at y = out_index -> T' of the CTA that owns row y
      if (C::kSplitSync) cluster_wait_exec();          // every CTA has read its S columns (and finished with its old T')
      else cluster_sync_exec();
      const int col = g.slot + (g.isB ? H : 0);
      if (C::kAsyncHandover && threadIdx.x == 0) handover_expect(hbar, C::kHandoverBytes);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int yy = out_index<C>(m, g.t);
        if (C::kAsyncHandover) st_cluster_async(buf + (yy % RPC) * PT + col, hbar, (uint32_t)(yy / RPC), cmake(v[m].x, -v[m].y));
        else st_cluster(buf + (yy % RPC) * PT + col, (uint32_t)(yy / RPC), cmake(v[m].x, -v[m].y));
      }
      if (C::kAsyncHandover) { handover_wait(hbar, hphase); hphase ^= 1; }
      else cluster_sync();
    } else if (hp == 1) {                              // physical row y, x = out_index (conjugate back, scale)
      conj16(v);
      if (phys >= PH_STORE_UV0) {                      // PROG_INVERT: (u, v) of a layer or (psi_0, psi_1) -> io.*_out
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
          const int i = g.y * N + out_index<C>(m, g.t);
          if (d0) d0[i] = v[m].x * s;
          if (d1) d1[i] = v[m].y * s;
        }
      } else if (phys == PH_EMIT) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int i = g.y * N + out_index<C>(m, g.t);
          const double q0 = v[m].x * s, q1 = v[m].y * s;
          P.q[i] = q0;
          P.q[NPIX + i] = q1;
          if (P.cnn_x) { P.cnn_x[i] = s64::div_by((float)q0, io.x_std[0], io.x_inv[0]); P.cnn_x[NPIX + i] = s64::div_by((float)q1, io.x_std[1], io.x_inv[1]); }
        }
      } else {
        const int z = phys == PH_PRODUCTS1 ? 1 : 0;
        const double* qz = P.q + z * NPIX + g.y * N;
        double U = T.Ubg[z];
        if (BUD) {                                     // budget products use the ANOMALY velocities: u f + i v f
          U = 0.0;
          if (phys == PH_ANOM0 || phys == PH_ANOM1) qz = P.q + (phys - PH_ANOM0) * NPIX + g.y * N;
          else qz = io.bud_scr + ((long long)member * 3 + (phys - PH_SCR_PROD0)) * NPIX + g.y * N;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double qq = qz[out_index<C>(m, g.t)];
          v[m] = cmake((v[m].x * s + U) * qq, (v[m].y * s) * qq);
        }
        out_to_in_naming<C>(v);
      }
    } else if (hp == 2) {                              // W_y(kx), kx = out_index -> T of the CTA that owns column slot(kx)
      if (C::kSplitSync) cluster_wait_exec();          // every CTA has read its T' rows
      else cluster_sync_exec();
      if (C::kAsyncHandover && threadIdx.x == 0) handover_expect(hbar, C::kHandoverBytes);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int kx = out_index<C>(m, g.t);
        const bool minus = kx >= H;                    // kx = N / 2 is the "minus" partner of slot 0 (packed columns)
        const int sg = kx < H ? kx : (N - kx) & (N - 1) & (H - 1);
        if (C::kAsyncHandover) st_cluster_async(buf + g.y * PS + (sg % SPC) + (minus ? SPC : 0), hbar, (uint32_t)(sg / SPC), v[m]);
        else st_cluster(buf + g.y * PS + (sg % SPC) + (minus ? SPC : 0), (uint32_t)(sg / SPC), v[m]);
      }
      if (C::kAsyncHandover) { handover_wait(hbar, hphase); hphase ^= 1; }
      else cluster_sync();
    } else {                                           // hp == 3: Ahat / Bhat / packed C at l = out_index -> S (local)
      __syncthreads();                                 // the CTA's T columns have been read
#pragma unroll
      for (int m = 0; m < 16; ++m) buf[out_index<C>(m, g.t) * PS + g.colS] = v[m];
      __syncthreads();
    }
  }
}

// prog: PROG_STEP, PROG_STEP_DQ, PROG_STEP_DQ_RAW, PROG_SET_Q, PROG_C2R, PROG_ADVECT, PROG_INVERT (spectral64.cuh).  One cluster of CL CTAs per member.
template <int N_, int G_, int CL_>
__global__ void __launch_bounds__((Cfg<N_, G_, CL_>::kThreads), (Cfg<N_, G_, CL_>::kCtasPerSm)) qg_step_cl_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io,
                                                            int prog, int members) {
  using C = Cfg<N_, G_, CL_>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + C::kBuf;
  uint64_t* hbar = reinterpret_cast<uint64_t*>(tw + C::N);
  uint32_t hphase = 0;
  for (int i = threadIdx.x; i < C::N; i += C::kThreads) tw[i] = T.tw[i];
  if (C::kAsyncHandover && threadIdx.x == 0) handover_init(hbar);
  const int rank = (int)cluster_rank();
  const int ncl = gridDim.x / C::CL, cl = blockIdx.x / C::CL;
  const bool with_dq = prog == PROG_STEP_DQ || prog == PROG_STEP_DQ_RAW;
  const bool demean = prog == PROG_STEP_DQ;
  const int nrounds = prog == PROG_C2R ? 1 : prog == PROG_SET_Q ? 2 : (with_dq ? 4 : 3);
  cluster_sync();                                      // all CTAs of the cluster are resident before any remote store
  for (int m = cl; m < members; m += ncl) {
#if SCL_PREFETCH
    {   // pull this member's arrays into L2 while the first transform runs: the pointwise phases then see L2 latency, not DRAM's
      const MemberPtrs<C> P(io, m);
      const int nspec = 2 * C::NN * (int)sizeof(cplx) / 128, nphys = 2 * C::NPIX * (int)sizeof(double) / 128;
      for (int i = rank * C::kThreads + threadIdx.x; i < nspec; i += C::CL * C::kThreads) {
        s64::prefetch_l2(reinterpret_cast<const char*>(P.qh) + 128 * i);
        if (prog != PROG_SET_Q && prog != PROG_C2R) {
          s64::prefetch_l2(reinterpret_cast<const char*>(P.d_p) + 128 * i);
          s64::prefetch_l2(reinterpret_cast<const char*>(P.d_pp) + 128 * i);
        }
      }
      if (prog != PROG_C2R)
        for (int i = rank * C::kThreads + threadIdx.x; i < nphys; i += C::CL * C::kThreads) {
          s64::prefetch_l2(reinterpret_cast<const char*>(P.q) + 128 * i);
          if (P.dq) s64::prefetch_l2(reinterpret_cast<const char*>(P.dq) + 128 * i);
        }
    }
#endif
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < nrounds; ++r) {
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
          case PW_UV0: pointwise_phase<C, PW_UV0, SCL_NB_UV>(T, io, m, rank, buf, demean); break;
          case PW_UV1: pointwise_phase<C, PW_UV1, SCL_NB_UV>(T, io, m, rank, buf, demean); break;
          case PW_PSI: pointwise_phase<C, PW_PSI, SCL_NB_UV>(T, io, m, rank, buf, demean); break;
          case PW_TEND0 | PW_UV1: pointwise_phase<C, PW_TEND0 | PW_UV1, SCL_NB_TEND>(T, io, m, rank, buf, demean); break;
          case PW_TEND1: pointwise_phase<C, PW_TEND1, SCL_NB_TEND>(T, io, m, rank, buf, demean); break;
          case PW_TEND1 | PW_UPDATE: pointwise_phase<C, PW_TEND1 | PW_UPDATE, SCL_NB_UPD>(T, io, m, rank, buf, demean); break;
          case PW_FORCING | PW_UPDATE: pointwise_phase<C, PW_FORCING | PW_UPDATE, SCL_NB_UPD>(T, io, m, rank, buf, demean); break;
          case PW_STORE_QH: pointwise_phase<C, PW_STORE_QH, SCL_NB_UV>(T, io, m, rank, buf, demean); break;
          default: pointwise_phase<C, PW_LOAD_QH, SCL_NB_UV>(T, io, m, rank, buf, demean); break;
        }
        __syncthreads();
      }
      if (rnd) round<C>(inv, phys, fwd, T, io, m, buf, tw, hbar, hphase);
      if (prog == PROG_INVERT) __syncthreads();        // an inverse-only round ends with reads of T': the next phase writes S over it
    }
  }
  cluster_sync();                                      // no CTA exits while a peer may still store into its shared memory
}

// ---- PROG_BUDGET on the cluster (spectral64.cuh budget_phase / qg_budget64_kernel; the per-point arithmetic is shared) -------
template <class C, int RD, int WR, int NB>
SCL_INL void budget_phase(const Tables& T, const StepIO& io, int member, int rank, cplx* S) {
  constexpr int NN = C::NN, NK = C::NK;
  constexpr int NMAIN = C::SPC * C::N / C::kThreads;
  constexpr int NIT = (NMAIN + 1 + NB - 1) / NB;
  constexpr bool kRead = RD != BR_NONE && RD != BR_DISS;
  const long long mo = (long long)member * 2 * NN;
  const cplx* qh = io.qh + mo;
  double* out = io.bud_out + (long long)member * kBudgetTerms * NN;
  cplx* tend = io.bud_tend + mo;
  const cplx *dp = io.d_p + mo, *dpp = io.d_pp + mo;
  const double dkw = T.kv[1], d1 = io.Hi_over_H[0], d2 = io.Hi_over_H[1];
#pragma unroll 1
  for (int b = 0; b < NIT; ++b) {
    int idx[NB], l[NB], k[NB], kk[NB];
    bool ok[NB];
    cplx q0[NB], q1[NB];
    double a00[NB], a01[NB], a10[NB], a11[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int n = b * NB + u;
      if (n < NMAIN) {
        const int i = threadIdx.x + C::kThreads * n;
        ok[u] = true;
        kk[u] = i % C::SPC;
        l[u] = i / C::SPC;
        k[u] = rank * C::SPC + kk[u];
      } else {                                    // the Nyquist column k = N / 2 (CTA 0, one point for the first N threads)
        ok[u] = n == NMAIN && rank == 0 && (int)threadIdx.x < C::N;
        l[u] = ok[u] ? (int)threadIdx.x : 0;
        k[u] = ok[u] ? C::H : rank * C::SPC + 1;
        kk[u] = ok[u] ? 0 : 1;
      }
      idx[u] = l[u] * NK + k[u];
      q0[u] = qh[idx[u]]; q1[u] = qh[NN + idx[u]];
      a00[u] = T.a[idx[u]]; a01[u] = T.a[NN + idx[u]]; a10[u] = T.a[2 * NN + idx[u]]; a11[u] = T.a[3 * NN + idx[u]];
    }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const double kv = dkw * (double)k[u], lv = dkw * (double)(l[u] < C::H ? l[u] : l[u] - C::N);
      const cplx p0 = cmake(a00[u] * q0[u].x + a01[u] * q1[u].x, a00[u] * q0[u].y + a01[u] * q1[u].y);
      const cplx p1 = cmake(a10[u] * q0[u].x + a11[u] * q1[u].x, a10[u] * q0[u].y + a11[u] * q1[u].y);
      cplx A = cmake(0.0, 0.0), B = cmake(0.0, 0.0);
      if (kRead) spec_read<C>(S, l[u], k[u], kk[u], A, B);
      if (RD != BR_NONE && ok[u])
        bud_read<RD>(T, io, out, tend, dp, dpp, NN, idx[u], kv, lv, q0[u], q1[u], p0, p1, a00[u], a01[u], a10[u], a11[u], A, B);
      if (WR != BWR_NONE && ok[u]) {
        cplx oA, oB;
        bud_build<WR>(p0, p1, kv, lv, d1, d2, oA, oB);
        spec_write<C>(S, l[u], k[u], kk[u], oA, oB);
      }
    }
  }
}

template <int N_, int G_, int CL_>
__global__ void __launch_bounds__((Cfg<N_, G_, CL_>::kThreads), (Cfg<N_, G_, CL_>::kCtasPerSm)) qg_budget_cl_kernel(const __grid_constant__ Tables T, const __grid_constant__ StepIO io,
                                                              int members) {
  using C = Cfg<N_, G_, CL_>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* buf = reinterpret_cast<cplx*>(smem_raw);
  cplx* tw = buf + C::kBuf;
  uint64_t* hbar = reinterpret_cast<uint64_t*>(tw + C::N);
  uint32_t hphase = 0;
  for (int i = threadIdx.x; i < C::N; i += C::kThreads) tw[i] = T.tw[i];
  if (C::kAsyncHandover && threadIdx.x == 0) handover_init(hbar);
  const int rank = (int)cluster_rank();
  const int ncl = gridDim.x / C::CL, cl = blockIdx.x / C::CL;
  const bool has_dq = io.dq != nullptr;
  cluster_sync();
  for (int m = cl; m < members; m += ncl) {
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
      int phys = PH_EMIT;
      bool inv = true, fwd = true, rnd = true;
      switch (r) {
        case 0: budget_phase<C, BR_NONE, BWR_XI, 2>(T, io, m, rank, buf); phys = PH_SCR_STORE01; fwd = false; break;
        case 1: budget_phase<C, BR_NONE, BWR_UV0, 2>(T, io, m, rank, buf); phys = PH_SCR_PROD0; break;
        case 2: budget_phase<C, BR_KEFLUX0, BWR_UV1, 2>(T, io, m, rank, buf); phys = PH_SCR_PROD1; break;
        case 3: budget_phase<C, BR_KEFLUX1, BWR_TAU, 2>(T, io, m, rank, buf); phys = PH_SCR_STORE2; fwd = false; break;
        case 4: budget_phase<C, BR_NONE, BWR_UVBT, 2>(T, io, m, rank, buf); phys = PH_SCR_PROD2; break;
        case 5: budget_phase<C, BR_APEFLUX, BWR_UV0, 2>(T, io, m, rank, buf); phys = PH_ANOM0; break;
        case 6: budget_phase<C, BR_ENS0, BWR_UV1, 2>(T, io, m, rank, buf); phys = PH_ANOM1; break;
        case 7: budget_phase<C, BR_ENS1, BWR_NONE, 2>(T, io, m, rank, buf); inv = false; phys = PH_LOAD_DQ; rnd = has_dq; break;
        default:
          if (has_dq) budget_phase<C, BR_PARAM_DISS, BWR_NONE, 1>(T, io, m, rank, buf);
          else budget_phase<C, BR_DISS, BWR_NONE, 1>(T, io, m, rank, buf);
          rnd = false;
          break;
      }
      __syncthreads();
      if (rnd) round<C, true>(inv, phys, fwd, T, io, m, buf, tw, hbar, hphase);
      if (!fwd) __syncthreads();                       // an inverse-only round ends with reads of T': the next phase writes S over it
    }
  }
  cluster_sync();
}

}  // namespace scl
}  // namespace qgb
