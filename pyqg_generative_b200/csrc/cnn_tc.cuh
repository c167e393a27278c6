// cnn_tc.cuh -- AndrewCNN forward on the 5th-generation tensor cores (sm_100a): implicit-GEMM circular convolutions
// issued as tcgen05.mma (cta_group::1, kind::f16 and kind::f8f6f4, M=128) with fp32 accumulators in TMEM, operands staged
// in shared memory by the TMA engine (cp.async.bulk[.tensor] + mbarrier complete_tx), warp-specialised persistent CTAs
// (one per SM).
//
// Reference semantics: pyqg_generative/tools/cnn_tools.py:79-98 (make_block: Conv2d circular 'same' -> ReLU ->
// BatchNorm2d), :125-176 (AndrewCNN), models/mean_var_model.py:14-17 (softplus head).
//
// Precision plan ("plain TF32 fails the 1e-3 bound"): operands are fp16 (11-bit significand, same as
// TF32, at twice the tensor rate).  Every layer but the second runs the 3-pass split  a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
// (a = a_hi + a_lo; error ~2^-22), fp32 accumulation throughout.  Layer 2 (128->64, 5x5, 75 % of the FLOPs) runs TWO passes,
// (a_hi + a_lo)*w_hi: with the shipped networks the rounding of its ACTIVATIONS to 11 bits alone costs up to 1.8e-3 relative
// (VAE decoder, GZ mean net; measured, profiles/r1_tc_precision.md) while rounding its weights costs 2-5e-4.  A single-pass
// variant (QGB_PREC_TC_FAST) is kept for networks where that is acceptable.  In the TMA-fed layers the low half a_lo travels
// as e4m3 (x 2^11) and multiplies an e4m3 copy of the weights (x 2^-11) in one kind::f8f6f4 MMA per 32 channels (see TcCfg).
// Weights are pre-scaled per layer by a power of two so they sit in the fp16 normal range; the epilogue undoes the scale.
//
// Data layout.  Activations live in HBM as  [image][C/32][ny+2p][nx+2p][32]  (fp16 hi plane, e4m3 lo plane): the innermost
// 64 B (32 B) are 32 consecutive channels of one pixel, the circular halo (p = padding of the CONSUMING layer) is
// materialised by the producing epilogue, so every tap of the consumer is a plain in-bounds box.  A CTA tile is
// 16*NH rows x 8T columns (T M-tiles of 128 accumulator rows, NH stacked output rows per accumulator row, see TcCfg).  The
// tile + halo of one 32-channel chunk is loaded ONCE by a 4-D TMA box with the 64-byte swizzle and serves all taps: a tap
// (vertical shift, column), K-step and M-tile are just a different start address of the SAME K-major SWIZZLE_64B canonical
// UMMA layout (8 pixel rows of 64 B per atom, the next row group at the stride byte offset).  Because a pixel row is 64 B,
// every tap shift keeps the 8x16 B core matrices bank-conflict free (the first version used 16 B pixel rows: shifted core
// matrices straddled two 128 B lines and the tensor pipe sat at 41 %, profiles/r1_history.md).
//
// Layer 1 (Cin = 4 or 2) is too thin for a K=16 MMA: it runs as a 1x1 convolution over its 5x5 im2col (K = 100 -> 128),
// which eight extra "builder" warps write straight into the shared-memory operand stages from the raw fp32 input
// (template parameter FUSE), so the im2col tensor never exists in HBM.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qgb200.h"
#include "cnn_tc_host.hpp"
#include "philox.cuh"

#ifndef QGB_TC_NW_OVERRIDE
#define QGB_TC_NW_OVERRIDE 0
#endif

namespace qgb {

// ------------------------------------------------------------------------------------------ PTX wrappers ----
namespace ptx {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA engine, non-tensor bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA tensor-tile load (3-D box) global -> shared (SASS: UTMALDG); ``tmap`` points at a __grid_constant__ CUtensorMap
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 inputs, fp32 accumulate; issued by ONE thread for the CTA (SASS: UTCHMMA)
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with e4m3 inputs (K = 32 per instruction): the low halves of the split-precision activations only need ~4 bits
__device__ __forceinline__ void mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
// one elected lane of a fully converged warp (same lane every time for the same mask)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred;
}
// 256-bit global store (SASS: STG.E.ENL2.256): one whole 32-byte sector per lane
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
// programmatic dependent launch: let the next kernel of the stream be scheduled early / wait for the previous one to finish
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
}  // namespace ptx

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) leading-dim byte offset>>4 (between the two 8-element K chunks), [32,46) stride-dim byte
// offset>>4 (between 8-row groups), [46,48) version = 1, [61,64) layout type = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format F16 = 0, K-major A and B,
// n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

enum { TC_OUT_HI = 0, TC_OUT_HILO = 1, TC_OUT_FINAL = 2 };

struct TcConvParams {
  const __half* in_hi; const __half* in_lo;  // [img][nch_in][HP][WP][32]
  int nch_in, HP, WP;
  const __half* w;                           // [chunk][tap column] weight stages, see tc_pack_layer
  float inv_wscale; int relu;
  __half* out_hi; __half* out_lo; int out_pad, out_nch;   // [img][out_nch][ny+2*out_pad][nx+2*out_pad][32]
  float* out_f32; long long out_bs; int out_c, softplus, accumulate;
  const float* x_f32; long long x_bs;          // fused layer-1 variant: raw network input (B, cin0, ny, nx) fp32
  int ny, nx, tiles_y, tiles_x, num_tiles;
  // direct layer 1 with the latent noise generated in the kernel (channels 2, 3 of a 4-channel input): Philox key material
  int noise; int noise_member0; unsigned long long noise_seed; const uint32_t* noise_draw;
  // last layer with the closure epilogue fused: dq = float64(y * y_std) * weight (models/cgan_regression.py:157-162) straight to the
  // forcing array (B, 2, ny, nx) instead of the fp32 network output
  double* out_dq; float dq_ys0, dq_ys1; double dq_weight;
};

// LO8: the low halves of the activations travel as e4m3 (value * 2^11, 1 byte per element, SWIZZLE_32B rows) and their pass
// a_lo x w_hi runs as ONE kind::f8f6f4 MMA per 32 channels against an e4m3 copy of the weights (w * 2^-11): the product only
// needs ~4 significant bits (it is 2^-11 of the result), costs half the shared-memory reads and half the HBM bytes.
template <int CIN, int COUT, int KS, int PASSES, int T, bool LO8 = false, int NH = 1>
struct TcCfg {
  static constexpr int NCHUNK = CIN / 32;
  static constexpr int TAPS = KS * KS;
  // PASSES 1: a_hi w_hi;  2: (a_hi + a_lo) w_hi;  3: a_hi w_hi + a_lo w_hi + a_hi w_lo
  static constexpr int PLANES = PASSES >= 2 ? 2 : 1;      // activation planes (hi [, lo])
  static constexpr int WPLANES = PASSES == 3 ? 2 : 1;     // fp16 weight planes
  // ROW STACKING (NH > 1).  With both operands in shared memory an M=128 MMA costs max(N/2, (128+N)/4) clocks: the 4 KB
  // A operand is re-read by every instruction, so N = 64 runs at 2/3 of the tensor rate and only N >= 128 reaches it
  // (profiles/r1_mma_rate_microbench.txt).  COUT is 64 or less, so N is widened with OUTPUT ROWS instead: an M-tile takes
  // every NH-th image row (the descriptor's stride-byte-offset is NH image rows), accumulator column block h holds output
  // row g*NH + h, and the MMA for vertical window shift s multiplies the shared A window by the weights of ALL taps
  // ky = s - h that are valid, concatenated along N (they are adjacent in shared memory because a weight stage stores a
  // tap COLUMN in descending ky).  KS + NH - 1 shifts replace KS*NH row taps; for the 5x5 layer with NH = 2 that is
  // 352 instead of 480 clocks per (tap column, K step) and the tensor pipe is busy 91 % of the time instead of 67 %.
  static constexpr int HY = 16 * NH + KS - 1;
  static constexpr int HX = 8 * T + KS - 1;
  static constexpr int NSHIFT = KS + NH - 1;
  static constexpr int A_BYTES = HY * HX * 64;                        // hi plane of one 32-channel chunk
  static constexpr int A_PLANE = (A_BYTES + 1023) / 1024 * 1024;      // swizzle atoms need aligned plane bases
  static constexpr int A_LO_BYTES = PLANES == 2 ? (LO8 ? HY * HX * 32 : A_BYTES) : 0;
  static constexpr int A_STAGE = A_PLANE + (A_LO_BYTES + 1023) / 1024 * 1024;
  static constexpr int A_TX = A_BYTES + A_LO_BYTES;
  // 3-pass layers with a narrow N concatenate [w_hi | w_lo] along N:  D[:, :COUT] += a_hi w_hi + a_lo w_hi,
  // D[:, COUT:] += a_hi w_lo  (two MMAs instead of three -> fewer shared-memory reads of the A operand)
  static constexpr bool NCAT = (PASSES == 3) && (COUT <= 32);
  static constexpr int NF = NCAT ? 2 * COUT : COUT;            // accumulator columns (= B rows) per stacked row
  static constexpr int WROWS = NCAT ? NF : WPLANES * COUT;     // fp16 B rows per (K chunk, tap)
  static constexpr int B_LBO = KS * WROWS;                     // 16-byte units between the 8-element K chunks
  static constexpr int B8_LBO = KS * NF;
  static constexpr int W16 = 4 * KS * WROWS * 16;              // fp16 part of one (chunk, tap column) weight stage
  static constexpr int W_STAGE = W16 + (LO8 ? 2 * KS * NF * 16 : 0);   // + e4m3 part [2][KS][NF][16 B]
  static constexpr int NW = QGB_TC_NW_OVERRIDE ? QGB_TC_NW_OVERRIDE : ((KS == 5) ? (PASSES == 2 ? 2 : 3) : 4);
  // thin layers: every weight stage of the layer fits the ring -> loaded once per CTA and kept (re-fetching them per tile
  // doubled the L2 -> SM traffic of the 32 -> 32 layers: 453 MB of weights against 481 MB of activations per launch)
  static constexpr bool W_RESIDENT = NCHUNK * KS <= NW;
  static constexpr int DCOLS = NH * NF;                        // TMEM columns per M-tile
  static constexpr int NACC = (4 * T * DCOLS <= 512) ? 4 : 2;    // accumulator ring: MMA of tile i+NACC waits for epilogue i
  static constexpr int NCOLS_USED = NACC * T * DCOLS;
  static constexpr int NCOLS = NCOLS_USED <= 32 ? 32 : NCOLS_USED <= 64 ? 64 : NCOLS_USED <= 128 ? 128 : NCOLS_USED <= 256 ? 256 : 512;
  // activation ring: as deep as shared memory allows (2..4 stages).  Thin layers (one 32-channel chunk per tile) spend
  // less time on a stage than a TMA round trip to HBM takes, so they need more than one load in flight.
  static constexpr int SMEM_FIXED = NW * W_STAGE + 1536 + 256 + 1024 + 6400;
  static constexpr int NA_FIT = (232448 - SMEM_FIXED) / A_STAGE;
  static constexpr int NA = NA_FIT >= 4 ? 4 : (NA_FIT >= 3 ? 3 : 2);
  static constexpr int SMEM = NA * A_STAGE + NW * W_STAGE + 1536 + 256 + 1024;   // + 6400 for the FUSE window
  static_assert(NCOLS_USED <= 512, "accumulators exceed TMEM");
  static_assert(CIN % 32 == 0 && COUT % 16 == 0, "bad channel counts");
  static_assert(NH == 1 || NH <= KS, "the first MMA of a tile must cover every stacked row");
  static_assert(NCAT || WPLANES == 1 || KS == 1, "separate w_lo pass is only laid out for 1x1 (layer 1)");
  static_assert(NH * NF <= 256, "MMA N exceeds 256");
};

__device__ __forceinline__ float softplus_f(float v) { return v > 20.f ? v : log1pf(expf(v)); }

// Activation tensor maps: the buffer [img*nch][HP][WP][32 x fp16] is a 4-D tensor (32, WP, HP, img*nch); one box
// (32, HX, HY, 1) with CU_TENSOR_MAP_SWIZZLE_64B = the tile + halo of one 32-channel chunk lands in shared memory as
// [HY][HX][64 B] with 16-byte chunks XOR-swizzled by address bits 7-8 -- the K-major SWIZZLE_64B UMMA layout.
struct TcMaps {
  CUtensorMap hi, lo;
};

// FUSE = 0: activations arrive by TMA.  FUSE = cin0 (4 or 2): layer 1 -- four extra warps BUILD the 5x5 im2col operand
// (K = 25*cin0 padded to CIN, fp16 hi/lo planes, SWIZZLE_64B layout) straight into the shared-memory stages from the raw
// fp32 network input, so the im2col tensor never exists in HBM.
template <int CIN, int COUT, int KS, int PASSES, int T, int OUTMODE, int FUSE = 0, int NH = 1>
__global__ void __launch_bounds__(FUSE ? 640 : 384, 1) conv_tc_kernel(const __grid_constant__ TcConvParams P,
                                                                      const __grid_constant__ TcMaps M,
                                                                      const __grid_constant__ TcEpi E) {
  static_assert(FUSE == 0 || (KS == 1 && T == 2 && PASSES == 3 && NH == 1), "fused im2col is the layer-1 configuration");
  constexpr bool LO8 = !FUSE && PASSES >= 2;
  using C = TcCfg<CIN, COUT, KS, PASSES, T, LO8, NH>;
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char* smem = tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sW = smem + C::NA * C::A_STAGE;
  float* sEpi = reinterpret_cast<float*>(sW + C::NW * C::W_STAGE);          // bias | scale | shift (only read when COUT > 32)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sEpi) + 1536);
  uint64_t* a_full = bars;            // [NA]
  uint64_t* a_empty = bars + C::NA;   // [NA]
  uint64_t* w_full = bars + 2 * C::NA;        // [NW]
  uint64_t* w_empty = bars + 2 * C::NA + C::NW;
  uint64_t* acc_full = bars + 2 * C::NA + 2 * C::NW;   // [NACC]
  uint64_t* acc_empty = acc_full + C::NACC;    // [NACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + C::NACC);
  float* s_x = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 256);   // FUSE: [cin0][20][20] input window

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool BREG = COUT <= 32;             // epilogue constants in registers
  if (!BREG) for (int i = threadIdx.x; i < COUT; i += blockDim.x) { sEpi[i] = E.b[i]; sEpi[COUT + i] = E.s[i]; sEpi[2 * COUT + i] = E.t[i]; }
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::NA; ++i) { ptx::mbar_init(&a_full[i], FUSE ? 256 : 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::NW; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < C::NACC; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 256); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, C::NCOLS);
  if (!FUSE && warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&M.hi);
    if (C::PLANES == 2) ptx::prefetch_tmap(&M.lo);
  }
  // Programmatic dependent launch: the layers of one forward pass are launched back to back with the stream-serialization
  // attribute; this CTA may have been scheduled while the previous layer's last CTAs are still running (its own set-up
  // above -- barriers, TMEM allocation, descriptor prefetch -- overlaps their tail) and must not touch activations before
  // the previous grid has completed and flushed.
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = P.tiles_y * P.tiles_x;

  if (FUSE && warp >= 12) {
    // ===================== A builders (layer 1): im2col of the raw input, written swizzled into the A stages =========
    // 8 warps, one pixel of the 16x16 tile per thread.  The input window is staged as [row][col][4] floats so one tap of a
    // pixel is ONE 16-byte shared-memory load; fp32 -> fp16 hi/lo uses the packed half2 conversions.
    constexpr int F = FUSE ? FUSE : 1;
    const int bt = threadIdx.x - 384;                    // 0..255
    const int py = bt >> 4, px = bt & 15;
    float4* s_x4 = reinterpret_cast<float4*>(s_x);
    uint32_t ia = 0;
    // window entries this thread stages: i = bt and bt + 256 (400 entries of [row][col] x 4 channels); the NEXT tile's
    // entries are fetched into registers before the current tile is built, hiding the global-memory latency
    auto fetch = [&](int tile, int i) -> float4 {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tile < P.num_tiles && i < 400) {
        const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
        const int y0 = (r / P.tiles_x) * 16, x0 = (r % P.tiles_x) * 16;
        const int rr = i / 20, col = i % 20;
        int sy = y0 + rr - 2, sx = x0 + col - 2;
        sy = sy < 0 ? sy + P.ny : (sy >= P.ny ? sy - P.ny : sy);
        sx = sx < 0 ? sx + P.nx : (sx >= P.nx ? sx - P.nx : sx);
        const float* src = P.x_f32 + (long long)img * P.x_bs + (long long)sy * P.nx + sx;
        const long long cs = (long long)P.ny * P.nx;
        v.x = src[0];
        v.y = src[cs];
        if (F > 2) { v.z = src[2 * cs]; v.w = src[3 * cs]; }
      }
      return v;
    };
    float4 w0 = fetch(blockIdx.x, bt), w1 = fetch(blockIdx.x, bt + 256);
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      ptx::named_bar_sync(1, 256);                       // the previous tile's window is no longer being read
      s_x4[bt] = w0;
      if (bt + 256 < 400) s_x4[bt + 256] = w1;
      ptx::named_bar_sync(1, 256);
      w0 = fetch(tile + gridDim.x, bt);
      w1 = fetch(tile + gridDim.x, bt + 256);
#pragma unroll
      for (int c = 0; c < C::NCHUNK; ++c, ++ia) {
        const uint32_t s = ia % C::NA, par = (ia / C::NA) & 1;
        ptx::mbar_wait(&a_empty[s], par ^ 1);
        unsigned char* stage = sA + s * C::A_STAGE;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // 8 K values of this 16-byte piece: K index kk = tap*F + channel
          float f[8];
          if (F == 4) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int tap = (c * 32 + j * 8) / 4 + h;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (tap < 25) v = s_x4[(py + tap / 5) * 20 + px + tap % 5];
              f[4 * h] = v.x; f[4 * h + 1] = v.y; f[4 * h + 2] = v.z; f[4 * h + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int tap = (c * 32 + j * 8) / 2 + h;
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (tap < 25) v = s_x4[(py + tap / 5) * 20 + px + tap % 5];
              f[2 * h] = v.x; f[2 * h + 1] = v.y;
            }
          }
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            const __half2 h2 = __floats2half2_rn(f[2 * e2], f[2 * e2 + 1]);
            const float2 back = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(f[2 * e2] - back.x, f[2 * e2 + 1] - back.y);
            hi[e2] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[e2] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          unsigned char* dst = stage + bt * 64 + ((j ^ ((bt >> 1) & 3)) << 4);   // 64-byte swizzle: chunk ^= address bits 7-8
          *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(dst + C::A_PLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        ptx::fence_proxy_async_smem();                   // generic-proxy stores -> visible to the tensor core (async proxy)
        ptx::mbar_arrive(&a_full[s]);
      }
    }
  } else if (!FUSE && warp == 0) {
    // ===================== A producer: tile + halo of one 32-channel chunk per stage (TMA bulk copies) ============
    uint32_t ia = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
      const int y0 = (r / P.tiles_x) * 16 * NH, x0 = (r % P.tiles_x) * 8 * T;
      for (int c = 0; c < C::NCHUNK; ++c, ++ia) {
        const uint32_t s = ia % C::NA, par = (ia / C::NA) & 1;
        ptx::mbar_wait(&a_empty[s], par ^ 1);
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(&a_full[s], C::A_TX);
          ptx::tma_load_4d(sA + s * C::A_STAGE, &M.hi, 0, x0, y0, img * P.nch_in + c, &a_full[s]);
          if (C::PLANES == 2) ptx::tma_load_4d(sA + s * C::A_STAGE + C::A_PLANE, &M.lo, 0, x0, y0, img * P.nch_in + c, &a_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ===================== W producer: one (chunk, tap column) weight slab per stage ================================
    if (lane == 0 && C::W_RESIDENT) {
      for (int cr = 0; cr < C::NCHUNK * KS; ++cr) {
        ptx::mbar_arrive_expect_tx(&w_full[cr], C::W_STAGE);
        ptx::bulk_g2s(sW + cr * C::W_STAGE, reinterpret_cast<const unsigned char*>(P.w) + (size_t)cr * C::W_STAGE, C::W_STAGE, &w_full[cr]);
      }
    } else if (lane == 0) {
      uint32_t iw = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        for (int cr = 0; cr < C::NCHUNK * KS; ++cr, ++iw) {
          const uint32_t s = iw % C::NW, par = (iw / C::NW) & 1;
          ptx::mbar_wait(&w_empty[s], par ^ 1);
          ptx::mbar_arrive_expect_tx(&w_full[s], C::W_STAGE);
          ptx::bulk_g2s(sW + s * C::W_STAGE, reinterpret_cast<const unsigned char*>(P.w) + (size_t)cr * C::W_STAGE,
                        C::W_STAGE, &w_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp stays converged, one elected lane drives the tensor core ==========
    // A: K-major SWIZZLE_64B (layout type 4): pixel rows of 64 B, 8-row groups (= NH image rows further) at the SBO
    // B: K-major no swizzle: weight stage [j][ky descending][WROWS][8]: 8 rows x 16 B per core matrix, next j at LBO
    constexpr uint32_t A_SBO = NH * C::HX * 4;                                  // in 16-byte units
    constexpr uint32_t B_LBO = C::B_LBO, B_SBO = 8;
    constexpr uint32_t a_hi32 = A_SBO | (1u << 14) | (4u << 29), b_hi32 = B_SBO | (1u << 14);   // upper descriptor words
    constexpr uint32_t a8_hi32 = (uint32_t)(NH * C::HX * 2) | (1u << 14) | (6u << 29);           // e4m3 plane: SWIZZLE_32B
    const uint32_t sA_u = ptx::smem_u32(sA) >> 4, sW_u = ptx::smem_u32(sW) >> 4;
    uint32_t ia = 0, iw = 0, it = 0;
    // The elected lane runs the WHOLE persistent loop (barrier waits included): no per-row elect / reconvergence, so the
    // tensor-pipe queue does not drain between tap columns.
    // Every mbarrier wait costs the issuing thread ~85 clk even when the barrier has completed, and the tensor-pipe queue
    // holds only a few hundred clocks of work: a per-tile timeline of the thin layers (2.6 k clk per tile) showed a ~650 clk
    // bubble between the last MMA of one tile and the first of the next.  So the waits for the NEXT tile's accumulator
    // and first activation stage are taken while the last tap column of the current tile is still queued, and resident
    // weight stages are only waited for on the first tile.
    bool ahead = false;     // acc_empty / a_full of the coming tile were already waited for
    if (ptx::elect_one_sync()) {
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it % C::NACC;
      if (!ahead) {
        ptx::mbar_wait(&acc_empty[as], ((it / C::NACC) & 1) ^ 1);
        ptx::tc_fence_after();
      }
      const uint32_t dbase = tmem_base + as * (T * C::DCOLS);
      for (int c = 0; c < C::NCHUNK; ++c, ++ia) {
        const uint32_t sa = ia % C::NA;
        if (!(ahead && c == 0)) {
          ptx::mbar_wait(&a_full[sa], (ia / C::NA) & 1);
          ptx::tc_fence_after();
        }
        if (c == 0) ahead = false;
#pragma unroll
        for (int dx = 0; dx < KS; ++dx, ++iw) {
          const uint32_t sw = C::W_RESIDENT ? (uint32_t)(c * KS + dx) : iw % C::NW;
          if (!C::W_RESIDENT || it == 0) {
            ptx::mbar_wait(&w_full[sw], C::W_RESIDENT ? 0u : (iw / C::NW) & 1);
            ptx::tc_fence_after();
          }
          // The 64 B swizzle is a pure function of the shared-memory ADDRESS bits (verified on B200: a non-zero 'matrix
          // base offset' gives wrong results), so shifted tap windows need no descriptor fix-up.
          const uint32_t a_col = sA_u + sa * (C::A_STAGE >> 4) + dx * 4;
          const uint32_t a_col8 = sA_u + sa * (C::A_STAGE >> 4) + (C::A_PLANE >> 4) + dx * 2;   // e4m3 lo plane
          const uint32_t b_col = sW_u + sw * (C::W_STAGE >> 4);
          const uint32_t b_col8 = b_col + (C::W16 >> 4);
#pragma unroll
          for (int si = 0; si < C::NSHIFT; ++si) {
            // vertical window shift s = ky + h.  The first shift issued covers every stacked row (it initialises the
            // accumulators), the others follow in ascending order.
            constexpr int S0 = KS - 1;
            const int s = si == 0 ? S0 : (si <= S0 ? si - 1 : si);
            const int h_min = s > KS - 1 ? s - (KS - 1) : 0, h_max = s < NH - 1 ? s : NH - 1, nv = h_max - h_min + 1;
            const int pos = (KS - 1) - (s - h_min);                 // weight block of ky = s - h_min (descending ky order)
            const uint32_t a0 = a_col + s * C::HX * 4;
            const uint32_t dcol = dbase + h_min * C::NF;
            const uint32_t idesc_n = make_idesc_f16(128, C::NF * nv);
            const uint32_t acc0 = (c | dx | si) ? 1u : 0u;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t b0 = b_col + 2 * ks * B_LBO + pos * C::WROWS;
              const uint64_t bdesc = ((uint64_t)b_hi32 << 32) | b0 | (B_LBO << 16);
              const uint32_t first = ks ? 1u : acc0;
#pragma unroll
              for (int t = 0; t < T; ++t) {
                const uint64_t adesc = ((uint64_t)a_hi32 << 32) | ((a0 + 32 * t + 2 * ks) & 0x3FFFu) | (1u << 16);
                ptx::mma_f16(dcol + t * C::DCOLS, adesc, bdesc, idesc_n, first);
              }
              if (PASSES >= 2 && !LO8) {        // fp16 low-half plane (layer 1, NH = 1)
#pragma unroll
                for (int t = 0; t < T; ++t) {
                  const uint64_t adesc_lo = ((uint64_t)a_hi32 << 32) | ((a0 + 32 * t + 2 * ks + (C::A_PLANE >> 4)) & 0x3FFFu) | (1u << 16);
                  ptx::mma_f16(dcol + t * C::DCOLS, adesc_lo, bdesc, make_idesc_f16(128, COUT), 1u);
                }
              }
              if (PASSES == 3 && !C::NCAT) {    // a_hi x w_lo as its own MMA (layer 1, NH = 1)
                const uint64_t bdesc_lo = ((uint64_t)b_hi32 << 32) | (b0 + COUT) | (B_LBO << 16);
#pragma unroll
                for (int t = 0; t < T; ++t) {
                  const uint64_t adesc = ((uint64_t)a_hi32 << 32) | ((a0 + 32 * t + 2 * ks) & 0x3FFFu) | (1u << 16);
                  ptx::mma_f16(dcol + t * C::DCOLS, adesc, bdesc_lo, make_idesc_f16(128, COUT), 1u);
                }
              }
            }
            if (LO8) {
              // a_lo (e4m3, SWIZZLE_32B rows of 32 channels) x w (e4m3, [2][ky][NF][16 B]): one K=32 MMA per M-tile.  With
              // N-concatenated weights the e4m3 rows of the w_lo half are zero (stacked rows must stay NF columns apart).
              const uint32_t a8 = a_col8 + s * C::HX * 2;
              const uint32_t n8 = (NH > 1 || !C::NCAT) ? C::NF * nv : COUT;
              const uint64_t bdesc8 = ((uint64_t)b_hi32 << 32) | (b_col8 + pos * C::NF) | ((uint32_t)C::B8_LBO << 16);
#pragma unroll
              for (int t = 0; t < T; ++t) {
                const uint64_t adesc8 = ((uint64_t)a8_hi32 << 32) | ((a8 + 16 * t) & 0x3FFFu) | (1u << 16);
                ptx::mma_f8(dcol + t * C::DCOLS, adesc8, bdesc8, make_idesc_f16(128, n8), 1u);
              }
            }
          }
          if (dx == KS - 1 && c == C::NCHUNK - 1 && tile + (int)gridDim.x < P.num_tiles) {
            // look ahead while the MMAs just issued execute (no dependence on this tile's commits: other ring slots)
            const uint32_t nas = (it + 1) % C::NACC, nsa = (ia + 1) % C::NA;
            ptx::mbar_wait(&acc_empty[nas], (((it + 1) / C::NACC) & 1) ^ 1);
            ptx::mbar_wait(&a_full[nsa], ((ia + 1) / C::NA) & 1);
            ptx::tc_fence_after();
            ahead = true;
          }
          if (!C::W_RESIDENT) ptx::tc_commit(&w_empty[sw]);               // weight column free once these MMAs have read it
          if (dx == KS - 1) {
            ptx::tc_commit(&a_empty[sa]);                                 // activation chunk free
            if (c == C::NCHUNK - 1) ptx::tc_commit(&acc_full[as]);        // tile complete -> epilogue
          }
        }
      }
    }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue: TMEM -> registers -> bias/ReLU/BN -> fp16 hi/lo (or fp32) -> HBM ==============
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;             // two warps share a lane quarter and split the (t, 16-column) items
    const int m = q * 32 + lane;                  // accumulator row = pixel of the M-tile: 16 row groups x 8 columns
    const int prow = m >> 3, pcol = m & 7;        // (row group g covers image rows g*NH .. g*NH + NH-1, one per column block)
    uint32_t it = 0;
    float bb[BREG ? COUT : 1], tt[BREG ? COUT : 1];
    if (BREG) {
#pragma unroll
      for (int i = 0; i < COUT; ++i) { bb[i] = E.b[i]; tt[i] = E.t[i]; }
    }
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
      const int y0 = (r / P.tiles_x) * 16 * NH, x0 = (r % P.tiles_x) * 8 * T;
      const uint32_t as = it % C::NACC;
      ptx::mbar_wait(&acc_full[as], (it / C::NACC) & 1);
      ptx::tc_fence_after();
      // An item is CW channels of one pixel: 32 where the output is an activation plane, so that every global store is a
      // whole 32-byte sector (hi: 64 B = two 256-bit stores, e4m3 lo: 32 B = one).  With 16-byte stores of half sectors the
      // epilogue alone ran at 2.7 TB/s and was the critical path of every layer but the second (profiles/r1_history.md).
      constexpr int CW = (COUT >= 32 && OUTMODE != TC_OUT_FINAL) ? 32 : 16;
      // The two warps of a lane quarter split the pixel blocks (M-tile t, stacked row h) when their number is even, else the
      // channel blocks; the channel-block loop is unrolled so the channel offset n0 is a compile-time register index.
      constexpr int NBW = COUT / CW, NPB = T * NH;
      constexpr bool SPLIT_PB = (NPB % 2 == 0) || (NBW % 2 != 0);
#pragma unroll 1
      for (int pb = SPLIT_PB ? half : 0; pb < NPB; pb += SPLIT_PB ? 2 : 1) {
      const int t = pb / NH, hrow = pb - t * NH;
#pragma unroll
      for (int nb = 0; nb < NBW; ++nb) {
        if (!SPLIT_PB && (nb & 1) != half) continue;
        constexpr int NB_STEP = SPLIT_PB ? 1 : 2;
        const bool last_item = (SPLIT_PB ? pb + 2 >= NPB : pb == NPB - 1) && (nb + NB_STEP >= NBW);
        const int n0 = nb * CW;
        const int x = x0 + 8 * t + pcol, y = y0 + prow * NH + hrow;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * (T * C::DCOLS) + t * C::DCOLS + hrow * C::NF;
        uint32_t rr[CW];
#pragma unroll
        for (int l = 0; l < CW / 16; ++l) ptx::tmem_ld16(taddr + n0 + 16 * l, reinterpret_cast<uint32_t(&)[16]>(rr[16 * l]));
        if (C::NCAT) {
          uint32_t r2[CW];
#pragma unroll
          for (int l = 0; l < CW / 16; ++l) ptx::tmem_ld16(taddr + COUT + n0 + 16 * l, reinterpret_cast<uint32_t(&)[16]>(r2[16 * l]));
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < CW; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) + __uint_as_float(r2[i]));
        }
        ptx::tmem_ld_wait();
        if (last_item) {                         // last read of this accumulator set by this thread: hand it back
          ptx::tc_fence_before();
          ptx::mbar_arrive(&acc_empty[as]);
        }
        float v[CW];
        if (BREG) {
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const float a = fmaf(__uint_as_float(rr[i]), P.inv_wscale, bb[n0 + i]);
            v[i] = P.relu ? fmaxf(a, 0.f) + tt[n0 + i] : a;
          }
        } else {
          const float4* eb = reinterpret_cast<const float4*>(sEpi + n0);
          const float4* es = reinterpret_cast<const float4*>(sEpi + COUT + n0);
          const float4* et = reinterpret_cast<const float4*>(sEpi + 2 * COUT + n0);
#pragma unroll
          for (int i4 = 0; i4 < CW / 4; ++i4) {
            const float4 b = eb[i4], sc = es[i4], sh = et[i4];
            const float a0 = __uint_as_float(rr[4 * i4 + 0]) * P.inv_wscale + b.x;
            const float a1 = __uint_as_float(rr[4 * i4 + 1]) * P.inv_wscale + b.y;
            const float a2 = __uint_as_float(rr[4 * i4 + 2]) * P.inv_wscale + b.z;
            const float a3 = __uint_as_float(rr[4 * i4 + 3]) * P.inv_wscale + b.w;
            v[4 * i4 + 0] = P.relu ? fmaxf(a0, 0.f) * sc.x + sh.x : a0;
            v[4 * i4 + 1] = P.relu ? fmaxf(a1, 0.f) * sc.y + sh.y : a1;
            v[4 * i4 + 2] = P.relu ? fmaxf(a2, 0.f) * sc.z + sh.z : a2;
            v[4 * i4 + 3] = P.relu ? fmaxf(a3, 0.f) * sc.w + sh.w : a3;
          }
        }
        if (OUTMODE == TC_OUT_FINAL) {
          for (int i = 0; i < CW; ++i) {
            const int n = n0 + i;
            if (n < P.out_c) {
              float a = v[i];
              if (P.softplus) a = softplus_f(a);
              if (P.out_dq) {
                P.out_dq[((long long)img * P.out_c + n) * P.ny * P.nx + (long long)y * P.nx + x] = (double)__fmul_rn(a, n ? P.dq_ys1 : P.dq_ys0) * P.dq_weight;
              } else {
                float* o = P.out_f32 + (long long)img * P.out_bs + ((long long)n * P.ny + y) * P.nx + x;
                *o = P.accumulate ? *o + a : a;
              }
            }
          }
        } else {
          const int op = P.out_pad, HPo = P.ny + 2 * op, WPo = P.nx + 2 * op;
          // circular halo: rows / columns within ``op`` of a border are also written to the wrapped positions
          int ys[2], xs[2], nys = 1, nxs = 1;
          ys[0] = y + op; xs[0] = x + op;
          if (y < op) ys[nys++] = y + op + P.ny; else if (y >= P.ny - op) ys[nys++] = y + op - P.ny;
          if (x < op) xs[nxs++] = x + op + P.nx; else if (x >= P.nx - op) xs[nxs++] = x + op - P.nx;
          uint32_t hi[CW / 2], lo8[CW / 4];
#pragma unroll
          for (int e = 0; e < CW / 2; ++e) {
            const float f0 = v[2 * e], f1 = v[2 * e + 1];
            const __half2 h2 = __floats2half2_rn(f0, f1);
            hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
            if (OUTMODE == TC_OUT_HILO) {
              // low half as e4m3 of (f - hi) * 2^11 (the consumer's e4m3 weights carry the 2^-11)
              const float2 back = __half22float2(h2);
              const unsigned short p = __nv_cvt_float2_to_fp8x2(make_float2((f0 - back.x) * 2048.f, (f1 - back.y) * 2048.f),
                                                                __NV_SATFINITE, __NV_E4M3);
              if (e & 1) lo8[e >> 1] |= (uint32_t)p << 16; else lo8[e >> 1] = p;
            }
          }
          const int chunk = n0 >> 5, within = n0 & 31;
          for (int a = 0; a < nys; ++a)
            for (int b = 0; b < nxs; ++b) {
              const long long pix = (((long long)img * P.out_nch + chunk) * HPo + ys[a]) * WPo + xs[b];
              __half* dh = P.out_hi + pix * 32 + within;
#pragma unroll
              for (int l = 0; l < CW / 16; ++l) ptx::st_global_v8(dh + 16 * l, &hi[8 * l]);
              if (OUTMODE == TC_OUT_HILO) {
                unsigned char* dl = reinterpret_cast<unsigned char*>(P.out_lo) + pix * 32 + within;
                if (CW == 32) ptx::st_global_v8(dl, lo8);
                else *reinterpret_cast<uint4*>(dl) = make_uint4(lo8[0], lo8[1], lo8[2], lo8[3]);
              }
            }
        }
      }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, C::NCOLS);
}

// ---------------------------------------------------------------------------------------- direct layer 1 ----
// Layer 1 (4 or 2 -> 128, 5x5) without an im2col operand.  The im2col variant above writes 512 B of shared memory per pixel
// (K = 100 padded to 128, hi and lo planes) and is bound by the LSU pipe (82 % busy, tensor pipe 36 %).  Here a pixel of the
// input window is ONE 32-byte row of 16 fp16 "channels"
//        [ a_hi (F) | a_lo (F) | a_hi (F) | 0 ... ]        F = real input channels, a = a_hi + a_lo
// and the weights of a tap are a 16 x 128 matrix with rows   [ w_hi (F) | w_hi (F) | w_lo (F) | 0 ... ],
// so ONE K = 16 tcgen05.mma per tap and M-tile accumulates the whole split-precision product a_hi w_hi + a_lo w_hi + a_hi w_lo,
// the tap being just a different start address of the same window (like every other layer).  25 MMAs of N = 128 per M-tile
// instead of 24 (same tensor time), 13 KB instead of 228 KB of builder traffic per 16 x 16 tile.
// Shared memory: window stages [20][20] x 32 B (SWIZZLE_32B rows), all 25 tap matrices resident (100 KB).
constexpr int kL1Win = 20, kL1AStage = 13312 /* 20*20*32 = 12800 -> 1024-aligned */, kL1NA = 4, kL1WBytes = 25 * 4096;
constexpr int kL1Smem = kL1NA * kL1AStage + kL1WBytes + 1536 + 256 + 1024;

template <int F, int OUTMODE>
__global__ void __launch_bounds__(512, 1) conv_l1_direct_kernel(const __grid_constant__ TcConvParams P, const __grid_constant__ TcEpi E) {
  constexpr int T = 2, COUT = 128, NACC = 2, DCOLS = 128, NCOLS = 512, HX = kL1Win;
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char* smem = tc_smem_raw + ((1024u - (ptx::smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sW = smem + kL1NA * kL1AStage;
  float* sEpi = reinterpret_cast<float*>(sW + kL1WBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(sEpi) + 1536);
  uint64_t* a_full = bars;                 // [NA]  128 builder arrivals
  uint64_t* a_empty = bars + kL1NA;        // [NA]
  uint64_t* w_full = bars + 2 * kL1NA;     // [1]
  uint64_t* acc_full = w_full + 1;         // [NACC]
  uint64_t* acc_empty = acc_full + NACC;   // [NACC] 256 epilogue arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NACC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) { sEpi[i] = E.b[i]; sEpi[COUT + i] = E.s[i]; sEpi[2 * COUT + i] = E.t[i]; }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kL1NA; ++i) { ptx::mbar_init(&a_full[i], 128); ptx::mbar_init(&a_empty[i], 1); }
    ptx::mbar_init(w_full, 1);
    for (int i = 0; i < NACC; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 256); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_slot, NCOLS);
  ptx::pdl_launch_dependents();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  ptx::pdl_wait();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = P.tiles_y * P.tiles_x;

  if (warp >= 12) {
    // ===================== window builders: 4 warps, fp32 -> [a_hi | a_lo | a_hi | 0] rows, 32-byte swizzle ==============
    // One thread per (window row, pixel quad): the 20-pixel window row x0-2 .. x0+17 is covered by the six aligned quads
    // x0-4 .. x0+19 (x0 is a multiple of 16, so no quad straddles the periodic wrap): 120 of the 128 builder threads load one
    // float4 per channel -- or GENERATE the two latent channels from the Philox counters of that quad, exactly what the latent
    // kernel would have written (closure.cuh) -- and write the rows of their (up to) four pixels.
    const int bt = threadIdx.x - 384;                      // 0..127
    const int wr = bt / 6, wq = bt - 6 * wr;               // window row 0..19 (bt < 120), quad 0..5
    uint32_t ia = 0;
    const long long cs = (long long)P.ny * P.nx;
    const bool gen = F == 4 && P.noise;
    const uint32_t draw = gen ? *P.noise_draw : 0u;
    auto fetch = [&](int tile, float4 (&f)[4]) {
#pragma unroll
      for (int c = 0; c < 4; ++c) f[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tile < P.num_tiles && bt < 120) {
        const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
        const int y0 = (r / P.tiles_x) * 16, x0 = (r % P.tiles_x) * 16;
        int sy = y0 + wr - 2, sx = x0 - 4 + 4 * wq;
        sy = sy < 0 ? sy + P.ny : (sy >= P.ny ? sy - P.ny : sy);
        sx = sx < 0 ? sx + P.nx : (sx >= P.nx ? sx - P.nx : sx);
        const float* src = P.x_f32 + (long long)img * P.x_bs + (long long)sy * P.nx + sx;
        f[0] = *reinterpret_cast<const float4*>(src);
        f[1] = *reinterpret_cast<const float4*>(src + cs);
        if (gen) {
          const uint32_t quad = (uint32_t)((sy * P.nx + sx) >> 2), mem = (uint32_t)(P.noise_member0 + img);
          float z[4];
          philox_normal4(P.noise_seed, mem, draw, 0u, quad, z);
          f[2] = make_float4(z[0], z[1], z[2], z[3]);
          philox_normal4(P.noise_seed, mem, draw, 1u, quad, z);
          f[3] = make_float4(z[0], z[1], z[2], z[3]);
        } else if (F == 4) {
          f[2] = *reinterpret_cast<const float4*>(src + 2 * cs);
          f[3] = *reinterpret_cast<const float4*>(src + 3 * cs);
        }
      }
    };
    float4 nxt[4];
    fetch(blockIdx.x, nxt);
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++ia) {
      float4 cur[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) cur[c] = nxt[c];
      fetch(tile + gridDim.x, nxt);
      const uint32_t s = ia % kL1NA, par = (ia / kL1NA) & 1;
      ptx::mbar_wait(&a_empty[s], par ^ 1);
      unsigned char* stage = sA + s * kL1AStage;
      if (bt < 120) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int wx = 4 * wq - 2 + j;                   // window column of pixel j of the quad
          if (wx >= 0 && wx < HX) {
            const float vals[4] = {j == 0 ? cur[0].x : j == 1 ? cur[0].y : j == 2 ? cur[0].z : cur[0].w,
                                   j == 0 ? cur[1].x : j == 1 ? cur[1].y : j == 2 ? cur[1].z : cur[1].w,
                                   j == 0 ? cur[2].x : j == 1 ? cur[2].y : j == 2 ? cur[2].z : cur[2].w,
                                   j == 0 ? cur[3].x : j == 1 ? cur[3].y : j == 2 ? cur[3].z : cur[3].w};
            __half row[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) row[k] = __float2half_rn(0.f);
#pragma unroll
            for (int c = 0; c < F; ++c) {
              const __half hi = __float2half_rn(vals[c]);
              row[c] = hi; row[F + c] = __float2half_rn(vals[c] - __half2float(hi)); row[2 * F + c] = hi;
            }
            const uint4* r4 = reinterpret_cast<const uint4*>(row);
            const int p = wr * HX + wx;
            unsigned char* dst = stage + p * 32;
            const int sw = (p >> 2) & 1;                    // 32-byte swizzle: 16-byte chunk ^= address bit 7
            *reinterpret_cast<uint4*>(dst + ((0 ^ sw) << 4)) = r4[0];
            *reinterpret_cast<uint4*>(dst + ((1 ^ sw) << 4)) = r4[1];
          }
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(&a_full[s]);
    }
  } else if (warp == 0) {
    // ===================== weights: all 25 tap matrices, once =========================================================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(w_full, kL1WBytes);
      for (int c = 0; c < 4; ++c)
        ptx::bulk_g2s(sW + c * (kL1WBytes / 4), reinterpret_cast<const unsigned char*>(P.w) + c * (kL1WBytes / 4), kL1WBytes / 4, w_full);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===================================================================================
    // A: K-major SWIZZLE_32B, pixel rows of 32 B, 8-row groups one image row (HX pixels) apart; B: K-major no swizzle,
    // tap matrix [2 k-chunks][128 n][8 halves]: core matrices 8 rows x 16 B, next k-chunk 2048 B further
    constexpr uint32_t a_hi32 = (uint32_t)(HX * 2) | (1u << 14) | (6u << 29);
    constexpr uint32_t b_hi32 = 8u | (1u << 14);
    constexpr uint32_t idesc = make_idesc_f16(128, 128);
    const uint32_t sA_u = ptx::smem_u32(sA) >> 4, sW_u = ptx::smem_u32(sW) >> 4;
    uint32_t ia = 0, it = 0;
    if (ptx::elect_one_sync()) {
      ptx::mbar_wait(w_full, 0);
      ptx::tc_fence_after();
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it, ++ia) {
        const uint32_t as = it % NACC, sa = ia % kL1NA;
        ptx::mbar_wait(&acc_empty[as], ((it / NACC) & 1) ^ 1);
        ptx::mbar_wait(&a_full[sa], (ia / kL1NA) & 1);
        ptx::tc_fence_after();
        const uint32_t dbase = tmem_base + as * (T * DCOLS);
        const uint32_t a_stage = sA_u + sa * (kL1AStage >> 4);
#pragma unroll 1
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const uint32_t a0 = a_stage + (dy * HX + dx) * 2;
            const uint64_t bdesc = ((uint64_t)b_hi32 << 32) | (sW_u + (dy * 5 + dx) * 256) | (128u << 16);
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const uint64_t adesc = ((uint64_t)a_hi32 << 32) | ((a0 + 16 * t) & 0x3FFFu) | (1u << 16);
              ptx::mma_f16(dbase + t * DCOLS, adesc, bdesc, idesc, (dy | dx) ? 1u : 0u);
            }
          }
        ptx::tc_commit(&a_empty[sa]);
        ptx::tc_commit(&acc_full[as]);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ===================== epilogue: TMEM -> bias / ReLU / BN -> fp16 hi (+ e4m3 lo) with the circular halo -> HBM =========
    const int q = warp & 3, half = (warp - 4) >> 2;
    const int m = q * 32 + lane, prow = m >> 3, pcol = m & 7;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
      const int y0 = (r / P.tiles_x) * 16, x0 = (r % P.tiles_x) * 16;
      const uint32_t as = it % NACC;
      ptx::mbar_wait(&acc_full[as], (it / NACC) & 1);
      ptx::tc_fence_after();
      constexpr int CW = 32, NBW = COUT / CW;
      const int t = half;                                  // the two warps of a lane quarter take one M-tile each
#pragma unroll
      for (int nb = 0; nb < NBW; ++nb) {
        const int n0 = nb * CW;
        const int x = x0 + 8 * t + pcol, y = y0 + prow;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * (T * DCOLS) + t * DCOLS;
        uint32_t rr[CW];
        ptx::tmem_ld16(taddr + n0, reinterpret_cast<uint32_t(&)[16]>(rr[0]));
        ptx::tmem_ld16(taddr + n0 + 16, reinterpret_cast<uint32_t(&)[16]>(rr[16]));
        ptx::tmem_ld_wait();
        if (nb == NBW - 1) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&acc_empty[as]);
        }
        float v[CW];
        const float4* eb = reinterpret_cast<const float4*>(sEpi + n0);
        const float4* es = reinterpret_cast<const float4*>(sEpi + COUT + n0);
        const float4* et = reinterpret_cast<const float4*>(sEpi + 2 * COUT + n0);
#pragma unroll
        for (int i4 = 0; i4 < CW / 4; ++i4) {
          const float4 b = eb[i4], sc = es[i4], sh = et[i4];
          const float a0 = __uint_as_float(rr[4 * i4 + 0]) * P.inv_wscale + b.x, a1 = __uint_as_float(rr[4 * i4 + 1]) * P.inv_wscale + b.y;
          const float a2 = __uint_as_float(rr[4 * i4 + 2]) * P.inv_wscale + b.z, a3 = __uint_as_float(rr[4 * i4 + 3]) * P.inv_wscale + b.w;
          v[4 * i4 + 0] = fmaxf(a0, 0.f) * sc.x + sh.x;
          v[4 * i4 + 1] = fmaxf(a1, 0.f) * sc.y + sh.y;
          v[4 * i4 + 2] = fmaxf(a2, 0.f) * sc.z + sh.z;
          v[4 * i4 + 3] = fmaxf(a3, 0.f) * sc.w + sh.w;
        }
        const int op = P.out_pad, HPo = P.ny + 2 * op, WPo = P.nx + 2 * op;
        int ys[2], xs[2], nys = 1, nxs = 1;
        ys[0] = y + op; xs[0] = x + op;
        if (y < op) ys[nys++] = y + op + P.ny; else if (y >= P.ny - op) ys[nys++] = y + op - P.ny;
        if (x < op) xs[nxs++] = x + op + P.nx; else if (x >= P.nx - op) xs[nxs++] = x + op - P.nx;
        uint32_t hi[CW / 2], lo8[CW / 4];
#pragma unroll
        for (int e = 0; e < CW / 2; ++e) {
          const float f0 = v[2 * e], f1 = v[2 * e + 1];
          const __half2 h2 = __floats2half2_rn(f0, f1);
          hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
          if (OUTMODE == TC_OUT_HILO) {
            const float2 back = __half22float2(h2);
            const unsigned short p8 = __nv_cvt_float2_to_fp8x2(make_float2((f0 - back.x) * 2048.f, (f1 - back.y) * 2048.f), __NV_SATFINITE, __NV_E4M3);
            if (e & 1) lo8[e >> 1] |= (uint32_t)p8 << 16; else lo8[e >> 1] = p8;
          }
        }
        const int chunk = n0 >> 5;
        for (int a = 0; a < nys; ++a)
          for (int b2 = 0; b2 < nxs; ++b2) {
            const long long pix = (((long long)img * P.out_nch + chunk) * HPo + ys[a]) * WPo + xs[b2];
            __half* dh = P.out_hi + pix * 32;
            ptx::st_global_v8(dh, &hi[0]);
            ptx::st_global_v8(dh + 16, &hi[8]);
            if (OUTMODE == TC_OUT_HILO) ptx::st_global_v8(reinterpret_cast<unsigned char*>(P.out_lo) + pix * 32, lo8);
          }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, NCOLS);
}

// ------------------------------------------------------------------------------------------------ host side ----
// (TcLayer / TcNet / TcWorkspace live in cnn_tc_host.hpp)
bool tc_l1_direct_enabled() {
  static int mode = -1;
  if (mode < 0) { const char* e1 = getenv("QGB_TC_L1"); mode = (e1 && std::strcmp(e1, "im2col") == 0) ? 0 : 1; }
  return mode == 1;
}
void tc_free_net(TcNet& n) {
  for (auto& L : n.layers) cudaFree(L.w);
  cudaFree(n.l2_fast.w);
  n.l2_fast = TcLayer();
  cudaFree(n.l1_direct.w);
  n.l1_direct = TcLayer();
  n.layers.clear();
  n.ready = false;
}
void tc_free_workspace(TcWorkspace& w) {
  for (int i = 0; i < 6; ++i) { cudaFree(w.buf[i]); w.buf[i] = nullptr; w.halves[i] = 0; }
}

// Pack one layer.  Weight stage = one (32-channel chunk, tap COLUMN kx): fp16 [4 j][ky descending][WROWS][8] (values scaled
// by 2^k; WROWS = [w_hi | w_lo] x cout_p for the N-concatenated layers, [plane][cout_p] otherwise) followed (lo8) by
// e4m3 [2 kc][ky descending][NF][16] = w * 2^k * 2^-11 (rows of the w_lo half stay zero).  Descending ky makes the
// weights of the row-stacked MMA (taps ky = s - h, h ascending) one contiguous B operand.  Epilogue vectors are padded.
inline bool tc_pack_layer(TcLayer& L, int cin_p, int cout_p, int ks, int passes, bool lo8, int real_cin, int real_cout,
                          const std::vector<float>& wdense /* [real_cout][real_cin][ks*ks] */, const float* bias,
                          const float* bnscale, const float* shift, int relu) {
  L.cin = cin_p; L.cout = cout_p; L.ks = ks; L.relu = relu; L.real_cout = real_cout; L.passes = passes;
  const int taps = ks * ks, planes = passes == 3 ? 2 : 1, nchunk = cin_p / 32;
  const bool ncat = passes == 3 && cout_p <= 32;
  const int nf = ncat ? 2 * cout_p : cout_p, wrows = ncat ? nf : planes * cout_p;
  float maxabs = 0.f;
  for (float v : wdense) maxabs = std::fmax(maxabs, std::fabs(v));
  int k = 0;
  if (maxabs > 0.f) {
    k = (int)std::floor(std::log2(16384.0 / (double)maxabs));
    if (k < 0) k = 0;
    if (k > 24) k = 24;
  }
  const float scale = std::ldexp(1.0f, k);
  L.inv_wscale = std::ldexp(1.0f, -k);
  const size_t w16 = (size_t)4 * ks * wrows * 16, w8 = lo8 ? (size_t)2 * ks * nf * 16 : 0, slab = w16 + w8;
  std::vector<unsigned char> pk((size_t)nchunk * ks * slab, 0);
  for (int c = 0; c < nchunk; ++c)
    for (int kx = 0; kx < ks; ++kx) {
      unsigned char* base = pk.data() + (size_t)(c * ks + kx) * slab;
      __half* b16 = reinterpret_cast<__half*>(base);
      unsigned char* b8 = base + w16;
      for (int ky = 0; ky < ks; ++ky) {
        const int pos = ks - 1 - ky, tap = ky * ks + kx;
        for (int j = 0; j < 4; ++j)
          for (int co = 0; co < cout_p; ++co)
            for (int e = 0; e < 8; ++e) {
              const int ci = c * 32 + j * 8 + e;
              float v = 0.f;
              if (ci < real_cin && co < real_cout) v = wdense[((size_t)co * real_cin + ci) * taps + tap] * scale;
              const __half h = __float2half_rn(v);
              b16[(((size_t)j * ks + pos) * wrows + co) * 8 + e] = h;
              if (planes == 2) b16[(((size_t)j * ks + pos) * wrows + cout_p + co) * 8 + e] = __float2half_rn(v - __half2float(h));
              if (lo8) b8[(((size_t)(ci % 32 / 16) * ks + pos) * nf + co) * 16 + (ci % 16)] =
                  (unsigned char)__nv_cvt_float_to_fp8(v * (1.0f / 2048.0f), __NV_SATFINITE, __NV_E4M3);
            }
      }
    }
  for (int i = 0; i < 128; ++i) {
    L.epi.b[i] = i < real_cout ? bias[i] : 0.f;
    L.epi.s[i] = (relu && i < real_cout) ? bnscale[i] : 1.f;
    L.epi.t[i] = (relu && i < real_cout) ? shift[i] : 0.f;
  }
  if (cudaMalloc(&L.w, pk.size()) != cudaSuccess) return false;
  if (cudaMemcpy(L.w, pk.data(), pk.size(), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  return true;
}

// Direct layer 1 (conv_l1_direct_kernel): tap matrix [2 k-chunks][128 n][8 halves] with K rows [w_hi (F) | w_hi (F) | w_lo (F) | 0]
inline bool tc_pack_l1_direct(TcLayer& L, int F, const float* w /* (128, F, 5, 5) */, const float* bias, const float* bnscale,
                              const float* shift) {
  L.cin = 16; L.cout = 128; L.ks = 5; L.relu = 1; L.real_cout = 128; L.passes = 3;
  float maxabs = 0.f;
  for (int i = 0; i < 128 * F * 25; ++i) maxabs = std::fmax(maxabs, std::fabs(w[i]));
  int k = 0;
  if (maxabs > 0.f) { k = (int)std::floor(std::log2(16384.0 / (double)maxabs)); if (k < 0) k = 0; if (k > 24) k = 24; }
  const float scale = std::ldexp(1.0f, k);
  L.inv_wscale = std::ldexp(1.0f, -k);
  std::vector<__half> pk((size_t)25 * 2 * 128 * 8, __float2half_rn(0.f));
  for (int tap = 0; tap < 25; ++tap)
    for (int n = 0; n < 128; ++n)
      for (int c = 0; c < F; ++c) {
        const float v = w[((size_t)n * F + c) * 25 + tap] * scale;
        const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
        const int rows[3] = {c, F + c, 2 * F + c};
        const __half vals[3] = {h, h, l};
        for (int j = 0; j < 3; ++j) pk[(((size_t)tap * 2 + rows[j] / 8) * 128 + n) * 8 + rows[j] % 8] = vals[j];
      }
  for (int i = 0; i < 128; ++i) { L.epi.b[i] = bias[i]; L.epi.s[i] = bnscale[i]; L.epi.t[i] = shift[i]; }
  if (cudaMalloc(&L.w, pk.size() * sizeof(__half)) != cudaSuccess) return false;
  return cudaMemcpy(L.w, pk.data(), pk.size() * sizeof(__half), cudaMemcpyHostToDevice) == cudaSuccess;
}

template <int F, int OUTMODE>
inline cudaError_t tc_launch_l1_direct(const TcConvParams& P, const TcEpi& E, int nsm, cudaStream_t st) {
  auto kern = conv_l1_direct_kernel<F, OUTMODE>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kL1Smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P.num_tiles < nsm ? P.num_tiles : nsm);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = kL1Smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, P, E);
}

// Accepts exactly the default AndrewCNN architecture: (4|2)->128 (5x5) ->64 (5x5) ->32 (3x3) -> 4 x [32->32 (3x3)] -> 2 (3x3)
int tc_pack_net(TcNet& n, int nlayers, const qgb_cnn_layer* L, std::string* err) {
  tc_free_net(n);
  static const int cin[8] = {0, 128, 64, 32, 32, 32, 32, 32}, cout[8] = {128, 64, 32, 32, 32, 32, 32, 2};
  static const int ks[8] = {5, 5, 3, 3, 3, 3, 3, 3};
  if (nlayers != 8 || (L[0].cin != 4 && L[0].cin != 2)) { *err = "not the default AndrewCNN architecture"; return 0; }
  for (int i = 0; i < 8; ++i)
    if ((i > 0 && L[i].cin != cin[i]) || L[i].cout != cout[i] || L[i].ksize != ks[i] || L[i].relu_bn != (i < 7)) {
      *err = "not the default AndrewCNN architecture";
      return 0;
    }
  n.cin0 = L[0].cin;
  n.kp = n.cin0 == 4 ? 128 : 64;
  n.layers.resize(8);
  // BatchNorm scale folding for the thin layers (COUT <= 32, epilogue constants in registers).  Layer i computes
  // y = relu(z) * s + t per channel (z = conv + bias).  With a = |s|, g = sign(s): relu(z) * s = g * relu(a z), so a scales
  // the layer's weight rows and bias, the epilogue stores  yhat = relu(a z) + g t  and the next layer's weights take the
  // sign, w'[co][ci] = w[co][ci] * g[ci]  (y = g * yhat).  Layers 1 and 2 keep scale and shift in the epilogue.
  auto folded = [&](int i) { return i >= 0 && i < 7 && cout[i] <= 32; };
  auto bn_abs = [&](int i, int c) { return folded(i) ? std::fabs(L[i].bn_scale[c]) : 1.f; };
  auto bn_sgn = [&](int i, int c) { return (folded(i) && L[i].bn_scale[c] < 0.f) ? -1.f : 1.f; };
  auto epi_scale = [&](int i) {
    std::vector<float> v(cout[i], 1.f);
    if (i < 7 && !folded(i)) for (int c = 0; c < cout[i]; ++c) v[c] = L[i].bn_scale[c];
    return v;
  };
  auto epi_shift = [&](int i) {
    std::vector<float> v(cout[i], 0.f);
    if (i < 7) for (int c = 0; c < cout[i]; ++c) v[c] = L[i].bn_shift[c] * bn_sgn(i, c);
    return v;
  };
  // layer 1 as a 1x1 convolution over the im2col tensor: K index = tap*cin0 + ch
  {
    const int K = 25 * n.cin0;
    std::vector<float> wd((size_t)128 * K), sf = epi_scale(0), tf = epi_shift(0);
    for (int co = 0; co < 128; ++co)
      for (int ch = 0; ch < n.cin0; ++ch)
        for (int tap = 0; tap < 25; ++tap) wd[(size_t)co * K + tap * n.cin0 + ch] = L[0].weight[((size_t)co * n.cin0 + ch) * 25 + tap];
    if (!tc_pack_layer(n.layers[0], n.kp, 128, 1, 3, false, K, 128, wd, L[0].bias, sf.data(), tf.data(), 1)) { *err = "cuda"; return QGB_ECUDA; }
    if (!tc_pack_l1_direct(n.l1_direct, n.cin0, L[0].weight, L[0].bias, sf.data(), tf.data())) { *err = "cuda"; return QGB_ECUDA; }
  }
  for (int i = 1; i < 8; ++i) {
    const int taps = ks[i] * ks[i];
    std::vector<float> wd((size_t)cout[i] * cin[i] * taps), bf(cout[i]), sf = epi_scale(i), tf = epi_shift(i);
    for (int co = 0; co < cout[i]; ++co) {
      for (int ci = 0; ci < cin[i]; ++ci)
        for (int tap = 0; tap < taps; ++tap) {
          const size_t idx = ((size_t)co * cin[i] + ci) * taps + tap;
          wd[idx] = L[i].weight[idx] * bn_sgn(i - 1, ci) * bn_abs(i, co);
        }
      bf[co] = L[i].bias[co] * bn_abs(i, co);
    }
    const int cout_p = i == 7 ? 16 : cout[i];
    const int passes = i == 1 ? 2 : 3;          // layer 2: (a_hi + a_lo) w_hi ; others: full split precision
    if (!tc_pack_layer(n.layers[i], cin[i], cout_p, ks[i], passes, true, cin[i], cout[i], wd, bf.data(), sf.data(), tf.data(), i < 7)) { *err = "cuda"; return QGB_ECUDA; }
    if (i == 1 && !tc_pack_layer(n.l2_fast, cin[i], cout_p, ks[i], 1, false, cin[i], cout[i], wd, bf.data(), sf.data(), tf.data(), 1)) { *err = "cuda"; return QGB_ECUDA; }
  }
  n.ready = true;
  return 0;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}
// 4-D map over an activation plane [nchunks][HP][WP][32 halves], box = (32, hx, hy, 1), 64-byte swizzle
inline bool tc_make_map(CUtensorMap* m, const __half* base, int WP, int HP, long long nchunks, int hx, int hy) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  cuuint64_t gdim[4] = {32, (cuuint64_t)WP, (cuuint64_t)HP, (cuuint64_t)nchunks};
  cuuint64_t gstr[3] = {64, (cuuint64_t)WP * 64, (cuuint64_t)HP * WP * 64};
  cuuint32_t box[4] = {32, (cuuint32_t)hx, (cuuint32_t)hy, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D map over an e4m3 lo plane [nchunks][HP][WP][32 bytes], box = (32, hx, hy, 1), 32-byte swizzle
inline bool tc_make_map8(CUtensorMap* m, const void* base, int WP, int HP, long long nchunks, int hx, int hy) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) return false;
  cuuint64_t gdim[4] = {32, (cuuint64_t)WP, (cuuint64_t)HP, (cuuint64_t)nchunks};
  cuuint64_t gstr[3] = {32, (cuuint64_t)WP * 32, (cuuint64_t)HP * WP * 32};
  cuuint32_t box[4] = {32, (cuuint32_t)hx, (cuuint32_t)hy, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CIN, int COUT, int KS, int PASSES, int T, int OUTMODE, int FUSE = 0, int NH = 1>
inline cudaError_t tc_launch(const TcConvParams& P, const TcEpi& E, int nimg, int nsm, cudaStream_t st) {
  constexpr bool LO8 = !FUSE && PASSES >= 2;
  using C = TcCfg<CIN, COUT, KS, PASSES, T, LO8, NH>;
  TcMaps M;
  if (FUSE) {
    std::memset(&M, 0, sizeof(M));
  } else if (!tc_make_map(&M.hi, P.in_hi, P.WP, P.HP, (long long)nimg * P.nch_in, C::HX, C::HY)) {
    return cudaErrorInvalidValue;
  } else if (C::PLANES == 2) {
    if (!tc_make_map8(&M.lo, P.in_lo, P.WP, P.HP, (long long)nimg * P.nch_in, C::HX, C::HY)) return cudaErrorInvalidValue;
  } else {
    M.lo = M.hi;
  }
  auto kern = conv_tc_kernel<CIN, COUT, KS, PASSES, T, OUTMODE, FUSE, NH>;
  // the opt-in shared-memory limit is a per-device (per-context) function attribute: track it per device ordinal
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM + (FUSE ? 6400 : 0));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  const int grid = P.num_tiles < nsm ? P.num_tiles : nsm;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(FUSE ? 640 : 384);
  cfg.dynamicSmemBytes = C::SMEM + (FUSE ? 6400 : 0);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, P, M, E);
}

// Tile shape of a layer: NH stacked rows (tile height 16*NH) x T M-tiles of 8 columns.  TMEM holds 2 x T x NH x NF columns.
struct TcTile { int nh, t; };
template <int COUT, int PASSES>
inline TcTile tc_pick_tile(int ny, int nx, bool thin = false) {
  constexpr int NF = (PASSES == 3 && COUT <= 32) ? 2 * COUT : COUT;
  static int force_nh = -1;
  if (force_nh < 0) { const char* e = getenv("QGB_TC_NH"); force_nh = e ? atoi(e) : 0; }
  if (force_nh != 1) {
    // layer 2: four stacked rows (N up to 256, 64 x 8 tiles) where the image height allows
    if (NF == 64 && PASSES != 3 && ny % 64 == 0 && force_nh != 2) return {4, 1};
    // 32 -> 32 layers: 32 x 8 tiles leave room for a 4-deep accumulator ring (measured 228 -> 208 us per 1024 images at 64^2)
    if (ny % 32 == 0 && PASSES == 3 && COUT == 32 && thin) return {2, 1};
    if (ny % 32 == 0 && nx % 16 == 0 && 2 * 2 * 2 * NF <= 512) return {2, 2};
    if (NF == 64 && PASSES != 3 && ny % 48 == 0) return {3, 1};
  }
  return {1, nx % 32 == 0 ? 4 : (nx % 24 == 0 ? 3 : 2)};
}
template <int CIN, int COUT, int KS, int PASSES, int OUTMODE>
inline cudaError_t tc_launch_T(TcTile tl, const TcConvParams& P, const TcEpi& E, int nimg, int nsm, cudaStream_t st) {
  if constexpr (PASSES == 3) {
    if (tl.nh == 2 && tl.t == 1) return tc_launch<CIN, COUT, KS, PASSES, 1, OUTMODE, 0, 2>(P, E, nimg, nsm, st);
  }
  if (tl.nh == 2) return tc_launch<CIN, COUT, KS, PASSES, 2, OUTMODE, 0, 2>(P, E, nimg, nsm, st);
  if constexpr (COUT == 64 && PASSES != 3) {
    if (tl.nh == 4) return tc_launch<CIN, COUT, KS, PASSES, 1, OUTMODE, 0, 4>(P, E, nimg, nsm, st);
    if (tl.nh == 3) return tc_launch<CIN, COUT, KS, PASSES, 1, OUTMODE, 0, 3>(P, E, nimg, nsm, st);
  }
  if (tl.t == 4) return tc_launch<CIN, COUT, KS, PASSES, 4, OUTMODE>(P, E, nimg, nsm, st);
  if (tl.t == 3) return tc_launch<CIN, COUT, KS, PASSES, 3, OUTMODE>(P, E, nimg, nsm, st);
  return tc_launch<CIN, COUT, KS, PASSES, 2, OUTMODE>(P, E, nimg, nsm, st);
}

int tc_forward(const TcNet& net, TcWorkspace& ws, const float* x, long long x_bs, float* y, long long y_bs,
                      int batch, int ny, int nx, int softplus, int accumulate, int nsm, cudaStream_t st, std::string* err,
                      bool fast_l2) {
  if (!net.ready) { *err = "tcgen05 path: network not packed"; return QGB_EUNSUPPORTED; }
  if (ny % 16 || nx % 16) { *err = "tcgen05 path needs ny and nx to be multiples of 16 (use precision='fp32')"; return QGB_EUNSUPPORTED; }
  // Images per launch: large launches amortise the persistent-CTA ramp/tail (measured on B200, 64^2: 128 -> 150 k,
  // 1024 -> 162 k member-steps/s); the workspace is capped at ~6 GB.
  static int max_chunk = 0;
  if (!max_chunk) { const char* e = getenv("QGB_TC_CHUNK"); max_chunk = e ? atoi(e) : 1024; if (max_chunk < 1) max_chunk = 1024; }
  const size_t per_img = (256 * (size_t)(ny + 4) * (nx + 4) + 128 * (size_t)(ny + 2) * (nx + 2)) * 2;
  int chunk = batch < max_chunk ? batch : max_chunk;
  while (chunk > 1 && (size_t)chunk * per_img > (6ull << 30)) chunk = (chunk + 1) / 2;
  // workspace: a0 (im2col, hi/lo), ping (<=128 ch, halo 2), pong (<=64 ch, halo 1)
  const size_t need[6] = {16, 16,   // (the im2col tensor of layer 1 is built on the fly in shared memory)
                          (size_t)chunk * 128 * (ny + 4) * (nx + 4), (size_t)chunk * 128 * (ny + 4) * (nx + 4),
                          (size_t)chunk * 64 * (ny + 2) * (nx + 2), (size_t)chunk * 64 * (ny + 2) * (nx + 2)};
  for (int i = 0; i < 6; ++i)
    if (ws.halves[i] < need[i]) {
      cudaFree(ws.buf[i]);
      ws.buf[i] = nullptr;
      if (cudaMalloc(&ws.buf[i], need[i] * sizeof(__half)) != cudaSuccess) { *err = "cudaMalloc failed (tc workspace)"; return QGB_ECUDA; }
      ws.halves[i] = need[i];
    }
  ws.last_launches = 0;
  __half *a0h = ws.buf[0], *a0l = ws.buf[1], *ping_h = ws.buf[2], *ping_l = ws.buf[3], *pong_h = ws.buf[4], *pong_l = ws.buf[5];
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int nb = batch - b0 < chunk ? batch - b0 : chunk;
    for (int li = 0; li < 8; ++li) {
      const TcLayer& L = (li == 1 && fast_l2) ? net.l2_fast : net.layers[li];
      TcConvParams P;
      P.w = L.w; P.inv_wscale = L.inv_wscale; P.relu = L.relu;
      P.ny = ny; P.nx = nx;
      P.out_f32 = nullptr; P.out_bs = 0; P.out_c = 0; P.softplus = 0; P.accumulate = 0;
      P.x_f32 = x + (long long)b0 * x_bs; P.x_bs = x_bs;
      P.out_hi = P.out_lo = nullptr; P.out_pad = 0; P.out_nch = 0;
      P.noise = 0; P.noise_member0 = 0; P.noise_seed = 0; P.noise_draw = nullptr;
      P.out_dq = nullptr; P.dq_ys0 = P.dq_ys1 = 1.f; P.dq_weight = 1.0;
      TcTile tl = {1, 2};                                        // layer 1 (fused im2col): 16 x 16 tiles
      if (li == 1) tl = fast_l2 ? tc_pick_tile<64, 1>(ny, nx) : tc_pick_tile<64, 2>(ny, nx);
      else if (li >= 2 && li < 7) tl = tc_pick_tile<32, 3>(ny, nx, L.cin == 32);
      else if (li == 7) tl = tc_pick_tile<16, 3>(ny, nx);
      P.tiles_y = ny / (16 * tl.nh);
      P.tiles_x = nx / (8 * tl.t);
      P.num_tiles = nb * P.tiles_y * P.tiles_x;
      const int pad = L.ks / 2;
      P.nch_in = L.cin / 32; P.HP = ny + 2 * pad; P.WP = nx + 2 * pad;
      // buffer rotation: a0 -> ping(hi) -> pong(hi,lo) -> ping(hi,lo) -> pong -> ping -> pong -> ping -> y
      if (li == 0) { P.in_hi = a0h; P.in_lo = a0l; }
      else if (li & 1) { P.in_hi = ping_h; P.in_lo = ping_l; }
      else { P.in_hi = pong_h; P.in_lo = pong_l; }
      if (li < 7) {
        if (li & 1) { P.out_hi = pong_h; P.out_lo = pong_l; } else { P.out_hi = ping_h; P.out_lo = ping_l; }
        P.out_pad = net.layers[li + 1].ks / 2;
        P.out_nch = L.cout / 32;
      } else {
        P.out_f32 = y + (long long)b0 * y_bs; P.out_bs = y_bs; P.out_c = L.real_cout; P.softplus = softplus; P.accumulate = accumulate;
        if (ws.dq_out && L.real_cout == 2 && !softplus && !accumulate) {
          P.out_dq = ws.dq_out + (long long)b0 * 2 * ny * nx; P.dq_ys0 = ws.dq_ys[0]; P.dq_ys1 = ws.dq_ys[1]; P.dq_weight = ws.dq_weight;
        }
      }
      const int pi = ws.prof ? ws.prof->start(8 * ws.prof_net + li, st) : -1;
      cudaError_t e;
      static int l1_mode = -1;     // QGB_TC_L1 = im2col selects the first-generation layer-1 kernel
      if (l1_mode < 0) { const char* e1 = getenv("QGB_TC_L1"); l1_mode = (e1 && std::strcmp(e1, "im2col") == 0) ? 0 : 1; }
      if (li == 0 && l1_mode == 1) {
        if (ws.noise_inkernel && net.cin0 == 4) { P.noise = 1; P.noise_member0 = ws.noise_member0 + b0; P.noise_seed = ws.noise_seed; P.noise_draw = ws.noise_draw; }
        P.w = net.l1_direct.w; P.inv_wscale = net.l1_direct.inv_wscale;
        P.tiles_y = ny / 16; P.tiles_x = nx / 16; P.num_tiles = nb * P.tiles_y * P.tiles_x;
        if (fast_l2) e = net.cin0 == 4 ? tc_launch_l1_direct<4, TC_OUT_HI>(P, net.l1_direct.epi, nsm, st) : tc_launch_l1_direct<2, TC_OUT_HI>(P, net.l1_direct.epi, nsm, st);
        else e = net.cin0 == 4 ? tc_launch_l1_direct<4, TC_OUT_HILO>(P, net.l1_direct.epi, nsm, st) : tc_launch_l1_direct<2, TC_OUT_HILO>(P, net.l1_direct.epi, nsm, st);
      } else if (li == 0) {
        if (fast_l2) e = net.kp == 128 ? tc_launch<128, 128, 1, 3, 2, TC_OUT_HI, 4>(P, L.epi, nb, nsm, st) : tc_launch<64, 128, 1, 3, 2, TC_OUT_HI, 2>(P, L.epi, nb, nsm, st);
        else e = net.kp == 128 ? tc_launch<128, 128, 1, 3, 2, TC_OUT_HILO, 4>(P, L.epi, nb, nsm, st) : tc_launch<64, 128, 1, 3, 2, TC_OUT_HILO, 2>(P, L.epi, nb, nsm, st);
      } else if (li == 1) e = fast_l2 ? tc_launch_T<128, 64, 5, 1, TC_OUT_HILO>(tl, P, L.epi, nb, nsm, st) : tc_launch_T<128, 64, 5, 2, TC_OUT_HILO>(tl, P, L.epi, nb, nsm, st);
      else if (li == 2) e = tc_launch_T<64, 32, 3, 3, TC_OUT_HILO>(tl, P, L.epi, nb, nsm, st);
      else if (li < 7) e = tc_launch_T<32, 32, 3, 3, TC_OUT_HILO>(tl, P, L.epi, nb, nsm, st);
      else e = tc_launch_T<32, 16, 3, 3, TC_OUT_FINAL>(tl, P, L.epi, nb, nsm, st);
      if (ws.prof) ws.prof->stop(pi, st, nb);
      ws.last_launches += 1;
      if (e != cudaSuccess) { *err = std::string("tcgen05 conv launch failed: ") + cudaGetErrorString(e); return QGB_ECUDA; }
    }
  }
  return 0;
}

}  // namespace qgb
