// tgemm.cuh -- fp32-accurate GEMM on the 5th-generation tensor cores (sm_100a) for the discriminator of the CGAN trainer.
//
//   C[i][j] (+ epilogue) = sum_k A(i, k) B(k, j),   A(i, k) = A[i sai + k sak],   B(k, j) = B[k sbk + j sbj]     (fp32 in, fp32 out)
//
// tcgen05.mma kind::tf32 (M = 128, N = 128, K = 8 per instruction, fp32 accumulators in TMEM) with the 3-term split
//   a b  ~=  a_hi b_hi + a_hi b_lo + a_lo b_hi,      x_hi = x with the low 13 mantissa bits cleared (exactly a TF32 number),  x_lo = x - x_hi
// (x_lo is exact in fp32; the tensor core truncates it to its own 10 bits, so the dropped terms are ~2^-21 of the product; a single
// TF32 pass is 2^-10).  Measured on the discriminator's forward pass (scripts/disc_accuracy.py, ndf = 64, K up to 8192) against float64:
// 1.0e-5 relative -- the FFMA kernel: 2.2e-6, torch fp32: 5.5e-7, plain TF32: ~1e-3 -- the remainder being the tensor core's truncating
// fp32 accumulation (see NACC).  The gradient parity of the trainer (1e-3 through a double backward) holds with it.  The three GEMM forms of a convolution layer (forward  col Wp^T, data gradient  dz Wp, weight gradient
// dz^T col) differ only in the strides.
//
// One CTA (256 threads, one per SM) computes a 128 x 128 tile.  Per stage of 32 k: every thread fetches its share of the A and B
// tiles from global memory into registers (16-byte loads where k is the unit-stride direction, else core-matrix-shaped scalar
// loads), splits them and writes hi / lo planes into shared memory in the canonical K-major no-swizzle UMMA layout
// (8 rows x 16 bytes core matrices, K-adjacent ones 128 bytes apart, 8-row groups 1024 bytes apart); after fence.proxy.async and a
// hand-over through an mbarrier a dedicated MMA warp issues the 12 MMAs of the stage (into one of four TMEM accumulators, see NACC) and
// commits them to the stage's "empty" barrier.  Two shared-memory stages, the
// global loads run two stages ahead in registers.  Epilogue: tcgen05.ld (32 lanes x 16 columns per warp), LeakyReLU / mask, global stores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qgb {
namespace tg {

constexpr int BM = 128, BN = 128, BK = 32;
// The tensor core adds into its fp32 accumulator with truncation: the error grows with the number of MMAs chained on one accumulator
// (measured 3.9e-5 relative on the discriminator's forward pass with one accumulator, K up to 8192).  Stages therefore rotate over NACC
// accumulators (all 512 TMEM columns) which the epilogue adds in fp32 registers.
constexpr int NACC = 4;
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB: one plane (hi or lo) of one operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;          // A_hi, A_lo, B_hi, B_lo
constexpr int SMEM_BYTES = 2 * STAGE_BYTES + 1024 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "TG_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra TG_DONE;\n\t"
      "bra TG_WAIT_LOOP;\n\t"
      "TG_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): [0,14) start >> 4, [16,30) leading-dimension byte
// offset >> 4 (between K-adjacent core matrices), [32,46) stride-dimension byte offset >> 4 (between 8-row groups), [46,48) version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((128u >> 4) & 0x3FFF) << 16) | ((uint64_t)((1024u >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}
// instruction descriptor: c_format F32 [4,6) = 1, a / b format TF32 = 2 at [7,10) / [10,13), K-major A and B, n >> 3 at [17,23),
// m >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// byte offset of element (row, k) inside a tile plane
__device__ __forceinline__ uint32_t cm_offset(int row, int k) { return (uint32_t)((row >> 3) * 1024 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4); }
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// One operand tile (128 rows x 32 k) of X(row, k) = X[row srow + k sk]: fetch into 16 registers, later split and store.
struct Loader {
  const float* base; long long srow, sk; int rows, vec;       // rows: valid rows from the tile origin; vec: 16-byte path usable
  __device__ __forceinline__ void fetch(float (&st)[16], int k0, int kend, int tid) const {
    if (vec) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int q = tid + 256 * r;
        const int row = (q >> 6) * 8 + (q & 7), k = (((q >> 3) & 3) + 4 * ((q >> 5) & 1)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows && k0 + k < kend) v = *reinterpret_cast<const float4*>(base + row * srow + (k0 + k));    // (K tail: multiples of 4)
        st[4 * r] = v.x; st[4 * r + 1] = v.y; st[4 * r + 2] = v.z; st[4 * r + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int q = tid + 256 * r;
        const int row = (q >> 8) * 8 + (q & 7), k = ((q >> 5) & 7) * 4 + ((q >> 3) & 3);
        st[r] = (row < rows && k0 + k < kend) ? base[row * srow + (k0 + k) * sk] : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(const float (&st)[16], unsigned char* hi, unsigned char* lo, int tid) const {
    if (vec) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int q = tid + 256 * r;
        const int row = (q >> 6) * 8 + (q & 7), k = (((q >> 3) & 3) + 4 * ((q >> 5) & 1)) * 4;
        float4 h, l;
        h.x = tf32_hi(st[4 * r]); h.y = tf32_hi(st[4 * r + 1]); h.z = tf32_hi(st[4 * r + 2]); h.w = tf32_hi(st[4 * r + 3]);
        l.x = st[4 * r] - h.x; l.y = st[4 * r + 1] - h.y; l.z = st[4 * r + 2] - h.z; l.w = st[4 * r + 3] - h.w;
        const uint32_t o = cm_offset(row, k);
        *reinterpret_cast<float4*>(hi + o) = h;
        *reinterpret_cast<float4*>(lo + o) = l;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int q = tid + 256 * r;
        const int row = (q >> 8) * 8 + (q & 7), k = ((q >> 5) & 7) * 4 + ((q >> 3) & 3);
        const float h = tf32_hi(st[r]);
        const uint32_t o = cm_offset(row, k);
        *reinterpret_cast<float*>(hi + o) = h;
        *reinterpret_cast<float*>(lo + o) = st[r] - h;
      }
    }
  }
};

// EPI 0: none, 1: LeakyReLU(0.2), 2: times the LeakyReLU slope of mask[i ldc + j].  blockIdx.z splits the k range into chunks of ksplit
// (partial results at C + z c_split).
constexpr int kThreads = 288;        // warps 0 - 7: producers (fetch, split, store) and epilogue; warp 8: MMA issue

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1) tgemm_kernel(const float* __restrict__ A, long long sai, long long sak,
                                                            const float* __restrict__ B, long long sbk, long long sbj, float* __restrict__ C,
                                                            long long ldc, int M, int N, int K, int ksplit, long long c_split,
                                                            const float* __restrict__ mask, int a_vec, int b_vec) {
  extern __shared__ unsigned char tg_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)tg_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * STAGE_BYTES);       // [2] planes of a stage written (256 producer arrivals)
  uint64_t* empty = full + 2;                                                 // [2] the MMAs that read a stage have completed (commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * ksplit, kend = min(K, kbeg + ksplit);
  const int nk = (kend - kbeg + BK - 1) / BK;
  if (tid == 0) { mbar_init(&full[0], 256); mbar_init(&full[1], 256); mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); fence_barrier_init(); }
  if (warp == 8) tmem_alloc(tmem_slot, NACC * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ---- MMA warp: one lane waits for a stage, issues its 12 MMAs and commits them to the stage's "empty" barrier ----
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it & 1;
        mbar_wait(&full[s], (uint32_t)((it / 2) & 1));
        tc_fence_after();
        const uint32_t base = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {
          const uint64_t ah = make_desc(base + j * 256), al = make_desc(base + TILE_BYTES + j * 256);
          const uint64_t bh = make_desc(base + 2 * TILE_BYTES + j * 256), bl = make_desc(base + 3 * TILE_BYTES + j * 256);
          const uint32_t acc = tmem + (uint32_t)((it % NACC) * BN);
          mma_tf32(acc, al, bh, kIdesc, (it >= NACC || j > 0) ? 1u : 0u);          // small terms first
          mma_tf32(acc, ah, bl, kIdesc, 1u);
          mma_tf32(acc, ah, bh, kIdesc, 1u);
        }
        tc_commit(&empty[s]);
      }
    }
  } else {
    // ---- producers: global loads two stages ahead in registers, split, planes of stage s, hand-over through the "full" barrier ----
    Loader la{A + (long long)i0 * sai, sai, sak, M - i0, a_vec};
    Loader lb{B + (long long)j0 * sbj, sbj, sbk, N - j0, b_vec};
    float sa0[16], sb0[16], sa1[16], sb1[16];
    la.fetch(sa0, kbeg, kend, tid);
    lb.fetch(sb0, kbeg, kend, tid);
    if (nk > 1) { la.fetch(sa1, kbeg + BK, kend, tid); lb.fetch(sb1, kbeg + BK, kend, tid); }
    auto stage = [&](int it, float (&sa)[16], float (&sb)[16]) {
      const int s = it & 1;
      unsigned char* st = smem + s * STAGE_BYTES;
      if (it >= 2) mbar_wait(&empty[s], (uint32_t)((it / 2 - 1) & 1));          // the MMAs that read this stage have completed
      la.store(sa, st, st + TILE_BYTES, tid);
      lb.store(sb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, tid);
      fence_proxy_async_smem();
      mbar_arrive(&full[s]);
      if (it + 2 < nk) {
        la.fetch(sa, kbeg + (it + 2) * BK, kend, tid);
        lb.fetch(sb, kbeg + (it + 2) * BK, kend, tid);
      }
    };
    for (int it = 0; it < nk; it += 2) {
      stage(it, sa0, sb0);
      if (it + 1 < nk) stage(it + 1, sa1, sb1);
    }
    if (nk > 0) mbar_wait(&empty[(nk - 1) & 1], (uint32_t)(((nk - 1) / 2) & 1));   // commits complete in order: the last one covers all
  }
  tc_fence_after();
  // epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (accumulator rows) and the column half w / 4
  float* Cz = C + (long long)blockIdx.z * c_split;
  const int gi = i0 + (warp & 3) * 32 + lane;
#pragma unroll 1
  for (int cb = 0; cb < (warp < 8 ? 4 : 0); ++cb) {
    const int col = (warp >> 2) * 64 + cb * 16;
    float sum[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) sum[e] = 0.f;
    for (int a = 0; a < NACC && a < nk; ++a) {                     // (uniform trip count: .sync.aligned loads)
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(a * BN + col), r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) sum[e] += __uint_as_float(r[e]);
    }
    if (gi < M) {
      if ((ldc & 3) == 0 && j0 + col + 16 <= N) {                   // 16-byte stores (and mask loads): 64 contiguous bytes per lane
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          float v[4] = {sum[e], sum[e + 1], sum[e + 2], sum[e + 3]};
          const long long o = gi * ldc + j0 + col + e;
          if (EPI == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = v[u] > 0.f ? v[u] : 0.2f * v[u];
          }
          if (EPI == 2) {
            const float4 m = *reinterpret_cast<const float4*>(mask + o);
            v[0] *= m.x > 0.f ? 1.f : 0.2f; v[1] *= m.y > 0.f ? 1.f : 0.2f; v[2] *= m.z > 0.f ? 1.f : 0.2f; v[3] *= m.w > 0.f ? 1.f : 0.2f;
          }
          *reinterpret_cast<float4*>(Cz + o) = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int gj = j0 + col + e;
          if (gj < N) {
            float v = sum[e];
            if (EPI == 1) v = v > 0.f ? v : 0.2f * v;
            if (EPI == 2) v *= mask[gi * ldc + gj] > 0.f ? 1.f : 0.2f;
            Cz[gi * ldc + gj] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, NACC * BN);
}

}  // namespace tg
}  // namespace qgb
