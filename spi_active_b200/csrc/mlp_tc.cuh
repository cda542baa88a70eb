// mlp_tc.cuh — the locomotion policy of the active-exploration rollout on the 5th-generation tensor cores, fp32-accurate.
//
// Reference: the actor is nn.Sequential(Linear(900,512), ELU, Linear(512,256), ELU, Linear(256,128), ELU, Linear(128,12))
// (spigym/agents/modules/modules.py:47-63, config/algo/ppo.yaml:32-40), evaluated in fp32 by torch once per control
// step for every env (agents/sysid/active_sysid.py:555-562): 1.25 MFLOP per env-step, 14 GFLOP per step at 1024 x 11
// envs — after the physics and the bookkeeping were fused this GEMM chain is ~3/4 of a step on cuBLAS' fp32 SIMT path.
//
// Here every Linear is a tcgen05.mma.kind::tf32 GEMM with the 3xTF32 split (x = hi + lo, both tf32:
// x.w ~= hi.hi + hi.lo + lo.hi, error ~2^-22 relative, i.e. fp32 grade — plain TF32 would be 2^-11):
//   * operands are PRE-SPLIT in global memory, K-major: the weights once at creation, the activations by the epilogue
//     of the layer (or of the observation kernel) that produces them, so the GEMM's producer is a pure copy:
//     cp.async 16-byte chunks global -> shared memory straight into the canonical no-swizzle K-major UMMA layout
//     (8-row x 16-byte core matrices), 3-stage ring, one mbarrier per stage armed by tcgen05.commit;
//   * one thread issues 12 MMAs (4 k-steps x 3 products) of 128 x 128 x 8 per stage into a 128-column fp32 TMEM tile;
//   * epilogue (thread = row): tcgen05.ld -> + bias -> ELU -> split -> hi / lo of the next layer's input; the last
//     hidden layer (128 wide = one tile) also applies the 128 -> 12 output layer in registers (fp32 FMA).
// Signed dot products make the tensor core's truncated accumulation a random walk (~1e-6 relative over K = 928), unlike
// the sum of squares of fim_tc.cuh, so the accumulator stays in TMEM for the whole K loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fim_tc.cuh"   // smem_u32, mbarrier / tcgen05 helpers, tf32_round

namespace mlptc {

constexpr int kTile = 128;                   // UMMA M = N
constexpr int kThreads = 128;
constexpr int kAccCols = 128;                // one fp32 accumulator tile
constexpr int kTmemCols = 2 * kAccCols;      // two accumulators, used alternately by groups of k-blocks
constexpr int kMaxOut = 16;
constexpr int kKAlign = 32;                  // the padded K of every operand is a multiple of this

// BK fp32 elements of K per stage; the canonical no-swizzle K-major tile: 8-row group g at g * SBO, 16-byte K chunk c at
// c * LBO, row r8 of the group at r8 * 16
template <int BK, int STAGES> struct Cfg {
  static constexpr int kChunks = BK / 4;
  static constexpr int kLBO = 128;
  static constexpr int kSBO = kChunks * 128;
  static constexpr int kOperandBytes = kTile * BK * 4;
  static constexpr int kStageBytes = 4 * kOperandBytes;        // A_hi, A_lo, W_hi, W_lo
  static constexpr int kSmemBytes = STAGES * kStageBytes + 256;
  // k-blocks per accumulator group: the tensor core's accumulation truncates (a ~3e-8 relative bias per MMA), so an
  // accumulator only ever holds ~48 MMAs before it is added to the fp32 running sums in registers
  static constexpr int kGroup = 128 / BK;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int lbo, int sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// K-major SWIZZLE_128B (cute::UMMA::LayoutType::SWIZZLE_128B = 2): rows are 128 bytes (32 fp32 of K), an 8-row group is
// 1024 contiguous bytes, the 16-byte chunk c of row r sits at chunk position c ^ (r & 7); SBO = 1024 between row groups,
// the leading-dimension field is 1 (unused: one swizzle atom spans the tile's K extent); a k-step of 8 fp32 advances the
// start address by 32 bytes inside the atom.  Tile bases are 1024-byte aligned (base_offset 0).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : expm1f(x); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

struct LayerArgs {
  const float* a_hi; const float* a_lo;    // [Mp, Kp] activations, split
  const float* w_hi; const float* w_lo;    // [N, Kp] weights (torch Linear layout), split
  const float* bias;                       // [N]
  int Kp, N;
  float* out_hi; float* out_lo; int out_stride;   // MODE 0: ELU(a w^T + b) split, [Mp, out_stride]
  const float* w_out; const float* b_out; int n_out; float* out; int M;   // MODE 1: + output layer -> out [M, n_out]
};

// one 128 x 128 tile of  A W^T  per CTA;  grid = (Mp / 128, N / 128)
template <int MODE, int BK, int STAGES>
__global__ void __launch_bounds__(kThreads) mlp_layer_kernel(const LayerArgs L) {
  using namespace fimtc;
  using C = Cfg<BK, STAGES>;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * C::kStageBytes);   // [STAGES] stage free, [2] accumulator full
  uint64_t* accbars = bars + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accbars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTile, n0 = blockIdx.y * kTile;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < STAGES + 2; s++) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // producer mapping: one warp instruction copies 8 rows x 64 contiguous bytes (full sectors; the 8 rows of a core matrix
  // land in 8 different bank groups)
  const int r8 = lane & 7, c4 = lane >> 3;
  const float* src[4] = {L.a_hi + (size_t)m0 * L.Kp, L.a_lo + (size_t)m0 * L.Kp, L.w_hi + (size_t)n0 * L.Kp,
                         L.w_lo + (size_t)n0 * L.Kp};
  auto load_stage = [&](int kb, int s) {
    const uint32_t stage = smem_u32(smem) + (uint32_t)s * C::kStageBytes;
    const int k0 = kb * BK;
    constexpr int kHalves = C::kChunks / 4;                  // 64-byte pieces per row
    constexpr int kIters = 16 * kHalves / 4;                 // (row group, piece) pairs per warp
#pragma unroll
    for (int op = 0; op < 4; op++) {
#pragma unroll
      for (int it = 0; it < kIters; it++) {
        const int pair = it * 4 + warp;
        const int g = pair / kHalves, half = pair % kHalves;
        const int row = g * 8 + r8, chunk = half * 4 + c4;
        const uint32_t dst = (BK == 32)
            ? stage + (uint32_t)op * C::kOperandBytes + (uint32_t)g * 1024 + (uint32_t)r8 * 128 + (uint32_t)((chunk ^ r8) * 16)
            : stage + (uint32_t)op * C::kOperandBytes + (uint32_t)g * C::kSBO + (uint32_t)chunk * C::kLBO + (uint32_t)r8 * 16;
        cp_async16(dst, src[op] + (size_t)row * L.Kp + k0 + chunk * 4);
      }
    }
  };

  // fp32 running sums of this thread's row (the promoted accumulator)
  float acc[kAccCols];
#pragma unroll
  for (int i = 0; i < kAccCols; i++) acc[i] = 0.f;
  auto flush = [&](int g) {   // acc += accumulator of k-block group g (complete once its commit has arrived)
    mbar_wait(smem_u32(accbars + (g & 1)), (uint32_t)((g >> 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int cb = 0; cb < kAccCols / 32; cb++) {
      uint32_t r[32];
      tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)((g & 1) * kAccCols + cb * 32), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 32; i++) acc[cb * 32 + i] += __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  const int n_blocks = L.Kp / BK;
  const int n_groups = (n_blocks + C::kGroup - 1) / C::kGroup;
  for (int kb = 0; kb < STAGES - 1; kb++) {
    if (kb < n_blocks) load_stage(kb, kb);
    cp_async_commit();
  }
  for (int kb = 0; kb < n_blocks; kb++) {
    const int s = kb % STAGES, g = kb / C::kGroup;
    const bool group_start = (kb % C::kGroup) == 0, group_end = ((kb + 1) % C::kGroup) == 0 || kb == n_blocks - 1;
    if (group_start && g >= 2) flush(g - 2);                       // its accumulator is about to be overwritten
    cp_async_wait<STAGES - 2>();                                    // this thread's chunks of block kb have landed
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // ... and are visible to the UMMA reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t stage = smem_u32(smem) + (uint32_t)s * C::kStageBytes;
      const uint32_t ah = stage, al = stage + C::kOperandBytes, wh = stage + 2 * C::kOperandBytes, wl = stage + 3 * C::kOperandBytes;
      const uint32_t tacc = tmem_d + (uint32_t)((g & 1) * kAccCols);
#pragma unroll
      for (int k = 0; k < BK / 8; k++) {
        const uint32_t acc_on = (!group_start || k > 0) ? 1u : 0u;
        if (BK == 32) {
          const uint32_t off = (uint32_t)k * 32;
          mma_tf32(tacc, make_desc_sw128(al + off), make_desc_sw128(wh + off), acc_on);
          mma_tf32(tacc, make_desc_sw128(ah + off), make_desc_sw128(wl + off), 1u);
          mma_tf32(tacc, make_desc_sw128(ah + off), make_desc_sw128(wh + off), 1u);
        } else {
          const uint32_t off = (uint32_t)k * 2 * C::kLBO;
          mma_tf32(tacc, make_desc(al + off, C::kLBO, C::kSBO), make_desc(wh + off, C::kLBO, C::kSBO), acc_on);
          mma_tf32(tacc, make_desc(ah + off, C::kLBO, C::kSBO), make_desc(wl + off, C::kLBO, C::kSBO), 1u);
          mma_tf32(tacc, make_desc(ah + off, C::kLBO, C::kSBO), make_desc(wh + off, C::kLBO, C::kSBO), 1u);
        }
      }
      umma_commit(smem_u32(bars + s));
      if (group_end) umma_commit(smem_u32(accbars + (g & 1)));
    }
    // refill the stage block kb-1 used with block kb + STAGES - 1 — AFTER block kb's MMAs were queued, so the tensor pipe
    // keeps running while this waits for block kb-1 to retire
    {
      const int nb = kb + STAGES - 1;
      if (nb < n_blocks) {
        if (kb >= 1) mbar_wait(smem_u32(bars + (nb % STAGES)), (uint32_t)(((kb - 1) / STAGES) & 1));
        load_stage(nb, nb % STAGES);
      }
      cp_async_commit();
    }
  }
  for (int g = (n_groups > 2 ? n_groups - 2 : 0); g < n_groups; g++) flush(g);

  // ---- epilogue: thread = row -----------------------------------------------------------------------------------------
  const int row = m0 + tid;
  if (MODE == 0) {
    float* oh = L.out_hi + (size_t)row * L.out_stride + n0;
    float* ol = L.out_lo + (size_t)row * L.out_stride + n0;
#pragma unroll
    for (int i = 0; i < kAccCols; i += 4) {
      float h[4];
#pragma unroll
      for (int j = 0; j < 4; j++) h[j] = elu(acc[i + j] + __ldg(L.bias + n0 + i + j));
      float4 vh, vl;
      vh.x = tf32_round(h[0]); vl.x = h[0] - vh.x;
      vh.y = tf32_round(h[1]); vl.y = h[1] - vh.y;
      vh.z = tf32_round(h[2]); vl.z = h[2] - vh.z;
      vh.w = tf32_round(h[3]); vl.w = h[3] - vh.w;
      *reinterpret_cast<float4*>(oh + i) = vh;
      *reinterpret_cast<float4*>(ol + i) = vl;
    }
  } else {
    float* w_s = reinterpret_cast<float*>(smem);      // the output layer's weights [n_out][128] (the stages are idle now)
    for (int i = tid; i < L.n_out * kTile; i += kThreads) w_s[i] = L.w_out[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kAccCols; i++) acc[i] = elu(acc[i] + __ldg(L.bias + n0 + i));
    for (int j = 0; j < L.n_out; j++) {
      float y = L.b_out[j];
#pragma unroll
      for (int i = 0; i < kAccCols; i++) y = fmaf(w_s[j * kTile + i], acc[i], y);
      if (row < L.M) L.out[(size_t)row * L.n_out + j] = y;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
}

// x [M, K] fp32 -> hi / lo [Mp, Kp] (rows >= M and columns >= K are left as they are: zero-initialised by the owner)
__global__ void split_kernel(const float* x, int M, int K, float* hi, float* lo, int Kp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * K) return;
  const int m = (int)(idx / K), k = (int)(idx - (size_t)m * K);
  const float v = x[idx];
  const float h = fimtc::tf32_round(v);
  hi[(size_t)m * Kp + k] = h;
  lo[(size_t)m * Kp + k] = v - h;
}

}  // namespace mlptc
