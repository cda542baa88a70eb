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
//     the global arrays are stored tile by tile as the SWIZZLE_128B K-major shared-memory image (tiled_layout.cuh) and a
//     stage is four contiguous 16 KB TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 3-stage ring with
//     full / empty mbarriers (tcgen05.commit releases a stage);
//   * a dedicated warp issues 12 MMAs (4 k-steps x 3 products) of 128 x 128 x 8 per stage into a 128-column fp32 TMEM tile;
//   * the tensor core's accumulation truncates (measured: a 1e-5 bias over K = 928), so two TMEM accumulators alternate
//     between groups of 4 k-blocks and are promoted into fp32 registers (tcgen05.ld + round-to-nearest add);
//   * epilogue (thread = row): + bias -> ELU -> split -> hi / lo of the next layer's input; the last hidden layer (128
//     wide = one tile) also applies the 128 -> 12 output layer in registers (fp32 FMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fim_tc.cuh"   // smem_u32, mbarrier / tcgen05 helpers, tf32_round
#include "tiled_layout.cuh"

namespace mlptc {

constexpr int kTile = 128;                   // UMMA M = N
constexpr int kEpilogue = 128;               // warps 0-3: accumulator promotion + epilogue (thread = row)
constexpr int kThreads = kEpilogue + 64;     // warp 4: MMA issuer, warp 5: bulk-copy producer (one elected lane each)
constexpr int kAccCols = 128;                // one fp32 accumulator tile
constexpr int kTmemCols = 2 * kAccCols;      // two accumulators, used alternately by groups of k-blocks
constexpr int kMaxOut = 16;
constexpr int kKAlign = 32;                  // the padded K of every operand is a multiple of this
constexpr int kBK = 32;                      // fp32 elements of K per stage = one 128-byte swizzle atom per row
constexpr int kStages = 3;
constexpr int kOperandBytes = kTile * kBK * 4;               // 16 KB = one tile of tiled_layout.cuh
constexpr int kStageBytes = 4 * kOperandBytes;               // A_hi, A_lo, W_hi, W_lo
constexpr int kSmemBytes = kStages * kStageBytes + 256;
// k-blocks per accumulator group: the tensor core's accumulation truncates (a ~3e-8 relative bias per MMA), so an
// accumulator only ever holds 48 MMAs before it is added to the fp32 running sums in registers
constexpr int kGroup = 4;

// K-major SWIZZLE_128B (cute::UMMA::LayoutType::SWIZZLE_128B = 2): rows are 128 bytes (32 fp32 of K), an 8-row group is
// 1024 contiguous bytes, the 16-byte chunk c of row r sits at chunk position c ^ (r & 7); SBO = 1024 between row groups,
// the leading-dimension field is 1 (unused: one swizzle atom spans the tile's K extent); a k-step of 8 fp32 advances the
// start address by 32 bytes inside the atom.  Tile bases are 1024-byte aligned (base_offset 0).  Measured: the no-swizzle
// canonical layout runs the same 128 x 128 x 8 MMA in ~256 cycles instead of ~65.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

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
  const float* a_hi; const float* a_lo;    // [Mp, Kp] activations, split, tiled_layout.cuh
  const float* w_hi; const float* w_lo;    // [N, Kp] weights (torch Linear layout [out, in]), split, tiled_layout.cuh
  const float* bias;                       // [N]
  int Kp, N;
  float* out_hi; float* out_lo; int out_stride;   // MODE 0: ELU(a w^T + b) split, tiled [Mp, out_stride]
  const float* w_out; const float* b_out; int n_out; float* out; int M;   // MODE 1: + output layer -> out [M, n_out]
};

// One 128 x 128 tile of  A W^T  per CTA;  grid = (Mp / 128, N / 128).  Warp-specialised, no CTA-wide barrier in the loop:
//   producer (1 lane)     wait empty[s] -> expect_tx(64 KB) -> 4 x cp.async.bulk of one contiguous 16 KB tile each -> full[s]
//   MMA issuer (1 lane)   wait full[s] -> 12 x tcgen05.mma (4 k-steps x {lo.hi, hi.lo, hi.hi}) -> tcgen05.commit -> empty[s]
//                         (+ accfull at the end of a group of 4 blocks; waits accempty before it reuses an accumulator)
//   promotion (128 thr)   wait accfull -> tcgen05.ld -> fp32 add into registers -> arrive accempty;  then the epilogue
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) mlp_layer_kernel(const LayerArgs L) {
  using namespace fimtc;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);   // [kStages] producer -> MMA
  uint64_t* empty = full + kStages;                                             // [kStages] MMA -> producer
  uint64_t* accfull = empty + kStages;                                          // [2] MMA -> promotion
  uint64_t* accempty = accfull + 2;                                             // [2] promotion -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTile, n0 = blockIdx.y * kTile;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kStages; s++) { mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(smem_u32(accfull + a), 1); mbar_init(smem_u32(accempty + a), kEpilogue); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  const int n_blocks = L.Kp / kBK;
  const int n_groups = (n_blocks + kGroup - 1) / kGroup;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == kEpilogue / 32 + 1) {
    // ===== bulk-copy producer =====
    if (lane == 0) {
      const size_t a_tile = (size_t)blockIdx.x * n_blocks * tiled::kTileFloats, w_tile = (size_t)blockIdx.y * n_blocks * tiled::kTileFloats;
      const float* src[4] = {L.a_hi + a_tile, L.a_lo + a_tile, L.w_hi + w_tile, L.w_lo + w_tile};
      for (int kb = 0; kb < n_blocks; kb++) {
        const int s = kb % kStages;
        if (kb >= kStages) mbar_wait(smem_u32(empty + s), (uint32_t)((kb / kStages - 1) & 1));
        const uint32_t bar = smem_u32(full + s), stage = smem_base + (uint32_t)s * kStageBytes;
        mbar_expect_tx(bar, kStageBytes);
#pragma unroll
        for (int op = 0; op < 4; op++)
          bulk_g2s(stage + (uint32_t)op * kOperandBytes, src[op] + (size_t)kb * tiled::kTileFloats, kOperandBytes, bar);
      }
    }
  } else if (warp == kEpilogue / 32) {
    // ===== MMA issuer =====
    if (lane == 0) {
      for (int kb = 0; kb < n_blocks; kb++) {
        const int s = kb % kStages, g = kb / kGroup;
        const bool group_start = (kb % kGroup) == 0, group_end = ((kb + 1) % kGroup) == 0 || kb == n_blocks - 1;
        if (group_start && g >= 2) mbar_wait(smem_u32(accempty + (g & 1)), (uint32_t)(((g - 2) >> 1) & 1));
        mbar_wait(smem_u32(full + s), (uint32_t)((kb / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage = smem_base + (uint32_t)s * kStageBytes;
        const uint64_t ah = make_desc_sw128(stage), al = make_desc_sw128(stage + kOperandBytes),
                       wh = make_desc_sw128(stage + 2 * kOperandBytes), wl = make_desc_sw128(stage + 3 * kOperandBytes);
        const uint32_t tacc = tmem_d + (uint32_t)((g & 1) * kAccCols);
#pragma unroll
        for (int k = 0; k < kBK / 8; k++) {
          const uint64_t off = (uint64_t)(k * 2);     // 32 bytes >> 4, inside the 14-bit start-address field
          mma_tf32(tacc, al + off, wh + off, (!group_start || k > 0) ? 1u : 0u);
          mma_tf32(tacc, ah + off, wl + off, 1u);
          mma_tf32(tacc, ah + off, wh + off, 1u);
        }
        umma_commit(smem_u32(empty + s));
        if (group_end) umma_commit(smem_u32(accfull + (g & 1)));
      }
    }
  } else {
    // ===== promotion + epilogue (thread = row) =====
    float acc[kAccCols];
#pragma unroll
    for (int i = 0; i < kAccCols; i++) acc[i] = 0.f;
    for (int g = 0; g < n_groups; g++) {   // acc += accumulator of group g, then hand the accumulator back to the MMA warp
      mbar_wait(smem_u32(accfull + (g & 1)), (uint32_t)((g >> 1) & 1));
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
      mbar_arrive(smem_u32(accempty + (g & 1)));
    }
    const int row = m0 + tid;
    if (MODE == 0) {
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
        const size_t o = tiled::offset(row, n0 + i, L.out_stride);      // 4 consecutive k = one 16-byte chunk
        *reinterpret_cast<float4*>(L.out_hi + o) = vh;
        *reinterpret_cast<float4*>(L.out_lo + o) = vl;
      }
    } else {
      // all MMAs have retired (the last accfull arrived), so the stages are free: stage the output layer's weights there
      float* w_s = reinterpret_cast<float*>(smem);      // [n_out][128]
      for (int i = tid; i < L.n_out * kTile; i += kEpilogue) w_s[i] = L.w_out[i];
      asm volatile("bar.sync 1, %0;" ::"n"(kEpilogue) : "memory");
#pragma unroll
      for (int i = 0; i < kAccCols; i++) acc[i] = elu(acc[i] + __ldg(L.bias + n0 + i));
      for (int j = 0; j < L.n_out; j++) {
        float y = L.b_out[j];
#pragma unroll
        for (int i = 0; i < kAccCols; i++) y = fmaf(w_s[j * kTile + i], acc[i], y);
        if (row < L.M) L.out[(size_t)row * L.n_out + j] = y;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
}

// x [M, K] fp32 row-major -> hi / lo in the tiled layout [Mp, Kp] (padding is left as it is: zero-initialised by the owner)
__global__ void split_kernel(const float* x, int M, int K, float* hi, float* lo, int Kp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * K) return;
  const int m = (int)(idx / K), k = (int)(idx - (size_t)m * K);
  const float v = x[idx];
  const float h = fimtc::tf32_round(v);
  const size_t o = tiled::offset(m, k, Kp);
  hi[o] = h;
  lo[o] = v - h;
}

// hi + lo in the tiled layout -> x [M, K] row-major (inspection / tests)
__global__ void unsplit_kernel(const float* hi, const float* lo, int M, int K, int Kp, float* x) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * K) return;
  const int m = (int)(idx / K), k = (int)(idx - (size_t)m * K);
  const size_t o = tiled::offset(m, k, Kp);
  x[idx] = hi[o] + lo[o];
}

}  // namespace mlptc
