// mlp_tc.cuh — the locomotion policy of the active-exploration rollout on the 5th-generation tensor cores, fp32-accurate.
//
// Reference: the actor is nn.Sequential(Linear(900,512), ELU, Linear(512,256), ELU, Linear(256,128), ELU, Linear(128,12))
// (spigym/agents/modules/modules.py:47-63, config/algo/ppo.yaml:32-40), evaluated in fp32 by torch once per control
// step for every env (agents/sysid/active_sysid.py:555-562): 1.25 MFLOP per env-step, 14 GFLOP per step at 1024 x 11
// envs — after the physics and the bookkeeping were fused this GEMM chain is ~3/4 of a step on cuBLAS' fp32 SIMT path.
//
// Here every Linear is a tcgen05.mma.kind::f16 GEMM on fp16 PAIRS — the "3 x FP16" split: x * s = hi + lo with hi = fp16(x s),
// lo = fp16(x s - hi) (s a power of two that keeps lo out of the fp16 subnormals; undone exactly in the epilogue), and
// x.w ~= hi.hi + hi.lo + lo.hi with exact fp16 x fp16 products accumulated in fp32: error ~2^-22 relative, i.e. fp32 grade
// (plain fp16 / TF32 would be 2^-11).  Against the 3 x TF32 split of r1_c this moves half the operand bytes (2 x 2 B instead of
// 2 x 4 B per value) and runs on the twice-as-fast f16 MMA — and L2 -> SM operand bytes and shared-memory operand reads are
// exactly what bounded that kernel (profiles/README.md):
//   * operands are PRE-SPLIT in global memory, K-major: the weights once at creation, the activations by the epilogue
//     of the layer (or of the observation kernel) that produces them, so the GEMM's producer is a pure copy:
//     the global arrays are stored tile by tile as the SWIZZLE_128B K-major shared-memory image (tiled_layout.cuh) and a
//     stage is four contiguous 16 KB TMA bulk copies (cp.async.bulk + mbarrier complete_tx), 3-stage ring with
//     full / empty mbarriers (tcgen05.commit releases a stage);
//   * a dedicated warp issues 8 MMAs per stage (4 k-steps of 16 x {a_hi [w_hi; w_lo] at N = 256, a_lo w_hi at N = 128}) into a
//     256-column fp32 TMEM tile whose two halves are added during the promotion;
//   * the tensor core's accumulation truncates (measured: a 1e-5 bias over K = 928), so two TMEM accumulators alternate
//     between groups of k-blocks and are promoted into fp32 registers (tcgen05.ld + round-to-nearest add);
//   * epilogue (thread = row x column half): x 1 / (s_a s_w) -> + bias -> ELU -> split -> hi / lo of the next layer's input,
//     assembled in shared memory and stored with bulk copies; the last hidden layer (128 wide = one tile) also applies the
//     128 -> 12 output layer in registers (fp32 FMA).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fim_tc.cuh"   // smem_u32, mbarrier / tcgen05 helpers
#include "tiled_layout.cuh"
#include "pdl.cuh"

namespace mlptc {

constexpr int kTile = 128;                   // UMMA M = N
constexpr int kEpilogue = 256;               // warps 0-7: accumulator promotion + epilogue; warp w owns rows 32 (w & 3) .. + 31 (its
                                             // TMEM lane quadrant) and columns 64 (w >> 2) .. + 63
constexpr int kHalfCols = 64;
constexpr int kThreads = kEpilogue + 64;     // warp 8: MMA issuer, warp 9: bulk-copy producer (one elected lane each)
constexpr int kAccCols = 256;                // one accumulator: columns [0,128) = a_hi w_hi + a_lo w_hi, [128,256) = a_hi w_lo
constexpr int kTmemCols = 2 * kAccCols;      // two accumulators, used alternately by groups of k-blocks (all of TMEM)
constexpr int kKAlign = 64;                  // the padded K of every operand is a multiple of this
constexpr int kBK = 64;                      // halves of K per stage = one 128-byte swizzle atom per row
constexpr float kActScale = 16.0f, kWeightScale = 256.0f;   // powers of two applied before the fp16 split (exact to undo)
constexpr int kStages = 3;
constexpr int kOperandBytes = kTile * kBK * 2;               // 16 KB = one tile of tiled_layout.cuh
constexpr int kStageBytes = 4 * kOperandBytes;               // A_hi, A_lo, W_hi, W_lo
constexpr int kMaxOut = 16;
// shared-memory tail behind the stages: barriers + TMEM slot (256 B) | bias of this n-tile (<= 256 floats) | output layer:
// weights [kMaxOut][128] (8 KB) + bias (64 B)
constexpr int kTailBias = 256, kTailWout = kTailBias + 1024, kTailBout = kTailWout + kMaxOut * kTile * 4;
constexpr int kSmemBytes = kStages * kStageBytes + kTailBout + 64;
// k-blocks per accumulator group: the tensor core's accumulation truncates (a ~3e-8 relative bias per MMA), so an
// accumulator only ever holds 32 MMAs (K = 256) before it is added to the fp32 running sums in registers
constexpr int kGroup = 4;

// K-major SWIZZLE_128B (cute::UMMA::LayoutType::SWIZZLE_128B = 2): rows are 128 bytes (64 halves of K), an 8-row group is
// 1024 contiguous bytes, the 16-byte chunk c of row r sits at chunk position c ^ (r & 7); SBO = 1024 between row groups,
// the leading-dimension field is 1 (unused: one swizzle atom spans the tile's K extent); a k-step of 16 halves advances the
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

// instruction descriptors (cute::UMMA::InstrDescriptor, see fim_tc.cuh): c_format F32 (1) at [4,6), a / b_format F16 (0) at
// [7,10) / [10,13), K-major operands, N >> 3 at [17,23), M >> 4 at [24,29); M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24); }
__device__ __forceinline__ void mma_f16_n(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
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

// ---- CTA pairs (cta_group::2): helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t v; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v)); return v; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// one MMA over both SMs of a pair: M = 256 (128 rows per CTA), B rows split between the two CTAs' shared memories
__device__ __forceinline__ void mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// tcgen05.commit of a pair's MMAs, the arrive delivered to the same mbarrier of both CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16_pair(int n) {   // M = 256 over the pair
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// TMA 1-D bulk copy shared -> global (bulk-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpilogue) : "memory"); }

// ELU without branches (the library expm1f compiles to divergent regions that serialise the 64 independent elements of an
// epilogue thread: measured 6.7 us of a 9 us epilogue).  x <= 0:  expm1(x) = Taylor degree 8 on (-0.35, 0] (truncation
// 2e-10), 2^(x log2 e) - 1 below (MUFU.EX2: 2 ulp of a value <= 0.7, i.e. <= 1.2e-7 absolute on a result of magnitude >= 0.3)
__device__ __forceinline__ float elu(float x) {
  float p = 2.4801587e-5f;                 // 1/8!
  p = fmaf(p, x, 1.9841270e-4f);           // 1/7!
  p = fmaf(p, x, 1.3888889e-3f);
  p = fmaf(p, x, 8.3333333e-3f);
  p = fmaf(p, x, 4.1666667e-2f);
  p = fmaf(p, x, 1.6666667e-1f);
  p = fmaf(p, x, 0.5f);
  p = fmaf(p * x, x, x);                   // x + x^2 (1/2 + x (1/6 + ...))
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  const float neg = (x > -0.35f) ? p : e - 1.0f;
  return (x > 0.f) ? x : neg;
}

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

// Development instrumentation, compiled in only with -DSPI_B200_MLP_DEV (tools/ and profiles/README.md r1_d): in-kernel
// timestamps of CTA (0, 0) (globaltimer ns: 0 start, 1 set-up done, 2 accumulators promoted = main loop done, 5 epilogue
// computed, 3 epilogue done, 4 exit) and the isolation switches L.dbg (bit 0: issue no MMAs, bit 1: copy no operands,
// bit 2: no accumulator promotion — results are garbage, only the timing means something).
#if defined(SPI_B200_MLP_DEV)
__device__ unsigned long long g_stamps[3][8];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MLP_STAMP(i) do { if (L.stamp >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) g_stamps[L.stamp][i] = gtimer(); } while (0)
#define MLP_DBG(bit) (L.dbg & (bit))
#else
#define MLP_STAMP(i) do { } while (0)
#define MLP_DBG(bit) 0
#endif

struct LayerArgs {
  const __half* a_hi; const __half* a_lo;  // [Mp, Kp] activations x kActScale, split, tiled_layout.cuh
  const __half* w_hi; const __half* w_lo;  // [N, Kp] weights (torch Linear layout [out, in]) x kWeightScale, split, tiled
  const float* bias;                       // [N]
  int Kp, N;
  __half* out_hi; __half* out_lo; int out_stride;   // MODE 0: ELU(a w^T + b) x kActScale split, tiled [Mp, out_stride]
  const float* w_out; const float* b_out; int n_out; float* out; int M;   // MODE 1: + output layer -> out [M, n_out]
  const int* rot; size_t rot_stride; int n_rot;   // device int selecting one of n_rot weight copies rot_stride elements apart, or null
  int stamp, dbg;   // development instrumentation (SPI_B200_MLP_DEV builds only)
};

// One 128 x BN tile of  A W^T  per CTA (BN = 256 for the wide layers, 128 otherwise);  grid = (Mp / 128, N / BN).
// Warp-specialised, no CTA-wide barrier in the loop:
//   producer (1 lane)     wait empty[s] -> expect_tx -> cp.async.bulk of contiguous 16 KB tiles (a_hi, a_lo, BN / 128 x w_hi,
//                         BN / 128 x w_lo) -> full[s]
//   MMA issuer (1 lane)   wait full[s] -> per k-step of 8:  BN = 256: a_hi w_hi, a_hi w_lo, a_lo w_hi as three N = 256 MMAs;
//                         BN = 128: a_hi [w_hi; w_lo] as ONE N = 256 MMA + a_lo w_hi at N = 128 (the halves are added in
//                         the promotion) -> tcgen05.commit -> empty[s]  (+ accfull at the end of a group of k-blocks;
//                         waits accempty before it reuses an accumulator)
//   promotion (256 thr)   wait accfull -> tcgen05.ld -> fp32 add into registers -> arrive accempty;  then the epilogue
// What bounds the loop (measured, profiles/README.md): L2 -> SM operand bytes at the chip-wide L2 throughput when every SM
// streams, then the shared-memory reads of the SS-mode MMAs; BN = 256 moves 25 % fewer bytes per flop than BN = 128.
// PAIR: two CTAs of a cluster (cta_group::2) compute a 256 x 256 tile: each CTA owns 128 rows of A and of the output,
// holds HALF of the 256 weight rows (so a stage is the same 64 KB as for BN = 128: three stages fit) and the leader's MMA
// lane issues M = 256 MMAs that read both shared memories — half the L2 -> SM operand bytes per flop of the 128 x 128 tiles.
template <int BN, bool PAIR = false> struct Cfg {
  static constexpr int kWTiles = PAIR ? 1 : BN / kTile;                       // 16 KB tiles per W operand and k-block IN THIS CTA
  static constexpr int kStageBytes = (2 + 2 * kWTiles) * kOperandBytes;       // a_hi, a_lo, w_hi[..], w_lo[..]
  static constexpr int kStages = (mlptc::kStages * mlptc::kStageBytes) / ((2 + 2 * kWTiles) * kOperandBytes);   // 192 KB of stages either way
  static constexpr int kCols = BN / 2;                                        // output columns per epilogue thread
};
static_assert(Cfg<128>::kStages == 3 && Cfg<256>::kStages == 2 && Cfg<256, true>::kStages == 3, "stage budget");

template <int MODE, int BN, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1) mlp_layer_kernel(const LayerArgs L) {
  using namespace fimtc;
  using C = Cfg<BN, PAIR>;
  static_assert(MODE == 0 || BN == kTile, "the fused output layer needs the whole hidden layer in one 128-wide tile");
  static_assert(!PAIR || (MODE == 0 && BN == 256), "CTA pairs: 256 x 256 tiles of a hidden layer");
  constexpr int kS = C::kStages, kSB = C::kStageBytes, kWT = C::kWTiles, kNC = C::kCols;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tail = smem + kStages * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);      // [kS] producer -> MMA
  uint64_t* empty = full + 3;                              // [kS] MMA -> producer
  uint64_t* accfull = empty + 3;                           // [2] MMA -> promotion
  uint64_t* accempty = accfull + 2;                        // [2] promotion -> MMA
  uint64_t* peer_full = accempty + 2;                      // [kS] PAIR, leader only: the peer CTA's stage has landed
  uint64_t* peer_accempty = peer_full + 3;                 // [2]  PAIR, leader only: the peer CTA has drained an accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_accempty + 2);
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;         // 0 = leader: issues the MMAs of the pair
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kTile, n0 = blockIdx.y * BN;
  pdl::trigger();          // the next kernel of the chain may be scheduled; it waits for this grid where it first reads memory
  MLP_STAMP(0);

  float* bias_s = reinterpret_cast<float*>(tail + kTailBias);     // [BN] bias of this n-tile
  float* wout_s = reinterpret_cast<float*>(tail + kTailWout);     // [n_out][128]
  float* bout_s = reinterpret_cast<float*>(tail + kTailBout);     // [n_out]
  if (warp == 0) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (tid == 32) {
    for (int s = 0; s < kS; s++) {
      mbar_init(smem_u32(full + s), 1); mbar_init(smem_u32(empty + s), 1); mbar_init(smem_u32(peer_full + s), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(smem_u32(accfull + a), 1); mbar_init(smem_u32(accempty + a), kEpilogue); mbar_init(smem_u32(peer_accempty + a), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync();        // the peer's barriers and TMEM exist before anything is signalled / issued across the pair
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  pdl::wait();             // set-up (TMEM, barriers) overlapped the predecessor's tail; its activations / ring head are read below
  MLP_STAMP(1);
  const int n_blocks = L.Kp / kBK;
  const int n_groups = (n_blocks + kGroup - 1) / kGroup;
  const uint32_t smem_base = smem_u32(smem);

  if (warp == kEpilogue / 32 + 1) {
    // ===== bulk-copy producer =====
    if (lane == 0) {
      size_t w_copy = 0;
      if (L.rot) { const int r = *L.rot; w_copy = (size_t)(r < 0 ? 0 : (r >= L.n_rot ? L.n_rot - 1 : r)) * L.rot_stride; }
      const size_t a_tile = (size_t)blockIdx.x * n_blocks * tiled::kTileElems;
      // first of this CTA's kWT 128-row weight tiles (a pair splits the 256 rows of its n-tile between its two CTAs)
      const size_t w_tile = w_copy + (size_t)(PAIR ? blockIdx.y * 2 + pair_rank : blockIdx.y * kWT) * n_blocks * tiled::kTileElems;
      for (int kb = 0; kb < n_blocks; kb++) {
        const int s = kb % kS;
        if (kb >= kS) mbar_wait(smem_u32(empty + s), (uint32_t)((kb / kS - 1) & 1));
        const uint32_t bar = smem_u32(full + s), stage = smem_base + (uint32_t)s * kSB;
        if (MLP_DBG(2)) { mbar_arrive(bar); continue; }
        mbar_expect_tx(bar, kSB);
        const size_t ko = (size_t)kb * tiled::kTileElems;
        bulk_g2s(stage, L.a_hi + a_tile + ko, kOperandBytes, bar);
        bulk_g2s(stage + kOperandBytes, L.a_lo + a_tile + ko, kOperandBytes, bar);
#pragma unroll
        for (int t = 0; t < kWT; t++) {
          const size_t wo = w_tile + (size_t)t * n_blocks * tiled::kTileElems + ko;
          bulk_g2s(stage + (uint32_t)(2 + t) * kOperandBytes, L.w_hi + wo, kOperandBytes, bar);
          bulk_g2s(stage + (uint32_t)(2 + kWT + t) * kOperandBytes, L.w_lo + wo, kOperandBytes, bar);
        }
      }
    }
  } else if (warp == kEpilogue / 32) {
    // ===== MMA issuer =====
    if (PAIR && lane == 0 && pair_rank != 0) {
      // peer CTA of a pair: it issues no MMAs; this lane only tells the leader when this CTA's stages have landed and its
      // accumulators are drained (the leader's MMAs read this CTA's shared memory and write its tensor memory)
      for (int kb = 0; kb < n_blocks; kb++) {
        const int s = kb % kS, g = kb / kGroup;
        if ((kb % kGroup) == 0 && g >= 2) {
          mbar_wait(smem_u32(accempty + (g & 1)), (uint32_t)(((g - 2) >> 1) & 1));
          mbar_arrive_remote(smem_u32(peer_accempty + (g & 1)), 0);
        }
        mbar_wait(smem_u32(full + s), (uint32_t)((kb / kS) & 1));
        mbar_arrive_remote(smem_u32(peer_full + s), 0);
      }
    } else if (lane == 0) {
      for (int kb = 0; kb < n_blocks; kb++) {
        const int s = kb % kS, g = kb / kGroup;
        const bool group_start = (kb % kGroup) == 0, group_end = ((kb + 1) % kGroup) == 0 || kb == n_blocks - 1;
        if (group_start && g >= 2) {
          mbar_wait(smem_u32(accempty + (g & 1)), (uint32_t)(((g - 2) >> 1) & 1));
          if (PAIR) mbar_wait_cluster(smem_u32(peer_accempty + (g & 1)), (uint32_t)(((g - 2) >> 1) & 1));
        }
        mbar_wait(smem_u32(full + s), (uint32_t)((kb / kS) & 1));
        if (PAIR) mbar_wait_cluster(smem_u32(peer_full + s), (uint32_t)((kb / kS) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t stage = smem_base + (uint32_t)s * kSB;
        // a W operand of BN rows = BN / 128 adjacent 16 KB tiles (8-row groups 1024 bytes apart throughout); for BN = 128 the
        // w_hi and w_lo tiles are adjacent too, i.e. ONE 256-row operand [w_hi; w_lo]
        const uint64_t ah = make_desc_sw128(stage), al = make_desc_sw128(stage + kOperandBytes),
                       wh = make_desc_sw128(stage + 2 * kOperandBytes), wl = make_desc_sw128(stage + (2 + kWT) * kOperandBytes);
        const uint32_t tacc = tmem_d + (uint32_t)((g & 1) * kAccCols);
#pragma unroll
        for (int k = 0; k < kBK / 16; k++) {
          if (MLP_DBG(1)) break;
          const uint64_t off = (uint64_t)(k * 2);     // 32 bytes >> 4, inside the 14-bit start-address field
          const uint32_t first = (!group_start || k > 0) ? 1u : 0u;
          if (PAIR) {              // M = 256 over the pair, N = 256: this CTA's 128 weight rows + the peer's
            mma_f16_pair(tacc, al + off, wh + off, idesc_f16_pair(256), first);
            mma_f16_pair(tacc, ah + off, wl + off, idesc_f16_pair(256), 1u);
            mma_f16_pair(tacc, ah + off, wh + off, idesc_f16_pair(256), 1u);
          } else if (BN == 256) {
            mma_f16_n(tacc, al + off, wh + off, idesc_f16(256), first);
            mma_f16_n(tacc, ah + off, wl + off, idesc_f16(256), 1u);
            mma_f16_n(tacc, ah + off, wh + off, idesc_f16(256), 1u);
          } else {
            mma_f16_n(tacc, ah + off, wh + off, idesc_f16(256), first);   // [0,128) += a_hi w_hi, [128,256) += a_hi w_lo
            mma_f16_n(tacc, al + off, wh + off, idesc_f16(128), 1u);      // [0,128) += a_lo w_hi
          }
        }
        if (PAIR) {
          umma_commit_pair(smem_u32(empty + s));
          if (group_end) umma_commit_pair(smem_u32(accfull + (g & 1)));
        } else {
          umma_commit(smem_u32(empty + s));
          if (group_end) umma_commit(smem_u32(accfull + (g & 1)));
        }
      }
    }
  } else {
    // ===== promotion + epilogue: thread = (row, column set) =====
    // BN = 128: thread (q, hf) owns output columns 64 hf + [0, 64) and adds the two accumulator halves;
    // BN = 256: it owns columns 64 hf + [0, 64) of BOTH 128-column halves of the tile (so that each half of the output can
    //           be staged in shared memory by all 256 threads), held as acc[half * 64 + i]
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;                     // row of the tile = TMEM lane
    const int c0 = hf * kHalfCols;
    // epilogue constants, fetched while the pipeline fills (ordered before their use by the epilogue barrier below)
    for (int i = tid; i < BN; i += kEpilogue) bias_s[i] = __ldg(L.bias + n0 + i);
    if (MODE == 1) {
      for (int i = tid; i < L.n_out * kTile; i += kEpilogue) wout_s[i] = __ldg(L.w_out + i);
      for (int i = tid; i < L.n_out; i += kEpilogue) bout_s[i] = __ldg(L.b_out + i);
    }
    epi_barrier();
    float acc[kNC];
#pragma unroll
    for (int i = 0; i < kNC; i++) acc[i] = 0.f;
    for (int g = 0; g < n_groups; g++) {   // acc += accumulator of group g, then hand the accumulator back to the MMA warp
      mbar_wait(smem_u32(accfull + (g & 1)), (uint32_t)((g >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int cb = 0; cb < 4; cb++) {     // 2 halves x 2 x 32 columns
        if (MLP_DBG(4)) break;
        uint32_t v[32];
        const int half = cb >> 1, sub = cb & 1;
        tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)((g & 1) * kAccCols + half * kTile + c0 + sub * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; i++) acc[(BN == 256 ? half * kHalfCols : 0) + sub * 32 + i] += __uint_as_float(v[i]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(accempty + (g & 1)));
    }
    MLP_STAMP(2);
    pdl::trigger_late();     // main loop done: the next layer's CTAs may be placed (on other SMs) and set up while this one runs its epilogue
    // every MMA has retired (the last accfull arrived) and every copy has landed: the stages are free
    if (MODE == 0) {
      // ELU(acc / (s_a s_w) + b) x s_a -> fp16 hi / lo, written as the next layer's operand tiles (tiled_layout.cuh): 128
      // output columns are 2 k-blocks x {hi, lo} = 4 tiles of 16 KB, assembled in shared memory, stored with 4 bulk copies
      constexpr float kInv = 1.0f / (kActScale * kWeightScale);
#pragma unroll
      for (int half = 0; half < BN / kTile; half++) {
        if (half > 0) epi_barrier();                 // the previous half has left shared memory (tid 0 waited for the reads)
#pragma unroll
        for (int i = 0; i < kHalfCols; i += 8) {
          const int col = c0 + i;                    // within this 128-column half
          const float* a = acc + half * kHalfCols + i;
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + half * kTile + col);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_s + half * kTile + col + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          __half2 ph[4], pl[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            // clamp into the finite fp16 range before the split (an activation >= 65504 / 16 would make hi = inf, lo = NaN;
            // ELU is bounded below by -1): the documented operating range is |activation| < 4000
            const float h0 = fminf(elu(fmaf(a[j], kInv, bb[j])) * kActScale, 65000.0f);
            const float h1 = fminf(elu(fmaf(a[j + 1], kInv, bb[j + 1])) * kActScale, 65000.0f);
            const __half2 hh = __floats2half2_rn(h0, h1);
            const float2 hf2 = __half22float2(hh);
            ph[j >> 1] = hh;
            pl[j >> 1] = __floats2half2_rn(h0 - hf2.x, h1 - hf2.y);
          }
          const int ob = col >> 6, ch = (col & 63) >> 3;
          unsigned char* dst = smem + ob * kOperandBytes + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(ph);
          *reinterpret_cast<uint4*>(dst + 2 * kOperandBytes) = *reinterpret_cast<const uint4*>(pl);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk-copy engine
        epi_barrier();
        if (half == 0) MLP_STAMP(5);
        if (tid == 0) {
          const size_t t0 = ((size_t)blockIdx.x * (size_t)(L.out_stride >> 6) + (size_t)((n0 + half * kTile) >> 6)) * tiled::kTileElems;
#pragma unroll
          for (int ob = 0; ob < 2; ob++) {
            bulk_s2g(L.out_hi + t0 + (size_t)ob * tiled::kTileElems, smem_base + (uint32_t)ob * kOperandBytes, kOperandBytes);
            bulk_s2g(L.out_lo + t0 + (size_t)ob * tiled::kTileElems, smem_base + (uint32_t)(2 + ob) * kOperandBytes, kOperandBytes);
          }
          bulk_commit_wait_read();
        }
      }
    } else {
      // last hidden layer (one n-tile) + the output layer: y[j] = b_out[j] + sum_i w_out[j][i] ELU(acc[i] + b[i]); each thread
      // contracts its 64 columns (4 independent partial sums), the two halves of a row meet in shared memory in a fixed order
      float* part = reinterpret_cast<float*>(smem);      // [128][kMaxOut]
#pragma unroll
      for (int i = 0; i < kHalfCols; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + i);
        constexpr float kInv = 1.0f / (kActScale * kWeightScale);
        acc[i] = elu(fmaf(acc[i], kInv, b4.x)); acc[i + 1] = elu(fmaf(acc[i + 1], kInv, b4.y));
        acc[i + 2] = elu(fmaf(acc[i + 2], kInv, b4.z)); acc[i + 3] = elu(fmaf(acc[i + 3], kInv, b4.w));
      }
      const int row = m0 + r;
      float yv[kMaxOut];
#pragma unroll
      for (int j = 0; j < kMaxOut; j++) {
        if (j >= L.n_out) break;
        float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
        const float4* w4 = reinterpret_cast<const float4*>(wout_s + j * kTile + c0);
#pragma unroll
        for (int i = 0; i < kHalfCols / 4; i++) {
          const float4 w = w4[i];
          y0 = fmaf(w.x, acc[4 * i], y0); y1 = fmaf(w.y, acc[4 * i + 1], y1);
          y2 = fmaf(w.z, acc[4 * i + 2], y2); y3 = fmaf(w.w, acc[4 * i + 3], y3);
        }
        const float y = (y0 + y1) + (y2 + y3);
        yv[j] = y;
        if (hf == 1) part[r * kMaxOut + j] = y;
      }
      epi_barrier();
      if (hf == 0 && row < L.M) {
#pragma unroll
        for (int j = 0; j < kMaxOut; j++)
          if (j < L.n_out) L.out[(size_t)row * L.n_out + j] = bout_s[j] + (yv[j] + part[r * kMaxOut + j]);
      }
    }
  }
  MLP_STAMP(3);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  MLP_STAMP(4);
  if (PAIR) {
    cluster_sync();                // neither CTA of the pair leaves (or frees tensor memory) while the other may still signal it
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  } else if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
  }
}

// the fp16 pair of x * scale
__device__ __forceinline__ void split_half(float x, float scale, __half* hi, __half* lo) {
  const float v = x * scale;
  const __half h = __float2half_rn(v);
  *hi = h;
  *lo = __float2half_rn(v - __half2float(h));
}

// x [M, K] fp32 row-major -> hi / lo in the tiled layout [Mp, Kp] (padding is left as it is: zero-initialised by the owner)
__global__ void split_kernel(const float* x, int M, int K, __half* hi, __half* lo, int Kp) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * K) return;
  const int m = (int)(idx / K), k = (int)(idx - (size_t)m * K);
  const size_t o = tiled::offset(m, k, Kp);
  split_half(x[idx], kActScale, hi + o, lo + o);
}

// (hi + lo) / scale in the tiled layout -> x [M, K] row-major (inspection / tests; exact to ~2^-22 relative)
__global__ void unsplit_kernel(const __half* hi, const __half* lo, int M, int K, int Kp, float* x) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)M * K) return;
  const int m = (int)(idx / K), k = (int)(idx - (size_t)m * K);
  const size_t o = tiled::offset(m, k, Kp);
  x[idx] = (__half2float(hi[o]) + __half2float(lo[o])) * (1.0f / kActScale);
}

}  // namespace mlptc
