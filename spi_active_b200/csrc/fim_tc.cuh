// fim_tc.cuh — accumulated Fisher-information contraction on the 5th-generation tensor cores.
//
// Reference: spigym/envs/sysid/active_sysid_openloop.py:402-426 computes, per main env and per control step,
//   J_t = [(root13_main - root13_aux_p) / delta  ||  (q_main - q_aux_p) / delta]  in R^{P x 25},   XtX = J_t J_t^T,
// and spigym/agents/sysid/active_sysid.py:528-600 sums trace(XtX) over the 1 248 steps of a rollout, scoring
// terminated groups with termination_rew = 0.  The full information matrix of a command trajectory is the dense
// contraction
//       FIM[m] = sum_t live[t,m] * J_t[m] J_t[m]^T  =  A_m A_m^T,     A_m in R^{P x (25 T)}
// i.e. a batched K-major "A A^T" GEMM with K = 25 T (31 200 at the reference's rollout length).
//
// Mapping (sm_100a): one CTA = 8 consecutive main envs x 16 parameter slots = 128 rows (UMMA M = N = 128).
//   * producer (all 128 threads, thread = row): load the recorded states of 2 control steps, form the finite
//     differences, split every fp32 value into tf32 hi + lo (3xTF32: fp32-accurate products) and store both tiles
//     in shared memory in the K-major SWIZZLE_128B UMMA layout (a row = 32 tf32 = one 128-byte swizzle atom);
//   * one elected thread issues tcgen05.mma.kind::tf32 (hi*hi, hi*lo, lo*hi) accumulating the 128 x 128 fp32 tile
//     in tensor memory; the same shared-memory tile is both the A and the B operand;  tcgen05.commit -> mbarrier
//     releases the stage back to the producers (2 stages: production of block k+1 overlaps the MMAs of block k);
//   * the tensor core accumulates in TMEM with truncation (measured: a 1 600-deep sum of squares drifts by 1.5e-5
//     relative), so every stage has its OWN accumulator holding one block's 24 MMAs only; when a stage is recycled
//     its diagonal 16 x 16 blocks (the only env-diagonal part of A A^T) are read back with tcgen05.ld and added to a
//     running sum in fp32 registers with round-to-nearest (the "promotion" trick of FP8 GEMMs);
//   * epilogue: running sums -> out_JtJ, trace by shuffles.
// The off-diagonal env blocks of the 128 x 128 tile are computed and dropped: UMMA has no M = 16 shape; the kernel is bound by
// the producer (global loads + difference / split / tile stores), not by the tensor pipe.
// r2 (profiles/README.md): one step per stage (64 KB of tiles -> 2 CTAs per SM), the steps cut into gridDim.y slices so that a
// 64-step chunk still fills the GPU, coalesced cooperative loads through a staging buffer with a one-step register prefetch.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fimtc {

constexpr int kRows = 128;            // UMMA M = N
constexpr int kEnvsPerCta = 8;
constexpr int kSlots = 16;            // parameter slots per env (P <= 16)
constexpr int kStateDim = 25;         // root13 + q12
constexpr int kKPerStep = 32;         // 25 padded to a multiple of the tf32 UMMA K (8)
constexpr int kStepsPerStage = 1;     // r2: one control step per stage -> 2 x 2 x 16 KB of tiles per CTA -> 2 CTAs per SM (TMEM-limited)
constexpr int kKPerStage = kKPerStep * kStepsPerStage;      // 32 floats = 128 bytes per row per stage = ONE 128-byte swizzle atom
constexpr int kChunksPerStage = kKPerStage / 4;             // 16-byte chunks along K (8)
constexpr int kGroupStride = 1024;                          // bytes between 8-row groups (SBO): 8 rows x 128 bytes
constexpr int kTileBytes = kRows * kKPerStage * 4;          // 16 KB
static_assert(kKPerStage * 4 == 128, "a tile row must be exactly one 128-byte swizzle atom");
constexpr int kStages = 2;
constexpr int kRawStride = (kSlots + 1) * kStateDim + 3;    // floats per env of the raw staging buffer (odd stride: the 16
                                                            // lanes of an env and the 2 envs of a warp spread over the banks)
constexpr int kRawBytes = kEnvsPerCta * kRawStride * 4;
constexpr int kMaxLoads = ((kSlots + 1) * kStateDim + 15) / 16;   // global loads per thread and step (27 at P = 16, 18 at P = 10)
constexpr int kSmemBytes = kStages * 2 * kTileBytes + kRawBytes + 64;   // hi + lo per stage, raw rows, barriers / tmem slot
constexpr int kTmemCols = 2 * kRows;   // one 128-column fp32 accumulator per stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout; same layout as mlp_tc.cuh):
//   [0,14) start address >> 4 | [16,30) leading byte offset field = 1 (unused: one swizzle atom spans the tile's K extent)
//   [32,46) stride byte offset >> 4 = 1024 >> 4 between 8-row groups | [46,48) version = 1 | [61,64) layout type = 2
// A row is 128 bytes (32 tf32 of K); the 16-byte chunk c of row r sits at chunk position c ^ (r & 7); a k-step of 8 tf32
// advances the start address by 32 bytes inside the atom.  r1 used the SWIZZLE_NONE canonical layout, whose 128 x 128 x 8 MMA
// takes ~256 cycles instead of ~65 (measured on the actor MLP, profiles/README.md r1_c) — 12 such MMAs per control step
// (~3 000 cycles) were what bounded this kernel, not its loads.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((kGroupStride >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b_format TF32 (2) at [7,10) /
// [10,13), K-major A and B (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct FimArgs {
  const float* hist;            // [T][M][P+1][25]  recorded (root13 origin-compensated, q12) of main + aux envs
  const unsigned char* live;    // [T][M] 1 = the group's step counts, or null
  int T, M, P;
  float inv_delta;
  int accumulate;
  float* out_JtJ;               // [M][P][P] or null
  float* out_trace;             // [M] or null
  // r2: the T control steps are cut into gridDim.y slices so that a 64-step chunk of 1 024 envs (128 env tiles) still fills
  // 148 SMs x 2 CTAs.  n_split > 1: slice y writes its partial sums to part_JtJ [n_split][M][P][P] / part_trace [n_split][M]
  // and fim_reduce_kernel adds the slices in a fixed order (deterministic) into out_*.
  int n_split;
  float* part_JtJ; float* part_trace;
};

__global__ void __launch_bounds__(kRows, 2) fim_contract_kernel(const FimArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = smem;                                         // [stage][hi, lo][kTileBytes]
  float* raw = reinterpret_cast<float*>(smem + kStages * 2 * kTileBytes);          // [8 envs][kRawStride] rows of one step
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * 2 * kTileBytes + kRawBytes);   // [kStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kStages);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < kStages; s++) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // this thread's row of the 128-row tile
  const int e = tid >> 4, p = tid & 15;
  const int m = blockIdx.x * kEnvsPerCta + e;
  const bool row_valid = (m < A.M) && (p < A.P);
  const int env_floats = (A.P + 1) * kStateDim;                        // contiguous floats of one env and step
  const size_t step_stride = (size_t)A.M * env_floats;
  const float* env_base = A.hist + (size_t)(m < A.M ? m : 0) * env_floats;
  float* raw_env = raw + e * kRawStride;
  // byte offset of this row in a tile, and its swizzle phase: chunk c of the row goes to chunk position c ^ (row & 7)
  const uint32_t row_off = (uint32_t)(tid >> 3) * kGroupStride + (uint32_t)(tid & 7) * 128;
  const uint32_t row_xor = (uint32_t)(tid & 7);
  // this CTA's slice of the control steps
  const int per = (A.T + A.n_split - 1) / A.n_split;
  const int t_begin = blockIdx.y * per, t_end = min(A.T, t_begin + per);

  // running fp32 sums of this thread's row: columns 32 * warp .. 32 * warp + 31 of the 128 x 128 tile
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = 0.f;
  auto flush = [&](int kb2) {   // add block kb2's accumulator (complete once its commit has arrived) to the running sums
    mbar_wait(smem_u32(bars + (kb2 & 1)), (uint32_t)((kb2 / kStages) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)((kb2 & 1) * kRows + warp * 32);
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] += __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the loads precede the MMAs that recycle the accumulator
  };

  // Producer, r2: the (P + 1) x 25 floats of an env and step are CONTIGUOUS in the history, so the 16 threads of an env fetch
  // them as coalesced 64-byte segments (18 loads per thread at P = 10 instead of 50 scalar loads of two 100-byte rows each),
  // one step ahead of their use (software prefetch into registers), park them in shared memory and form the differences
  // from there.  The 16 threads of an env sit in one warp: __syncwarp orders the staging buffer.
  float pre[kMaxLoads];
  unsigned char pre_live = 1;
  auto prefetch = [&](int t) {
    const bool on = (m < A.M) && (t < t_end);
    pre_live = (on && A.live) ? A.live[(size_t)t * A.M + m] : (unsigned char)1;
    const float* src = env_base + (size_t)t * step_stride;
#pragma unroll
    for (int k = 0; k < kMaxLoads; k++) {
      const int idx = p + 16 * k;
      pre[k] = (on && idx < env_floats) ? __ldg(src + idx) : 0.f;
    }
  };
  prefetch(t_begin);
  const int n_blocks = t_end - t_begin;
  for (int kb = 0; kb < n_blocks; kb++) {
    const int s = kb & 1;
    const int t = t_begin + kb;
    __syncwarp();                                  // the previous step's reads of the staging rows are done
#pragma unroll
    for (int k = 0; k < kMaxLoads; k++) {
      const int idx = p + 16 * k;
      if (idx < env_floats) raw_env[idx] = pre[k];
    }
    const bool live_now = pre_live != 0;           // fetched one step ahead too: a dependent global load per step was exposed
    prefetch(t + 1);                               // in flight during the conversion below
    __syncwarp();
    if (kb >= kStages) flush(kb - kStages);        // also guarantees the MMAs that read this stage's tiles are done
    unsigned char* hi_tile = tiles + (size_t)(s * 2) * kTileBytes;
    unsigned char* lo_tile = hi_tile + kTileBytes;
    const bool on = row_valid && live_now;
    float j[kKPerStep];
    if (on) {
      const float* auxp = raw_env + (p + 1) * kStateDim;
#pragma unroll
      for (int d = 0; d < kStateDim; d++) j[d] = (raw_env[d] - auxp[d]) * A.inv_delta;
    } else {
#pragma unroll
      for (int d = 0; d < kStateDim; d++) j[d] = 0.f;
    }
#pragma unroll
    for (int d = kStateDim; d < kKPerStep; d++) j[d] = 0.f;
#pragma unroll
    for (int c = 0; c < kKPerStep / 4; c++) {
      float4 h, l;
      h.x = tf32_round(j[4 * c + 0]); l.x = j[4 * c + 0] - h.x;
      h.y = tf32_round(j[4 * c + 1]); l.y = j[4 * c + 1] - h.y;
      h.z = tf32_round(j[4 * c + 2]); l.z = j[4 * c + 2] - h.z;
      h.w = tf32_round(j[4 * c + 3]); l.w = j[4 * c + 3] - h.w;
      const uint32_t off = row_off + (((uint32_t)c ^ row_xor) << 4);
      *reinterpret_cast<float4*>(hi_tile + off) = h;
      *reinterpret_cast<float4*>(lo_tile + off) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the UMMA reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t hi_addr = smem_u32(hi_tile), lo_addr = smem_u32(lo_tile);
      const uint32_t tacc = tmem_d + (uint32_t)(s * kRows);
      // small cross terms first, then the hi * hi products: fewer truncated additions at full magnitude
#pragma unroll
      for (int k = 0; k < kKPerStage / 8; k++) {     // a k-step of 8 tf32 = 32 bytes inside the swizzle atom
        const uint64_t dh = make_desc(hi_addr + (uint32_t)k * 32);
        const uint64_t dl = make_desc(lo_addr + (uint32_t)k * 32);
        mma_tf32(tacc, dh, dl, k > 0 ? 1u : 0u);
        mma_tf32(tacc, dl, dh, 1u);
      }
#pragma unroll
      for (int k = 0; k < kKPerStage / 8; k++) {
        const uint64_t dh = make_desc(hi_addr + (uint32_t)k * 32);
        mma_tf32(tacc, dh, dh, 1u);
      }
      umma_commit(smem_u32(bars + s));
    }
  }
  for (int kb2 = (n_blocks > kStages ? n_blocks - kStages : 0); kb2 < n_blocks; kb2++) flush(kb2);

  const bool upper = (lane >> 4) != 0;
  const bool split = A.n_split > 1;
  float* dst_jtj = split ? A.part_JtJ + (size_t)blockIdx.y * A.M * A.P * A.P : A.out_JtJ;
  float* dst_trace = split ? A.part_trace + (size_t)blockIdx.y * A.M : A.out_trace;
  const bool accumulate = !split && A.accumulate;
  float diag = 0.f;
#pragma unroll
  for (int q = 0; q < kSlots; q++) {
    const float v = upper ? acc[16 + q] : acc[q];
    if (q == p) diag = v;
    if (row_valid && q < A.P && dst_jtj) {
      float* o = dst_jtj + ((size_t)m * A.P + p) * A.P + q;
      *o = accumulate ? (*o + v) : v;
    }
  }
  if (!row_valid) diag = 0.f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) diag += __shfl_xor_sync(0xffffffffu, diag, o);   // sum over the env's 16 slots
  if (p == 0 && m < A.M && dst_trace) dst_trace[m] = accumulate ? (dst_trace[m] + diag) : diag;

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kTmemCols) : "memory");
}

// slices of the control steps -> out (fixed order: deterministic), n = M P P (JtJ) or M (trace)
__global__ void fim_reduce_kernel(const float* part, int n_split, size_t n, int accumulate, float* out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = accumulate ? out[i] : 0.f;
  for (int s = 0; s < n_split; s++) v += part[(size_t)s * n + i];
  out[i] = v;
}

}  // namespace fimtc
