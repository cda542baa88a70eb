// rollout_ws.cuh — the warp-specialised fused rollout kernel (see go2_ws.cuh for the role arithmetic and
// DESIGN.md §4.2 for the mapping).  One CTA = 32 (candidate, segment) rollouts x 5 warps:
//   4 leg warps (lane = rollout) + 1 base warp.  The base role sits on the HIGHEST warp id of the CTA: 5-warp CTAs then spread
//   over the 4 sub-partitions so that each hosts exactly one base warp and four leg warps of four different CTAs (rotating the
//   roles with blockIdx clusters base warps: 8.6 % slower, profiles/README.md).
// Per integrator sub-step:   legs: [torques of a new physics step] phase 1 -> pair sums -> smem[27] | barrier [A] |
//   base: sum, 6x6 solve -> a0 -> smem[6] | barrier [B1] | legs: phase 2, overlapped with base: integrate, publish
//   R / v0 / pz -> smem[16] | barrier [B2] | base: bias force of the next sub-step.
// The kernel is bound by the latency of this round, not by issue bandwidth (tools/ws_timeline.py).
#pragma once
#include "go2_ws.cuh"
#include "pdl.cuh"

namespace ws {

constexpr int kWsWarps = 5;
constexpr int kWsThreads = 32 * kWsWarps;
constexpr int kWsRollouts = 32;

struct WsArgs {
  ModelK M;
  const float* params; int C, P; ParamIdsK ids;
  const float* seg_init; const float* seg_actions; const float* seg_target; const float* seg_gains;
  const unsigned char* seg_mask;
  int S, H, decimation, motor_model; unsigned flags;
  int n_cta_per_cand;
  int C_grid;          // number of candidate rows of the grid (C; 1 in paired mode)
  int tail_lanes;      // 0: every candidate owns ceil(S / 32) CTAs (the last one padded).  L' > 0 (dense packing): a candidate owns its
                       // S / 32 full CTAs only; the S % 32 left-over segments of 32 / L' candidates share one TAIL CTA, each in its own
                       // aligned group of L' lanes (L' = the power of two >= S % 32).  S = 1730: 2 left-over segments, 16 candidates
                       // per tail CTA, 54.06 instead of 55 CTAs per candidate
  int rotate_roles;
  int paired;          // 1: rollout r uses candidate row r AND segment r (one env per rollout; C == 1 for the grid)
  float* partial;      // [C][n_cta_per_cand][3]
  float* per_seg;      // [C][S][3] or null
  int* bad;            // [C]
  float* out_states;   // [C][S][H][37] (RECORD)
  const unsigned char* zero_mask;   // paired mode: [S] 1 = this env's action is replaced by 0 (terminated group), or null
#if defined(SPI_WS_PROFILE)         // dev builds only (tools/ws_timeline.py): clock64 stamps of every barrier of a window of CTAs
  long long* prof; int prof_blk0, prof_nblk;
#endif
};

#if defined(SPI_WS_PROFILE)
constexpr int kProfSlots = 256;
struct WsProf {
  long long* p; int n;
  __device__ __forceinline__ void init(const WsArgs& A, int lane, int warp) {
    const int b = (int)blockIdx.x - A.prof_blk0;
    p = (A.prof && lane == 0 && b >= 0 && b < A.prof_nblk) ? A.prof + ((size_t)b * kWsWarps + warp) * kProfSlots : nullptr;
    n = 1;
    if (p) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p[0] = sm; }
  }
  // BAR.SYNC.DEFER_BLOCKING lets a warp run ahead of an incomplete barrier until its next shared-memory access, and ptxas may
  // move a clock read past arithmetic: every stamp is therefore PREDICATED on a value (v == v) that the event of interest
  // produces — the data loaded behind the barrier, or the last result of a computation
  __device__ __forceinline__ void stamp(float v) {
    long long t = 0;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.f32 p, %1, %1;\n\t@p mov.u64 %0, %%clock64;\n\t}" : "+l"(t) : "f"(v) : "memory");
    if (p && n < kProfSlots) p[n] = t;
    n++;
  }
};
#define WS_PROF_INIT(warp) WsProf prof_; prof_.init(A, lane, warp)
#define WS_STAMP(v) prof_.stamp(v)
#else
#define WS_PROF_INIT(warp)
#define WS_STAMP(v)
#endif

// The roles exchange their per-sub-step data as float4 (LDS.128 / STS.128, lane stride 16 bytes: conflict-free): 7 + 6
// shared-memory instructions per leg and 28 + 6 + 4 for the base instead of 27 + 22 and 108 + 22 + 12 scalar ones — the
// exchange was 7 % of all issued instructions of a kernel that is bound by issue slots (profiles/README.md).
constexpr int kLegVec = (kLegOut + 3) / 4;     // 7: the 27 leg -> base floats, padded
constexpr int kBcStateVec = 4;                 // R(9) v0(6) pz(1) = logical bc[6 .. 22)
constexpr int kBcA0Vec = 2;                    // a0(6) + 2 pad   = logical bc[0 .. 6)
struct WsSmem {
  float4 part[4][kLegVec][32];                     // leg -> base, per sub-step
  float4 bc[kBcStateVec + kBcA0Vec][32];           // base -> legs, per sub-step: [0,4) state of the sub-step, [4,6) a0
  float ej[4][32];                                 // per-leg squared joint error (end of rollout)
  int finite[4][32];
#if defined(SPI_WS_CALF_HARMONIC)
  alignas(16) float calf[4][64];                   // per leg: harmonic coefficients of the calf's share of the thigh inertia
#endif
};
static_assert(kBaseOut - kBcR == 4 * kBcStateVec && kBcR == 6 && kBcA0 == 0, "bc packing");

// a shared-memory read the scheduler may not move across a barrier or merge with an earlier register copy
__device__ __forceinline__ float4 ws_lds_volatile(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void ws_load_state(const WsSmem& sm, int lane, float* bc) {     // bc[6 .. 22) <- shared
#pragma unroll
  for (int v = 0; v < kBcStateVec; v++) {
    const float4 t = sm.bc[v][lane];
    bc[kBcR + 4 * v] = t.x; bc[kBcR + 4 * v + 1] = t.y; bc[kBcR + 4 * v + 2] = t.z; bc[kBcR + 4 * v + 3] = t.w;
  }
}
__device__ __forceinline__ void ws_store_state(WsSmem& sm, int lane, const float* bc) {    // shared <- bc[6 .. 22)
#pragma unroll
  for (int v = 0; v < kBcStateVec; v++)
    sm.bc[v][lane] = make_float4(bc[kBcR + 4 * v], bc[kBcR + 4 * v + 1], bc[kBcR + 4 * v + 2], bc[kBcR + 4 * v + 3]);
}
__device__ __forceinline__ void ws_load_a0(const WsSmem& sm, int lane, float* a0) {
  const float4 t0 = sm.bc[kBcStateVec][lane], t1 = sm.bc[kBcStateVec + 1][lane];
  a0[0] = t0.x; a0[1] = t0.y; a0[2] = t0.z; a0[3] = t0.w; a0[4] = t1.x; a0[5] = t1.y;
}
__device__ __forceinline__ void ws_store_a0(WsSmem& sm, int lane, const float* a0) {
  sm.bc[kBcStateVec][lane] = make_float4(a0[0], a0[1], a0[2], a0[3]);
  sm.bc[kBcStateVec + 1][lane] = make_float4(a0[4], a0[5], 0.f, 0.f);
}


__device__ __forceinline__ bool finite_acc(float acc) { return acc == 0.f; }  // NaN/Inf * 0 = NaN

// CTA-wide barrier reached from role-specific code paths: a named barrier with an explicit thread count
// (every role executes the same number of ws_barrier() calls).
__device__ __forceinline__ void ws_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kWsThreads) : "memory"); }

// Named-barrier plan: barrier 1 = the CTA's 160 threads, 2 / 3 = the two leg pairs (64 threads each).
// (Round 2 also measured two 32-rollout groups per 10-warp CTA whose leg phases 1 were kept from overlapping by a token on
// named barriers — 43.6 ms vs 41.4 ms, profiles/README.md r2; removed again, git history: commit 6ac9449.)
struct WsBars {
  __device__ __forceinline__ void cta() const { ws_barrier(); }
  __device__ __forceinline__ void pair_arrive(int leg) const {
    if (leg < 2) asm volatile("bar.arrive 2, 64;" ::: "memory");
    else asm volatile("bar.arrive 3, 64;" ::: "memory");
  }
  __device__ __forceinline__ void pair_sync(int leg) const {
    if (leg < 2) asm volatile("bar.sync 2, 64;" ::: "memory");
    else asm volatile("bar.sync 3, 64;" ::: "memory");
  }
};

// LEG is a warp-uniform run-time value (one code copy for the four legs: four template copies overflow the
// instruction cache — profiles/README.md, experiment ws-templated); the leg's constants are then fetched through
// the constant bank with a uniform offset.
// MOTOR >= 0: the motor model is a compile-time constant (2 * SPI_MOTOR_* + tanh-before-clip): the throughput kernel is
// instantiated per model so that the per-physics-step torque block carries no dispatch (a chain of uniform compares / branches);
// MOTOR < 0: run-time dispatch on A.motor_model / A.flags.
template <bool RECORD, int MOTOR>
__device__ __forceinline__ void ws_leg_role(const WsArgs& A, WsSmem& sm, const WsBars bars, int lane, const int LEG,
                                            int c, int seg, bool active) {
  const SimK& S = A.M.sim;
  const LegK& L = A.M.leg[LEG];
  // this leg's motor parameters (act2tau_scalar uses one gain for every joint) and k2 = 2 log2(e) / gain
  float motor[3], motor_k2[3];
  {
    float m3[3] = {20.0f, 20.0f, 20.0f};
    if (A.params) {
      const float* row = A.params + (size_t)c * A.P;
      for (int p = 0; p < A.ids.n; p++) {
        const int id = A.ids.id[p];
        if (id == SPI_PARAM_MOTOR_HIP) m3[0] = row[p];
        else if (id == SPI_PARAM_MOTOR_THIGH) m3[1] = row[p];
        else if (id == SPI_PARAM_MOTOR_CALF) m3[2] = row[p];
      }
    }
#pragma unroll
    for (int j = 0; j < 3; j++) { motor[j] = m3[j]; motor_k2[j] = kTwoLog2e * rcp_fast(m3[j]); }
  }
  const float hip_scale = (A.flags & SPI_FLAG_HIP_HALF) ? 0.5f : 1.0f;
  LegState s;
  float kp[3], kd[3];
  {
    // plain loads: in the paired RECORD launch of spi_b200_env_step the output rows alias seg_init, and ld.global.nc must not
    // be used on memory the same kernel writes
    const float* row = A.seg_init + (size_t)seg * SPI_STATE_DIM;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      s.q[j] = *(row + 13 + 3 * LEG + j);
      s.qd[j] = *(row + 25 + 3 * LEG + j);
      kp[j] = A.seg_gains ? __ldg(A.seg_gains + (size_t)seg * 24 + 3 * LEG + j) : A.M.kp[3 * LEG + j];
      kd[j] = A.seg_gains ? __ldg(A.seg_gains + (size_t)seg * 24 + 12 + 3 * LEG + j) : A.M.kd[3 * LEG + j];
    }
  }
  const float h = S.dt / (float)S.nsub;
  const float* act_row = A.seg_actions + (size_t)seg * A.H * 12 + 3 * LEG;
  LegKeep K;
  const bool zero_act = A.zero_mask && A.zero_mask[seg] != 0;
  // (Measured and dropped, profiles/README.md r2: evaluating the torques / sin / cos of the next sub-step in the TAIL of the
  // previous one, in front of [B2] — the legs hardly wait there, so it only delays the barrier: 39.9 ms vs 39.2 ms; parking the
  // PD bias / gains / motor gains in shared memory: 39.2 - 39.4 ms, and the extra 4 KB per CTA keep a rollout CTA from sharing an
  // SM with an actor-MLP CTA in the pipelined exploration rollout; 5 CTAs per SM at 72 registers: 40.8 ms, 3 at 126: 43.1 ms.)
#if defined(SPI_WS_CALF_HARMONIC)
  sm.calf[LEG][lane] = L.calfA2[lane];
  sm.calf[LEG][lane + 32] = L.calfA2[lane + 32];
  __syncwarp();
  const float* calf_coef = sm.calf[LEG];
#endif
  WS_PROF_INIT(LEG);
  bars.cta();   // [S0] the base role has published R / v0 / pz of the initial state
  for (int k = 0; k < A.H; k++) {
    float pdb[3];        // kp (a s + q_default): what the PD law needs of the control step
    {
      float act[3];
#pragma unroll
      for (int j = 0; j < 3; j++)
        act[j] = zero_act ? 0.f : fminf(fmaxf(__ldg(act_row + 12 * k + j), -S.action_clip), S.action_clip);
      // the next control step's actions: into L1 now, so that their load is a hit (L2: ~800 cycles under load)
      if (k + 1 < A.H) asm volatile("prefetch.global.L1 [%0];" ::"l"(act_row + 12 * (k + 1)));
      leg_pd_bias(S, L, act, kp, hip_scale, pdb);
    }
    for (int d = 0; d < A.decimation; d++) {
      // PD law + clip + motor model, once per physics step (go2_ws.cuh: leg_torques_t)
      float tau[3];
      if constexpr (MOTOR >= 0) leg_torques_t<MOTOR / 2, (MOTOR & 1) != 0>(L, pdb, s.q, s.qd, kp, kd, motor, motor_k2, tau);
      else leg_torques_dispatch(L, pdb, s.q, s.qd, kp, kd, motor, motor_k2, A.motor_model, A.flags, tau);
      for (int n = 0; n < S.nsub; n++) {
        float bc[kBaseOut];
        ws_load_state(sm, lane, bc);
        WS_STAMP(bc[kBcPz]);   // 0: [B2] released, the state of the sub-step has arrived: phase 1 starts
        float out[4 * kLegVec];
        out[4 * kLegVec - 1] = 0.f;
        leg_angles(s.q, K);
#if defined(SPI_WS_CALF_HARMONIC)
        leg_phase1_core(S, L, bc, s, tau, K, out, nullptr, calf_coef);
#else
        leg_phase1_core(S, L, bc, s, tau, K, out, nullptr);
#endif
        WS_STAMP(out[26]);   // 1: phase 1 computed
        // the four legs are summed pairwise: the even leg of a pair publishes and signals (bar.arrive on the pair's own named
        // barrier, 64 threads), the odd leg waits for it, adds its own contribution and publishes the pair sum — the base
        // role then only adds two vectors in its serial section ((p0 + p1) + (p2 + p3), the same order as before)
        if ((LEG & 1) == 0) {
#pragma unroll
          for (int v = 0; v < kLegVec; v++)
            sm.part[LEG][v][lane] = make_float4(out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
          bars.pair_arrive(LEG);
        } else {
          bars.pair_sync(LEG);
#pragma unroll
          for (int v = 0; v < kLegVec; v++) {
            const float4 p = sm.part[LEG - 1][v][lane];
            sm.part[LEG][v][lane] = make_float4(p.x + out[4 * v], p.y + out[4 * v + 1], p.z + out[4 * v + 2], p.w + out[4 * v + 3]);
#if defined(SPI_WS_PROFILE)
            if (v == kLegVec - 1) out[26] += p.z;   // (dev builds) the stamp below waits for the partner's data
#endif
          }
        }
        WS_STAMP(out[26]);   // 2: published (odd legs: the partner's vector has arrived and is added)
        bars.cta();   // [A]  the pair sums are in shared memory
        bars.cta();   // [B1] the base role has published a0
        ws_load_a0(sm, lane, bc + kBcA0);
        WS_STAMP(bc[kBcA0]);   // 3: [B1] released, a0 has arrived
        leg_phase2(L, bc, K, s, h);
        WS_STAMP(s.q[2]);   // 4: phase 2 done
        bars.cta();   // [B2] the base role has published R / v0 / pz of the new state
      }
    }
    if (RECORD && active) {
      float* row = A.out_states + (((size_t)(A.paired ? 0 : c) * A.S + seg) * A.H + k) * SPI_STATE_DIM;
#pragma unroll
      for (int j = 0; j < 3; j++) { row[13 + 3 * LEG + j] = s.q[j]; row[25 + 3 * LEG + j] = s.qd[j]; }
    }
  }
  if (RECORD) return;
  // scripts/eval.py:292 — this leg's share of the squared joint-position error
  const float* tgt = A.seg_target + (size_t)seg * SPI_TARGET_DIM;
  float ej = 0.f, acc = 0.f;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const float e = s.q[j] - __ldg(tgt + 7 + 3 * LEG + j);
    ej += e * e;
    acc += s.q[j] * 0.f + s.qd[j] * 0.f;
  }
  sm.ej[LEG][lane] = ej;
  sm.finite[LEG][lane] = finite_acc(acc) ? 1 : 0;
  bars.cta();   // [C]
}

template <bool RECORD>
__device__ __forceinline__ void ws_base_role(const WsArgs& A, WsSmem& sm, const WsBars bars, int lane, int c,
                                             int cta_in_cand, int seg, bool active, bool group_live, int tail_lp) {
  const SimK& S = A.M.sim;
  BaseInertia B;
  {
    float motor_unused[3];
    apply_candidate(A.M, A.params ? A.params + (size_t)c * A.P : nullptr, A.ids, A.flags, B, motor_unused);
  }
  BaseState s;
  {
    const float* row = A.seg_init + (size_t)seg * SPI_STATE_DIM;
#pragma unroll
    for (int i = 0; i < 3; i++) { s.p[i] = *(row + i); s.v[i] = *(row + 7 + i); s.w[i] = *(row + 10 + i); }
#pragma unroll
    for (int i = 0; i < 4; i++) s.quat[i] = *(row + 3 + i);
  }
  float bc[kBaseOut];
#pragma unroll
  for (int i = 0; i < 6; i++) bc[i] = 0.f;
  base_publish(s, bc);
  ws_store_state(sm, lane, bc);
  float pb[6];
  WS_PROF_INIT(4);
  bars.cta();   // [S0] the legs have published their joint state
  const float h = S.dt / (float)S.nsub;
  for (int k = 0; k < A.H; k++) {
    for (int d = 0; d < A.decimation; d++) {
      for (int n = 0; n < S.nsub; n++) {
        // velocity-product bias of this sub-step, while the legs run their phase 1 (v0 comes back through shared memory: a
        // dependency the scheduler cannot hoist in front of [B2], where it would delay the legs' release)
        {                                              // logical bc[15 .. 21) = floats 9 .. 14 of the packed state
          const float4 t2 = ws_lds_volatile(&sm.bc[2][lane]), t3 = ws_lds_volatile(&sm.bc[3][lane]);
          bc[kBcV0] = t2.y; bc[kBcV0 + 1] = t2.z; bc[kBcV0 + 2] = t2.w; bc[kBcV0 + 3] = t3.x; bc[kBcV0 + 4] = t3.y; bc[kBcV0 + 5] = t3.z;
        }
        base_bias(B, bc, pb);
        WS_STAMP(pb[0]);   // 0: bias force done, waiting for [A]
        bars.cta();   // [A]
        float legsum[4 * kLegVec];
#pragma unroll
        for (int v = 0; v < kLegVec; v++) {
          const float4 p01 = sm.part[1][v][lane], p23 = sm.part[3][v][lane];      // pair sums (written by the odd legs)
          legsum[4 * v] = p01.x + p23.x;
          legsum[4 * v + 1] = p01.y + p23.y;
          legsum[4 * v + 2] = p01.z + p23.z;
          legsum[4 * v + 3] = p01.w + p23.w;
        }
        WS_STAMP(legsum[0]);   // 1: [A] released, the pair sums have arrived
        float a0[6];
        base_solve(B, legsum, pb, a0);
        ws_store_a0(sm, lane, a0);
        WS_STAMP(a0[5]);   // 2: solved
        bars.cta();   // [B1] the legs start their acceleration pass
        // keep the integration BEHIND the barrier: it only needs registers, so ptxas would otherwise schedule it between
        // the a0 stores and the barrier and delay the legs by ~90 instructions.  Reading a0 back from shared memory is a
        // dependency the scheduler cannot move across bar.sync (6 LDS, off the critical path).
        {
          const float4 t0 = ws_lds_volatile(&sm.bc[kBcStateVec][lane]), t1 = ws_lds_volatile(&sm.bc[kBcStateVec + 1][lane]);
          a0[0] = t0.x; a0[1] = t0.y; a0[2] = t0.z; a0[3] = t0.w; a0[4] = t1.x; a0[5] = t1.y;
        }
        WS_STAMP(a0[0]);   // 3: [B1] released
        base_advance(S, a0, s, h, bc);
        ws_store_state(sm, lane, bc);
        WS_STAMP(bc[kBcR]);   // 4: advanced
        bars.cta();   // [B2]
      }
    }
    if (RECORD) {
      if (active) {
        float* row = A.out_states + (((size_t)(A.paired ? 0 : c) * A.S + seg) * A.H + k) * SPI_STATE_DIM;
#pragma unroll
        for (int i = 0; i < 3; i++) { row[i] = s.p[i]; row[7 + i] = s.v[i]; row[10 + i] = s.w[i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) row[3 + i] = s.quat[i];
      }
    }
  }
  if (RECORD) return;
  // scripts/eval.py:287-292 — L2 errors of the final state
  const float* tgt = A.seg_target + (size_t)seg * SPI_TARGET_DIM;
  float ep = 0.f, eq = 0.f, acc = 0.f;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float e = s.p[i] - __ldg(tgt + i);
    ep += e * e;
    acc += s.p[i] * 0.f + s.v[i] * 0.f + s.w[i] * 0.f;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float e = s.quat[i] - __ldg(tgt + 3 + i);
    eq += e * e;
    acc += s.quat[i] * 0.f;
  }
  bars.cta();   // [C]
  const float ej = (sm.ej[0][lane] + sm.ej[1][lane]) + (sm.ej[2][lane] + sm.ej[3][lane]);
  const bool ok = finite_acc(acc) && (sm.finite[0][lane] & sm.finite[1][lane] & sm.finite[2][lane] & sm.finite[3][lane]);
  float err[3] = {sqrtf(ep), sqrtf(eq), sqrtf(ej)};
  if (!ok && active) atomicOr(A.bad + c, 1);
  if (A.per_seg && active) {
    float* o = A.per_seg + ((size_t)c * A.S + seg) * 3;
    o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
  }
  const bool counts = active && (A.seg_mask ? (A.seg_mask[seg] != 0) : true);
  if (tail_lp > 0) {
    // tail CTA: group j = lane / L' holds the left-over segments of one candidate.  The butterfly over the L' lanes of a group is the
    // tail of the 32-lane butterfly the padded launch runs on the same values (its first steps only add the zeros of the idle
    // lanes), so the partial sum has the same bits
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float v = counts ? err[i] : 0.f;
      for (int o = tail_lp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      err[i] = v;
    }
    if ((lane & (tail_lp - 1)) == 0 && group_live) {
      float* o = A.partial + ((size_t)c * A.n_cta_per_cand + cta_in_cand) * 3;
      o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float v = counts ? err[i] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    err[i] = v;
  }
  if (lane == 0 && group_live) {
    float* o = A.partial + ((size_t)c * A.n_cta_per_cand + cta_in_cand) * 3;
    o[0] = err[0]; o[1] = err[1]; o[2] = err[2];
  }
}

// blockIdx = one group of 32 rollouts, warps 0..3 = legs, warp 4 = base.
template <bool RECORD, int MOTOR = -1>
__device__ __forceinline__ void rollout_ws_body(const WsArgs& A) {
  __shared__ WsSmem sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int role = __shfl_sync(0xffffffffu, A.rotate_roles ? (warp + blockIdx.x) % kWsWarps : warp, 0);  // warp-uniform
  int cg = (int)(blockIdx.x / A.n_cta_per_cand);
  int cta_in_cand = (int)(blockIdx.x - (unsigned)cg * A.n_cta_per_cand);
  int seg_raw = cta_in_cand * kWsRollouts + lane;
  bool active = seg_raw < A.S, group_live = true;
  int tail_lp = 0;
  if (A.tail_lanes) {                          // dense packing (WsArgs::tail_lanes)
    const int n_full = A.S / kWsRollouts;
    const unsigned first_tail = (unsigned)A.C * (unsigned)n_full;
    if (blockIdx.x < first_tail) {             // one of the candidate's full CTAs
      cg = (int)(blockIdx.x / (unsigned)n_full);
      cta_in_cand = (int)(blockIdx.x - (unsigned)cg * (unsigned)n_full);
      seg_raw = cta_in_cand * kWsRollouts + lane;
      active = true;
    } else {                                   // a tail CTA: 32 / L' candidates, L' lanes each
      tail_lp = A.tail_lanes;
      const int in_group = lane & (tail_lp - 1);
      const int cand = (int)(blockIdx.x - first_tail) * (kWsRollouts / tail_lp) + lane / tail_lp;
      group_live = cand < A.C;
      cg = group_live ? cand : A.C - 1;
      cta_in_cand = n_full;
      seg_raw = n_full * kWsRollouts + in_group;
      active = group_live && seg_raw < A.S;
    }
  }
  const int seg = seg_raw < A.S ? seg_raw : A.S - 1;
  const int c = A.paired ? seg : cg;
  const WsBars bars{};
  if (role < 4) ws_leg_role<RECORD, MOTOR>(A, sm, bars, lane, role, c, seg, active);
  else ws_base_role<RECORD>(A, sm, bars, lane, c, cta_in_cand, seg, active, group_live, tail_lp);
}

template <bool RECORD, int MINB, int MOTOR = -1>
__global__ void __launch_bounds__(kWsThreads, MINB) rollout_ws_kernel(const __grid_constant__ WsArgs A) {
  pdl::trigger();
  pdl::wait();             // (no-ops unless launched with programmatic stream serialisation: spi_b200_env_step)
  rollout_ws_body<RECORD, MOTOR>(A);
}

}  // namespace ws
