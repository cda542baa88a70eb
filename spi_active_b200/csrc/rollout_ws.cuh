// rollout_ws.cuh — the warp-specialised fused rollout kernel (see go2_ws.cuh for the role arithmetic and
// DESIGN.md §4.2 for the mapping).  One CTA = 32 (candidate, segment) rollouts x 5 warps:
//   4 leg warps (lane = rollout) + 1 base warp.  The base role sits on the HIGHEST warp id of the CTA: it is the
//   critical path between the two leg phases and the sub-partition arbiter favours high warp ids (measured:
//   +8.6 % over rotating the roles with blockIdx, profiles/README.md).  5-warp CTAs spread over the 4
//   sub-partitions by themselves.
// Per integrator sub-step:   legs: phase 1 -> smem[27] | barrier | base: sum, 6x6 solve, integrate ->
//   a0 -> smem[6] | barrier | legs: phase 2, overlapped with base: integrate, publish R / v0 / pz -> smem[16],
//   bias force of the next sub-step | barrier.
#pragma once
#include "go2_ws.cuh"

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
  int rotate_roles;
  int token_mode;
  int paired;          // 1: rollout r uses candidate row r AND segment r (one env per rollout; C == 1 for the grid)
  float* partial;      // [C][n_cta_per_cand][3]
  float* per_seg;      // [C][S][3] or null
  int* bad;            // [C]
  float* out_states;   // [C][S][H][37] (RECORD)
  const unsigned char* zero_mask;   // paired mode: [S] 1 = this env's action is replaced by 0 (terminated group), or null
};

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

// Named-barrier plan.  HALVES == 1 (one 32-rollout group per CTA): barrier 1 = the group's 160 threads, 2 / 3 = the leg pairs.
// HALVES == 2 (two groups per CTA, DESIGN.md 4.2 "anti-phase pairing"): group g uses 1 + 3 g, 2 + 3 g, 3 + 3 g, and the two
// groups hand a token back and forth on barriers 7 / 8 (256 = the 8 leg warps) so that their leg phases 1 never overlap:
// resident CTAs of this kernel otherwise fall into lock-step (all legs in phase 1, then all waiting for their base role),
// which left 22 % of the issue slots empty although 2.1 warps per scheduler were eligible on average (profiles/README.md r2).
template <int HALVES> struct WsBars {
  int g;
  int mode;   // experiment switch: 2 = token on every sub-step, 1 = first sub-step only (initial stagger), 0 = none
  __device__ __forceinline__ void cta() const {
    if (HALVES == 1) ws_barrier();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + 3 * g), "n"(kWsThreads) : "memory");
  }
  __device__ __forceinline__ void pair_arrive(int leg) const {
    if (HALVES == 1) {
      if (leg < 2) asm volatile("bar.arrive 2, 64;" ::: "memory");
      else asm volatile("bar.arrive 3, 64;" ::: "memory");
    } else asm volatile("bar.arrive %0, 64;" ::"r"(2 + 3 * g + (leg >> 1)) : "memory");
  }
  __device__ __forceinline__ void pair_sync(int leg) const {
    if (HALVES == 1) {
      if (leg < 2) asm volatile("bar.sync 2, 64;" ::: "memory");
      else asm volatile("bar.sync 3, 64;" ::: "memory");
    } else asm volatile("bar.sync %0, 64;" ::"r"(2 + 3 * g + (leg >> 1)) : "memory");
  }
  // phase-1 token: group 0 owns it first.  wait = before phase 1 (n = index of the sub-step), pass = after phase 1.
  __device__ __forceinline__ void token_wait(int n) const {
    if (HALVES == 1 || mode == 0) return;
    if (mode == 1) { if (g == 1 && n == 0) asm volatile("bar.sync 7, 256;" ::: "memory"); return; }
    if (g == 0) { if (n > 0) asm volatile("bar.sync 8, 256;" ::: "memory"); }
    else asm volatile("bar.sync 7, 256;" ::: "memory");
  }
  __device__ __forceinline__ void token_pass(int n, int n_total) const {
    if (HALVES == 1 || mode == 0) return;
    if (mode == 1) { if (g == 0 && n == 0) asm volatile("bar.arrive 7, 256;" ::: "memory"); return; }
    if (g == 0) asm volatile("bar.arrive 7, 256;" ::: "memory");
    else if (n + 1 < n_total) asm volatile("bar.arrive 8, 256;" ::: "memory");
  }
};

// LEG is a warp-uniform run-time value (one code copy for the four legs: four template copies overflow the
// instruction cache — profiles/README.md, experiment ws-templated); the leg's constants are then fetched through
// the constant bank with a uniform offset.
template <bool RECORD, int HALVES>
__device__ __forceinline__ void ws_leg_role(const WsArgs& A, WsSmem& sm, const WsBars<HALVES> bars, int lane, const int LEG,
                                            int c, int seg, bool active) {
  const SimK& S = A.M.sim;
  const LegK& L = A.M.leg[LEG];
  // this leg's motor parameters (act2tau_scalar uses one gain for every joint)
  float motor[3];
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
    motor[0] = m3[0]; motor[1] = m3[1]; motor[2] = m3[2];
    if (A.motor_model == SPI_MOTOR_SCALAR) { motor[1] = m3[0]; motor[2] = m3[0]; }
  }
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
  bars.cta();   // [S0] the base role has published R / v0 / pz of the initial state
  const int n_sub_total = A.H * A.decimation * S.nsub;
  int i_sub = 0;
  for (int k = 0; k < A.H; k++) {
    float act[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
      act[j] = zero_act ? 0.f : fminf(fmaxf(__ldg(act_row + 12 * k + j), -S.action_clip), S.action_clip);
    for (int d = 0; d < A.decimation; d++) {
      float tau[3];
      leg_torques(S, L, act, s.q, s.qd, kp, kd, motor, A.motor_model, A.flags, tau);
      for (int n = 0; n < S.nsub; n++, i_sub++) {
        bars.token_wait(i_sub);      // HALVES == 2: the other group's legs have finished their phase 1
        float bc[kBaseOut];
        ws_load_state(sm, lane, bc);
        float out[4 * kLegVec];
        out[4 * kLegVec - 1] = 0.f;
        leg_phase1(S, L, bc, s, tau, K, out, nullptr);
        bars.token_pass(i_sub, n_sub_total);
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
          }
        }
        bars.cta();   // [A]  the pair sums are in shared memory
        bars.cta();   // [B1] the base role has published a0
        ws_load_a0(sm, lane, bc + kBcA0);
        leg_phase2(L, bc, K, s, h);
        bars.cta();   // [B2] the base role has published R / v0 / pz of the new state
      }
    }
    if (RECORD) {
      if (active) {
        float* row = A.out_states + (((size_t)(A.paired ? 0 : c) * A.S + seg) * A.H + k) * SPI_STATE_DIM;
#pragma unroll
        for (int j = 0; j < 3; j++) { row[13 + 3 * LEG + j] = s.q[j]; row[25 + 3 * LEG + j] = s.qd[j]; }
      }
      bars.cta();   // [R] keeps the barrier count of the base role (which writes its rows here)
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

template <bool RECORD, int HALVES>
__device__ __forceinline__ void ws_base_role(const WsArgs& A, WsSmem& sm, const WsBars<HALVES> bars, int lane, int c,
                                             int cta_in_cand, int seg, bool active, bool group_live) {
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
  base_bias(B, bc, pb);
  bars.cta();   // [S0]
  const float h = S.dt / (float)S.nsub;
  for (int k = 0; k < A.H; k++) {
    for (int d = 0; d < A.decimation; d++) {
      for (int n = 0; n < S.nsub; n++) {
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
        float a0[6];
        base_solve(B, legsum, pb, a0);
        ws_store_a0(sm, lane, a0);
        bars.cta();   // [B1] the legs start their acceleration pass
        // keep the integration BEHIND the barrier: it only needs registers, so ptxas would otherwise schedule it between
        // the a0 stores and the barrier and delay the legs by ~90 instructions.  Reading a0 back from shared memory is a
        // dependency the scheduler cannot move across bar.sync (6 LDS, off the critical path).
        {
          const float4 t0 = ws_lds_volatile(&sm.bc[kBcStateVec][lane]), t1 = ws_lds_volatile(&sm.bc[kBcStateVec + 1][lane]);
          a0[0] = t0.x; a0[1] = t0.y; a0[2] = t0.z; a0[3] = t0.w; a0[4] = t1.x; a0[5] = t1.y;
        }
        base_advance(S, a0, s, h, bc);
        ws_store_state(sm, lane, bc);
        bars.cta();   // [B2]
        // velocity-product bias of the next sub-step: after the barrier, so that it overlaps the legs' phase 1 instead
        // of delaying their release (re-read through shared memory for the same reason as a0 above)
        {                                              // logical bc[15 .. 21) = floats 9 .. 14 of the packed state
          const float4 t2 = ws_lds_volatile(&sm.bc[2][lane]), t3 = ws_lds_volatile(&sm.bc[3][lane]);
          bc[kBcV0] = t2.y; bc[kBcV0 + 1] = t2.z; bc[kBcV0 + 2] = t2.w; bc[kBcV0 + 3] = t3.x; bc[kBcV0 + 4] = t3.y; bc[kBcV0 + 5] = t3.z;
        }
        base_bias(B, bc, pb);
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
      bars.cta();   // [R]
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

// HALVES == 1: blockIdx = one group of 32 rollouts, warps 0..3 = legs, warp 4 = base.
// HALVES == 2: blockIdx = two consecutive groups; warps 0..3 / 4..7 = the legs of group 0 / 1 (leg i of both groups shares a
// sub-partition, and the phase-1 token makes them take turns on it), warps 8 / 9 = the two base roles (highest warp ids: the
// arbiter favours them).  An odd group count leaves the last CTA's second group without work: it replays the last group with
// every write suppressed (the token protocol needs both groups).
template <bool RECORD, int HALVES>
__device__ __forceinline__ void rollout_ws_body(const WsArgs& A) {
  __shared__ WsSmem sm_all[HALVES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int role, g;
  if (HALVES == 1) {
    role = __shfl_sync(0xffffffffu, A.rotate_roles ? (warp + blockIdx.x) % kWsWarps : warp, 0);  // warp-uniform
    g = 0;
  } else {
    const int w = __shfl_sync(0xffffffffu, warp, 0);
    role = (w < 8) ? (w & 3) : 4;
    g = (w < 8) ? (w >> 2) : (w - 8);
  }
  const long long n_groups = (long long)A.C_grid * A.n_cta_per_cand;
  long long grp = (long long)blockIdx.x * HALVES + g;
  const bool group_live = grp < n_groups;
  if (!group_live) grp = n_groups - 1;
  const int cg = (int)(grp / A.n_cta_per_cand);
  const int cta_in_cand = (int)(grp - (long long)cg * A.n_cta_per_cand);
  const int seg_raw = cta_in_cand * kWsRollouts + lane;
  const bool active = group_live && seg_raw < A.S;
  const int seg = seg_raw < A.S ? seg_raw : A.S - 1;
  const int c = A.paired ? seg : cg;
  WsSmem& sm = sm_all[g];
  const WsBars<HALVES> bars{g, A.token_mode};
  if (role < 4) ws_leg_role<RECORD, HALVES>(A, sm, bars, lane, role, c, seg, active);
  else ws_base_role<RECORD, HALVES>(A, sm, bars, lane, c, cta_in_cand, seg, active, group_live);
}

template <bool RECORD, int MINB>
__global__ void __launch_bounds__(kWsThreads, MINB) rollout_ws_kernel(const __grid_constant__ WsArgs A) {
  rollout_ws_body<RECORD, 1>(A);
}

// two anti-phased groups per CTA (evaluation only; the RECORD / paired launches are latency-bound single waves)
template <int MINB>
__global__ void __launch_bounds__(2 * kWsThreads, MINB) rollout_ws2_kernel(const __grid_constant__ WsArgs A) {
  rollout_ws_body<false, 2>(A);
}

}  // namespace ws
